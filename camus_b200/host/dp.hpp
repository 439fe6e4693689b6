// Host-side consumer of the hot path: scorers, the level-1 network DP, traceback and the
// extended-newick writer. These stay on the host in the drop-in (SURVEY.md §2 "OUT OF SCOPE",
// §8(f) rank 1); they are restated here in C++ only so that results can be carried through to
// the networks the reference's golden files hold.
//
// Mirrors: internal/score/scorers.go:47-144, internal/infer/main_dp.go:40-291,
// internal/infer/helpers.go:9-110, internal/infer/traceback.go:6-57, internal/graphs/net.go:38-101.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "treedata.hpp"

namespace camus {

// ---- scorer options (score/scorers.go:12-38) ----
enum class ScoreMode : int { Max = 0, Norm = 1, Sym = 2 };
struct ScorerOpts {  // scorers.go:25-29
  int nGTrees = 0;
  double alpha = 0;
  bool asSet = false;
};
using ScoreOption = std::function<void(ScorerOpts&)>;
inline ScoreOption AsSet(bool v) {
  return [v](ScorerOpts& o) { o.asSet = v; };
}
inline ScoreOption WithNGtrees(int n) {
  return [n](ScorerOpts& o) {
    if (n <= 0) throw CamusError(Err::InvalidScorerOption, "invalid scorer option, number of gene trees must be positive, but is " + std::to_string(n));
    o.nGTrees = n;
  };
}
inline ScoreOption WithAlpha(double a) {
  return [a](ScorerOpts& o) {
    if (a <= 0 || a > 1)
      throw CamusError(Err::InvalidScorerOption, "invalid scorer option, alpha must be in greater than zero and less than or equal to one, but is " + std::to_string(a));
    o.alpha = a;
  };
}

// score.ParseScorer (scorers.go:12-16)
inline bool ParseScorer(const std::string& s, ScoreMode& m) {
  if (s == "max") m = ScoreMode::Max;
  else if (s == "norm") m = ScoreMode::Norm;
  else if (s == "sym") m = ScoreMode::Sym;
  else return false;
  return true;
}

struct Branch {
  int u = 0, w = 0;
  bool collide(const Branch& o) const { return u == o.u || u == o.w || w == o.u || w == o.w; }
};

// ------------------------------------------------------------------------------------------
// Scorers (scorers.go). They own the two integer tables produced by the hot path and turn
// them into the DP's score type. All float math is here, on the host, exactly as in Go.
// ------------------------------------------------------------------------------------------
struct ScoreTables {
  int n = 0;
  std::vector<uint64_t> totals;     // [u*n + w]  quartettotals.go:19-22
  std::vector<uint64_t> penalties;  // [u*n + w]  penalty.go:11-29 (norm / sym only)
  bool as_set = false;
  uint64_t n_survivors = 0;  // len(qCounts)                 treedata.go:305-307
  uint64_t sum_counts = 0;   // exact sum; truncated to u32  treedata.go:297-303
};

struct MaximizeScorer {
  using S = uint64_t;
  const ScoreTables* t;
  S calc(int u, int w) const { return t->totals[(size_t)u * t->n + w]; }  // scorers.go:62-64
};
struct NormalizedScorer {
  using S = double;
  const ScoreTables* t;
  int n_gtree;
  S calc(int u, int w) const {  // scorers.go:101-103
    size_t i = (size_t)u * t->n + w;
    return double(t->totals[i]) / (double(n_gtree) * double(t->penalties[i]));
  }
};
struct SymDiffScorer {
  using S = double;
  const ScoreTables* t;
  int n_gtree;  // the reference never sets it (infer.go:84-85) -> 0
  double alpha;
  S calc(int u, int w) const {  // scorers.go:142-144
    size_t i = (size_t)u * t->n + w;
    return 2 * double(t->totals[i]) - alpha * double(t->penalties[i]) * double(n_gtree);
  }
};

// QuartetTotals.PercentQuartetSat (quartettotals.go:25-40), with the uint32 truncation of
// TotalNumQuartets / TotalNumUniqueQuartets (treedata.go:297-307).
inline double percent_quartet_sat(const ScoreTables& t, const std::vector<Branch>& branches) {
  uint64_t sum = 0;
  for (const Branch& b : branches) sum += t.totals[(size_t)b.u * t.n + b.w];
  uint32_t denom = t.as_set ? (uint32_t)t.n_survivors : (uint32_t)t.sum_counts;
  return 100 * double(sum) / double(denom);
}

// ------------------------------------------------------------------------------------------
// helpers.go
// ------------------------------------------------------------------------------------------
template <class S>
inline bool best_split(const std::vector<S>& l, const std::vector<S>& r, int k, int& kl, int& kr) {
  if ((int)l.size() + (int)r.size() - 2 < k) return false;
  S best{};
  bool found = false;
  kl = kr = -1;
  int lo = std::max(0, k - ((int)r.size() - 1)), hi = std::min(k, (int)l.size() - 1);
  for (int i = lo; i <= hi; ++i) {
    S cur = l[i] + r[k - i];
    if (cur > best || !found) {
      best = cur;
      kl = i;
      kr = k - i;
      found = true;
    }
  }
  return true;
}

template <class S>
inline void solve_all_splits(const std::vector<S>& a, const std::vector<S>& b, int k,
                             std::vector<std::array<int, 2>>& sol, std::vector<S>& sc) {
  sol.clear();
  sc.clear();
  for (int i = 0; i <= k; ++i) {
    int l, r;
    if (!best_split(a, b, i, l, r)) break;
    sol.push_back({l, r});
    sc.push_back(a[l] + b[r]);
  }
}

template <class S>
inline bool four_way_best_split(const std::vector<S>* lists[4], int k, int idx[4]) {
  int combined = 0;
  for (int i = 0; i < 4; ++i) combined += (int)lists[i]->size();
  if (combined - 4 < k) return false;
  thread_local std::vector<std::array<int, 2>> s1, s2;  // scratch, reused across the ~1e9 calls
  thread_local std::vector<S> c1, c2;
  solve_all_splits(*lists[0], *lists[1], k, s1, c1);
  solve_all_splits(*lists[2], *lists[3], k, s2, c2);
  int i1, i2;
  if (!best_split(c1, c2, k, i1, i2)) throw CamusError(Err::Internal, "did not find expected split");
  idx[0] = s1[i1][0];
  idx[1] = s1[i1][1];
  idx[2] = s2[i2][0];
  idx[3] = s2[i2][1];
  return true;
}

// ------------------------------------------------------------------------------------------
// traceback.go, as an index arena instead of Go pointers. A "trace ref" names the entry
// Traceback[v][k]; Go's `&dp.Traceback[v][k]` points at exactly that (immutable) slot.
// ------------------------------------------------------------------------------------------
struct TraceArena {
  struct Trace {
    bool cycle = false;
    int prev0 = -1, prev1 = -1;    // noCycleTrace.prevs
    int pathW = -1, pathU = -1;    // cycleTraceNode ids
    int wDown = -1, uDown = -1;    // trace ids
    Branch branch;
  };
  struct CTN {
    int sib = -1;  // trace id
    int p = -1;    // CTN id
  };
  std::vector<Trace> traces;
  std::vector<CTN> ctns;

  int add(const Trace& t) {
    traces.push_back(t);
    return (int)traces.size() - 1;
  }
  int add_ctn(int p, int sib) {
    ctns.push_back({sib, p});
    return (int)ctns.size() - 1;
  }
  void traceback(int id, std::vector<Branch>& out) const {
    const Trace& t = traces[id];
    if (!t.cycle) {  // noCycleTrace.traceback
      if (t.prev0 < 0) return;
      traceback(t.prev0, out);
      traceback(t.prev1, out);
      return;
    }
    traceback(t.wDown, out);  // cycleTrace.traceback
    out.push_back(t.branch);
    if (t.uDown >= 0) traceback(t.uDown, out);
    if (t.pathU >= 0) trace_up(t.pathU, out);
    if (t.pathW >= 0) trace_up(t.pathW, out);
  }
  void trace_up(int c, std::vector<Branch>& out) const {
    traceback(ctns[c].sib, out);
    if (ctns[c].p >= 0) trace_up(ctns[c].p, out);
  }
};

// Minimal persistent pool for the DP's widest loop (scoreEdgesAcross over all u). The reference
// runs its DP on one goroutine; results here are combined in the reference's iteration order, so
// the thread count never changes a tie-break.
class WorkerPool {
 public:
  explicit WorkerPool(int n) {
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size() + 1; }
  // f(item, worker) for item in [0, n); dynamic scheduling; returns when all items are done
  void run(int n, const std::function<void(int, int)>& f) {
    if (workers_.empty() || n <= 1) {
      for (int i = 0; i < n; ++i) f(i, 0);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &f;
      n_ = n;
      next_.store(0);
      pending_ = (int)workers_.size();
      ++gen_;
    }
    cv_.notify_all();
    for (int i; (i = next_.fetch_add(1)) < n;) f(i, 0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop() {
    static std::atomic<int> ids{1};
    const int me = ids.fetch_add(1);
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int, int)>* f;
      int n;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        f = fn_;
        n = n_;
      }
      if (f)
        for (int i; (i = next_.fetch_add(1)) < n;) (*f)(i, me);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int, int)>* fn_ = nullptr;
  int n_ = 0, pending_ = 0;
  std::atomic<int> next_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// val[i] = max(val[i], c[i] + cu): the inner loop of the integer across-scan below. Cloned per ISA
// and dispatched at load time (the library is built on one machine and runs on another).
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default"), optimize("O3")))
#endif
static void maxplus_accumulate(uint64_t* __restrict__ val, const uint64_t* __restrict__ c, uint64_t cu, int n) {
  for (int i = 0; i < n; ++i) {
    const uint64_t t = c[i] + cu;
    val[i] = t > val[i] ? t : val[i];
  }
}

// Optional GPU offload of the across-scan (include/camus_b200.h, "DP offload"): plain C function
// pointers into libcamus_b200.so, so that this library does not link CUDA. Field order = the order
// camus_host_run_dp_accel documents.
struct DpAccelApi {
  int (*create)(void** out, int device, int score_is_f64, int32_t n_nodes, const int32_t* subtree_end, const int32_t* depth, const void* edge_scores);
  void (*destroy)(void* dp);
  const char* (*last_error)(const void* dp);
  int (*set_node)(void* dp, int32_t node, int32_t len, const void* scores);
  int (*set_vertex)(void* dp, int32_t v, int32_t n_items, const int32_t* item_u, const int32_t* item_sub);
  int (*set_path_columns)(void* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols);
  int (*across)(void* dp, int32_t prev_k, void* best);
  uint64_t (*launches)(const void* dp);
  int (*set_conv_columns)(void* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols);
  int (*across_maxplus)(void* dp, int32_t prev_k, void* best);
};
struct DpAccelBest {  // = camus_b200_dp_best
  int32_t found, nan_seen;
  uint64_t score_bits;
  int32_t u, w, cycle_length;
  int32_t split[4];
  int32_t first_is_nan, first_u, first_w;
  int32_t first_split[4];
  uint64_t first_score_bits;
};
static_assert(sizeof(DpAccelBest) == 80, "layout of camus_b200_dp_best");
struct DpAccel {
  const DpAccelApi* api = nullptr;
  void* handle = nullptr;
  uint64_t min_work = 0;   // steps below this many (pairs x list entries) stay on the host
  bool verify = false;     // CAMUS_DP_VERIFY=1: run the host scan as well and compare the winner
  uint64_t steps = 0, steps_host = 0, nan_fallbacks = 0;
  double s_columns = 0, s_across = 0;  // seconds inside set_vertex + set_path_columns / inside across
  double s_create = 0;                 // edge table + camus_b200_dp_create
  void check(int rc, const char* what) const {
    if (rc != 0) throw CamusError(Err::Internal, std::string("DP offload: ") + what + ": " + (api->last_error(handle) ? api->last_error(handle) : "?"));
  }
};

struct DPResults {
  std::vector<std::vector<Branch>> branches;  // per k = 1..m
  std::vector<double> qsat;
  std::vector<double> root_scores_f;  // DP[root][k] as double (for logging / tests)
  std::vector<uint64_t> root_scores_u;
  uint64_t accel_steps = 0, accel_steps_host = 0, accel_nan_fallbacks = 0, accel_launches = 0;
  double accel_s_columns = 0, accel_s_across = 0, accel_s_create = 0, accel_s_destroy = 0;
};

// ------------------------------------------------------------------------------------------
// main_dp.go
// ------------------------------------------------------------------------------------------
template <class Scorer>
class DP {
  using S = typename Scorer::S;
  // best (u -> w) of one across-scan as a plain value: score, nodes, the four split indices
  struct Desc {
    bool ok = false;
    S score{};
    int u = 0, w = 0;
    int idx[4] = {0, 0, 0, 0};
  };

 public:
  DP(const TreeData& td, const Scorer& sc, int nthreads = 1, DpAccel* accel = nullptr)
      : td_(td), sc_(sc), n_(td.n()), dp_(n_), tb_(n_), pool_(std::max(1, nthreads)), accel_(accel) {}

  void run() {  // RunDP, main_dp.go:90-104 (gotree PostOrder = children in order, then node)
    std::vector<int> order;
    order.reserve(n_);
    post_order(0, order);
    for (int v : order) {
      if (!td_.tip(v)) {
        solve(v);
      } else {
        dp_[v].assign(1, S{});
        TraceArena::Trace t;
        tb_[v].assign(1, arena_.add(t));
      }
    }
  }

  ~DP() {
    if (profile_)
      fprintf(stderr, "[dp] cycleDP.update %.3f s, scoreEdgesDown %.3f s, item lists %.3f s, across scan %.3f s, acceptance %.3f s\n", t_update_, t_down_,
              t_items_, t_scan_, t_accept_);
  }
  const std::vector<S>& root_scores() const { return dp_[0]; }
  std::vector<Branch> traceback(int k) const {
    std::vector<Branch> out;
    arena_.traceback(tb_[0][k], out);
    return out;
  }

 private:
  const TreeData& td_;
  const Scorer& sc_;
  int n_;
  std::vector<std::vector<S>> dp_;
  std::vector<std::vector<int>> tb_;
  TraceArena arena_;
  WorkerPool pool_;
  DpAccel* accel_ = nullptr;
  // CAMUS_DP_PROFILE=1: seconds per part of the DP, printed by the destructor
  bool profile_ = getenv("CAMUS_DP_PROFILE") && getenv("CAMUS_DP_PROFILE")[0] == '1';
  double t_update_ = 0, t_down_ = 0, t_items_ = 0, t_scan_ = 0, t_accept_ = 0, t_noedge_ = 0;
  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  int accel_v_ = -1, accel_cols_ = 0;  // vertex whose items are on the device, path columns already there
  int accel_conv_cols_ = 0;            // max-plus form: conv columns already there
  // CAMUS_DP_ACCEL_LITERAL=1: integer scores take the literal device scan too (A/B, tests)
  bool accel_literal_ = getenv("CAMUS_DP_ACCEL_LITERAL") && getenv("CAMUS_DP_ACCEL_LITERAL")[0] == '1';
  std::vector<uint64_t> col_buf_;
  // CAMUS_DP_LITERAL=1: literal FourWayBestSplit scan for integer scores too (A/B, tests)
  bool fast_enabled_ = !(getenv("CAMUS_DP_LITERAL") && getenv("CAMUS_DP_LITERAL")[0] == '1');
  // cycleDP of the vertex being solved (main_dp.go:30-34)
  std::vector<std::vector<S>> cs_;
  std::vector<std::vector<int>> ct_;
  std::vector<std::pair<int, int>> items_;  // (u, other subtree) in the reference's visiting order
  int items_v_ = -1;                        // vertex items_ was built for
  std::vector<Desc> descs_;
  // integer scores only: conv_[j][x] = max over i + i' = j of cs_[x][i] + dp_[x][i'] (see fast_across)
  std::vector<std::vector<S>> conv_;
  static constexpr bool kIntegerScores = std::is_same_v<S, uint64_t>;

  void post_order(int v, std::vector<int>& out) const {
    std::vector<std::pair<int, int>> st{{v, 0}};
    while (!st.empty()) {
      int x = st.back().first, i = st.back().second;
      if (td_.tip(x) || i == 2) {
        out.push_back(x);
        st.pop_back();
      } else {
        st.back().second = i + 1;
        st.push_back({td_.children[x][i], 0});
      }
    }
  }
  template <class F>
  void subtree_pre(int v, F&& f) const {  // helpers.go:37-47
    int end = td_.subtree_end[v];
    for (int x = v; x < end; ++x) f(x);  // ids are preorder, so the subtree is an id interval
  }
  template <class F>
  void subtree_post_helper(int cur, int other, F&& f) const {  // helpers.go:26-35
    std::vector<int> order;
    post_order(cur, order);
    for (int x : order) f(x, other);
  }

  void cdp_update(int v, int prevK) {  // main_dp.go:40-72
    subtree_pre(v, [&](int cur) {
      if (prevK == 0) {
        cs_[cur].clear();
        ct_[cur].clear();
      }
      cs_[cur].push_back(S{});
      ct_[cur].push_back(-1);
      if (cur == v) return;
      int p = td_.parent[cur];
      if (p == v) return;
      int sib = td_.sibling(cur);
      int pK, sK;
      if (!best_split(cs_[p], dp_[sib], prevK, pK, sK)) return;
      cs_[cur][prevK] = cs_[p][pK] + dp_[sib][sK];
      ct_[cur][prevK] = arena_.add_ctn(ct_[p][pK], tb_[sib][sK]);
    });
  }

  struct Cand {
    bool ok = false;
    S score{};
    int trace = -1;
    Branch br;
  };

  Cand score_edges_down(int v, int prevK) {  // main_dp.go:219-244
    Cand best;
    subtree_pre(v, [&](int w) {
      if (!should_calc_edge(v, w, td_)) return;
      S edge = sc_.calc(v, w);
      int wp, wd;
      if (!best_split(cs_[w], dp_[w], prevK, wp, wd)) return;
      S score = edge + cs_[w][wp] + dp_[w][wd];
      if (score > best.score || !best.ok) {
        TraceArena::Trace t;
        t.cycle = true;
        t.pathW = ct_[w][wp];
        t.wDown = tb_[w][wd];
        t.branch = {v, w};
        best.trace = arena_.add(t);
        best.br = t.branch;
        best.score = score;
        best.ok = true;
      }
    });
    return best;
  }

  // scoreEdgesAcross (main_dp.go:247-287) without side effects: the best w for this u as a plain
  // descriptor; the trace is materialised by the caller only for a candidate it accepts.
  Desc score_edges_across(int u, int sub, int prevK) const {
    Desc best;
    const int end = td_.subtree_end[sub];
    for (int w = sub; w < end; ++w) {  // SubtreePreOrder(sub): ids are preorder
      S edge = sc_.calc(u, w);
      const std::vector<S>* lists[4] = {&cs_[w], &cs_[u], &dp_[w], &dp_[u]};
      int idx[4];
      if (!four_way_best_split(lists, prevK, idx)) continue;
      S score = edge + cs_[w][idx[0]] + cs_[u][idx[1]] + dp_[w][idx[2]] + dp_[u][idx[3]];
      if (score > best.score || !best.ok) {
        best.ok = true;
        best.score = score;
        best.u = u;
        best.w = w;
        for (int i = 0; i < 4; ++i) best.idx[i] = idx[i];
      }
    }
    return best;
  }
  // ---- integer scores: the across-scan as a max-plus product -----------------------------------
  // FourWayBestSplit maximises cs_w[a] + cs_u[b] + dp_w[c] + dp_u[d] over a + b + c + d = prevK,
  // pairing (cs_w, cs_u) and (dp_w, dp_u) and re-solving every split for every (u, w, k): O(k^2)
  // per pair and step. Max-plus convolution is associative, so for exact (integer) scores the
  // same maximum is max_j conv_w[j] + conv_u[prevK - j] with the per-node conv_x = cs_x (*) dp_x,
  // of which every step only adds the entry j = prevK: O(k) per pair and step, vectorised over w.
  // Only the VALUE is taken from this form; the split indices of a candidate the replay accepts
  // come from the reference's own FourWayBestSplit (materialise), so tie-breaks are untouched.
  // Float scorers keep the literal scan: their sums are order dependent.
  void conv_update(int v, int prevK) {
    if ((int)conv_.size() <= prevK) conv_.resize(prevK + 1);
    std::vector<S>& c = conv_[prevK];
    if ((int)c.size() != n_) c.assign(n_, S{});
    const int end = td_.subtree_end[v];
    for (int x = v; x < end; ++x) {
      const std::vector<S>&a = cs_[x], &b = dp_[x];
      const int lo = std::max(0, prevK - ((int)b.size() - 1)), hi = std::min(prevK, (int)a.size() - 1);
      S best{};
      for (int i = lo; i <= hi; ++i) best = std::max(best, a[i] + b[prevK - i]);
      c[x] = best;
    }
  }
  Desc fast_across(int u, int sub, int prevK) const {
    Desc best;
    const int end = td_.subtree_end[sub], len = end - sub;
    thread_local std::vector<S> val;
    val.assign(len, S{});
    for (int j = 0; j <= prevK; ++j)
      maxplus_accumulate(reinterpret_cast<uint64_t*>(val.data()), reinterpret_cast<const uint64_t*>(conv_[j].data() + sub),
                         (uint64_t)conv_[prevK - j][u], len);
    for (int i = 0; i < len; ++i) {
      const S score = sc_.calc(u, sub + i) + val[i];
      if (score > best.score || !best.ok) {
        best.ok = true;
        best.score = score;
        best.u = u;
        best.w = sub + i;
      }
    }
    best.idx[0] = -1;  // split indices are resolved on acceptance
    return best;
  }

  Cand materialise(Desc d) {
    Cand c;
    if (!d.ok) return c;
    if (d.idx[0] < 0) {
      const std::vector<S>* lists[4] = {&cs_[d.w], &cs_[d.u], &dp_[d.w], &dp_[d.u]};
      const int prevK = (int)cs_[d.w].size() - 1;
      if (!four_way_best_split(lists, prevK, d.idx)) throw CamusError(Err::Internal, "fast across-scan: infeasible split");
      if (sc_.calc(d.u, d.w) + cs_[d.w][d.idx[0]] + cs_[d.u][d.idx[1]] + dp_[d.w][d.idx[2]] + dp_[d.u][d.idx[3]] != d.score)
        throw CamusError(Err::Internal, "fast across-scan: score mismatch");
    }
    TraceArena::Trace t;
    t.cycle = true;
    t.pathW = ct_[d.w][d.idx[0]];
    t.pathU = ct_[d.u][d.idx[1]];
    t.wDown = tb_[d.w][d.idx[2]];
    t.uDown = tb_[d.u][d.idx[3]];
    t.branch = {d.u, d.w};
    c.trace = arena_.add(t);
    c.br = t.branch;
    c.score = d.score;
    c.ok = true;
    return c;
  }

  // ---- GPU offload of the across-scan (DpAccel) ---------------------------------------------------
  static uint64_t raw_of(S v) {
    uint64_t r;
    static_assert(sizeof(S) == 8, "8-byte scores");
    memcpy(&r, &v, 8);
    return r;
  }
  static S score_of(uint64_t r) {
    S v;
    memcpy(&v, &r, 8);
    return v;
  }
  // d = the across-scan's winner over the candidates whose score is not NaN (d.ok = false: none);
  // first = the first u's candidate when it scores NaN (include/camus_b200.h, camus_b200_dp_across)
  void accel_across(int v, int prevK, Desc& d, Desc& first) {
    DpAccel& a = *accel_;
    const auto t0 = std::chrono::steady_clock::now();
    const int end = td_.subtree_end[v], cnt = end - v;
    if (accel_v_ != v) {
      std::vector<int32_t> iu(items_.size()), is(items_.size());
      for (size_t i = 0; i < items_.size(); ++i) {
        iu[i] = items_[i].first;
        is[i] = items_[i].second;
      }
      a.check(a.api->set_vertex(a.handle, v, (int32_t)items_.size(), iu.data(), is.data()), "set_vertex");
      accel_v_ = v;
      accel_cols_ = 0;
      accel_conv_cols_ = 0;
    }
    bool maxplus = false;
    if constexpr (kIntegerScores) maxplus = fast_enabled_ && !accel_literal_ && a.api->across_maxplus != nullptr;
    if (maxplus) {
      // integer scores: conv columns instead of path columns (conv_update has run for every step of this vertex)
      if (accel_conv_cols_ <= prevK) {
        const int k0 = accel_conv_cols_, k1 = prevK + 1;
        col_buf_.resize((size_t)(k1 - k0) * cnt);
        for (int j = k0; j < k1; ++j)
          for (int x = v; x < end; ++x) col_buf_[(size_t)(j - k0) * cnt + (x - v)] = raw_of(conv_[j][x]);
        a.check(a.api->set_conv_columns(a.handle, k0, k1, v, cnt, col_buf_.data()), "set_conv_columns");
        accel_conv_cols_ = k1;
      }
    } else if (accel_cols_ <= prevK) {
      const int k0 = accel_cols_, k1 = prevK + 1;
      col_buf_.resize((size_t)(k1 - k0) * cnt);
      for (int j = k0; j < k1; ++j)
        for (int x = v; x < end; ++x) col_buf_[(size_t)(j - k0) * cnt + (x - v)] = raw_of(cs_[x][j]);
      a.check(a.api->set_path_columns(a.handle, k0, k1, v, cnt, col_buf_.data()), "set_path_columns");
      accel_cols_ = k1;
    }
    DpAccelBest b;
    const auto t1 = std::chrono::steady_clock::now();
    if (maxplus) a.check(a.api->across_maxplus(a.handle, prevK, &b), "across_maxplus");
    else a.check(a.api->across(a.handle, prevK, &b), "across");
    const auto t2 = std::chrono::steady_clock::now();
    a.s_columns += std::chrono::duration<double>(t1 - t0).count();
    a.s_across += std::chrono::duration<double>(t2 - t1).count();
    ++a.steps;
    if (b.nan_seen) ++a.nan_fallbacks;  // statistic: steps in which some pair scored NaN
    d = Desc{};
    d.ok = b.found != 0;
    if (d.ok) {
      d.score = score_of(b.score_bits);
      d.u = b.u;
      d.w = b.w;
      for (int i = 0; i < 4; ++i) d.idx[i] = b.split[i];
    }
    first = Desc{};
    first.ok = b.first_is_nan != 0;
    if (first.ok) {
      first.score = score_of(b.first_score_bits);
      first.u = b.first_u;
      first.w = b.first_w;
      for (int i = 0; i < 4; ++i) first.idx[i] = b.first_split[i];
    }
  }
  // CAMUS_DP_VERIFY=1: the literal host scan over the same items, folded in the reference's order
  // from the same starting candidate; the offloaded step must end with the same one
  Cand host_scan_from(Cand best, int bestLen, int prevK) {
    for (const auto& it : items_) {
      Desc c = score_edges_across(it.first, it.second, prevK);
      if (!c.ok) continue;
      int len = cycle_length(c.u, c.w, td_);
      if (c.score > best.score || !best.ok || (c.score == best.score && len <= bestLen)) {
        best = materialise(c);
        bestLen = len;
      }
    }
    return best;
  }
  void verify_same(const Cand& h, const Cand& g, int v, int prevK) const {
    bool same = h.ok == g.ok;
    if (same && h.ok) {
      const auto &a = arena_.traces[h.trace], &b = arena_.traces[g.trace];
      same = raw_of(h.score) == raw_of(g.score) && h.br.u == g.br.u && h.br.w == g.br.w && a.pathW == b.pathW && a.pathU == b.pathU &&
             a.wDown == b.wDown && a.uDown == b.uDown;
    }
    if (!same)
      throw CamusError(Err::Internal, "DP offload: device and host across-scan disagree at v " + std::to_string(v) + " prevK " +
                                          std::to_string(prevK) + ": host (" + std::to_string(h.br.u) + "," + std::to_string(h.br.w) +
                                          ") device (" + std::to_string(g.br.u) + "," + std::to_string(g.br.w) + ")");
  }

  Cand score_add_edge_k(int v, int k) {  // main_dp.go:178-216
    int prevK = k - 1;
    Cand best;
    int bestLen = 0;
    double t0 = now();
    cdp_update(v, prevK);
    t_update_ += now() - t0;
    auto consider = [&](const Cand& c) {
      if (!c.ok) return;
      int len = cycle_length(c.br.u, c.br.w, td_);
      if (c.score > best.score || !best.ok || (c.score == best.score && len <= bestLen)) {
        best = c;
        bestLen = len;
      }
    };
    t0 = now();
    for (int c : td_.children[v]) {
      if (td_.tip(c)) continue;
      consider(score_edges_down(v, prevK));
    }
    t_down_ += now() - t0;
    t0 = now();
    // SubtreePostOrder(v, ...) (helpers.go:9-24): every u of the first child's subtree in post-order
    // against the second child's subtree, then the other way round. The per-u scans are
    // independent, so they run on the pool; acceptance below replays the reference's order.
    int c0 = td_.children[v][0], c1 = td_.children[v][1];
    if (items_v_ != v) {  // the item list depends on v only
      items_v_ = v;
      items_.clear();
      std::vector<int> order;
      post_order(c0, order);
      for (int u : order) items_.push_back({u, c1});
      order.clear();
      post_order(c1, order);
      for (int u : order) items_.push_back({u, c0});
    }
    descs_.assign(items_.size(), Desc{});
    t_items_ += now() - t0;
    t0 = now();
    const size_t work = (size_t)(td_.subtree_end[c0] - c0) * (size_t)(td_.subtree_end[c1] - c1) * (size_t)(prevK + 1);
    bool fast = false;
    if constexpr (kIntegerScores) fast = fast_enabled_;
    if (fast) conv_update(v, prevK);  // every step of the vertex: the device's max-plus form and the host's both read it
    if (accel_ && !items_.empty()) {
      // float scores pay (k + 1)^2 additions per pair on the host, integer ones (max-plus form) k + 1
      const uint64_t cost = kIntegerScores && fast_enabled_ ? (uint64_t)work : (uint64_t)work * (uint64_t)(prevK + 1);
      if (cost >= accel_->min_work) {
        Desc d, first;
        accel_across(v, prevK, d, first);
        Cand host_best;
        if (accel_->verify) host_best = host_scan_from(best, bestLen, prevK);
        if (!best.ok && first.ok) {
          // the first u's candidate scores NaN and nothing was accepted before it: it is taken
          // (bestCycleTrace == nil) and no later `>` or `==` against a NaN replaces it
          best = materialise(first);
        } else if (d.ok) {
          int len = cycle_length(d.u, d.w, td_);
          if (d.score > best.score || !best.ok || (d.score == best.score && len <= bestLen)) {
            best = materialise(d);
            bestLen = len;
          }
        }
        if (accel_->verify) verify_same(host_best, best, v, prevK);
        t_scan_ += now() - t0;
        return best;
      } else {
        ++accel_->steps_host;
      }
    }
    auto body = [&](int i, int) {
      if constexpr (kIntegerScores) {
        if (fast) {
          descs_[i] = fast_across(items_[i].first, items_[i].second, prevK);
          return;
        }
      }
      descs_[i] = score_edges_across(items_[i].first, items_[i].second, prevK);
    };
    if (work > 20000 && pool_.size() > 1) {
      pool_.run((int)items_.size(), body);
    } else {
      for (int i = 0; i < (int)items_.size(); ++i) body(i, 0);
    }
    t_scan_ += now() - t0;
    t0 = now();
    // same acceptance test as consider(), folded over plain descriptors: only the candidate the loop ENDS
    // with gets its split indices and trace (the reference allocates one per acceptance; the intermediate
    // ones are unreachable garbage). CycleLength (an LCA walk) only where the test needs it.
    const Desc* win = nullptr;
    S top = best.score;
    bool have = best.ok;
    for (const Desc& d : descs_) {
      if (!d.ok) continue;
      const bool better = d.score > top || !have;
      if (!better && !(d.score == top)) continue;
      int len = cycle_length(d.u, d.w, td_);
      if (better || len <= bestLen) {
        win = &d;
        top = d.score;
        have = true;
        bestLen = len;
      }
    }
    if (win) best = materialise(*win);
    t_accept_ += now() - t0;
    return best;
  }

  void solve(int v) {  // main_dp.go:130-166
    int l = td_.children[v][0], r = td_.children[v][1];
    std::vector<S> scores{dp_[l][0] + dp_[r][0]};
    TraceArena::Trace t0;
    t0.prev0 = tb_[l][0];
    t0.prev1 = tb_[r][0];
    std::vector<int> traces{arena_.add(t0)};
    if (cs_.empty()) {
      cs_.resize(n_);
      ct_.resize(n_);
    }
    for (int k = 1;; ++k) {
      S score{};
      int bt = -1;
      int lK, rK;
      if (best_split(dp_[l], dp_[r], k, lK, rK)) {  // scoreNoAddEdgeK, main_dp.go:169-174
        score = dp_[l][lK] + dp_[r][rK];
        TraceArena::Trace t;
        t.prev0 = tb_[l][lK];
        t.prev1 = tb_[r][rK];
        bt = arena_.add(t);
      }
      Cand e = score_add_edge_k(v, k);
      if (e.ok && e.score > score) {
        score = e.score;
        bt = e.trace;
      }
      if (bt < 0 || scores[k - 1] >= score) break;
      scores.push_back(score);
      traces.push_back(bt);
      if (k == n_ * n_) throw CamusError(Err::Internal, "runaway loop");
    }
    dp_[v] = std::move(scores);
    tb_[v] = std::move(traces);
    if (accel_) {
      static_assert(sizeof(S) == 8, "8-byte scores");
      accel_->check(accel_->api->set_node(accel_->handle, v, (int32_t)dp_[v].size(), dp_[v].data()), "set_node");
    }
  }
};

// ------------------------------------------------------------------------------------------
// graphs.MakeNetwork + Network.Newick (net.go:38-101). The branch order (and so the #Hk
// numbering) comes from Go's slices.SortFunc with a comparator that is not a total order; for
// <= 12 branches Go runs a plain insertion sort, restated here. Above 12 Go switches to pdqsort
// and the numbering is NOT reproduced (SURVEY.md §8c caution) -- the branch set still is.
// ------------------------------------------------------------------------------------------
inline std::string make_network_newick(const TreeData& td, std::vector<Branch> branches) {
  auto cmp = [&](const Branch& a, const Branch& b) {
    if (a.collide(b)) {
      if (td.under(a.u, b.u) || td.under(a.u, b.w) || td.under(a.w, b.u) || td.under(a.w, b.w)) return -1;
      return 1;
    }
    return 0;
  };
  for (size_t i = 1; i < branches.size(); ++i)
    for (size_t j = i; j > 0 && cmp(branches[j], branches[j - 1]) < 0; --j) std::swap(branches[j], branches[j - 1]);
  Tree t = td.tree;  // td.Clone()
  auto graft = [&](int tip, int child) {  // gotree GraftTipOnEdge on the edge above `child`
    int p = t.nodes[child].parent;
    int m = t.add_node();
    t.nodes[m].parent = p;
    auto& pc = t.nodes[p].children;
    *std::find(pc.begin(), pc.end(), child) = m;
    t.nodes[m].children = {tip, child};
    t.nodes[tip].parent = m;
    t.nodes[child].parent = m;
    return m;
  };
  for (size_t i = 0; i < branches.size(); ++i) {
    std::string h = "#H" + std::to_string(i + 1);
    int r = t.add_node();
    t.nodes[r].name = h;
    graft(r, td.id_to_node[branches[i].u]);
    r = t.add_node();
    t.nodes[r].name = "####";
    int m = graft(r, td.id_to_node[branches[i].w]);
    t.nodes[m].name = h;
  }
  std::string nwk = t.newick();
  auto replace_all = [](std::string s, const std::string& a, const std::string& b) {
    size_t p = 0;
    while ((p = s.find(a, p)) != std::string::npos) {
      s.replace(p, a.size(), b);
      p += b.size();
    }
    return s;
  };
  nwk = replace_all(nwk, "####,", "");
  nwk = replace_all(nwk, ",####", "");
  return nwk;
}


// infer.Infer after Preprocess + Scorer.Init (infer.go:79-95) and DP.collateResults
// (main_dp.go:106-127): run the DP over precomputed tables.
inline DPResults run_dp(const TreeData& td, const ScoreTables& tabs, ScoreMode mode, int n_gtrees, double alpha,
                        int nthreads = 0, const DpAccelApi* api = nullptr, int device = 0, int64_t min_work = -1) {
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  DPResults res;
  DpAccel accel;
  // the across-scan on the GPU: the whole edge-score table goes to the device once
  auto open_accel = [&](const auto& sc) -> DpAccel* {
    if (!api) return nullptr;
    const auto tc0 = std::chrono::steady_clock::now();
    using S = typename std::decay_t<decltype(sc)>::S;
    const int n = td.n();
    std::vector<S> edge((size_t)n * n);
    for (int u = 0; u < n; ++u)
      for (int w = 0; w < n; ++w) edge[(size_t)u * n + w] = sc.calc(u, w);
    std::vector<int32_t> end(td.subtree_end.begin(), td.subtree_end.end()), dep(td.depths.begin(), td.depths.end());
    accel.api = api;
    int rc = api->create(&accel.handle, device, std::is_same_v<S, double> ? 1 : 0, n, end.data(), dep.data(), edge.data());
    if (rc != 0) {
      const char* m = api->last_error(nullptr);
      throw CamusError(Err::Internal, std::string("DP offload: create: ") + (m ? m : "?"));
    }
    accel.s_create = std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count();
    accel.min_work = min_work >= 0 ? (uint64_t)min_work : 30000;  // measured at 1000 taxa: 1e5 .. 8e3 are within 10 % of each other
    if (const char* e = getenv("CAMUS_DP_ACCEL_MIN_WORK")) accel.min_work = strtoull(e, nullptr, 10);
    accel.verify = getenv("CAMUS_DP_VERIFY") && getenv("CAMUS_DP_VERIFY")[0] == '1';
    return &accel;
  };
  struct Closer {
    DpAccel& a;
    void close() {
      if (a.handle) a.api->destroy(a.handle);
      a.handle = nullptr;
    }
    ~Closer() { close(); }
  } closer{accel};
  auto collate = [&](auto& dp) {
    const auto& rs = dp.root_scores();
    int m = (int)rs.size() - 1;
    for (int k = 0; k <= m; ++k) {
      res.root_scores_f.push_back(double(rs[k]));
      if constexpr (std::is_same_v<std::decay_t<decltype(rs[k])>, uint64_t>) res.root_scores_u.push_back(rs[k]);
    }
    for (int k = 1; k <= m; ++k) {
      res.branches.push_back(dp.traceback(k));
      res.qsat.push_back(percent_quartet_sat(tabs, res.branches.back()));
    }
  };
  if (mode == ScoreMode::Max) {
    MaximizeScorer sc{&tabs};
    DP<MaximizeScorer> dp(td, sc, nthreads, open_accel(sc));
    dp.run();
    collate(dp);
  } else if (mode == ScoreMode::Norm) {
    NormalizedScorer sc{&tabs, n_gtrees};
    DP<NormalizedScorer> dp(td, sc, nthreads, open_accel(sc));
    dp.run();
    collate(dp);
  } else {
    SymDiffScorer sc{&tabs, 0, alpha};  // NGTree is never set for sym (infer.go:84-85)
    DP<SymDiffScorer> dp(td, sc, nthreads, open_accel(sc));
    dp.run();
    collate(dp);
  }
  if (accel.handle) {
    res.accel_steps = accel.steps;
    res.accel_steps_host = accel.steps_host;
    res.accel_nan_fallbacks = accel.nan_fallbacks;
    res.accel_launches = api->launches(accel.handle);
    res.accel_s_columns = accel.s_columns;
    res.accel_s_across = accel.s_across;
    res.accel_s_create = accel.s_create;
    const auto td0 = std::chrono::steady_clock::now();
    closer.close();
    res.accel_s_destroy = std::chrono::duration<double>(std::chrono::steady_clock::now() - td0).count();
  }
  return res;
}

}  // namespace camus
