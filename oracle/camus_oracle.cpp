// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path (camus_b200/, include/) may
// import, link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker / reported baseline.
//
// CPU restatement of the CAMUS hot path (jsdoublel/camus, /root/reference) at the same boundary
// as the C-ABI in include/camus_b200.h: flat constraint tree + flat gene trees in, packed
// quartet counts / n x n edge tables out. Two modes:
//   literal = the reference's own algorithms, function by function (cited below);
//   fast    = algebraically equivalent forms (four-point LCA-depth test, per-quartet scatter)
//             used to check mid-size inputs; property-tested against literal in tests/.
//
// PARITY PIN: the Go reference cannot be built here (no Go toolchain; gotree v0.4.5 not
// vendored -- SURVEY.md §8c). The literal mode is pinned by the reference's own known-answer
// tests and golden files (tests/test_oracle_golden.py): 5/5 TestProcessQuartets cases, the
// TestQuartetsFromTree / TestQuartetsTotal / CycleLength / ShouldCalcEdge / penalty KATs, 12/12
// TestInfer networks and all 17 golden networks of TestInfer_Large. gotree's
// tree.Quartets(false, cb) (go.mod:7, called at internal/graphs/quartet.go:119) is restated from
// its published algorithm: for every internal edge of the unrooted tree, every (pair of tips on
// one side) x (pair on the other). Not pinned by any reference test: -s support collapse,
// inputs above 22 taxa, the uint32 wrap of TotalNumQuartets ("parity unpinned" for those).
#include <algorithm>
#include <array>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using Quartet = uint64_t;
constexpr int kTaxaShift = 15, kTopoShift = 60;           // graphs/quartet.go:17-18
constexpr uint64_t kTaxaMask = (1u << kTaxaShift) - 1;    // graphs/quartet.go:20
constexpr uint8_t kTopo1 = 0b1100, kTopo2 = 0b1010, kTopo3 = 0b0110;  // graphs/quartet.go:23-25
// graphs/quartet.go:27-31: iota continues through the const block, so the values are 9,10,11;
// only identity matters.
constexpr int Qneq = 9, Qeq = 10, Qdiff = 11;

// graphs/quartet.go:93-109
uint8_t sort_taxa(int16_t arr[4]) {
  uint8_t topo = 0b0011;
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 4; j++)
      if (arr[i] > arr[j]) {
        uint8_t bi = (topo >> i) & 1, bj = (topo >> j) & 1;
        if (bi != bj) topo ^= uint8_t((1 << i) | (1 << j));
        std::swap(arr[i], arr[j]);
      }
  return topo;
}
// graphs/quartet.go:77-89
uint8_t set_topology(int16_t taxa[4]) {
  uint8_t topo = sort_taxa(taxa);
  if (topo % 2 != 0) topo ^= 0b1111;
  return topo;
}
// graphs/quartet.go:67-74
Quartet make_quartet(const int16_t taxa[4], uint8_t topo) {
  uint64_t q = 0;
  for (int i = 0; i < 4; i++) q |= uint64_t(uint16_t(taxa[i])) << (kTaxaShift * i);
  q |= uint64_t(topo) << kTopoShift;
  return q;
}
inline uint8_t q_topology(Quartet q) { return uint8_t((q >> kTopoShift) & 0xF); }      // quartet.go:151-153
inline uint16_t q_taxon(Quartet q, int i) { return uint16_t((q >> (kTaxaShift * i)) & kTaxaMask); }  // :155-157
// graphs/quartet.go:169-177
std::array<Quartet, 3> all_quartets(Quartet q) {
  uint64_t base = q & ~(uint64_t(0xF) << kTopoShift);
  return {base | (uint64_t(kTopo1) << kTopoShift), base | (uint64_t(kTopo2) << kTopoShift),
          base | (uint64_t(kTopo3) << kTopoShift)};
}

// map[Quartet]uint32 of the reference. Open addressing with linear probing (Go's maps are
// bucketed open-addressing tables too; a node-based std::unordered_map would make the CPU
// baseline several times slower than the Go code it stands for). Key 0 (NilQuartet) = empty.
class QMap {
 public:
  struct KV {
    Quartet first;
    uint32_t second;
  };
  struct iterator {
    const QMap* m;
    size_t i;
    KV operator*() const { return KV{m->keys_[i], m->vals_[i]}; }
    iterator& operator++() {
      ++i;
      while (i < m->keys_.size() && m->keys_[i] == 0) ++i;
      return *this;
    }
    bool operator!=(const iterator& o) const { return i != o.i; }
  };
  iterator begin() const {
    iterator it{this, 0};
    while (it.i < keys_.size() && keys_[it.i] == 0) ++it.i;
    return it;
  }
  iterator end() const { return iterator{this, keys_.size()}; }
  size_t size() const { return count_; }
  void clear() {
    keys_.clear();
    vals_.clear();
    count_ = 0;
  }
  uint32_t& operator[](Quartet k) {
    if ((count_ + 1) * 10 > keys_.size() * 7) grow();
    size_t i = slot(k);
    while (keys_[i] != 0 && keys_[i] != k) i = (i + 1) & mask();
    if (keys_[i] == 0) {
      keys_[i] = k;
      vals_[i] = 0;
      ++count_;
    }
    return vals_[i];
  }
  const uint32_t* find(Quartet k) const {
    if (keys_.empty()) return nullptr;
    size_t i = slot(k);
    while (keys_[i] != 0) {
      if (keys_[i] == k) return &vals_[i];
      i = (i + 1) & mask();
    }
    return nullptr;
  }
  size_t count(Quartet k) const { return find(k) ? 1 : 0; }
  uint32_t at(Quartet k) const { return *find(k); }
  void erase(Quartet k) {  // backward-shift deletion
    if (keys_.empty()) return;
    size_t i = slot(k);
    while (keys_[i] != 0 && keys_[i] != k) i = (i + 1) & mask();
    if (keys_[i] == 0) return;
    --count_;
    size_t j = i;
    for (;;) {
      j = (j + 1) & mask();
      if (keys_[j] == 0) break;
      size_t h = slot(keys_[j]);
      if (((j - h) & mask()) >= ((j - i) & mask())) {
        keys_[i] = keys_[j];
        vals_[i] = vals_[j];
        i = j;
      }
    }
    keys_[i] = 0;
  }

 private:
  std::vector<Quartet> keys_;
  std::vector<uint32_t> vals_;
  size_t count_ = 0;
  size_t mask() const { return keys_.size() - 1; }
  size_t slot(Quartet k) const {
    uint64_t z = k;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return size_t(z ^ (z >> 31)) & mask();
  }
  void grow() {
    size_t cap = keys_.empty() ? 64 : keys_.size() * 2;
    std::vector<Quartet> ok(cap, 0);
    std::vector<uint32_t> ov(cap, 0);
    ok.swap(keys_);
    ov.swap(vals_);
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i] != 0) {
        size_t j = slot(ok[i]);
        while (keys_[j] != 0) j = (j + 1) & mask();
        keys_[j] = ok[i];
        vals_[j] = ov[i];
      }
  }
};

struct FlatTree {  // rooted tree given by parent pointers; children in index order
  int n = 0;
  std::vector<int> parent, taxon;
  std::vector<std::vector<int>> children;
  int root = -1;
  void build() {
    children.assign(n, {});
    for (int i = 0; i < n; i++) {
      if (parent[i] < 0) root = i;
      else children[parent[i]].push_back(i);
    }
  }
};

// gotree tree.Quartets(false, cb) restated (see header). Emits every (pair | pair) over every
// internal edge; duplicates across edges are emitted again, exactly like the enumerator the
// reference drives (2.6-3.2x duplicate emissions). Taxon ids are constraint
// tip indices (graphs.MapIDsFromConstTree, quartet.go:131-149, already applied by the caller).
template <class F>
void enumerate_quartets_literal(const FlatTree& t, F&& cb, uint64_t* emissions = nullptr) {
  // leaves below each node
  std::vector<std::vector<int16_t>> below(t.n);
  std::vector<int> order;  // preorder via stack, then reverse for postorder
  order.reserve(t.n);
  std::vector<int> st{t.root};
  while (!st.empty()) {
    int x = st.back();
    st.pop_back();
    order.push_back(x);
    for (int c : t.children[x]) st.push_back(c);
  }
  std::vector<int16_t> all;
  for (int i = (int)order.size() - 1; i >= 0; i--) {
    int x = order[i];
    if (t.children[x].empty()) below[x].push_back((int16_t)t.taxon[x]);
    else
      for (int c : t.children[x]) below[x].insert(below[x].end(), below[c].begin(), below[c].end());
  }
  all = below[t.root];
  std::vector<char> in(1 << 15, 0);
  uint64_t em = 0;
  // tre.UnRoot() (quartet.go:113): a degree-2 root disappears and its two edges become one.
  int skip = -1;
  if (t.children[t.root].size() == 2) skip = t.children[t.root][1];
  for (int x : order) {
    if (x == t.root || x == skip || t.children[x].empty()) continue;  // internal edges only
    const auto& L = below[x];
    if (L.size() < 2 || all.size() - L.size() < 2) continue;
    for (int16_t a : L) in[a] = 1;
    std::vector<int16_t> R;
    for (int16_t a : all)
      if (!in[a]) R.push_back(a);
    for (int16_t a : L) in[a] = 0;
    for (size_t i = 0; i < L.size(); i++)
      for (size_t j = i + 1; j < L.size(); j++)
        for (size_t k = 0; k < R.size(); k++)
          for (size_t l = k + 1; l < R.size(); l++) {
            int16_t ids[4] = {L[i], L[j], R[k], R[l]};  // graphs.QuartetFromTreeQ, quartet.go:126-129
            uint8_t topo = set_topology(ids);
            cb(make_quartet(ids, topo));
            em++;
          }
  }
  if (emissions) *emissions += em;
}

// Appendix B of SURVEY.md: four-point test on LCA depths (fast mode).
template <class F>
void enumerate_quartets_fast(const FlatTree& t, F&& cb) {
  std::vector<int> depth(t.n, 0), order;
  std::vector<int> st{t.root};
  while (!st.empty()) {
    int x = st.back();
    st.pop_back();
    order.push_back(x);
    for (auto it = t.children[x].rbegin(); it != t.children[x].rend(); ++it) {
      depth[*it] = depth[x] + 1;
      st.push_back(*it);
    }
  }
  std::vector<int> leaves;  // DFS order
  for (int x : order)
    if (t.children[x].empty()) leaves.push_back(x);
  int m = (int)leaves.size();
  if (m < 4) return;
  // gap[k] = depth of LCA(leaf k, leaf k+1): the node after leaf k in preorder hangs off that LCA
  std::vector<int> pos(t.n, -1);
  for (int i = 0; i < (int)order.size(); i++) pos[order[i]] = i;
  std::vector<int> gap(m - 1);
  for (int k = 0; k + 1 < m; k++) gap[k] = depth[t.parent[order[pos[leaves[k]] + 1]]];
  std::vector<int> D((size_t)m * m, 0);
  for (int p = 0; p < m; p++) {
    int mn = 1 << 30;
    for (int q = p + 1; q < m; q++) {
      mn = std::min(mn, gap[q - 1]);
      D[(size_t)p * m + q] = D[(size_t)q * m + p] = mn;
    }
  }
  std::vector<int> byid(m);
  for (int i = 0; i < m; i++) byid[i] = i;
  std::sort(byid.begin(), byid.end(), [&](int x, int y) { return t.taxon[leaves[x]] < t.taxon[leaves[y]]; });
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++)
      for (int k = j + 1; k < m; k++)
        for (int l = k + 1; l < m; l++) {
          int a = byid[i], b = byid[j], c = byid[k], d = byid[l];
          int s1 = D[(size_t)a * m + b] + D[(size_t)c * m + d];
          int s2 = D[(size_t)a * m + c] + D[(size_t)b * m + d];
          int s3 = D[(size_t)a * m + d] + D[(size_t)b * m + c];
          int mx = std::max(s1, std::max(s2, s3));
          if ((s1 == mx) + (s2 == mx) + (s3 == mx) != 1) continue;
          uint8_t topo = s1 == mx ? kTopo1 : (s2 == mx ? kTopo2 : kTopo3);
          int16_t ids[4] = {(int16_t)t.taxon[leaves[a]], (int16_t)t.taxon[leaves[b]],
                            (int16_t)t.taxon[leaves[c]], (int16_t)t.taxon[leaves[d]]};
          cb(make_quartet(ids, topo));
        }
}

// prep.Threshold.Keep (quartetfilter.go:67-74)
bool threshold_keep(double thresh, uint32_t c0, uint32_t c1, uint32_t c2) {
  uint32_t c[3] = {c0, c1, c2};
  std::sort(c, c + 3);
  uint32_t sum = c[0] + c[1];
  return uint32_t(thresh * double(sum)) < c[1] - c[0];
}

// prep.filterQuartets (quartetfilter.go:76-96), iterating a snapshot of the keys and skipping
// the ones already deleted (Go's map iteration never yields a deleted key).
void filter_quartets(QMap& qc, int mode, double thresh) {
  std::vector<Quartet> keys;
  keys.reserve(qc.size());
  for (auto kv :qc) keys.push_back(kv.first);
  auto get = [&](Quartet q) -> uint32_t {
    const uint32_t* v = qc.find(q);
    return v ? *v : 0;
  };
  for (Quartet q : keys) {
    if (!qc.count(q)) continue;
    auto qs = all_quartets(q);
    uint32_t c0 = get(qs[0]), c1 = get(qs[1]), c2 = get(qs[2]);
    // slices.SortFunc on 3 elements is an insertion sort => stable
    std::stable_sort(qs.begin(), qs.end(), [&](Quartet a, Quartet b) { return get(a) < get(b); });
    if (!threshold_keep(thresh, c0, c1, c2)) {
      qc.erase(qs[0]);
      qc.erase(qs[1]);
      continue;
    }
    if (mode == 2) qc.erase(qs[0]);
  }
}

struct Dense;  // dense.hpp

struct Ctx {
  std::string err;
  std::shared_ptr<Dense> dense;  // LCA-depth tables of the dense mode, built on first use
  // constraint tree, ids as given (reference order)
  int n = 0, n_taxa = 0;
  std::vector<int> parent, child0, child1, tip_index, depth, tip_to_node;
  std::vector<uint64_t> leaves_below;
  std::vector<std::vector<uint64_t>> leafset;  // bitsets by tip index (treedata.go:107-127)
  std::vector<std::vector<int>> lca;           // treedata.go:130-166 (all pairs)
  std::vector<FlatTree> genes;
  QMap raw, q;  // raw = processQuartets output; q = after filter + constraint removal
  bool counted = false;
  std::vector<std::vector<Quartet>> qset;  // treedata.go:197-222
  bool qset_valid = false;
  uint64_t literal_emissions = 0;

  bool in_leafset(int node, int tip) const { return (leafset[node][tip >> 6] >> (tip & 63)) & 1; }
  bool under(int n1, int n2) const { return lca[n1][n2] == n1 && n1 != n2; }  // treedata.go:291-294
};

// graphs.MakeTreeData pieces (treedata.go:27-51)
void build_tree_tables(Ctx& c) {
  int n = c.n, words = (c.n_taxa + 63) / 64;
  c.depth.assign(n, 0);
  c.leaves_below.assign(n, 0);
  c.leafset.assign(n, std::vector<uint64_t>(words, 0));
  c.tip_to_node.assign(c.n_taxa, -1);
  // order nodes so that parents come first
  std::vector<int> order;
  int root = -1;
  for (int i = 0; i < n; i++)
    if (c.parent[i] < 0) root = i;
  std::vector<int> st{root};
  while (!st.empty()) {
    int x = st.back();
    st.pop_back();
    order.push_back(x);
    if (c.child0[x] >= 0) {
      c.depth[c.child0[x]] = c.depth[x] + 1;
      c.depth[c.child1[x]] = c.depth[x] + 1;
      st.push_back(c.child1[x]);
      st.push_back(c.child0[x]);
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    int x = order[i];
    if (c.child0[x] < 0) {
      c.leaves_below[x] = 1;
      int t = c.tip_index[x];
      c.leafset[x][t >> 6] |= 1ull << (t & 63);
      c.tip_to_node[t] = x;
    } else {
      c.leaves_below[x] = c.leaves_below[c.child0[x]] + c.leaves_below[c.child1[x]];
      for (int w = 0; w < words; w++) c.leafset[x][w] = c.leafset[c.child0[x]][w] | c.leafset[c.child1[x]][w];
    }
  }
  // all-pairs LCA (the reference does this in Theta(n^3), treedata.go:130-166; same table by climbing)
  c.lca.assign(n, std::vector<int>(n, 0));
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      int x = a, y = b;
      while (x != y) {
        if (c.depth[x] >= c.depth[y]) x = c.parent[x];
        else y = c.parent[y];
      }
      c.lca[a][b] = x;
    }
}

// graphs.mapQuartetsToVertices (treedata.go:197-222)
// Built lazily (first literal totals / lookup): the reference runs it inside MakeTreeData at the
// end of Preprocess; keeping it out of count_filter lets bench.py time the count stage alone.
void map_quartets_to_vertices(Ctx& c) {
  if (c.qset_valid) return;
  c.qset_valid = true;
  c.qset.assign(c.n, {});
  for (int v = 0; v < c.n; v++)
    for (auto kv :c.q) {
      int found = 0;
      for (int i = 0; i < 4; i++) found += c.in_leafset(v, q_taxon(kv.first, i));
      if (found >= 3) c.qset[v].push_back(kv.first);
    }
}

// score.CycleLength / ShouldCalcEdge (quartettotals.go:63-74)
int cycle_length(const Ctx& c, int u, int w) {
  int v = c.lca[u][w];
  int len = (c.depth[u] - c.depth[v]) + (c.depth[w] - c.depth[v]) + 1;
  if (v == u) len += 1;
  return len;
}
bool should_calc_edge(const Ctx& c, int u, int w) {
  return !c.under(w, u) && cycle_length(c, u, w) > 3 && u != 0 && w != 0;
}
// score.getWSubtree (quartettotals.go:95-104)
int get_w_subtree(const Ctx& c, int u, int w, int v) {
  if (u == v) return v;
  if (c.under(c.child0[v], w) || w == c.child0[v]) return c.child0[v];
  return c.child1[v];
}

// score.QuartetScore (quartettotals.go:107-157) with its helpers (:160-209), line by line.
int quartet_score(const Ctx& c, Quartet q, int u, int w, int v, int wSub) {
  // uniqueTaxaBelowNodeFromQ (:160-171)
  const uint16_t Max16 = 0xFFFF;
  uint16_t bottom = Max16;
  int bi = -1;
  bool unique = true;
  for (int i = 0; i < 4; i++) {
    uint16_t t = q_taxon(q, i);
    if (c.in_leafset(w, t) && bottom == Max16) {
      bottom = t;
      bi = i;
    } else if (c.in_leafset(w, t)) {
      unique = false;
      break;
    }
  }
  if (!unique || bottom == Max16) return Qdiff;
  int cycleNodes[4];
  int taxaToLCA[4][2];
  for (int i = 0; i < 4; i++) {
    uint16_t t = q_taxon(q, i);
    int tID = c.tip_to_node[t];
    int l;
    if (!c.in_leafset(v, t)) l = 0;
    else if (c.in_leafset(wSub, t) || c.in_leafset(u, bottom)) l = c.lca[w][tID];
    else l = c.lca[u][tID];
    cycleNodes[i] = l;
    taxaToLCA[i][0] = t;
    taxaToLCA[i][1] = l;
  }
  for (int i = 0; i < 4; i++)  // dups (:185-194)
    for (int j = i + 1; j < 4; j++)
      if (cycleNodes[i] == cycleNodes[j]) return Qdiff;
  // neighborTaxaQ (:174-182)
  uint16_t neighbor = 0;
  {
    uint8_t b = (q_topology(q) >> bi) % 2;
    for (int j = 0; j < 4; j++)
      if (j != bi && (q_topology(q) >> j) % 2 == b) {
        neighbor = q_taxon(q, j);
        break;
      }
  }
  int lcaDepths[4][2];
  for (int i = 0; i < 4; i++) {
    lcaDepths[i][0] = cycleNodes[i];
    lcaDepths[i][1] = c.depth[cycleNodes[i]];
  }
  auto get = [](int mp[4][2], int k) {  // stackMap.get: first match (:203-209)
    for (int i = 0; i < 4; i++)
      if (mp[i][0] == k) return mp[i][1];
    return -1;
  };
  int minW = c.n_taxa, maxU = -1;
  uint16_t bestTaxa = 0;
  bool taxaInU = false;
  for (int i = 0; i < 4; i++) {
    uint16_t t = q_taxon(q, i);
    int d = get(lcaDepths, get(taxaToLCA, t));
    if (!taxaInU && (c.in_leafset(wSub, t) && d < minW)) {
      minW = d;
      bestTaxa = t;
    } else if (!c.in_leafset(wSub, t) && d > maxU) {
      taxaInU = true;
      maxU = d;
      bestTaxa = t;
    }
  }
  return bestTaxa == neighbor ? Qeq : Qneq;
}

// score.quartetsTotal (quartettotals.go:78-93)
uint64_t quartets_total(const Ctx& c, int u, int w, bool asSet) {
  int v = c.lca[u][w];
  uint64_t total = 0;
  int wSub = get_w_subtree(c, u, w, v);
  for (Quartet q : c.qset[v])
    if (quartet_score(c, q, u, w, v, wSub) == Qeq) total += asSet ? 1 : uint64_t(c.q.at(q));
  return total;
}

// score.collectNodesUpPath / getNumTaxaUnderNodes / calculatePenalty (penalty.go:33-86)
void collect_up(const Ctx& c, int start, int end, std::vector<uint64_t>& subsets) {
  subsets.push_back(c.leaves_below[start]);
  int cur = start, parentID = -1;
  while (parentID != end) {
    parentID = c.parent[cur];
    if (c.child0[parentID] == cur) subsets.push_back(c.leaves_below[c.child1[parentID]]);
    else subsets.push_back(c.leaves_below[c.child0[parentID]]);
    cur = parentID;
  }
}
uint64_t calculate_penalty(const Ctx& c, int u, int w) {
  std::vector<uint64_t> s;
  int v = c.lca[u][w];
  if (u == v) {
    collect_up(c, w, c.parent[v], s);
    s.push_back(uint64_t(c.n_taxa) - c.leaves_below[v]);
  } else {
    collect_up(c, w, v, s);
    collect_up(c, u, v, s);
    s.push_back(uint64_t(c.n_taxa) - c.leaves_below[v]);
  }
  uint64_t coe[4] = {1, 0, 0, 0};
  for (size_t i = 1; i < s.size(); i++)
    for (int j = 3; j > 0; j--) coe[j] += s[i] * coe[j - 1];
  return s[0] * coe[3];
}

// ---------------------------------------------------------------------------------------------
// fast mode of the edge totals: per-quartet scatter (restated in the
// (w-point, junction J, component) form DESIGN.md describes). Requires preorder node ids.
// ---------------------------------------------------------------------------------------------
bool ids_are_preorder(const Ctx& c) {
  if (c.n == 0 || c.parent[0] >= 0) return false;
  for (int i = 0; i < c.n; i++)
    if (c.child0[i] >= 0 && c.child0[i] != i + 1) return false;
  return true;
}

// One surviving quartet's contribution to the difference table diff[w][u] (n rows of n + 1).
struct ScatterAux {
  std::vector<int> sz;  // subtree sizes (preorder ids: subtree(x) = [x, x + sz[x]))
  explicit ScatterAux(const Ctx& c) : sz(c.n, 1) {
    for (int i = c.n - 1; i > 0; i--) sz[c.parent[i]] += sz[i];
  }
};
// leaf[i] = node of the quartet's i-th taxon (sorted by tip index), Lm[i][j] = LCA node of leaf i
// and leaf j, topo = the quartet's topology nibble. Preorder node ids (subtree(x) = [x, x + sz[x])).
inline void scatter_core(const Ctx& c, const ScatterAux& aux, const int leaf[4], const int Lm[4][4], uint8_t topo,
                         int64_t wgt, int64_t* diff) {
  const int n = c.n;
  const std::vector<int>& sz = aux.sz;
  auto deeper = [&](int x, int y) { return c.depth[x] >= c.depth[y] ? x : y; };
  for (int bi = 0; bi < 4; bi++) {
    int xi = -1;
    for (int j = 0; j < 4; j++)
      if (j != bi && ((topo >> j) & 1) == ((topo >> bi) & 1)) xi = j;
    int o[2], k = 0;
    for (int j = 0; j < 4; j++)
      if (j != bi && j != xi) o[k++] = j;
    const int b = leaf[bi], x = leaf[xi];
    // m: lowest ancestor of b shared with another quartet taxon
    const int m = deeper(deeper(Lm[bi][xi], Lm[bi][o[0]]), Lm[bi][o[1]]);
    // junction of {x,y,z} and the component holding x
    const int J = deeper(deeper(Lm[xi][o[0]], Lm[xi][o[1]]), Lm[o[0]][o[1]]);
    int lo, hi;
    bool complement = false;
    if (x > J && x < J + sz[J]) {  // x below J: the child subtree of J that holds x
      int ch = (x >= c.child1[J]) ? c.child1[J] : c.child0[J];
      lo = ch;
      hi = ch + sz[ch];
    } else {  // x outside subtree(J): everything outside, plus J itself
      complement = true;
      lo = J + 1;
      hi = J + sz[J];
    }
    for (int s = 0; s < 2; s++) {
      int wp = s == 0 ? b : m;
      int64_t v = s == 0 ? wgt : -wgt;
      int64_t* row = &diff[(size_t)wp * (n + 1)];
      if (!complement) {
        row[lo] += v;
        row[hi] -= v;
      } else {
        row[0] += v;
        row[lo] -= v;
        row[hi] += v;
      }
    }
  }
}
inline void scatter_quartet(const Ctx& c, const ScatterAux& aux, Quartet q, int64_t wgt, int64_t* diff) {
  int leaf[4], Lm[4][4];
  for (int i = 0; i < 4; i++) leaf[i] = c.tip_to_node[q_taxon(q, i)];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) Lm[i][j] = c.lca[leaf[i]][leaf[j]];
  scatter_core(c, aux, leaf, Lm, q_topology(q), wgt, diff);
}
// prefix over u, then subtree sum over w (children have larger ids), then the validity mask
void scatter_finish(const Ctx& c, const std::vector<int64_t>& diff, uint64_t* out) {
  const int n = c.n;
  std::vector<int64_t> val((size_t)n * n, 0);
  for (int w = 0; w < n; w++) {
    int64_t acc = 0;
    for (int u = 0; u < n; u++) {
      acc += diff[(size_t)w * (n + 1) + u];
      val[(size_t)w * n + u] = acc;
    }
  }
  for (int w = n - 1; w > 0; w--) {
    int p = c.parent[w];
    for (int u = 0; u < n; u++) val[(size_t)p * n + u] += val[(size_t)w * n + u];
  }
  for (int u = 0; u < n; u++)
    for (int w = 0; w < n; w++) out[(size_t)u * n + w] = should_calc_edge(c, u, w) ? (uint64_t)val[(size_t)w * n + u] : 0;
}

void totals_scatter(const Ctx& c, bool asSet, uint64_t* out) {
  ScatterAux aux(c);
  std::vector<int64_t> diff((size_t)c.n * (c.n + 1), 0);  // diff[w][u], u in [0, n]
  for (auto kv : c.q) scatter_quartet(c, aux, kv.first, asSet ? 1 : (int64_t)kv.second, diff.data());
  scatter_finish(c, diff, out);
}

template <class F>
void parallel_for(int n, int nthreads, F&& f) {
  if (nthreads <= 1 || n <= 1) {
    for (int i = 0; i < n; i++) f(i, 0);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&, t] {
      for (;;) {
        int i = next.fetch_add(1);
        if (i >= n) break;
        f(i, t);
      }
    });
  for (auto& x : th) x.join();
}

#include "dense.hpp"

}  // namespace

extern "C" {

struct oracle_ctx : Ctx {};

oracle_ctx* oracle_create(void) { return new oracle_ctx(); }
void oracle_destroy(oracle_ctx* c) { delete c; }
const char* oracle_last_error(const oracle_ctx* c) { return c->err.c_str(); }

int oracle_set_constraint(oracle_ctx* c, int32_t n_nodes, int32_t n_taxa, const int32_t* parent,
                          const int32_t* child0, const int32_t* child1, const int32_t* tip_index) {
  c->n = n_nodes;
  c->n_taxa = n_taxa;
  c->parent.assign(parent, parent + n_nodes);
  c->child0.assign(child0, child0 + n_nodes);
  c->child1.assign(child1, child1 + n_nodes);
  c->tip_index.assign(tip_index, tip_index + n_nodes);
  build_tree_tables(*c);
  c->dense.reset();
  c->genes.clear();
  c->raw.clear();
  c->q.clear();
  c->counted = false;
  return 0;
}

int oracle_add_gene_trees(oracle_ctx* c, int32_t n_trees, const int64_t* node_off, const int32_t* parent,
                          const int32_t* taxon) {
  for (int g = 0; g < n_trees; g++) {
    FlatTree t;
    t.n = int(node_off[g + 1] - node_off[g]);
    t.parent.assign(parent + node_off[g], parent + node_off[g + 1]);
    t.taxon.assign(taxon + node_off[g], taxon + node_off[g + 1]);
    t.build();
    c->genes.push_back(std::move(t));
  }
  c->dense.reset();
  return 0;
}

// makeTreeDataWithQuartets of the reference's score tests (quartettotals_test.go): install a
// surviving quartet table directly.
int oracle_set_quartets(oracle_ctx* c, uint64_t n, const uint64_t* keys, const uint32_t* counts) {
  c->q.clear();
  for (uint64_t i = 0; i < n; i++) c->q[keys[i]] = counts[i];
  c->counted = true;
  c->qset_valid = false;
  return 0;
}

// quartets of ONE flat tree as a set (graphs.QuartetsFromTree, quartet.go:112-123)
int oracle_tree_quartets(int32_t n_nodes, const int32_t* parent, const int32_t* taxon, int literal,
                         uint64_t* keys, uint64_t cap, uint64_t* n_out) {
  FlatTree t;
  t.n = n_nodes;
  t.parent.assign(parent, parent + n_nodes);
  t.taxon.assign(taxon, taxon + n_nodes);
  t.build();
  QMap m;
  if (literal) enumerate_quartets_literal(t, [&](Quartet q) { m[q] = 1; });
  else enumerate_quartets_fast(t, [&](Quartet q) { m[q] = 1; });
  std::vector<Quartet> ks;
  for (auto kv :m) ks.push_back(kv.first);
  std::sort(ks.begin(), ks.end());
  *n_out = ks.size();
  for (uint64_t i = 0; i < ks.size() && i < cap; i++) keys[i] = ks[i];
  return 0;
}

// prep.processQuartets (preprocess.go:71-124) + filterQuartets (:48-50) + constraint removal
// (:51-57). literal=1 uses the bipartition enumerator; 0 the four-point form.
int oracle_count_filter(oracle_ctx* c, int qmode, double threshold, int literal, int nthreads,
                        uint64_t* n_survivors, uint64_t* sum_counts) {
  if (nthreads < 1) nthreads = 1;
  int G = (int)c->genes.size();
  std::vector<QMap> local(nthreads);
  std::vector<uint64_t> em(nthreads, 0);
  parallel_for(G, nthreads, [&](int g, int t) {
    QMap tq;  // treeQuartets[...] = 1 (quartet.go:119-121): a set per tree
    if (literal) enumerate_quartets_literal(c->genes[g], [&](Quartet q) { tq[q] = 1; }, &em[t]);
    else enumerate_quartets_fast(c->genes[g], [&](Quartet q) { tq[q] = 1; });
    for (auto kv :tq) local[t][kv.first] += kv.second;  // shards[..].counts[q] += c (:105-110)
  });
  c->raw.clear();
  for (auto& m : local)
    for (auto kv :m) c->raw[kv.first] += kv.second;  // merge (:117-122)
  for (auto e : em) c->literal_emissions += e;
  c->q = c->raw;
  if (qmode != 0) filter_quartets(c->q, qmode, threshold);
  // gr.QuartetsFromTree(tre.Clone(), tre) and delete (:51-57)
  FlatTree ct;
  ct.n = c->n;
  ct.parent = c->parent;
  ct.taxon = c->tip_index;
  ct.build();
  // children order must follow child0/child1, not index order
  for (int i = 0; i < c->n; i++)
    if (c->child0[i] >= 0) ct.children[i] = {c->child0[i], c->child1[i]};
  auto del = [&](Quartet q) { c->q.erase(q); };
  if (literal) enumerate_quartets_literal(ct, del);
  else enumerate_quartets_fast(ct, del);
  c->counted = true;
  c->qset_valid = false;
  uint64_t sum = 0;
  for (auto kv :c->q) sum += kv.second;
  if (n_survivors) *n_survivors = c->q.size();
  if (sum_counts) *sum_counts = sum;
  return 0;
}

uint64_t oracle_literal_emissions(const oracle_ctx* c) { return c->literal_emissions; }

static int export_map(const QMap& m, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_out) {
  std::vector<std::pair<Quartet, uint32_t>> v;
  v.reserve(m.size());
  for (auto kv : m) v.emplace_back(kv.first, kv.second);
  std::sort(v.begin(), v.end());
  *n_out = v.size();
  for (uint64_t i = 0; i < v.size() && i < cap; i++) {
    keys[i] = v[i].first;
    counts[i] = v[i].second;
  }
  return 0;
}
// survivors (after filter + removal), sorted by key
int oracle_export_quartets(const oracle_ctx* c, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_out) {
  return export_map(c->q, keys, counts, cap, n_out);
}
// processQuartets output before filter / removal, sorted by key
int oracle_export_raw_counts(const oracle_ctx* c, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_out) {
  return export_map(c->raw, keys, counts, cap, n_out);
}

// Raw counts of given 4-taxon sets (tip indices): per gene tree the four-point test on LCA depths
// (SURVEY.md Appendix B), LCAs by climbing. out[3*q + t], t in the order 0b1100, 0b1010, 0b0110 for
// the taxa sorted by tip index. Lets the tests check sampled cells at sizes where no table fits.
int oracle_count_cells(oracle_ctx* c, int32_t n_query, const int32_t* quads, uint32_t* out) {
  std::memset(out, 0, sizeof(uint32_t) * 3 * n_query);
  for (const FlatTree& t : c->genes) {
    std::vector<int> depth(t.n, 0), leaf_of(c->n_taxa, -1);
    for (int i = 0; i < t.n; i++) {
      if (t.parent[i] >= 0) depth[i] = depth[t.parent[i]] + 1;  // preorder: parents first
      if (t.taxon[i] >= 0) leaf_of[t.taxon[i]] = i;
    }
    auto D = [&](int x, int y) {
      while (x != y) {
        if (depth[x] >= depth[y]) x = t.parent[x];
        else y = t.parent[y];
      }
      return depth[x];
    };
    for (int q = 0; q < n_query; q++) {
      int s[4] = {quads[4 * q], quads[4 * q + 1], quads[4 * q + 2], quads[4 * q + 3]};
      std::sort(s, s + 4);
      int l[4];
      bool ok = true;
      for (int i = 0; i < 4; i++) {
        l[i] = leaf_of[s[i]];
        ok = ok && l[i] >= 0;
      }
      if (!ok) continue;
      int s1 = D(l[0], l[1]) + D(l[2], l[3]), s2 = D(l[0], l[2]) + D(l[1], l[3]), s3 = D(l[0], l[3]) + D(l[1], l[2]);
      int mx = std::max(s1, std::max(s2, s3));
      if ((s1 == mx) + (s2 == mx) + (s3 == mx) != 1) continue;
      out[3 * q + (s1 == mx ? 0 : (s2 == mx ? 1 : 2))]++;
    }
  }
  return 0;
}

// score.CalculateQuartetTotals (quartettotals.go:43-61). literal=1: per-(u,w) QuartetScore loop
// over td.Quartets(v); 0: scatter form. out[u*n+w], 0 where !ShouldCalcEdge.
int oracle_quartet_totals(oracle_ctx* c, int as_set, int literal, int nthreads, uint64_t* out) {
  if (!c->counted) {
    c->err = "quartets not initialised";
    return 1;
  }
  int n = c->n;
  if (!literal) {
    if (!ids_are_preorder(*c)) {
      c->err = "fast totals need preorder node ids";
      return 1;
    }
    totals_scatter(*c, as_set != 0, out);
    return 0;
  }
  std::memset(out, 0, sizeof(uint64_t) * n * n);
  map_quartets_to_vertices(*c);
  parallel_for(n, std::max(1, nthreads), [&](int u, int) {
    for (int w = 0; w < n; w++)
      if (should_calc_edge(*c, u, w)) out[(size_t)u * n + w] = quartets_total(*c, u, w, as_set != 0);
  });
  return 0;
}

// score.CalculateEdgePenalties (penalty.go:11-29)
int oracle_edge_penalties(oracle_ctx* c, uint64_t* out) {
  int n = c->n;
  std::memset(out, 0, sizeof(uint64_t) * n * n);
  for (int u = 0; u < n; u++)
    for (int w = 0; w < n; w++)
      if (should_calc_edge(*c, u, w)) out[(size_t)u * n + w] = calculate_penalty(*c, u, w);
  return 0;
}

// score.calculatePenalty for one pair, without the ShouldCalcEdge mask (penalty_test.go calls it so)
uint64_t oracle_calculate_penalty(const oracle_ctx* c, int u, int w) { return calculate_penalty(*c, u, w); }

int oracle_quartet_score(oracle_ctx* c, uint64_t q, int u, int w) {
  int v = c->lca[u][w];
  return quartet_score(*c, q, u, w, v, get_w_subtree(*c, u, w, v));
}
int oracle_should_calc_edge(const oracle_ctx* c, int u, int w) { return should_calc_edge(*c, u, w); }
int oracle_cycle_length(const oracle_ctx* c, int u, int w) { return cycle_length(*c, u, w); }
int oracle_lca(const oracle_ctx* c, int a, int b) { return c->lca[a][b]; }
uint64_t oracle_num_quartets_at(oracle_ctx* c, int v) {
  map_quartets_to_vertices(*c);
  return c->qset[v].size();
}
// pack four constraint tip indices with the pairing (t0,t1 | t2,t3) (quartet.go:67-89)
uint64_t oracle_pack_quartet(int t0, int t1, int t2, int t3) {
  int16_t ids[4] = {(int16_t)t0, (int16_t)t1, (int16_t)t2, (int16_t)t3};
  uint8_t topo = set_topology(ids);
  return make_quartet(ids, topo);
}
int oracle_threshold_keep(double t, uint32_t c0, uint32_t c1, uint32_t c2) { return threshold_keep(t, c0, c1, c2); }

// ---- dense mode (dense.hpp): the whole path without a quartet map, for the large parity cases ----
// stats[8] = n_survivors, sum_counts, n_raw_keys, sum_raw, hash_raw, hash_surv (0 unless digests), resolutions, cells
int oracle_dense_run(oracle_ctx* c, int qmode, double threshold, int as_set, int nthreads, int digests, uint64_t* out, uint64_t* stats) {
  DenseStats s;
  int rc = dense_run(*c, qmode, threshold, as_set != 0, nthreads, out, &s, digests != 0);
  if (rc) return rc;
  const uint64_t v[8] = {s.n_survivors, s.sum_counts, s.n_raw_keys, s.sum_raw, s.hash_raw, s.hash_surv, s.resolutions, s.cells};
  if (stats) std::memcpy(stats, v, sizeof(v));
  return 0;
}
// literal quartetsTotal (quartettotals.go:78-93) of pairs[(u, w)] that all have LCA v, any input size
int oracle_dense_literal_at(oracle_ctx* c, int v, int qmode, double threshold, int as_set, int nthreads, int n_pairs,
                            const int32_t* pairs, uint64_t* out, uint64_t* n_quartets_at_v) {
  return dense_literal_at(*c, v, qmode, threshold, as_set != 0, nthreads, n_pairs, pairs, out, n_quartets_at_v);
}
// filterQuartets on one 4-taxon set, visiting its keys in the given order (a permutation of 0 1 2)
void oracle_filter_cell(uint32_t* counts, int qmode, double threshold, const int32_t* order) {
  const int o[3] = {order[0], order[1], order[2]};
  filter_cell_literal(counts, qmode, threshold, o);
}
// prep.filterQuartets (quartetfilter.go:76-96) on an explicit map: keys/counts in, survivors out (sorted by key)
int oracle_filter_quartets(uint64_t n, const uint64_t* keys, const uint32_t* counts, int qmode, double threshold,
                           uint64_t* out_keys, uint32_t* out_counts, uint64_t* n_out) {
  QMap m;
  for (uint64_t i = 0; i < n; i++) m[keys[i]] = counts[i];
  filter_quartets(m, qmode, threshold);
  return export_map(m, out_keys, out_counts, n, n_out);
}
// digest of a (key, count) list as dense_run accumulates it
uint64_t oracle_mix_key(uint64_t key, uint32_t count) { return mix_key(key, count); }
int oracle_leaves_below(const oracle_ctx* c, int v) { return (int)c->leaves_below[v]; }
int oracle_depth(const oracle_ctx* c, int v) { return c->depth[v]; }

// score.ReticulationScore (score.go:25-72), literal: MakeTreeData on the network tree as parsed
// (general arity: #H tips and the unary #H nodes stay in), then per gene tree every EMISSION of
// gotree's enumerator -- duplicates included, the reference does not dedupe here (score.go:43) --
// is scored against every reticulation (u, w, v, wSub). Node ids: parents before children, 0 = root.
// totals[g * n_ret + r] = emissions with QuartetScore != Qdiff, supported = those == Qeq.
int oracle_reticulation_counts(int32_t n_nodes, const int32_t* net_parent, const int32_t* net_tip_index, int32_t n_tips,
                               int32_t n_ret, const int32_t* uwvs, int32_t n_trees, const int64_t* node_off,
                               const int32_t* parent, const int32_t* taxon, uint64_t* totals, uint64_t* supported) {
  Ctx c;
  c.n = n_nodes;
  c.n_taxa = n_tips;  // td.NLeaves (treedata.go:49)
  c.parent.assign(net_parent, net_parent + n_nodes);
  c.depth.assign(n_nodes, 0);
  int words = (n_tips + 63) / 64;
  c.leafset.assign(n_nodes, std::vector<uint64_t>(words, 0));
  c.tip_to_node.assign(n_tips, -1);
  for (int i = 1; i < n_nodes; i++) {
    if (net_parent[i] < 0 || net_parent[i] >= i) return 1;
    c.depth[i] = c.depth[net_parent[i]] + 1;
  }
  for (int i = n_nodes - 1; i >= 0; i--) {
    int t = net_tip_index[i];
    if (t >= 0) {
      c.leafset[i][t >> 6] |= 1ull << (t & 63);
      c.tip_to_node[t] = i;
    }
    if (i > 0)
      for (int w = 0; w < words; w++) c.leafset[net_parent[i]][w] |= c.leafset[i][w];
  }
  c.lca.assign(n_nodes, std::vector<int>(n_nodes, 0));
  for (int a = 0; a < n_nodes; a++)
    for (int b = 0; b < n_nodes; b++) {
      int x = a, y = b;
      while (x != y) {
        if (c.depth[x] >= c.depth[y]) x = c.parent[x];
        else y = c.parent[y];
      }
      c.lca[a][b] = x;
    }
  for (int g = 0; g < n_trees; g++) {
    FlatTree t;
    t.n = int(node_off[g + 1] - node_off[g]);
    t.parent.assign(parent + node_off[g], parent + node_off[g + 1]);
    t.taxon.assign(taxon + node_off[g], taxon + node_off[g + 1]);
    t.build();
    uint64_t* tot = totals + (size_t)g * n_ret;
    uint64_t* sup = supported + (size_t)g * n_ret;
    for (int r = 0; r < n_ret; r++) tot[r] = sup[r] = 0;
    enumerate_quartets_literal(t, [&](Quartet q) {
      for (int r = 0; r < n_ret; r++) {
        int comp = quartet_score(c, q, uwvs[4 * r], uwvs[4 * r + 1], uwvs[4 * r + 2], uwvs[4 * r + 3]);
        if (comp != Qdiff) tot[r]++;
        if (comp == Qeq) sup[r]++;
      }
    });
  }
  return 0;
}

}  // extern "C"
