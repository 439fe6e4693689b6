// extern "C" surface of the host-side mirror (libcamus_host.so): parsing + flattening on one
// side of the hot path, scorers + DP + network writer on the other. Used by the Python tests,
// bench.py and the `camus` CLI. Contains NO quartet counting / scoring: that is the C-ABI in
// include/camus_b200.h (CUDA) -- or, in tests only, the oracle.
#include <cstring>
#include <memory>

#include "dp.hpp"
#include "dp_mock.hpp"
#include "retic.hpp"

using namespace camus;

namespace {
thread_local std::string g_err;
int fail(const CamusError& e) {
  g_err = e.what();
  return (int)e.code;
}
}  // namespace

struct camus_host_problem {
  TreeData td;
  ConstraintFlat cf;
  GeneFlat gf;
  int n_gene_trees = 0;
  bool warned_missing = false;
  double pct_no_support = 0;
  std::string constraint_newick, constraint_newick_scratch;
};

// score.ReticulationScore inputs (internal/score/score.go:25-72), flattened for
// camus_b200_reticulation_counts
struct camus_host_network {
  NetworkData nd;
  std::vector<uint32_t> records;
  RetGeneFlat gf;
  std::vector<int32_t> net_parent, net_tip_index, uwvs;
  std::string fmt;
};

struct camus_host_results {
  DPResults res;
  std::vector<std::string> newicks;
};

extern "C" {

const char* camus_host_last_error(void) { return g_err.c_str(); }

// prep.ReadInputFiles + host half of prep.Preprocess (io.go:81-170, preprocess.go:26-44,87-100).
// gene_format: 0 newick, 1 nexus. Returns 0 or a camus::Err code.
int camus_host_load(const char* constraint_text, const char* gene_text, int gene_format, double min_support,
                    camus_host_problem** out) {
  try {
    auto p = std::make_unique<camus_host_problem>();
    Tree tre = read_constraint_text(constraint_text);
    GeneTrees gts;
    if (gene_text && *gene_text)
      gts = gene_format == 1 ? read_gene_trees_nexus(gene_text) : read_gene_trees_newick(gene_text);
    p->td = make_tree_data(std::move(tre));
    p->cf = flatten_constraint(p->td);
    p->pct_no_support = gts.trees.empty() ? 0 : percent_no_support(gts.trees);
    p->gf = flatten_gene_trees(gts.trees, p->td, min_support, [&](const std::string&) { p->warned_missing = true; });
    p->n_gene_trees = (int)gts.trees.size();
    p->constraint_newick = p->td.tree.newick();
    *out = p.release();
    return 0;
  } catch (const CamusError& e) {
    return fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
    return (int)Err::Internal;
  }
}

void camus_host_free(camus_host_problem* p) { delete p; }

void camus_host_dims(const camus_host_problem* p, int32_t* n_nodes, int32_t* n_taxa, int32_t* n_trees,
                     int64_t* total_gene_nodes) {
  *n_nodes = p->cf.n_nodes;
  *n_taxa = p->cf.n_taxa;
  *n_trees = p->gf.n_trees();
  *total_gene_nodes = (int64_t)p->gf.parent.size();
}
const int32_t* camus_host_constraint_parent(const camus_host_problem* p) { return p->cf.parent.data(); }
const int32_t* camus_host_constraint_child0(const camus_host_problem* p) { return p->cf.child0.data(); }
const int32_t* camus_host_constraint_child1(const camus_host_problem* p) { return p->cf.child1.data(); }
const int32_t* camus_host_constraint_tip_index(const camus_host_problem* p) { return p->cf.tip_index.data(); }
const int64_t* camus_host_gene_node_off(const camus_host_problem* p) { return p->gf.node_off.data(); }
const int32_t* camus_host_gene_parent(const camus_host_problem* p) { return p->gf.parent.data(); }
const int32_t* camus_host_gene_taxon(const camus_host_problem* p) { return p->gf.taxon.data(); }
int camus_host_warned_missing(const camus_host_problem* p) { return p->warned_missing; }
double camus_host_percent_no_support(const camus_host_problem* p) { return p->pct_no_support; }
const char* camus_host_constraint_newick(const camus_host_problem* p) { return p->constraint_newick.c_str(); }
const char* camus_host_tip_name(const camus_host_problem* p, int tip_index) {
  return (tip_index >= 0 && tip_index < p->td.n_leaves) ? p->td.tip_names[tip_index].c_str() : "";
}
const char* camus_host_node_name(const camus_host_problem* p, int id) {
  return (id >= 0 && id < p->td.n()) ? p->td.tree.nodes[p->td.id_to_node[id]].name.c_str() : "";
}
int camus_host_should_calc_edge(const camus_host_problem* p, int u, int w) { return should_calc_edge(u, w, p->td); }
int camus_host_cycle_length(const camus_host_problem* p, int u, int w) { return cycle_length(u, w, p->td); }

// accel != NULL: the DP's across-scan runs on the GPU (include/camus_b200.h "DP offload"). accel =
// ten function pointers, in this order: camus_b200_dp_create, _destroy, _last_error, _set_node,
// _set_vertex, _set_path_columns, _across, _launches, _set_conv_columns, _across_maxplus. min_work < 0 = default threshold below which a
// step stays on the host.
int camus_host_run_dp_accel(const camus_host_problem* p, int mode, const uint64_t* totals, const uint64_t* penalties,
                            int as_set, uint64_t n_survivors, uint64_t sum_counts, int n_gtrees, double alpha,
                            const void* const* accel, int device, int64_t min_work, camus_host_results** out) {
  try {
    ScoreTables tabs;
    tabs.n = p->td.n();
    size_t nn = (size_t)tabs.n * tabs.n;
    tabs.totals.assign(totals, totals + nn);
    if (penalties) tabs.penalties.assign(penalties, penalties + nn);
    else tabs.penalties.assign(nn, 0);
    tabs.as_set = as_set != 0;
    tabs.n_survivors = n_survivors;
    tabs.sum_counts = sum_counts;
    auto r = std::make_unique<camus_host_results>();
    // CAMUS_HOST_THREADS: worker threads of the DP's across-scan (0 / unset = all cores); the
    // result does not depend on it
    const char* e = getenv("CAMUS_HOST_THREADS");
    DpAccelApi api{};
    if (accel) {
      static_assert(sizeof(DpAccelApi) == 10 * sizeof(void*), "ten function pointers");
      memcpy(&api, accel, sizeof api);
    }
    r->res = run_dp(p->td, tabs, (ScoreMode)mode, n_gtrees, alpha, e ? atoi(e) : 0, accel ? &api : nullptr, device, min_work);
    // one network per k, independent of each other (each clones the tree and grafts its k branches): on all host threads
    {
      const size_t m = r->res.branches.size();
      r->newicks.assign(m, std::string());
      const size_t nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), m > 8 ? m : 1);
      std::atomic<size_t> next{0};
      std::vector<std::thread> th;
      auto work = [&] {
        for (size_t i; (i = next.fetch_add(1)) < m;) r->newicks[i] = make_network_newick(p->td, r->res.branches[i]);
      };
      for (size_t t = 1; t < nt; ++t) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
    }
    *out = r.release();
    return 0;
  } catch (const CamusError& e) {
    return fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
    return (int)Err::Internal;
  }
}
// infer.Infer after Scorer.Init (infer.go:79-95): DP + traceback + MakeNetwork over the two
// integer tables the hot path produced. mode: 0 max, 1 norm, 2 sym. penalties may be NULL for max.
int camus_host_run_dp(const camus_host_problem* p, int mode, const uint64_t* totals, const uint64_t* penalties,
                      int as_set, uint64_t n_survivors, uint64_t sum_counts, int n_gtrees, double alpha,
                      camus_host_results** out) {
  return camus_host_run_dp_accel(p, mode, totals, penalties, as_set, n_survivors, sum_counts, n_gtrees, alpha, nullptr, 0, -1, out);
}
// offload statistics of the last run: [0] steps on the device, [1] steps kept on the host, [2] steps that saw a NaN score,
// [3] kernel launches, [4] microseconds uploading items / path columns, [5] microseconds inside camus_b200_dp_across,
// [6] microseconds building + uploading the edge table and creating the handle, [7] microseconds destroying it
void camus_host_results_accel_stats(const camus_host_results* r, uint64_t* out6) {
  out6[0] = r->res.accel_steps;
  out6[1] = r->res.accel_steps_host;
  out6[2] = r->res.accel_nan_fallbacks;
  out6[3] = r->res.accel_launches;
  out6[4] = (uint64_t)(r->res.accel_s_columns * 1e6);
  out6[5] = (uint64_t)(r->res.accel_s_across * 1e6);
  out6[6] = (uint64_t)(r->res.accel_s_create * 1e6);
  out6[7] = (uint64_t)(r->res.accel_s_destroy * 1e6);
}
// the CPU stand-in of the camus_b200_dp_* entry points (dp_mock.hpp): tests without a GPU only
const void* const* camus_host_dp_mock_api() { return reinterpret_cast<const void* const*>(camus::dp_mock::api()); }
void camus_host_results_free(camus_host_results* r) { delete r; }
int camus_host_results_count(const camus_host_results* r) { return (int)r->res.branches.size(); }
// writes 2*k ints (u0,w0,u1,w1,...) for the k-branch solution (k = 1..count), traceback order
int camus_host_results_branches(const camus_host_results* r, int k, int32_t* uw) {
  if (k < 1 || k > (int)r->res.branches.size()) return -1;
  const auto& b = r->res.branches[k - 1];
  for (size_t i = 0; i < b.size(); ++i) {
    uw[2 * i] = b[i].u;
    uw[2 * i + 1] = b[i].w;
  }
  return (int)b.size();
}
double camus_host_results_qsat(const camus_host_results* r, int k) { return r->res.qsat[k - 1]; }
double camus_host_results_root_score(const camus_host_results* r, int k) { return r->res.root_scores_f[k]; }
const char* camus_host_results_newick(const camus_host_results* r, int k) { return r->newicks[k - 1].c_str(); }

// graphs.MakeNetwork(td, branches).Newick() (graphs/net.go:38-101) for k given branches (u0,w0,u1,w1,...).
const char* camus_host_make_network(camus_host_problem* p, int k, const int32_t* uw) {
  try {
    std::vector<Branch> b(k);
    for (int i = 0; i < k; ++i) {
      b[i].u = uw[2 * i];
      b[i].w = uw[2 * i + 1];
    }
    p->constraint_newick_scratch = make_network_newick(p->td, b);
    return p->constraint_newick_scratch.c_str();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// score.ParseScorer + the option funcs infer.Infer passes (infer.go:79-88, scorers.go:12-38):
// 0 ok, -1 unknown score mode (camus.go reports it), else camus::Err (8 = ErrInvalidScorerOption).
int camus_host_check_scorer_options(const char* mode, int n_gtrees, double alpha) {
  try {
    ScoreMode m;
    if (!ParseScorer(mode, m)) return -1;
    ScorerOpts o;
    if (m == ScoreMode::Norm) WithNGtrees(n_gtrees)(o);
    if (m == ScoreMode::Sym) WithAlpha(alpha)(o);
    return 0;
  } catch (const CamusError& e) {
    return fail(e);
  }
}

// ---- score.ReticulationScore: host half (network parsing, level-1 check, per-tip records) ----
// Returns 0, a camus::Err code, 11 (ErrNoReticulations) or 12 (ErrNotLevel1).
int camus_host_network_load(const char* network_text, const char* gene_text, int gene_format, camus_host_network** out) {
  try {
    auto p = std::make_unique<camus_host_network>();
    p->nd = make_network_data(read_constraint_text(network_text));
    p->records = reticulation_records(p->nd);
    GeneTrees gts;
    if (gene_text && *gene_text)
      gts = gene_format == 1 ? read_gene_trees_nexus(gene_text) : read_gene_trees_newick(gene_text);
    p->gf = flatten_gene_trees_for_network(gts.trees, p->nd);
    for (int i = 0; i < p->nd.n; ++i) {
      p->net_parent.push_back(p->nd.parent[i]);
      const Node& x = p->nd.tree.nodes[p->nd.node_of_id[i]];
      p->net_tip_index.push_back(x.tip() ? x.tip_index : -1);
    }
    for (size_t r = 0; r < p->nd.labels.size(); ++r) {
      p->uwvs.push_back(p->nd.u[r]);
      p->uwvs.push_back(p->nd.w[r]);
      p->uwvs.push_back(p->nd.v[r]);
      p->uwvs.push_back(p->nd.wsub[r]);
    }
    *out = p.release();
    return 0;
  } catch (const RetError& e) {
    g_err = e.what();
    return e.code;
  } catch (const CamusError& e) {
    return fail(e);
  } catch (const std::exception& e) {
    g_err = e.what();
    return (int)Err::Internal;
  }
}
void camus_host_network_free(camus_host_network* p) { delete p; }
void camus_host_network_dims(const camus_host_network* p, int32_t* n_nodes, int32_t* n_tips, int32_t* n_ret, int32_t* n_trees,
                             int64_t* total_gene_nodes) {
  *n_nodes = p->nd.n;
  *n_tips = p->nd.n_tips;
  *n_ret = (int32_t)p->nd.labels.size();
  *n_trees = (int32_t)p->gf.node_off.size() - 1;
  *total_gene_nodes = (int64_t)p->gf.parent.size();
}
const uint32_t* camus_host_network_records(const camus_host_network* p) { return p->records.data(); }
const int64_t* camus_host_network_gene_node_off(const camus_host_network* p) { return p->gf.node_off.data(); }
const int32_t* camus_host_network_gene_parent(const camus_host_network* p) { return p->gf.parent.data(); }
const int32_t* camus_host_network_gene_taxon(const camus_host_network* p) { return p->gf.taxon.data(); }
const int32_t* camus_host_network_parent(const camus_host_network* p) { return p->net_parent.data(); }
const int32_t* camus_host_network_tip_index(const camus_host_network* p) { return p->net_tip_index.data(); }
const int32_t* camus_host_network_uwvs(const camus_host_network* p) { return p->uwvs.data(); }  // 4 ints per label: u, w, v, wSub
const char* camus_host_network_label(const camus_host_network* p, int r) { return p->nd.labels[r].c_str(); }
const char* camus_host_network_node_name(const camus_host_network* p, int id) {
  return (id >= 0 && id < p->nd.n) ? p->nd.tree.nodes[p->nd.node_of_id[id]].name.c_str() : "";
}
// strconv.FormatFloat(supported / total, 'f', -1, 64) as prep.WriteRetScoresToCSV prints it (io.go:315-340)
const char* camus_host_format_ratio(camus_host_network* p, uint64_t supported, uint64_t total) {
  p->fmt = format_ratio(supported, total);
  return p->fmt.c_str();
}

}  // extern "C"
