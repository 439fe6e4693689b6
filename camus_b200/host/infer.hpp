// Host-side mirror of the reference's package interfaces around the hot path, with the hot path
// itself delegated to the CUDA C-ABI (include/camus_b200.h). Names, argument meaning and error
// behaviour follow the Go code so that this file reads like it:
//
//   prep.SetQuartetFilterOptions   internal/prep/quartetfilter.go:18-29
//   prep.Preprocess                internal/prep/preprocess.go:26-62
//   score.Scorer / Init / options  internal/score/scorers.go:12-144
//   infer.MakeInferOptions / Infer internal/infer/infer.go:19-96
//
// What differs by design: Preprocess returns a TreeData whose quartet tables live on the device
// (behind the camus_b200 context) instead of map[Quartet]uint32 + per-node quartet lists.
#pragma once
#include <memory>
#include <thread>

#include "camus_b200.h"
#include "dp.hpp"

namespace camus {

using LogFn = std::function<void(const std::string&)>;

// ---- device handle (RAII over the C-ABI) ------------------------------------------------------
class Device {
 public:
  explicit Device(int ordinal = 0) : Device(std::vector<int>{ordinal}) {}
  // one handle over several GPUs of this process (camus_b200_create with n_gpus > 1)
  explicit Device(const std::vector<int>& ordinals) {
    if (camus_b200_create(&ctx_, (int)ordinals.size(), ordinals.data()) != 0)
      throw CamusError(Err::Device, std::string("camus_b200: ") + camus_b200_last_error(nullptr));
  }
  ~Device() { camus_b200_destroy(ctx_); }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  camus_b200* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != 0) throw CamusError(Err::Device, std::string("camus_b200: ") + camus_b200_last_error(ctx_));
  }

 private:
  camus_b200* ctx_ = nullptr;
};

// ---- prep -------------------------------------------------------------------------------------
struct QuartetFilterOptions {
  int mode = 0;          // QMode: 0 off, 1 non-restrictive, 2 restrictive (quartetfilter.go:33-38)
  double threshold = 0;  // Threshold in [0, 1]
  bool QuartetFilterOff() const { return mode == 0; }
};

inline QuartetFilterOptions SetQuartetFilterOptions(int mode, double threshold) {
  if (mode < 0 || mode > 2) throw CamusError(Err::TypeOutRange, "quartet mode " + std::to_string(mode) + " is out of type range");
  if (!(threshold >= 0 && threshold <= 1)) throw CamusError(Err::TypeOutRange, "threshold " + std::to_string(threshold) + " is out of type range");
  return QuartetFilterOptions{mode, threshold};
}

// TreeData plus the device-resident quartet tables.
struct DeviceTreeData {
  TreeData td;
  std::shared_ptr<Device> dev;
  uint64_t n_quartets = 0;   // len(qCounts)              (TotalNumUniqueQuartets, treedata.go:305-307)
  uint64_t sum_counts = 0;   // exact; Go truncates to u32 (TotalNumQuartets, treedata.go:297-303)
  int n_gene_trees = 0;
};

// prep.Preprocess. `as_set_hint` is the one addition: the Go caller (infer.Infer) knows the score
// weight before it preprocesses, and passing it lets one fused GPU pass serve both calls.
inline DeviceTreeData Preprocess(Tree tre, std::vector<Tree>& geneTrees, int /*nprocs*/, QuartetFilterOptions opts,
                                 double minSupp, std::shared_ptr<Device> dev, const LogFn& log = nullptr,
                                 bool as_set_hint = false) {
  DeviceTreeData out;
  out.td = make_tree_data(std::move(tre));  // RemoveSingleNodes, ids, tip index, rooted / binary checks
  if (!geneTrees.empty()) {
    double pct = percent_no_support(geneTrees);
    if (pct != 0 && minSupp != 0 && log) {
      char buf[128];
      snprintf(buf, sizeof buf, "WARNING: %.2f%% of gene tree edges do not have support values", pct);
      log(buf);
    }
  }
  if (log) log("reading quartets from gene trees");
  ConstraintFlat cf = flatten_constraint(out.td);
  GeneFlat gf = flatten_gene_trees(geneTrees, out.td, minSupp, log);
  dev->check(camus_b200_set_constraint(dev->get(), cf.n_nodes, cf.n_taxa, cf.parent.data(), cf.child0.data(),
                                       cf.child1.data(), cf.tip_index.data()));
  dev->check(camus_b200_add_gene_trees(dev->get(), gf.n_trees(), gf.node_off.data(), gf.parent.data(), gf.taxon.data()));
  dev->check(camus_b200_set_score_hint(dev->get(), as_set_hint));
  dev->check(camus_b200_count_filter(dev->get(), opts.mode, opts.threshold, &out.n_quartets, &out.sum_counts));
  if (log) {
    log(std::to_string(geneTrees.size()) + " gene trees provided, containing " + std::to_string(out.n_quartets) +
        " quartets not in the constraint tree");
    log("analyzing constraint tree");
  }
  out.dev = std::move(dev);
  out.n_gene_trees = (int)geneTrees.size();
  return out;
}

// ---- score ------------------------------------------------------------------------------------
// ScorerOpts, AsSet / WithNGtrees / WithAlpha and ParseScorer (scorers.go:12-38) live in dp.hpp

// Scorer.Init for all three scorers (scorers.go:51-60, 82-99, 123-140): fills the integer tables
// from the device. The float formulas stay in dp.hpp's *Scorer::calc.
inline ScoreTables ScorerInit(ScoreMode mode, DeviceTreeData& dtd, int /*nprocs*/, const std::vector<ScoreOption>& opts,
                              ScorerOpts* resolved = nullptr, const LogFn& log = nullptr) {
  ScorerOpts o;
  for (auto& f : opts) f(o);
  ScoreTables t;
  t.n = dtd.td.n();
  t.as_set = o.asSet;
  t.n_survivors = dtd.n_quartets;
  t.sum_counts = dtd.sum_counts;
  t.totals.assign((size_t)t.n * t.n, 0);
  if (log) log("calculating edge scores");
  dtd.dev->check(camus_b200_quartet_totals(dtd.dev->get(), o.asSet, t.totals.data()));
  if (mode != ScoreMode::Max) {
    if (log) log("calculating penalties");
    t.penalties.assign((size_t)t.n * t.n, 0);
    dtd.dev->check(camus_b200_edge_penalties(dtd.dev->get(), t.penalties.data()));
  }
  if (resolved) *resolved = o;
  return t;
}

// ---- infer ------------------------------------------------------------------------------------
struct InferOptions {  // infer.go:19-26
  int NProcs = 0;
  QuartetFilterOptions QuartetOpts;
  double MinSupport = 0;
  ScoreMode ScoreModeV = ScoreMode::Max;
  bool AsSet = false;
  double Alpha = 0.1;
};

inline int setNProcs(int nprocs, const LogFn& log) {  // infer.go:54-66
  int maxProcs = (int)std::max(1u, std::thread::hardware_concurrency());
  if (nprocs > maxProcs) {
    if (log) log(std::to_string(nprocs) + " is greater than available processes (" + std::to_string(maxProcs) + "); limit set to " + std::to_string(maxProcs));
    return maxProcs;
  }
  if (nprocs <= 0) {
    if (log) log("number of processes not set; defaulting to " + std::to_string(maxProcs) + " processes");
    return maxProcs;
  }
  return nprocs;
}

inline InferOptions MakeInferOptions(int nprocs, QuartetFilterOptions q, double minSupport, ScoreMode mode, bool asSet,
                                     double alpha, const LogFn& log = nullptr) {  // infer.go:40-52
  if (q.QuartetFilterOff() && asSet && log) log("WARNING: using -asSet without quartet filtering is not recommended");
  return InferOptions{setNProcs(nprocs, log), q, minSupport, mode, asSet, alpha};
}

struct InferResults {
  DeviceTreeData tree;
  DPResults dp;
  ScoreTables tables;
};

// infer.Infer (infer.go:70-96)
inline InferResults Infer(Tree tre, std::vector<Tree>& geneTrees, const InferOptions& opts, std::shared_ptr<Device> dev,
                          const LogFn& log = nullptr) {
  auto t0 = std::chrono::steady_clock::now();
  if (log) {
    log("running infer...");
    log("beginning data preprocessing");
  }
  InferResults r;
  const bool asSet = opts.ScoreModeV == ScoreMode::Sym ? true : opts.AsSet;  // infer.go:84-85
  try {
    r.tree = Preprocess(std::move(tre), geneTrees, opts.NProcs, opts.QuartetOpts, opts.MinSupport, std::move(dev), log, asSet);
  } catch (const CamusError& e) {
    throw CamusError(e.code, std::string("preprocess error: ") + e.what());
  }
  std::vector<ScoreOption> so;
  switch (opts.ScoreModeV) {
    case ScoreMode::Max: so = {AsSet(opts.AsSet)}; break;
    case ScoreMode::Norm: so = {AsSet(opts.AsSet), WithNGtrees((int)geneTrees.size())}; break;
    case ScoreMode::Sym: so = {AsSet(true), WithAlpha(opts.Alpha)}; break;
  }
  ScorerOpts resolved;
  r.tables = ScorerInit(opts.ScoreModeV, r.tree, opts.NProcs, so, &resolved, log);
  if (log) log("preprocessing finished, beginning dp algorithm");
  r.dp = run_dp(r.tree.td, r.tables, opts.ScoreModeV, resolved.nGTrees, resolved.alpha, opts.NProcs);
  if (log) {
    log(std::to_string(r.dp.branches.size()) + " edges identified");
    log("beginning traceback");
    for (size_t k = 0; k < r.dp.branches.size(); ++k) {
      char buf[160];
      if (opts.ScoreModeV == ScoreMode::Max)
        snprintf(buf, sizeof buf, "dp scored %llu at root with %zu edges", (unsigned long long)r.dp.root_scores_u[k + 1], k + 1);
      else
        snprintf(buf, sizeof buf, "dp scored %g at root with %zu edges", r.dp.root_scores_f[k + 1], k + 1);
      log(buf);
      snprintf(buf, sizeof buf, "%f percent of quartets satisfied", r.dp.qsat[k]);
      log(buf);
    }
    char buf[96];
    snprintf(buf, sizeof buf, "done. took %f seconds.", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    log(buf);
  }
  return r;
}

}  // namespace camus
