#!/usr/bin/env python
"""Developer tool: time camus_b200_reticulation_counts (score.ReticulationScore, score.go:25-72) on a
network inferred by the product itself, next to the literal CPU oracle on a bounded sample.

  python tools/retic_times.py [--config 2] [--taxa N] [--trees G] [--k 5] [--cpu-trees 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from camus_b200 import synth  # noqa: E402
from camus_b200.cabi import CamusB200  # noqa: E402
from camus_b200.hostlib import NetworkProblem, Problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--taxa", type=int, default=None)
ap.add_argument("--trees", type=int, default=None)
ap.add_argument("--k", type=int, default=5, help="reticulations of the network to score")
ap.add_argument("--cpu-trees", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

s, c = synth.make_config(a.config, a.taxa, a.trees)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
g = CamusB200(0).load(prob)
ns, sc = g.count_filter(c["qmode"], c["threshold"])
res = prob.run_dp(g.quartet_totals(False), None, "max", False, ns, sc)
k = min(a.k, len(res.newicks))
net = NetworkProblem(res.newicks[k - 1], s.genes, s.fmt)
best = None
for _ in range(a.reps):
    t0 = time.perf_counter()
    tot, sup = g.reticulation_counts(net)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
n_sets = net.n_tips * (net.n_tips - 1) * (net.n_tips - 2) * (net.n_tips - 3) // 24
out = {"taxa": prob.n_taxa, "network_tips": net.n_tips, "gene_trees": net.n_trees, "reticulations": net.n_ret,
       "gpu_s": round(best, 4), "sets_x_trees_x_reticulations_per_s": n_sets * net.n_trees * net.n_ret / best,
       "comparable_emissions": int(tot.sum()), "supporting_emissions": int(sup.sum())}
if a.cpu_trees > 0:
    from oracle import oracle as orc  # the checker, timed beside the GPU path

    class Sample:
        pass
    sm = Sample()
    for f in ("parent", "tip_index", "uwvs", "n_nodes", "n_tips", "n_ret"):
        setattr(sm, f, getattr(net, f))
    n = a.cpu_trees
    sm.n_trees = n
    sm.gene_node_off = net.gene_node_off[: n + 1]
    sm.gene_parent = net.gene_parent[: net.gene_node_off[n]]
    sm.gene_taxon = net.gene_taxon[: net.gene_node_off[n]]
    t0 = time.perf_counter()
    ot, os_ = orc.reticulation_counts(sm)
    dtc = time.perf_counter() - t0
    out.update({"cpu_trees": n, "cpu_s": round(dtc, 3), "cpu_s_per_tree": dtc / n, "gpu_s_per_tree": best / net.n_trees,
                "sample_equal": bool(np.array_equal(ot, tot[:n]) and np.array_equal(os_, sup[:n]))})
print(json.dumps(out))
g.close()
