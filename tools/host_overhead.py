#!/usr/bin/env python
"""Developer tool: wall time of each C-ABI call against the device time it reports (host overhead)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from camus_b200 import synth
from camus_b200.cabi import CamusB200
from camus_b200.hostlib import Problem
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
s, c = synth.make_config(cfg)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
g = CamusB200(0).load(prob)
g.set_score_hint(c["as_set"])
out = np.empty((prob.n_nodes, prob.n_nodes), np.uint64)
g.count_filter(c["qmode"], c["threshold"]); g.quartet_totals(c["as_set"], out)
acc = {}
def t(name, f):
    t0 = time.perf_counter(); r = f(); acc[name] = acc.get(name, 0) + (time.perf_counter() - t0); return r
R = 20
dev = 0
for _ in range(R):
    t("clear", g.clear_gene_trees)
    t("add_gene_trees", lambda: g.add_gene_trees(prob.gene_node_off, prob.gene_parent, prob.gene_taxon))
    t("count_filter", lambda: g.count_filter(c["qmode"], c["threshold"]))
    ms, _ = g.stage_ms(); dev += sum(ms.values())
    t("quartet_totals", lambda: g.quartet_totals(c["as_set"], out))
    t("run_resident", lambda: g.run_resident(c["qmode"], c["threshold"], c["as_set"], out))
print(f"config {cfg} PASSES={os.environ.get('CAMUS_B200_PASSES','-')}: per call ms:", {k: round(v / R * 1e3, 3) for k, v in acc.items()}, "device ms in count_filter:", round(dev / R, 3))
