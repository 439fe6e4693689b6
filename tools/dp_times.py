#!/usr/bin/env python
"""Developer tool: seconds of the level-1 network DP (infer/main_dp.go) per score mode, host scan
against the GPU across-scan, on the tables the GPU hot path produces for one configuration.

  python tools/dp_times.py --config 3 [--modes max,norm,sym] [--host-cap 120]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from camus_b200 import synth  # noqa: E402
from camus_b200.cabi import CamusB200  # noqa: E402
from camus_b200.hostlib import Problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=5)
ap.add_argument("--taxa", type=int, default=None)
ap.add_argument("--trees", type=int, default=None)
ap.add_argument("--modes", default="max,norm,sym")
ap.add_argument("--host", default="max,norm,sym", help="modes for which the host-only DP is timed too")
a = ap.parse_args()

s, c = synth.make_config(a.config, a.taxa, a.trees)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
g = CamusB200(0).load(prob)
out = {"config": a.config, "taxa": prob.n_taxa, "trees": prob.n_trees, "host_cores": os.cpu_count(), "modes": {}}
for mode in a.modes.split(","):
    as_set = c["as_set"] or mode == "sym"
    g.set_score_hint(as_set)
    ns, sc = g.count_filter(c["qmode"], c["threshold"])
    tot = g.quartet_totals(as_set)
    pen = g.edge_penalties() if mode != "max" else None
    rec = {}
    t0 = time.perf_counter()
    dev = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5, gpu=0)
    rec["gpu_scan_s"] = time.perf_counter() - t0
    rec["networks"] = len(dev.newicks)
    rec["accel"] = dev.accel
    if mode in a.host.split(","):
        t0 = time.perf_counter()
        host = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5)
        rec["host_scan_s"] = time.perf_counter() - t0
        rec["identical"] = host.branches == dev.branches and host.newicks == dev.newicks and host.root_scores == dev.root_scores
    out["modes"][mode] = rec
    print(mode, json.dumps(rec), flush=True)
print(json.dumps(out))
