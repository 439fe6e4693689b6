#!/usr/bin/env python
"""Developer tool: per-stage device times of the resident hot path for one configuration.

  python tools/stage_times.py [--config 2] [--taxa N] [--trees G] [--qmode Q] [--thr T] [--reps 5]

CAMUS_B200_DEBUG=1 skips the score scatter, =2 the whole fused epilogue (timing decomposition
only: results are then wrong on purpose). CAMUS_B200_UNFUSED=1 selects the table pipeline.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from camus_b200 import synth  # noqa: E402
from camus_b200.cabi import CamusB200  # noqa: E402
from camus_b200.hostlib import Problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--taxa", type=int, default=None)
ap.add_argument("--trees", type=int, default=None)
ap.add_argument("--qmode", type=int, default=None)
ap.add_argument("--thr", type=float, default=None)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--drop-frac", type=float, default=None, help="every gene tree misses this fraction of the taxa (general / packed kernel)")
ap.add_argument("--devices", type=int, default=1, help="GPUs of this process driven by ONE handle (camus_b200_create n_gpus)")
a = ap.parse_args()

s, c = synth.make_config(a.config, a.taxa, a.trees, a.drop_frac)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
q = c["qmode"] if a.qmode is None else a.qmode
t = c["threshold"] if a.thr is None else a.thr
g = (CamusB200(devices=list(range(a.devices))) if a.devices > 1 else CamusB200(0)).load(prob)
g.set_score_hint(c["as_set"])
ns, sc = g.count_filter(q, t)
import time  # noqa: E402
acc = {}
wall = 0.0
for _ in range(a.reps):
    t0 = time.perf_counter()
    g.run_resident(q, t, c["as_set"])
    wall += (time.perf_counter() - t0) * 1e3 / a.reps
    ms, nl = g.stage_ms()
    for k, v in ms.items():
        acc[k] = acc.get(k, 0.0) + v / a.reps
R = g.workload()["resolutions"]
tot = sum(acc.values())
print(f"devices={a.devices} wall={wall:.3f}ms/step  {R / (wall * 1e-3):.3e} quartets/s")
print(f"N={prob.n_taxa} G={prob.n_trees} q={q} t={t} survivors={ns} launches={nl} debug={os.environ.get('CAMUS_B200_DEBUG', '0')}")
print("  " + "  ".join(f"{k}={v:.3f}ms" for k, v in acc.items() if v > 0) + f"  total={tot:.3f}ms  {R / tot / 1e9:.1f} Gquartets/ms-equivalent={R / (tot * 1e-3):.3e}/s")
