#!/usr/bin/env python
"""Developer tool: per-share device times of a multi-rank split, measured on ONE GPU.

  python tools/share_times.py [--config 3] [--world 8] [--reps 2]

Every share of `camus_b200_set_partition(r, world)` is run through the resident hot path one after
the other on cuda:0 (the shares of a real run are independent until the all-reduce of the junction
table), so the prep / count time of each rank of an N-GPU run, the balance of the split and the sum
check against the unpartitioned run cost 1 GPU instead of N. CAMUS_B200_PARTITION=chunk|class picks
the scheme (layout.cuh).
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
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--taxa", type=int, default=None)
ap.add_argument("--trees", type=int, default=None)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--skip-whole", action="store_true")
a = ap.parse_args()

s, c = synth.make_config(a.config, a.taxa, a.trees)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
q, t, as_set = c["qmode"], c["threshold"], c["as_set"]
g = CamusB200(0).load(prob)
g.set_score_hint(as_set)
whole = g.count_filter(q, t)
rows = []


def run(label):
    best = None
    for _ in range(a.reps):
        t0 = time.perf_counter()
        _, ns, sc = g.run_resident(q, t, as_set, None, partial_only=True)
        wall = (time.perf_counter() - t0) * 1e3
        ms, nl = g.stage_ms()
        if best is None or wall < best["wall_ms"]:
            best = {"share": label, "wall_ms": round(wall, 3), "prep_ms": round(ms["prep"], 3), "count_ms": round(ms["count"], 3),
                    "launches": int(nl), "n_survivors": ns, "sum_counts": sc}
    try:  # device memory in use by this process (the bit-planes dominate)
        import torch
        free_b, total_b = torch.cuda.mem_get_info(0)
        best["device_used_gb"] = round((total_b - free_b) / 1e9, 2)
    except Exception:
        pass
    rows.append(best)
    print(json.dumps(best), flush=True)
    return best


if not a.skip_whole:
    run("1/1")
ns = sc = 0
for r in range(a.world):
    g.set_partition(r, a.world)
    b = run(f"{r}/{a.world}")
    ns, sc = ns + b["n_survivors"], sc + b["sum_counts"]
shares = [x for x in rows if x["share"] != "1/1"]
print(json.dumps({
    "world": a.world, "scheme": os.environ.get("CAMUS_B200_PARTITION", "default"), "sum_matches_whole": (ns, sc) == tuple(whole),
    "max_wall_ms": max(x["wall_ms"] for x in shares), "max_prep_ms": max(x["prep_ms"] for x in shares),
    "max_count_ms": max(x["count_ms"] for x in shares), "mean_count_ms": round(sum(x["count_ms"] for x in shares) / len(shares), 3),
    "whole_wall_ms": rows[0]["wall_ms"] if not a.skip_whole else None}))
g.close()
