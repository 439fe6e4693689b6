#!/usr/bin/env python
"""Generates tests/golden/synthetic/vectors.json: digests of the hot path's outputs on seeded
synthetic inputs (camus_b200.synth), computed by the CPU oracle in LITERAL mode (the reference's
algorithms restated, oracle/camus_oracle.cpp). The committed file pins both the oracle (drift
check, tests/test_synthetic_golden.py on CPU) and the CUDA path (same test, -m gpu) on shapes the
reference's own fixtures do not reach: 30-50 taxa, missing taxa, support collapse, NEXUS, -asSet.

  python tests/golden/make_synthetic.py        # rewrites vectors.json (run in the build container)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = [
    # name, synth.make kwargs, min_support, qmode, threshold, as_set
    ("complete_30x60_q2", dict(n_taxa=30, n_trees=60, config_id=101), 0.0, 2, 0.5, False),
    ("missing_36x50_q1", dict(n_taxa=36, n_trees=50, config_id=102, drop_frac=0.15), 0.0, 1, 0.3, False),
    ("nexus_support_40x40_q2_asset", dict(n_taxa=40, n_trees=40, config_id=103, support=True, drop_frac=0.1, nexus=True), 0.7, 2, 0.5, True),
    ("config1_50x200_q2", dict(n_taxa=50, n_trees=200, config_id=1), 0.0, 2, 0.5, False),
    ("q0_33x33", dict(n_taxa=33, n_trees=33, config_id=104, drop_frac=0.05), 0.0, 0, 0.0, False),
]


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def digest(backend, prob, qmode, thr, as_set):
    """backend: an object with the C-ABI call sequence (oracle.Oracle or cabi.CamusB200), already loaded."""
    ns, sc = backend.count_filter_plain(qmode, thr)
    keys, counts = backend.export_quartets()
    tot = backend.totals_plain(as_set)
    pen = backend.edge_penalties()
    return {"n_survivors": int(ns), "sum_counts": int(sc), "quartets_sha256": sha(keys.astype(np.uint64), counts.astype(np.uint32)),
            "totals_sha256": sha(tot.astype(np.uint64)), "totals_sum": int(tot.sum()), "penalties_sha256": sha(pen.astype(np.uint64))}


class OracleBackend:
    def __init__(self, prob, literal=True):
        from oracle import oracle as orc
        self.o = orc.Oracle().load(prob)
        self.literal = literal

    def count_filter_plain(self, qmode, thr):
        return self.o.count_filter(qmode, thr, self.literal)

    def export_quartets(self):
        return self.o.export_quartets()

    def totals_plain(self, as_set):
        return self.o.quartet_totals(as_set, self.literal)

    def edge_penalties(self):
        return self.o.edge_penalties()


def make_problem(kwargs, min_support):
    from camus_b200 import synth
    from camus_b200.hostlib import Problem
    s = synth.make(**kwargs)
    return Problem(s.constraint, s.genes, s.fmt, min_support)


if __name__ == "__main__":
    out = {}
    for name, kw, ms, qmode, thr, as_set in CASES:
        prob = make_problem(kw, ms)
        out[name] = digest(OracleBackend(prob, literal=True), prob, qmode, thr, as_set)
        print(name, out[name]["n_survivors"], out[name]["totals_sum"])
    with open(os.path.join(HERE, "synthetic", "vectors.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
