#!/usr/bin/env python
"""Writes tests/golden/synthetic/config_digests.json: n_survivors, sum_counts and digests of the
edge tables for the synthetic configurations of BASELINE.json.

  python tools/make_config_digests.py            configs 1, 2, 4 (256 trees), 5 from the ORACLE (dense mode, CPU)
  python tools/make_config_digests.py --gpu      adds config 3 (1000 x 1000) from the CUDA path: no CPU
                                                 algorithm can produce that table; the entry is marked
                                                 source = "gpu" and tests/test_gpu_parity_large.py checks
                                                 sampled entries of the same run against the literal
                                                 QuartetScore loop before comparing digests.
bench.py asserts the digests of the configuration it runs; the GPU tests assert all of them.
"""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from camus_b200 import synth  # noqa: E402
from camus_b200.hostlib import Problem  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "synthetic", "config_digests.json")


def table_digest(tab):
    return "%08x-%016x" % (zlib.crc32(np.ascontiguousarray(tab).tobytes()), int(tab.sum(dtype=np.uint64)))


def main():
    d = json.load(open(OUT)) if os.path.exists(OUT) else {}
    if "--gpu" in sys.argv:
        from camus_b200.cabi import CamusB200
        s, c = synth.make_config(3)
        prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
        gpu = CamusB200(0).load(prob)
        ns, sc = gpu.count_filter(c["qmode"], c["threshold"])
        tab = gpu.quartet_totals(False)
        gpu.set_score_hint(True)
        gpu.count_filter(c["qmode"], c["threshold"])
        tab_set = gpu.quartet_totals(True)
        d["config3"] = {"n_survivors": ns, "sum_counts": sc, "quartet_totals": table_digest(tab),
                        "quartet_totals_as_set": table_digest(tab_set)}
        d["_source"]["config3"] = "gpu (sampled entries checked against the literal QuartetScore loop by tests/test_gpu_parity_large.py)"
    else:
        from oracle import oracle as orc
        d.setdefault("_source", {})
        for cid, kw in ((1, {}), (2, {}), (4, {"n_trees": 256}), (5, {})):
            s, c = synth.make_config(cid, **kw)
            prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
            o = orc.Oracle().load(prob)
            st, tab = o.dense_run(c["qmode"], c["threshold"], False, digests=False)
            _, tab_set = o.dense_run(c["qmode"], c["threshold"], True, digests=False)
            key = "config%d" % cid + ("_%dtrees" % kw["n_trees"] if kw else "")
            d[key] = {"n_survivors": st["n_survivors"], "sum_counts": st["sum_counts"], "quartet_totals": table_digest(tab),
                      "quartet_totals_as_set": table_digest(tab_set)}
            d["_source"][key] = "oracle dense mode (CPU)"
            print(key, d[key])
    json.dump(d, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
