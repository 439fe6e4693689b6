#!/usr/bin/env python
"""Developer tool: device memory a single share of an N-way split allocates (fresh context).
  python tools/share_memory.py [--config 3] [--world 8] [--rank 0]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from camus_b200 import synth  # noqa: E402
from camus_b200.cabi import CamusB200  # noqa: E402
from camus_b200.hostlib import Problem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=0)
a = ap.parse_args()
s, c = synth.make_config(a.config, None, None)
prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
torch.cuda.init()
free0, total = torch.cuda.mem_get_info(0)
g = CamusB200(0)
g.set_partition(a.rank, a.world)
g.load(prob)
g.set_score_hint(c["as_set"])
ns, sc = g.count_filter(c["qmode"], c["threshold"])
free1, _ = torch.cuda.mem_get_info(0)
print(json.dumps({"config": a.config, "share": f"{a.rank}/{a.world}", "scheme": os.environ.get("CAMUS_B200_PARTITION", "default"),
                  "device_gb_allocated_by_the_share": round((free0 - free1) / 1e9, 2), "n_survivors": ns}))
g.close()
