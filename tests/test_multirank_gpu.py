"""Multi-process GPU runs with the collective owned by the library (camus_b200_comm_init: NCCL
all-reduce of the partial junction tables inside libcamus_b200.so). Needs two GPUs
(`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, sys.argv[1])
rank, world, tmp = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
from camus_b200 import synth
from camus_b200.cabi import CamusB200
from camus_b200.hostlib import Problem
idf = os.path.join(tmp, "nccl_id")
if rank == 0:
    uid = CamusB200.comm_unique_id()
    open(idf + ".tmp", "wb").write(uid)
    os.rename(idf + ".tmp", idf)
else:
    for _ in range(600):
        if os.path.exists(idf):
            break
        time.sleep(0.1)
    uid = open(idf, "rb").read()
g = CamusB200(rank)
g.comm_init(uid, rank, world)
res = {}
for name, s, ms in (("a", synth.make(70, 90, 61, drop_frac=0.1), 0.0), ("b", synth.make(96, 64, 62), 0.0)):
    prob = Problem(s.constraint, s.genes, s.fmt, ms)
    g.load(prob)
    ns, sc = g.count_filter(2, 0.5)
    tab = g.quartet_totals(False)
    tab2 = g.quartet_totals(False)          # idempotent: the table is reduced once per scoring pass
    out, ns3, sc3 = g.run_resident(2, 0.5, True)
    assert np.array_equal(tab, tab2) and (ns3, sc3) == (ns, sc)
    res[name + "_scal"] = np.array([ns, sc], np.uint64)
    res[name + "_tab"] = tab
    res[name + "_set"] = out
np.savez(os.path.join(tmp, "rank%d.npz" % rank), **res)
g.close()
"""


def _n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_library_owned_collective_two_processes():
    from camus_b200 import synth
    from camus_b200.cabi import CamusB200
    from camus_b200.hostlib import Problem
    world = min(_n_devices(), 4)
    with tempfile.TemporaryDirectory() as tmp:
        procs = [subprocess.Popen([sys.executable, "-c", WORKER, ROOT, str(r), str(world), tmp], stderr=subprocess.PIPE, text=True)
                 for r in range(world)]
        errs = [p.communicate(timeout=600)[1] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(e[-2000:] for e in errs)
        ranks = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
    single = CamusB200(0)
    for name, s in (("a", synth.make(70, 90, 61, drop_frac=0.1)), ("b", synth.make(96, 64, 62))):
        prob = Problem(s.constraint, s.genes)
        single.load(prob)
        ns, sc = single.count_filter(2, 0.5)
        tab = single.quartet_totals(False)
        single.set_score_hint(True)
        single.count_filter(2, 0.5)
        tab_set = single.quartet_totals(True)
        single.set_score_hint(False)
        for r in ranks:  # every rank holds the global results
            assert tuple(int(x) for x in r[name + "_scal"]) == (ns, sc)
            assert np.array_equal(r[name + "_tab"], tab) and np.array_equal(r[name + "_set"], tab_set)
    single.close()
