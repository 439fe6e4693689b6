"""DP offload: the across-scan of infer/main_dp.go:247-287 on the GPU
(camus_b200_dp_* in include/camus_b200.h) against the host mirror of the same loop, which is itself
pinned on the reference's 12 TestInfer cases and 17 golden networks (tests/test_oracle_golden.py).
Bar: identical -- the same branches in the same traceback order, the same DP[root][k] to the last
bit (float64 for norm / sym), the same extended-newick strings."""
import os

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem, read_text
from oracle import oracle as orc
from tests.golden import ref_kats as K

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


def _tables(prob, qmode, thr, as_set):
    o = orc.Oracle().load(prob)
    st, tot = o.dense_run(qmode, thr, as_set, True, False)
    return tot, o.edge_penalties(), st["n_survivors"], st["sum_counts"]


def _same(a, b):
    assert a.branches == b.branches
    assert a.newicks == b.newicks
    assert np.array_equal(np.array(a.root_scores, np.float64).view(np.uint64), np.array(b.root_scores, np.float64).view(np.uint64))
    assert np.array_equal(np.array(a.qsat, np.float64).view(np.uint64), np.array(b.qsat, np.float64).view(np.uint64))


@pytest.mark.parametrize("qmode,thr,mode,alpha,nnets,golden", K.INFER_LARGE)
def test_golden_networks_through_the_device_scan(qmode, thr, mode, alpha, nnets, golden, monkeypatch):
    """The reference's own dataset and golden network files, every step of the DP on the device
    (min_work = 0) with the host scan run beside it (CAMUS_DP_VERIFY=1 throws on any difference in
    winner, score bits or split indices)."""
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    prob = Problem.from_files(os.path.join(REF, "infer_constraint.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))
    as_set = mode == "sym"
    tot, pen, ns, sc = _tables(prob, qmode, thr, as_set)
    res = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, alpha, gpu=0, min_work=0)
    want = [l.strip() for l in read_text(os.path.join(REF, golden)).split("\n") if l.strip()]
    assert res.newicks == want
    assert res.accel["device_steps"] > 0 and res.accel["host_steps"] == 0


@pytest.mark.parametrize("name,tre,genes,nedges,network", K.INFER)
def test_infer_kats_through_the_device_scan(name, tre, genes, nedges, network, monkeypatch):
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    prob = Problem(tre, "\n".join(genes))
    tot, pen, ns, sc = _tables(prob, 0, 0.0, False)
    res = prob.run_dp(tot, None, "max", False, ns, sc, gpu=0, min_work=0)
    assert len(res.branches) == nedges and res.newicks[-1] == network


@pytest.mark.parametrize("taxa,trees,seed", [(40, 60, 7), (90, 120, 8), (160, 80, 9)])
@pytest.mark.parametrize("mode", ["max", "norm", "sym"])
def test_synthetic_device_scan_equals_host_scan(taxa, trees, seed, mode, monkeypatch):
    """Every step on the device, then the default threshold (large steps only), against the host-only
    DP: identical results in all three score modes (ties are frequent in max mode)."""
    s = synth.make(taxa, trees, seed)
    prob = Problem(s.constraint, s.genes)
    as_set = mode == "sym"
    tot, pen, ns, sc = _tables(prob, 2, 0.5, as_set)
    host = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5)
    assert len(host.branches) > 0
    if taxa <= 90:
        monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    dev = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5, gpu=0, min_work=0)
    _same(host, dev)
    assert dev.accel["device_steps"] > 0
    monkeypatch.delenv("CAMUS_DP_VERIFY", raising=False)
    mixed = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5, gpu=0)
    _same(host, mixed)


def test_nan_scores_follow_the_reference_comparisons(monkeypatch):
    """norm mode scores the length-3 cycle between v's two children as 0 / 0 = NaN (scoreEdgesAcross
    does not call ShouldCalcEdge). The reference's `>` and `==` give a NaN a definite fate (kept only
    as the very first candidate); the device scan reproduces it -- also with EVERY score NaN."""
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    s = synth.make(30, 40, 11)
    prob = Problem(s.constraint, s.genes)
    tot, pen, ns, sc = _tables(prob, 0, 0.0, False)
    for p in (pen, np.zeros_like(pen)):  # real penalties; all-zero penalties: every edge score NaN or +Inf
        host = prob.run_dp(tot, p, "norm", False, ns, sc, prob.n_trees, 0.5)
        dev = prob.run_dp(tot, p, "norm", False, ns, sc, prob.n_trees, 0.5, gpu=0, min_work=0)
        _same(host, dev)
        assert dev.accel["nan_steps"] > 0  # steps that saw a NaN score


def test_argument_errors():
    import ctypes as C
    from camus_b200 import cabi
    L = cabi.lib()
    L.camus_b200_dp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.camus_b200_dp_last_error.argtypes = [C.c_void_p]
    L.camus_b200_dp_last_error.restype = C.c_char_p
    h = C.c_void_p()
    assert L.camus_b200_dp_create(C.byref(h), 0, 0, 0, None, None, None) != 0
    assert b"dp_create" in L.camus_b200_dp_last_error(None)
    end = np.array([3, 2, 3], np.int32)
    dep = np.array([0, 1, 1], np.int32)
    edge = np.zeros(9, np.uint64)
    assert L.camus_b200_dp_create(C.byref(h), 99, 0, 3, end.ctypes.data, dep.ctypes.data, edge.ctypes.data) != 0
    assert L.camus_b200_dp_create(C.byref(h), 0, 0, 3, end.ctypes.data, dep.ctypes.data, edge.ctypes.data) == 0
    L.camus_b200_dp_across.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    best = (C.c_uint64 * 8)()
    assert L.camus_b200_dp_across(h, 0, best) != 0  # no vertex set
    assert b"vertex" in L.camus_b200_dp_last_error(h)
    L.camus_b200_dp_set_node.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    assert L.camus_b200_dp_set_node(h, 7, 1, edge.ctypes.data) != 0
    L.camus_b200_dp_destroy.argtypes = [C.c_void_p]
    L.camus_b200_dp_destroy(h)


@pytest.mark.parametrize("taxa,trees,seed", [(60, 80, 17), (160, 80, 9)])
def test_integer_scores_literal_and_maxplus_forms_agree(taxa, trees, seed, monkeypatch):
    """max mode on the device: the max-plus form (camus_b200_dp_across_maxplus, default) and the literal scan
    (CAMUS_DP_ACCEL_LITERAL=1) against the host DP; every step offloaded, host scan beside it where affordable."""
    s = synth.make(taxa, trees, seed)
    prob = Problem(s.constraint, s.genes)
    tot, pen, ns, sc = _tables(prob, 2, 0.5, False)
    host = prob.run_dp(tot, None, "max", False, ns, sc)
    if taxa <= 60:
        monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    fast = prob.run_dp(tot, None, "max", False, ns, sc, gpu=0, min_work=0)
    monkeypatch.setenv("CAMUS_DP_ACCEL_LITERAL", "1")
    lit = prob.run_dp(tot, None, "max", False, ns, sc, gpu=0, min_work=0)
    _same(host, fast)
    _same(host, lit)
    assert fast.accel["device_steps"] > 0 and lit.accel["device_steps"] > 0
