"""Host side of the DP offload (camus_b200/host/dp.hpp) without a GPU: the `-m "not gpu"` suite drives it through a
CPU stand-in of the camus_b200_dp_* entry points (camus_b200/host/dp_mock.hpp) that, like the kernels, evaluates
every pair literally and reduces with the CLOSED FORM of the reference's sequential acceptance (main_dp.go:199-214):
first maximum over w in any visiting order; across the u maximum score / smallest CycleLength / last u; NaN scores
kept only as the very first candidate. CAMUS_DP_VERIFY=1 runs the literal loop beside every step. The CUDA kernels
themselves are checked in tests/test_dp_offload.py (GPU)."""
import os

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem, read_text
from oracle import oracle as orc
from tests.golden import ref_kats as K

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


def _tables(prob, qmode, thr, as_set):
    o = orc.Oracle().load(prob)
    st, tot = o.dense_run(qmode, thr, as_set, True, False)
    return tot, o.edge_penalties(), st["n_survivors"], st["sum_counts"]


def _same(a, b):
    assert a.branches == b.branches and a.newicks == b.newicks
    assert np.array_equal(np.array(a.root_scores, np.float64).view(np.uint64), np.array(b.root_scores, np.float64).view(np.uint64))
    assert np.array_equal(np.array(a.qsat, np.float64).view(np.uint64), np.array(b.qsat, np.float64).view(np.uint64))


@pytest.mark.parametrize("qmode,thr,mode,alpha,nnets,golden", K.INFER_LARGE)
def test_golden_networks_through_the_offload_interface(qmode, thr, mode, alpha, nnets, golden, monkeypatch):
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    prob = Problem.from_files(os.path.join(REF, "infer_constraint.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))
    as_set = mode == "sym"
    tot, pen, ns, sc = _tables(prob, qmode, thr, as_set)
    res = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, alpha, gpu="mock", min_work=0)
    want = [l.strip() for l in read_text(os.path.join(REF, golden)).split("\n") if l.strip()]
    assert res.newicks == want
    assert res.accel["device_steps"] > 0 and res.accel["host_steps"] == 0


@pytest.mark.parametrize("name,tre,genes,nedges,network", K.INFER)
def test_infer_kats_through_the_offload_interface(name, tre, genes, nedges, network, monkeypatch):
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    prob = Problem(tre, "\n".join(genes))
    tot, pen, ns, sc = _tables(prob, 0, 0.0, False)
    res = prob.run_dp(tot, None, "max", False, ns, sc, gpu="mock", min_work=0)
    assert len(res.branches) == nedges and res.newicks[-1] == network


@pytest.mark.parametrize("taxa,trees,seed", [(24, 40, 3), (45, 60, 4)])
@pytest.mark.parametrize("mode", ["max", "norm", "sym"])
def test_synthetic_offloaded_dp_equals_host_dp(taxa, trees, seed, mode, monkeypatch):
    s = synth.make(taxa, trees, seed)
    prob = Problem(s.constraint, s.genes)
    as_set = mode == "sym"
    tot, pen, ns, sc = _tables(prob, 2, 0.5, as_set)
    host = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5)
    assert len(host.branches) > 0
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    _same(host, prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5, gpu="mock", min_work=0))
    monkeypatch.delenv("CAMUS_DP_VERIFY")
    mixed = prob.run_dp(tot, pen, mode, as_set, ns, sc, prob.n_trees, 0.5, gpu="mock", min_work=400)  # small steps stay on the host
    _same(host, mixed)
    assert mixed.accel["host_steps"] > 0 and mixed.accel["device_steps"] + mixed.accel["host_steps"] > 10


def test_nan_scores_keep_the_reference_semantics(monkeypatch):
    """norm mode: 0 / 0 on the length-3 cycle between v's children; then EVERY edge score NaN or +Inf."""
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    s = synth.make(30, 40, 11)
    prob = Problem(s.constraint, s.genes)
    tot, pen, ns, sc = _tables(prob, 0, 0.0, False)
    for p in (pen, np.zeros_like(pen)):
        host = prob.run_dp(tot, p, "norm", False, ns, sc, prob.n_trees, 0.5)
        dev = prob.run_dp(tot, p, "norm", False, ns, sc, prob.n_trees, 0.5, gpu="mock", min_work=0)
        _same(host, dev)
        assert dev.accel["nan_steps"] > 0


def test_ties_everywhere(monkeypatch):
    """A constant edge table: every comparison of the acceptance loop is a tie, so the result is decided by the
    tie-breaks alone (first w, smallest cycle length, last u)."""
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    s = synth.make(26, 10, 21)
    prob = Problem(s.constraint, s.genes)
    n = prob.n_nodes
    tot = np.full((n, n), 7, np.uint64)
    pen = np.full((n, n), 3, np.uint64)
    for mode in ("max", "norm", "sym"):
        host = prob.run_dp(tot, pen, mode, False, 100, 1000, prob.n_trees, 0.5)
        dev = prob.run_dp(tot, pen, mode, False, 100, 1000, prob.n_trees, 0.5, gpu="mock", min_work=0)
        _same(host, dev)


def test_integer_scores_literal_and_maxplus_forms_agree(monkeypatch):
    """max mode: the offload takes the max-plus form by default (conv columns, split indices of the winner from the
    reference's FourWayBestSplit); CAMUS_DP_ACCEL_LITERAL=1 sends integers through the literal scan as well."""
    monkeypatch.setenv("CAMUS_DP_VERIFY", "1")
    s = synth.make(40, 50, 31)
    prob = Problem(s.constraint, s.genes)
    tot, pen, ns, sc = _tables(prob, 2, 0.5, False)
    host = prob.run_dp(tot, None, "max", False, ns, sc)
    fast = prob.run_dp(tot, None, "max", False, ns, sc, gpu="mock", min_work=0)
    monkeypatch.setenv("CAMUS_DP_ACCEL_LITERAL", "1")
    lit = prob.run_dp(tot, None, "max", False, ns, sc, gpu="mock", min_work=0)
    _same(host, fast)
    _same(host, lit)
