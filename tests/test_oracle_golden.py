"""Pins the CPU oracle (literal mode) on every known-answer test and golden file the reference
holds for the hot path, then checks the oracle's fast mode against literal."""
import os
from collections import Counter

import numpy as np
import pytest

from camus_b200.hostlib import CamusError, Problem, read_text
from oracle import oracle as orc
from tests.golden import ref_kats as K

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


def qkey(prob, quad):
    a, b, c, d = (prob.tip_id(x) for x in quad)
    return orc.pack_quartet(a, b, c, d)


def survivors(o):
    keys, counts = o.export_quartets()
    return {int(k): int(c) for k, c in zip(keys, counts)}


# ---------------------------------------------------------------- packing (quartet.go:67-109)
def test_new_quartet_and_compare():
    prob = Problem("(((a,c),(b,d)),f);", "")
    q = qkey(prob, ("a", "b", "c", "d"))
    taxa, topo = orc.unpack_quartet(q)
    assert taxa == [0, 1, 2, 3] and topo == 0b1100  # a,b on taxon-0's side -> bits 2,3 set
    # every ordering / side order of the same split packs to the same key
    for quad in [("b", "a", "d", "c"), ("c", "d", "a", "b"), ("d", "c", "b", "a")]:
        assert qkey(prob, quad) == q
    for q1, q2, res in K.COMPARE:
        k1, k2 = qkey(prob, q1), qkey(prob, q2)
        t1, p1 = orc.unpack_quartet(k1)
        t2, p2 = orc.unpack_quartet(k2)
        got = "diff" if t1 != t2 else ("eq" if (p1 ^ p2) in (0, 0b1111) else "neq")
        assert got == res
    assert orc.unpack_quartet(qkey(prob, ("a", "c", "b", "d")))[1] == 0b1010
    assert orc.unpack_quartet(qkey(prob, ("a", "d", "b", "c")))[1] == 0b0110


@pytest.mark.parametrize("name,tre,expected", K.QUARTETS_FROM_TREE)
@pytest.mark.parametrize("literal", [True, False])
def test_quartets_from_tree(name, tre, expected, literal):
    prob = Problem(tre, tre)  # the tree against itself
    off = prob.gene_node_off
    keys = orc.tree_quartets(prob.gene_parent[off[0]:off[1]], prob.gene_taxon[off[0]:off[1]], literal)
    assert sorted(int(k) for k in keys) == sorted(qkey(prob, q) for q in expected)


# ------------------------------------------- counting + filter (preprocess_test.go:130-283)
@pytest.mark.parametrize("name,tre,qmode,thr,genes,expected", K.PROCESS_QUARTETS)
@pytest.mark.parametrize("literal", [True, False])
def test_process_quartets(name, tre, qmode, thr, genes, expected, literal):
    prob = Problem(tre, "\n".join(genes))
    o = orc.Oracle(2).load(prob)
    o.count_filter(qmode, thr, literal)
    want = Counter(qkey(prob, q) for q in expected)
    assert survivors(o) == dict(want)


@pytest.mark.parametrize("name,tre,genes,err", K.PREPROCESS_ERRORS)
def test_preprocess_errors(name, tre, genes, err):
    if err is None:
        Problem(tre, "\n".join(genes))
    else:
        with pytest.raises(CamusError) as e:
            Problem(tre, "\n".join(genes))
        assert e.value.name == err


def test_threshold_keep():
    # quartetfilter.go:67-74 compares the two least frequent topologies, truncating t*sum to uint32
    assert orc.threshold_keep(0.0, 5, 1, 9)
    assert not orc.threshold_keep(0.0, 3, 3, 9)
    assert not orc.threshold_keep(1.0, 0, 7, 9)
    assert orc.threshold_keep(0.5, 1, 4, 100)       # uint32(2.5)=2 < 3
    assert not orc.threshold_keep(0.5, 1, 3, 100)   # uint32(2.0)=2 < 2 false


# --------------------------------------------------------- tree tables (treedata_test.go)
def test_tree_data():
    prob = Problem(K.TREEDATA_TREE, "")
    o = orc.Oracle(1).load(prob)
    for anc, pairs in K.TREEDATA_LCA.items():
        for a, b in pairs:
            assert o.lca(prob.node_id(a), prob.node_id(b)) == prob.node_id(anc)
            assert o.lca(prob.node_id(b), prob.node_id(a)) == prob.node_id(anc)
    quad, sets = K.TREEDATA_QSETS
    o.set_quartets([qkey(prob, quad)], [1])
    for lab, cnt in sets.items():  # the reference only asserts inclusion (treedata_test.go:194-219)
        assert o.num_quartets_at(prob.node_id(lab)) >= cnt
    # treedata.go:213: ">= 3 taxa below the node" -> A,B,C under b already qualifies
    assert [o.num_quartets_at(prob.node_id(x)) for x in "abcr"] == [0, 1, 1, 1]
    assert prob.node_id("r") == 0  # preorder ids, root = 0


# ------------------------------------------------------------- edges (quartettotals_test.go)
def test_should_calc_edge_and_cycle_length():
    prob = Problem(K.SHOULD_CALC_TREE, "")
    o = orc.Oracle(1).load(prob)
    for u, w, exp in K.SHOULD_CALC_EDGE:
        assert o.should_calc_edge(prob.node_id(u), prob.node_id(w)) == exp
        assert prob.should_calc_edge(prob.node_id(u), prob.node_id(w)) == exp
    for u, w, exp in K.CYCLE_LENGTH:
        assert o.cycle_length(prob.node_id(u), prob.node_id(w)) == exp
        assert prob.cycle_length(prob.node_id(u), prob.node_id(w)) == exp


def _oracle_with(tree_and_q):
    tre, qs = tree_and_q
    prob = Problem(tre, "")
    o = orc.Oracle(2).load(prob)
    o.set_quartets([qkey(prob, q) for q, _ in qs], [c for _, c in qs])
    return prob, o


@pytest.mark.parametrize("literal", [True, False])
def test_quartets_total(literal):
    for which in ("A", "B"):
        prob, o = _oracle_with(K.TOTALS_TREE_A if which == "A" else K.TOTALS_TREE_B)
        tot = o.quartet_totals(False, literal)
        for w_, _, u, w, want in K.QUARTETS_TOTAL:
            if w_ == which:
                assert tot[prob.node_id(u), prob.node_id(w)] == want
        # quartettotals_test.go:12-69: zero wherever !ShouldCalcEdge, some positive entry
        n = prob.n_nodes
        for u in range(n):
            for w in range(n):
                if not o.should_calc_edge(u, w):
                    assert tot[u, w] == 0
        assert (tot > 0).any()


def test_quartet_score_enum():
    names = {"eq": orc.QEQ, "neq": orc.QNEQ, "diff": orc.QDIFF}
    for tre, quad, u, w, want in K.QUARTET_SCORE:
        prob = Problem(tre, "")
        o = orc.Oracle(1).load(prob)
        q = qkey(prob, quad)
        o.set_quartets([q], [1])
        assert o.quartet_score(q, prob.node_id(u), prob.node_id(w)) == names[want]


def test_penalties():
    prob = Problem(K.PENALTY_TREE, "")
    o = orc.Oracle(1).load(prob)
    pen = o.edge_penalties()
    for u, w, want in K.PENALTIES:
        ui, wi = prob.node_id(u), prob.node_id(w)
        assert o.calculate_penalty(ui, wi) == want
        # the table only holds pairs passing ShouldCalcEdge (penalty.go:21-23)
        assert pen[ui, wi] == (want if o.should_calc_edge(ui, wi) else 0)


# ------------------------------------------------------- end to end (infer_test.go:17-301)
def _infer(prob, qmode, thr, mode, alpha, literal=True):
    o = orc.Oracle().load(prob)
    n_surv, sum_counts = o.count_filter(qmode, thr, literal)
    as_set = mode == "sym"  # infer.go:84-85: sym always runs AsSet(true)
    tot = o.quartet_totals(as_set, literal)
    pen = o.edge_penalties() if mode != "max" else None
    return prob.run_dp(tot, pen, mode, as_set, n_surv, sum_counts, prob.n_trees, alpha), tot, (n_surv, sum_counts)


@pytest.mark.parametrize("name,tre,genes,nedges,network", K.INFER)
def test_infer_small(name, tre, genes, nedges, network):
    prob = Problem(tre, "\n".join(genes))
    res, _, _ = _infer(prob, 0, 0.0, "max", 0.0)
    assert len(res.branches) == nedges
    for i, b in enumerate(res.branches):
        assert len(b) == i + 1
    assert res.newicks[-1] == network


@pytest.fixture(scope="module")
def large_problem():
    return Problem.from_files(os.path.join(REF, "infer_constraint.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))


@pytest.mark.parametrize("qmode,thr,mode,alpha,nnets,golden", K.INFER_LARGE)
def test_infer_large(large_problem, qmode, thr, mode, alpha, nnets, golden):
    prob = large_problem
    assert prob.n_taxa == 22 and prob.n_trees == 1123
    res, tot, (n_surv, sum_counts) = _infer(prob, qmode, thr, mode, alpha)
    want = [l.strip() for l in read_text(os.path.join(REF, golden)).split("\n") if l.strip()]
    assert len(want) == nnets
    assert len(res.newicks) == nnets
    assert res.newicks == want
    # survivor counts reported in SURVEY.md §4 (probe) for this dataset
    if qmode == 0:
        assert n_surv == 14423
    else:
        assert n_surv == 1150
    if mode == "max":
        roots = [int(x) for x in res.root_scores]
        assert roots == ([0, 162870, 210460, 221365, 224988, 228156] if qmode == 0 else [0, 92680, 115242, 116736, 117507])


@pytest.mark.parametrize("qmode,thr", [(0, 0.0), (1, 0.3), (2, 0.5), (2, 0.0)])
def test_fast_equals_literal_on_dataset(large_problem, qmode, thr):
    prob = large_problem
    a = orc.Oracle().load(prob)
    b = orc.Oracle().load(prob)
    ra = a.count_filter(qmode, thr, True)
    rb = b.count_filter(qmode, thr, False)
    assert ra == rb
    for x, y in zip(a.export_raw_counts(), b.export_raw_counts()):
        assert np.array_equal(x, y)
    for x, y in zip(a.export_quartets(), b.export_quartets()):
        assert np.array_equal(x, y)
    for as_set in (False, True):
        assert np.array_equal(a.quartet_totals(as_set, True), b.quartet_totals(as_set, False))
    assert a.literal_emissions() > 0


def test_dp_is_thread_count_invariant(monkeypatch):
    """The host DP runs scoreEdgesAcross on a thread pool and replays the reference's visiting
    order when it accepts candidates: 1 thread and all threads give identical networks."""
    from camus_b200 import synth
    s = synth.make(90, 50, 3)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    ns, sc = o.count_filter(2, 0.5, False)
    tot, pen = o.quartet_totals(False, False), o.edge_penalties()
    out = {}
    for nt in ("1", "3", "0"):
        monkeypatch.setenv("CAMUS_HOST_THREADS", nt)
        out[nt] = [prob.run_dp(tot, pen if m != "max" else None, m, False, ns, sc, prob.n_trees, 0.3) for m in ("max", "norm", "sym")]
    for nt in ("3", "0"):
        for a, b in zip(out["1"], out[nt]):
            assert a.newicks == b.newicks and a.branches == b.branches and a.qsat == b.qsat
            assert [repr(x) for x in a.root_scores] == [repr(x) for x in b.root_scores]
    assert len(out["1"][0].newicks) >= 3


@pytest.mark.parametrize("kind", ["oracle", "sparse_small", "dense_ties", "wide"])
def test_dp_maxplus_scan_equals_literal_scan(monkeypatch, kind):
    """For integer scores (max mode) the across-scan takes the candidate VALUES from a max-plus
    product and the split indices from the reference's FourWayBestSplit (dp.hpp, fast_across);
    CAMUS_DP_LITERAL=1 runs the literal O(k^2) scan. Same branches, same order, same networks: on
    an oracle table and on synthetic score tables built for ties (the DP accepts any u64 table)."""
    from camus_b200 import synth
    rng = np.random.default_rng(11)
    if kind == "oracle":
        s = synth.make(64, 20, 5)
        prob = Problem(s.constraint, s.genes)
        o = orc.Oracle().load(prob)
        ns, sc = o.count_filter(0, 0.0, False)
        tables = [o.quartet_totals(False, False), o.quartet_totals(True, False)]
    else:
        s = synth.make({"sparse_small": 120, "dense_ties": 48, "wide": 150}[kind], 1, 9)
        prob = Problem(s.constraint, s.genes)
        n = prob.n_nodes
        ns, sc = 1000, 100000
        tables = []
        for rep in range(2):
            if kind == "sparse_small":  # few scoring edges, tiny equal weights
                t = (rng.random((n, n)) < 0.02) * rng.integers(1, 3, (n, n))
            elif kind == "dense_ties":  # every edge scores 0..2
                t = rng.integers(0, 3, (n, n))
            else:  # a few heavy edges over a floor of ones
                t = np.ones((n, n), np.int64) + (rng.random((n, n)) < 0.001) * rng.integers(1, 10**9, (n, n))
            tables.append(np.ascontiguousarray(t, np.uint64))
    for tot in tables:
        out = {}
        for lit in ("1", "0"):
            monkeypatch.setenv("CAMUS_DP_LITERAL", lit)
            out[lit] = prob.run_dp(tot, None, "max", False, ns, sc, 1, 0.3)
        a, b = out["1"], out["0"]
        assert a.branches == b.branches and a.newicks == b.newicks and a.qsat == b.qsat
        assert [repr(x) for x in a.root_scores] == [repr(x) for x in b.root_scores]
        assert len(a.newicks) >= 3
