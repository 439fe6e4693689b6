"""GPU parity: the CUDA path, called through the C-ABI (include/camus_b200.h), against the oracle
and the reference's golden vectors. Bit-exact: everything on this path is integer."""
import os
from collections import Counter

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem, read_text
from oracle import oracle as orc
from tests.golden import ref_kats as K

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


@pytest.fixture(scope="module")
def gpu():
    from camus_b200.cabi import CamusB200
    g = CamusB200(0)
    yield g
    g.close()


def qkey(prob, quad):
    return orc.pack_quartet(*(prob.tip_id(x) for x in quad))


def as_dict(kc):
    return {int(k): int(c) for k, c in zip(*kc)}


def compare_all(gpu, prob, qmode, thr, literal=False, as_sets=(False, True), penalties=True):
    """count/filter scalars, raw counts, survivors, totals (both weights), penalties vs the oracle."""
    o = orc.Oracle().load(prob)
    want = o.count_filter(qmode, thr, literal)
    gpu.load(prob)
    got = gpu.count_filter(qmode, thr)
    assert got == want
    gk, gc = gpu.export_quartets(raw=False)
    ok, oc = o.export_quartets()
    assert np.array_equal(gk, ok) and np.array_equal(gc, oc)
    for as_set in as_sets:
        assert np.array_equal(gpu.quartet_totals(as_set), o.quartet_totals(as_set, literal)), f"totals as_set={as_set}"
    gk, gc = gpu.export_quartets(raw=True)
    ok, oc = o.export_raw_counts()
    assert np.array_equal(gk, ok) and np.array_equal(gc, oc)
    # totals must survive the raw export (recount + refilter)
    assert np.array_equal(gpu.quartet_totals(False), o.quartet_totals(False, literal))
    if penalties:
        assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())
    return o


@pytest.mark.parametrize("qmode,thr", [(0, 0.0), (2, 0.5)])
def test_fused_and_table_pipelines_agree(gpu, qmode, thr, monkeypatch):
    """count_filter runs the fused kernel (counts never leave registers); CAMUS_B200_UNFUSED=1
    selects the K2 -> K3 -> K4 pipeline over a cell table in HBM. Same results, both vs the oracle."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_UNFUSED", "1")
    table = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_UNFUSED")
    for s, ms in ((synth.make(60, 100, 4, support=True, drop_frac=0.1, nexus=True), 0.7), (synth.make_config(1)[0], 0.0)):
        prob = Problem(s.constraint, s.genes, s.fmt, ms)
        o = orc.Oracle().load(prob)
        want = o.count_filter(qmode, thr, False)
        for ctx in (gpu, table):
            ctx.load(prob)
            for hint in (False, True):
                ctx.set_score_hint(hint)
                assert ctx.count_filter(qmode, thr) == want
                for as_set in (True, False):
                    assert np.array_equal(ctx.quartet_totals(as_set), o.quartet_totals(as_set, False))
        gpu.set_score_hint(False)
        ms_f, _ = gpu.stage_ms()
        ms_t, _ = table.stage_ms()
        assert ms_f["filter"] == 0 and ms_f["score"] == 0  # fused: one kernel
    table.close()


def test_packed_and_three_register_general_kernels_agree(gpu, monkeypatch):
    """Incomplete / polytomous gene trees: up to 1023 of them take k_count<kModePacked> (three 10-bit
    counters in one register, 3 CTAs per SM), more -- or CAMUS_B200_NO_PACKED=1 -- the 3-register
    general kernel. Same survivors and tables from both, every filter mode, against the oracle; and
    the 1023 / 1024 tree boundary (a count of 1023 must not carry into the next field)."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_NO_PACKED", "1")
    wide = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_PACKED")
    cases = [(synth.make(41, 20, 51, drop_frac=0.1), 0.0), (synth.make(66, 300, 53, drop_frac=0.1), 0.0),
             (synth.make(58, 90, 54, support=True, drop_frac=0.1, nexus=True), 0.7), (synth.make(30, 33, 55, drop_frac=0.2), 0.0)]
    for s, ms in cases:
        prob = Problem(s.constraint, s.genes, s.fmt, ms)
        o = orc.Oracle().load(prob)
        for qmode, thr in ((0, 0.0), (1, 0.3), (2, 0.5)):
            want = o.count_filter(qmode, thr, False)
            for as_set in (False, True):
                ref = o.quartet_totals(as_set, False)
                for ctx in (gpu, wide):
                    ctx.load(prob)
                    ctx.set_score_hint(as_set)
                    assert ctx.count_filter(qmode, thr) == want
                    assert np.array_equal(ctx.quartet_totals(as_set), ref)
    gpu.set_score_hint(False)
    # boundary: 1023 identical trees + one with a missing taxon (packed), then one more tree (3-register)
    base = synth.make(20, 1, 56)
    one = base.genes.strip()
    dropped = Problem(base.constraint, synth.make(20, 1, 56, drop_frac=0.3).genes)
    for n_same in (1022, 1023):
        prob = Problem(base.constraint, "\n".join([one] * n_same) + "\n")
        for ctx in (gpu, wide):
            ctx.load(prob)
            ctx.add_gene_trees(dropped.gene_node_off, dropped.gene_parent, dropped.gene_taxon)
        got = gpu.count_filter(0, 0.0)
        assert got == wide.count_filter(0, 0.0)
        gk, gc = gpu.export_quartets(raw=True)
        assert int(gc.max()) >= n_same
        assert np.array_equal(gpu.quartet_totals(False), wide.quartet_totals(False))
    wide.close()


def test_plane_builders_agree(gpu, monkeypatch):
    """Three builders of the rooted-triple bit-planes: k_triple_planes_bs (bit-sliced depths, the default; slices
    straight from the trees, or from the depth table with CAMUS_B200_NO_DIRECT=1),
    k_triple_planes_h2 (packed fp16 mask compares, NaN = absent; CAMUS_B200_NO_SLICED=1) and the integer
    ballot kernel trees deeper than 2046 fall back to (CAMUS_B200_NO_HALF=1). All against the oracle."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_NO_HALF", "1")
    integer = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_HALF")
    monkeypatch.setenv("CAMUS_B200_NO_SLICED", "1")
    fp16 = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_SLICED")
    monkeypatch.setenv("CAMUS_B200_NO_DIRECT", "1")  # slices made from the depth table (inputs above 6400 taxa)
    via_table = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_DIRECT")
    for s, ms in ((synth.make(52, 70, 8, support=True, drop_frac=0.12, nexus=True), 0.6), (synth.make(37, 40, 9), 0.0)):
        prob = Problem(s.constraint, s.genes, s.fmt, ms)
        o = orc.Oracle().load(prob)
        want = o.count_filter(0, 0.0, False)
        for ctx in (gpu, via_table, fp16, integer):
            ctx.load(prob)
            assert ctx.count_filter(0, 0.0) == want
            for a, b in zip(ctx.export_quartets(raw=True), o.export_raw_counts()):
                assert np.array_equal(a, b)
            assert np.array_equal(ctx.quartet_totals(False), o.quartet_totals(False, False))
    integer.close()
    fp16.close()
    via_table.close()


def _deepen(genes: str, pad: int) -> str:
    """wrap the first leaf of every newick line in `pad` unifurcations: same quartets, deeper tree"""
    import re
    out = []
    for line in genes.strip().splitlines():
        m = re.search(r"t\d{4}", line)
        out.append(line[:m.start()] + "(" * pad + m.group(0) + ")" * pad + line[m.end():])
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("pad", [0, 25, 40, 60, 300, 1500, 2100])
def test_sliced_plane_builder_depth_classes_general(gpu, pad):
    """k_triple_planes_bs is compiled for 5, 6, 7, 9 and 11 depth bits (the smallest that holds depth + 1 is
    picked per input); deeper trees (> 2046) take the integer kernel. Incomplete trees with polytomies
    (presence masks, three compared planes), 70 trees = two full groups + 6."""
    s = synth.make(26, 70, 77, support=True, drop_frac=0.15)
    prob = Problem(s.constraint, _deepen(s.genes, pad), s.fmt, 0.5)
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    assert gpu.count_filter(0, 0.0) == o.count_filter(0, 0.0, False)
    for a, b in zip(gpu.export_quartets(raw=True), o.export_raw_counts()):
        assert np.array_equal(a, b)
    assert np.array_equal(gpu.quartet_totals(False), o.quartet_totals(False, False))
    assert gpu.count_filter(2, 0.4) == o.count_filter(2, 0.4, False)
    assert np.array_equal(gpu.quartet_totals(False), o.quartet_totals(False, False))


@pytest.mark.parametrize("n_taxa,n_trees", [(40, 70), (90, 45), (140, 40), (516, 2)])
def test_sliced_plane_builder_depth_classes_complete(gpu, n_taxa, n_trees):
    """The same for complete binary trees (third plane derived as the complement): caterpillar gene trees over
    shuffled taxa are as deep as they are wide (40 -> 6 bits, 90 -> 7, 140 -> 9, 516 -> 11), mixed with random trees."""
    s = synth.make(n_taxa, n_trees, 78)
    rng = np.random.default_rng(n_taxa)
    lines = s.genes.strip().splitlines()
    for i in range(0, n_trees, 2):
        order = rng.permutation(n_taxa)
        cat = "t%04d" % order[0]
        for k in order[1:]:
            cat = "(" + cat + ",t%04d)" % k
        lines[i] = cat + ";"
    prob = Problem(s.constraint, "\n".join(lines) + "\n")
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    for qmode, thr in ((0, 0.0), (2, 0.5)) if n_taxa < 200 else ((2, 0.5),):  # the oracle needs 17 s per pass at 516 taxa
        got = gpu.count_filter(qmode, thr)
        tab = gpu.quartet_totals(False)
        st, want = o.dense_run(qmode, thr, False, digests=False)
        assert got == (st["n_survivors"], st["sum_counts"])
        assert np.array_equal(tab, want)


def test_regular_box_epilogue(gpu, monkeypatch):
    """Boxes whose three consecutive segment pairs have one LCA each are scored from marginal sums
    (kernels_count.cuh, "regular boxes"); CAMUS_B200_NO_REGULAR=1 sends every box through the
    survivor queue. Both against the oracle, on a balanced constraint (every strict box regular),
    a random one and a caterpillar (no regular box)."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_NO_REGULAR", "1")
    queue_only = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_REGULAR")

    def balanced(lo, hi):
        if hi - lo == 1:
            return "t%04d" % lo
        mid = (lo + hi) // 2
        return "(" + balanced(lo, mid) + "," + balanced(mid, hi) + ")"

    s = synth.make(64, 70, 21, drop_frac=0.05)
    cat = "t0000"
    for i in range(1, 64):
        cat = "(" + cat + ",t%04d)" % i
    for constraint in (balanced(0, 64) + ";", s.constraint, cat + ";"):
        prob = Problem(constraint, s.genes)
        o = orc.Oracle().load(prob)
        for qmode, thr in ((0, 0.0), (1, 0.3), (2, 0.5)):
            want = o.count_filter(qmode, thr, False)
            for ctx in (gpu, queue_only):
                ctx.load(prob)
                for as_set in (False, True):
                    ctx.set_score_hint(as_set)
                    assert ctx.count_filter(qmode, thr) == want
                    assert np.array_equal(ctx.quartet_totals(as_set), o.quartet_totals(as_set, False))
    gpu.set_score_hint(False)
    queue_only.close()


def _unroot(nwk: str) -> str:
    """((X,Y),Z); -> (X,Y,Z);  (plain newick without lengths; a leaf first child is rotated away)"""
    body = nwk.strip().rstrip(";")[1:-1]
    depth, cut = 0, -1
    for i, ch in enumerate(body):
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            cut = i
            break
    first, rest = body[:cut], body[cut + 1:]
    if not first.startswith("("):
        first, rest = rest, first
    if not first.startswith("("):
        return nwk
    return "(" + first[1:-1] + "," + rest + ");"


def test_complete_tree_specialisation(gpu, monkeypatch):
    """When every gene tree holds all taxa and is binary the count kernel derives the third
    topology count as G - c0 - c1 (k_count<.., true>). Same results as the general kernel
    (CAMUS_B200_NO_COMPLETE=1), and one incomplete / unresolved tree switches it off."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_NO_COMPLETE", "1")
    general = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_NO_COMPLETE")
    s = synth.make(41, 77, 31)  # complete binary trees, 77 = 2 full groups + 13
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    for qmode, thr in ((0, 0.0), (2, 0.4)):
        want = o.count_filter(qmode, thr, False)
        for ctx in (gpu, general):
            ctx.load(prob)
            assert ctx.count_filter(qmode, thr) == want
            assert np.array_equal(ctx.quartet_totals(False), o.quartet_totals(False, False))
            for a, b in zip(ctx.export_quartets(raw=True), o.export_raw_counts()):
                assert np.array_equal(a, b)
    # unrooted input (trifurcating root): every quartet is still resolved (count kernel stays
    # specialised) but the rooted triples across the three root subtrees are not (plane builder
    # must not derive the third plane)
    lines = [_unroot(l) for l in s.genes.strip().split("\n")]
    assert sum(l.count("(") == 41 - 2 for l in lines) > 60  # really trifurcating now
    prob3 = Problem(s.constraint, "\n".join(lines[:40]) + "\n")
    o3 = orc.Oracle().load(prob3)
    want = o3.count_filter(2, 0.4, False)
    for ctx in (gpu, general):
        ctx.load(prob3)
        assert ctx.count_filter(2, 0.4) == want
        assert np.array_equal(ctx.quartet_totals(False), o3.quartet_totals(False, False))
        for x, y in zip(ctx.export_quartets(raw=True), o3.export_raw_counts()):
            assert np.array_equal(x, y)
    # add one star tree and one tree with a missing taxon: the specialisation must switch itself off
    labs = prob.tip_names
    extra = "(" + ",".join(labs) + ");\n(" + ",".join(labs[:-1]) + ");\n"
    prob2 = Problem(s.constraint, s.genes + extra)
    o2 = orc.Oracle().load(prob2)
    want = o2.count_filter(1, 0.2, False)
    gpu.load(prob2)
    assert gpu.count_filter(1, 0.2) == want
    assert np.array_equal(gpu.quartet_totals(False), o2.quartet_totals(False, False))
    for a, b in zip(gpu.export_quartets(raw=True), o2.export_raw_counts()):
        assert np.array_equal(a, b)
    general.close()


# ------------------------------------------------------------------ reference KATs on the GPU
@pytest.mark.parametrize("name,tre,qmode,thr,genes,expected", K.PROCESS_QUARTETS)
def test_process_quartets(gpu, name, tre, qmode, thr, genes, expected):
    prob = Problem(tre, "\n".join(genes))
    gpu.load(prob)
    gpu.count_filter(qmode, thr)
    assert as_dict(gpu.export_quartets()) == dict(Counter(qkey(prob, q) for q in expected))


def test_quartets_total_kats(gpu):
    # quartettotals_test.go:121-153; the quartet tables are produced by gene trees with missing taxa
    for which, (tre, qs) in (("A", K.TOTALS_TREE_A), ("B", K.TOTALS_TREE_B)):
        genes = []
        for (a, b, c, d), cnt in qs:
            genes += ["((%s,%s),(%s,%s));" % (a, b, c, d)] * cnt
        prob = Problem(tre, "\n".join(genes))
        gpu.load(prob)
        gpu.count_filter(0, 0.0)
        tot = gpu.quartet_totals(False)
        for w_, _, u, w, want in K.QUARTETS_TOTAL:
            if w_ == which:
                assert tot[prob.node_id(u), prob.node_id(w)] == want
        for u in range(prob.n_nodes):
            for w in range(prob.n_nodes):
                if not prob.should_calc_edge(u, w):
                    assert tot[u, w] == 0
        assert (tot > 0).any()


def test_constraint_tables_vs_oracle(gpu):
    """camus_b200_constraint_tables = graphs.calcLCAs / calcDepths / countLeavesBelow
    (treedata.go:130-194), which the Go side would otherwise compute in Theta(n^3)."""
    for s in (synth.make(57, 3, 71), synth.make(200, 2, 72), synth.make(9, 2, 73)):
        prob = Problem(s.constraint, s.genes)
        o = orc.Oracle().load(prob)
        gpu.load(prob)
        lca, depth, below = gpu.constraint_tables()
        n = prob.n_nodes
        assert np.array_equal(lca, np.array([[o.lca(u, w) for w in range(n)] for u in range(n)], np.int32))
        assert [int(x) for x in depth] == [o.depth(v) for v in range(n)]
        assert [int(x) for x in below] == [o.leaves_below(v) for v in range(n)]
        assert np.array_equal(lca, lca.T) and (np.diag(lca) == np.arange(n)).all()


def test_penalty_kats(gpu):
    prob = Problem(K.PENALTY_TREE, "")
    gpu.load(prob)
    o = orc.Oracle(1).load(prob)
    assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())
    prob = Problem(K.TOTALS_TREE_B[0], "")
    gpu.load(prob)
    o = orc.Oracle(1).load(prob)
    pen = gpu.edge_penalties()
    assert np.array_equal(pen, o.edge_penalties())
    assert (pen > 0).any()


def _infer_gpu(gpu, prob, qmode, thr, mode, alpha):
    gpu.load(prob)
    n_surv, sum_counts = gpu.count_filter(qmode, thr)
    as_set = mode == "sym"
    tot = gpu.quartet_totals(as_set)
    pen = gpu.edge_penalties() if mode != "max" else None
    return prob.run_dp(tot, pen, mode, as_set, n_surv, sum_counts, prob.n_trees, alpha), n_surv


@pytest.mark.parametrize("name,tre,genes,nedges,network", K.INFER)
def test_infer_small(gpu, name, tre, genes, nedges, network):
    prob = Problem(tre, "\n".join(genes))
    res, _ = _infer_gpu(gpu, prob, 0, 0.0, "max", 0.0)
    assert len(res.branches) == nedges
    assert res.newicks[-1] == network


@pytest.mark.parametrize("qmode,thr,mode,alpha,nnets,golden", K.INFER_LARGE)
def test_infer_large_golden(gpu, qmode, thr, mode, alpha, nnets, golden):
    prob = Problem.from_files(os.path.join(REF, "infer_constraint.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))
    res, n_surv = _infer_gpu(gpu, prob, qmode, thr, mode, alpha)
    want = [l.strip() for l in read_text(os.path.join(REF, golden)).split("\n") if l.strip()]
    assert res.newicks == want
    assert n_surv == (14423 if qmode == 0 else 1150)


@pytest.mark.parametrize("qmode,thr", [(0, 0.0), (1, 0.3), (2, 0.5), (2, 0.0), (1, 1.0)])
def test_dataset_vs_literal_oracle(gpu, qmode, thr):
    prob = Problem.from_files(os.path.join(REF, "infer_constraint.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))
    compare_all(gpu, prob, qmode, thr, literal=True)


# ----------------------------------------------------------------- synthetic configurations
@pytest.mark.parametrize("qmode,thr", [(0, 0.0), (1, 0.25), (2, 0.5)])
def test_config1_vs_oracle(gpu, qmode, thr):
    s, c = synth.make_config(1)
    prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
    compare_all(gpu, prob, qmode, thr)


def test_config1_literal_and_networks(gpu):
    s, c = synth.make_config(1)
    prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
    o = compare_all(gpu, prob, c["qmode"], c["threshold"], literal=True, as_sets=(False,))
    res, _ = _infer_gpu(gpu, prob, c["qmode"], c["threshold"], "max", 0.0)
    ns, sc = o.count_filter(c["qmode"], c["threshold"], True)
    ref = prob.run_dp(o.quartet_totals(False, True), None, "max", False, ns, sc)
    assert res.newicks == ref.newicks and res.qsat == ref.qsat and len(res.newicks) >= 1


def test_config4_like_polytomies_missing_taxa_nexus(gpu):
    s = synth.make(60, 100, 4, support=True, drop_frac=0.1, nexus=True)
    prob = Problem(s.constraint, s.genes, s.fmt, 0.7)
    assert prob.warned_missing
    compare_all(gpu, prob, 2, 0.5, literal=True)
    compare_all(gpu, prob, 0, 0.0)


def test_config5_like_sym_norm(gpu):
    s = synth.make(48, 150, 5)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    ns, sc = o.count_filter(2, 0.5, False)
    for mode in ("sym", "norm"):
        res, _ = _infer_gpu(gpu, prob, 2, 0.5, mode, 0.5)
        as_set = mode == "sym"
        ref = prob.run_dp(o.quartet_totals(as_set, False), o.edge_penalties(), mode, as_set, ns, sc, prob.n_trees, 0.5)
        assert res.newicks == ref.newicks
        # float scores come from identical integer tables through identical host code: exact
        assert res.root_scores == ref.root_scores or all(
            abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(res.root_scores, ref.root_scores))


@pytest.mark.parametrize("n_taxa,n_trees", [(4, 3), (5, 40), (8, 33), (9, 31), (17, 64), (23, 70)])
def test_small_shapes(gpu, n_taxa, n_trees):
    s = synth.make(n_taxa, n_trees, 100 + n_taxa, drop_frac=0.15 if n_taxa > 6 else 0.0)
    prob = Problem(s.constraint, s.genes)
    compare_all(gpu, prob, 0, 0.0, literal=True)
    compare_all(gpu, prob, 2, 0.3)


def test_deep_caterpillars_and_star(gpu):
    labs = ["t%04d" % i for i in range(30)]
    cat = labs[0]
    for l in labs[1:]:
        cat = "(%s,%s)" % (cat, l)
    rev = labs[-1]
    for l in reversed(labs[:-1]):
        rev = "(%s,%s)" % (rev, l)
    star = "(" + ",".join(labs) + ")"
    half = "((" + ",".join(labs[:15]) + "),(" + ",".join(labs[15:]) + "))"
    genes = [cat + ";", rev + ";", star + ";", half + ";", rev + ";"] * 8
    s = synth.make(30, 1, 7)
    prob = Problem(s.constraint, "\n".join(genes))
    compare_all(gpu, prob, 0, 0.0, literal=True)
    prob = Problem(cat + ";", "\n".join(genes))
    compare_all(gpu, prob, 1, 0.0, literal=True)


def test_no_gene_trees_and_reuse(gpu):
    s = synth.make(12, 5, 9)
    prob = Problem(s.constraint, "")
    gpu.load(prob)
    assert gpu.count_filter(2, 0.5) == (0, 0)
    assert not gpu.quartet_totals(False).any()
    # the same context is reused for a different problem afterwards
    prob = Problem(s.constraint, s.genes)
    compare_all(gpu, prob, 0, 0.0, literal=True)


def test_query_cells_probe(gpu):
    """camus_b200_query_cells against the exported raw table (GPU) and the oracle, polytomies and
    missing taxa included."""
    s = synth.make(45, 70, 12, support=True, drop_frac=0.1, nexus=True)
    prob = Problem(s.constraint, s.genes, s.fmt, 0.7)
    gpu.load(prob)
    gpu.count_filter(0, 0.0)
    rng = np.random.default_rng(3)
    quads = np.array([rng.choice(45, 4, replace=False) for _ in range(2000)], np.int32)
    got = gpu.query_cells(quads)
    o = orc.Oracle().load(prob)
    assert np.array_equal(got, o.count_cells(quads))
    raw = as_dict(gpu.export_quartets(raw=True))
    for q, g in zip(quads[:300], got[:300]):
        t = sorted(int(x) for x in q)
        base = sum(t[i] << (15 * i) for i in range(4))
        assert [raw.get(base | (tp << 60), 0) for tp in (0b1100, 0b1010, 0b0110)] == [int(x) for x in g]


def test_bad_inputs_are_errors_not_crashes(gpu):
    from camus_b200.cabi import CamusB200Error
    s = synth.make(10, 2, 3)
    prob = Problem(s.constraint, s.genes)
    gpu.load(prob)
    with pytest.raises(CamusB200Error):
        gpu.count_filter(3, 0.5)
    with pytest.raises(CamusB200Error):
        gpu.count_filter(1, 1.5)
    bad_parent = prob.gene_parent.copy()
    bad_parent[1] = 5  # parent after child
    with pytest.raises(CamusB200Error):
        gpu.add_gene_trees(prob.gene_node_off, bad_parent, prob.gene_taxon)
    bad_taxon = prob.gene_taxon.copy()
    bad_taxon[bad_taxon >= 0] = 0  # duplicate taxon
    with pytest.raises(CamusB200Error):
        gpu.add_gene_trees(prob.gene_node_off, prob.gene_parent, bad_taxon)
    with pytest.raises(CamusB200Error):
        gpu.set_constraint(prob.parent[:-1], prob.child0[:-1], prob.child1[:-1], prob.tip_index[:-1], prob.n_taxa)


@pytest.mark.parametrize("world", [2, 5])
def test_partitions_on_one_gpu(gpu, world):
    """camus_b200_set_partition: every rank's box range, run one after the other on this GPU, gives
    the survivors the host-side mirror (camus_b200.partition) assigns to it; partial junction
    tables add up to the single-rank table."""
    import torch
    from camus_b200 import partition as part
    s = synth.make(44, 90, 21, drop_frac=0.05)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    o.count_filter(2, 0.3, False)
    keys, counts = o.export_quartets()
    want_tot = o.quartet_totals(False, False)
    owner = part.owner_of_quartets(keys, part.dfs_rank_of_tip(prob.parent, prob.child0, prob.child1, prob.tip_index), world)
    acc = None
    try:
        for r in range(world):
            gpu.set_partition(r, world)
            gpu.load(prob)
            ns, sc = gpu.count_filter(2, 0.3)
            assert (ns, sc) == (int((owner == r).sum()), int(counts[owner == r].sum()))
            gk, gc = gpu.export_quartets()
            assert np.array_equal(gk, keys[owner == r]) and np.array_equal(gc, counts[owner == r])
            gpu.quartet_totals_partial(False)
            z = gpu.partial_tensor().clone()
            acc = z if acc is None else acc + z
        # last rank's context receives the summed table, as after an all-reduce
        gpu.partial_tensor().copy_(acc)
        torch.cuda.synchronize()
        assert np.array_equal(gpu.quartet_totals_finish(), want_tot)
    finally:
        gpu.set_partition(0, 1)


@pytest.mark.parametrize("world", [4, 8])
def test_class_partitions_on_one_gpu(gpu, world, monkeypatch):
    """The class scheme of camus_b200_set_partition (4 and 8 shares; forced here below its default
    size): every share, run one after the other on this GPU with ONLY its own tile classes built,
    yields the survivors the mirror assigns to it, and the partial tables add up to the oracle's.
    query_cells afterwards rebuilds every tile."""
    import torch
    from camus_b200 import partition as part
    monkeypatch.setenv("CAMUS_B200_PARTITION", "class")
    s = synth.make(100, 60, 23, drop_frac=0.05)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    o.count_filter(2, 0.3, False)
    keys, counts = o.export_quartets()
    want_tot = o.quartet_totals(False, False)
    owner = part.owner_of_quartets(keys, part.dfs_rank_of_tip(prob.parent, prob.child0, prob.child1, prob.tip_index), world)
    assert len(set(owner.tolist())) == world
    acc = None
    try:
        for r in range(world):
            gpu.set_partition(r, world)
            if r % 2 == 0:
                gpu.load(prob)  # odd shares reuse the resident trees: the planes are rebuilt for the new classes
            ns, sc = gpu.count_filter(2, 0.3)
            assert (ns, sc) == (int((owner == r).sum()), int(counts[owner == r].sum()))
            if r in (0, world - 1):
                gk, gc = gpu.export_quartets()
                assert np.array_equal(gk, keys[owner == r]) and np.array_equal(gc, counts[owner == r])
            gpu.quartet_totals_partial(False)
            z = gpu.partial_tensor().clone()
            acc = z if acc is None else acc + z
        gpu.partial_tensor().copy_(acc)
        torch.cuda.synchronize()
        assert np.array_equal(gpu.quartet_totals_finish(), want_tot)
        quads = np.array([[0, 1, 2, 3], [5, 17, 40, 63], [99, 8, 57, 6], [90, 91, 92, 93]], np.int32)
        got = gpu.query_cells(quads)
        gpu.set_partition(0, 1)
        gpu.load(prob)
        gpu.count_filter(2, 0.3)
        assert np.array_equal(got, gpu.query_cells(quads))
    finally:
        gpu.set_partition(0, 1)


@pytest.mark.parametrize("n_taxa", [6, 17, 41])
def test_class_partition_tiny_inputs(gpu, n_taxa, monkeypatch):
    """Class scheme forced on inputs with 1, 3 and 6 segments: some of the 8 shares own no box (and
    build no tile) at all; the shares still add up to the oracle's table and survivors."""
    import torch
    monkeypatch.setenv("CAMUS_B200_PARTITION", "class")
    s = synth.make(n_taxa, 30, 100 + n_taxa, drop_frac=0.1)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    want = o.count_filter(1, 0.2, False)
    want_tot = o.quartet_totals(True, False)
    acc, ns, sc = None, 0, 0
    try:
        for r in range(8):
            gpu.set_partition(r, 8)
            gpu.load(prob)
            a, b = gpu.count_filter(1, 0.2)
            ns, sc = ns + a, sc + b
            gpu.quartet_totals_partial(True)
            z = gpu.partial_tensor().clone()
            acc = z if acc is None else acc + z
        assert (ns, sc) == want
        gpu.partial_tensor().copy_(acc)
        torch.cuda.synchronize()
        assert np.array_equal(gpu.quartet_totals_finish(), want_tot)
    finally:
        gpu.set_partition(0, 1)


def test_default_class_scheme_shares_sum_to_the_whole(gpu, monkeypatch):
    """264 taxa = 33 segments: 8 shares take the class scheme by default (complete binary trees,
    the complete-tree kernels). Shares run one after the other through run_resident; scalars and
    junction tables add up to the unpartitioned run's."""
    import torch
    from camus_b200 import partition as part
    monkeypatch.delenv("CAMUS_B200_PARTITION", raising=False)
    assert part.use_class_scheme(264, 8)
    s = synth.make(264, 40, 29)
    prob = Problem(s.constraint, s.genes)
    gpu.set_partition(0, 1)
    gpu.load(prob)
    want = gpu.count_filter(2, 0.5)
    want_tot = gpu.quartet_totals(False)
    acc, ns, sc = None, 0, 0
    try:
        for r in range(8):
            gpu.set_partition(r, 8)
            if r == 0:
                gpu.count_filter(2, 0.5)
            _, a, b = gpu.run_resident(2, 0.5, False, partial_only=True)
            ns, sc = ns + a, sc + b
            z = gpu.partial_tensor().clone()
            acc = z if acc is None else acc + z
        assert (ns, sc) == want
        gpu.partial_tensor().copy_(acc)
        torch.cuda.synchronize()
        assert np.array_equal(gpu.quartet_totals_finish(), want_tot)
    finally:
        gpu.set_partition(0, 1)


# ------------------------------------------------- mid / full size: oracle + size-free properties
def test_mid_size_vs_oracle(gpu):
    s = synth.make(96, 100, 11)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    want = o.count_filter(2, 0.5, False)
    gpu.load(prob)
    assert gpu.count_filter(2, 0.5) == want
    for as_set in (False, True):
        assert np.array_equal(gpu.quartet_totals(as_set), o.quartet_totals(as_set, False))
    assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())


def test_config3_taxa_scale_properties(gpu):
    """1000 taxa (BASELINE.json configs[2] taxon count; 96 gene trees to keep it to seconds):
    10.7 M boxes. Exact agreement with the oracle on 4000 sampled 4-taxon sets (four-point test on
    the host vs camus_b200_query_cells), linearity over tree subsets, order independence,
    agreement of the complete-tree and general kernels."""
    s, c = synth.make_config(3, n_trees=96)
    prob = Problem(s.constraint, s.genes)
    off, par, tax = prob.gene_node_off, prob.gene_parent, prob.gene_taxon
    gpu.load(prob)
    ns_all, sum_all = gpu.count_filter(0, 0.0)
    tot_all = gpu.quartet_totals(False)
    assert gpu.workload()["resolutions"] == 96 * (1000 * 999 * 998 * 997 // 24)
    # sampled cells of the 1000-taxon bit-plane table against the oracle's four-point counts
    rng = np.random.default_rng(7)
    quads = np.array([rng.choice(1000, 4, replace=False) for _ in range(4000)], np.int32)
    quads[:64, :] = np.arange(64)[:, None] + np.array([0, 1, 2, 3])  # neighbouring tips: diagonal boxes
    o = orc.Oracle().load(prob)
    assert np.array_equal(gpu.query_cells(quads), o.count_cells(quads))

    def sub(lo, hi):
        return off[lo:hi + 1] - off[lo], par[off[lo]:off[hi]], tax[off[lo]:off[hi]]

    acc_sum, acc_tot = 0, np.zeros_like(tot_all)
    for lo, hi in ((0, 40), (40, 96)):
        gpu.clear_gene_trees()
        gpu.add_gene_trees(*sub(lo, hi))
        _, sm = gpu.count_filter(0, 0.0)
        acc_sum += sm
        acc_tot += gpu.quartet_totals(False)
    assert acc_sum == sum_all and np.array_equal(acc_tot, tot_all)
    # a star tree switches to the general kernel and adds nothing
    gpu.clear_gene_trees()
    gpu.add_gene_trees(*sub(50, 96))
    gpu.add_gene_trees(*sub(0, 50))
    star = Problem(s.constraint, "(" + ",".join(prob.tip_names) + ");")
    gpu.add_gene_trees(star.gene_node_off, star.gene_parent, star.gene_taxon)
    assert gpu.count_filter(0, 0.0) == (ns_all, sum_all)
    assert np.array_equal(gpu.quartet_totals(False), tot_all)
    # filtered run: every row/column masked by ShouldCalcEdge is zero, totals shrink
    gpu.load(prob)
    ns2, sum2 = gpu.count_filter(2, 0.5)
    tot2 = gpu.quartet_totals(False)
    assert 0 < ns2 < ns_all and sum2 < sum_all and (tot2 <= tot_all).all() and tot2.any()
    assert not tot2[0].any() and not tot2[:, 0].any()
    # the two scoring epilogues (marginal sums on regular boxes vs survivor queue everywhere) are
    # independent code paths: identical tables at 1000 taxa, weighted and -asSet
    import os
    from camus_b200.cabi import CamusB200
    os.environ["CAMUS_B200_NO_REGULAR"] = "1"
    try:
        queue_only = CamusB200(0)
    finally:
        del os.environ["CAMUS_B200_NO_REGULAR"]
    queue_only.load(prob)
    assert queue_only.count_filter(2, 0.5) == (ns2, sum2)
    assert np.array_equal(queue_only.quartet_totals(False), tot2)
    gpu.set_score_hint(True)
    queue_only.set_score_hint(True)
    assert gpu.count_filter(2, 0.5) == queue_only.count_filter(2, 0.5)
    assert np.array_equal(gpu.quartet_totals(True), queue_only.quartet_totals(True))
    gpu.set_score_hint(False)
    queue_only.close()


def test_config2_full_size_properties(gpu):
    """200 taxa x 1000 trees (BASELINE.json configs[1]): linearity in the gene trees, order
    independence, and the counting identity sum(raw) = R for complete binary trees."""
    s, c = synth.make_config(2)
    prob = Problem(s.constraint, s.genes)
    off, par, tax = prob.gene_node_off, prob.gene_parent, prob.gene_taxon
    gpu.load(prob)
    ns_all, sum_all = gpu.count_filter(0, 0.0)
    tot_all = gpu.quartet_totals(False)
    R = gpu.workload()["resolutions"]
    assert R == 1000 * (200 * 199 * 198 * 197 // 24)

    def sub(lo, hi):
        o = off[lo:hi + 1] - off[lo]
        return o, par[off[lo]:off[hi]], tax[off[lo]:off[hi]]

    parts = []
    for lo, hi in ((0, 333), (333, 1000)):
        gpu.clear_gene_trees()
        gpu.add_gene_trees(*sub(lo, hi))
        ns, sm = gpu.count_filter(0, 0.0)
        parts.append((sm, gpu.quartet_totals(False)))
    assert parts[0][0] + parts[1][0] == sum_all
    assert np.array_equal(parts[0][1] + parts[1][1], tot_all)
    # reversed tree order, added in two calls
    gpu.clear_gene_trees()
    gpu.add_gene_trees(*sub(500, 1000))
    gpu.add_gene_trees(*sub(0, 500))
    assert gpu.count_filter(0, 0.0) == (ns_all, sum_all)
    assert np.array_equal(gpu.quartet_totals(False), tot_all)
    # every complete binary tree resolves every 4-set exactly once: the counts the constraint
    # removal drops are the only difference between sum(survivors) and R
    gpu.load(prob)
    ns2, sum2 = gpu.count_filter(2, 0.5)
    assert 0 < ns2 <= ns_all and sum2 <= sum_all < R
    tot2 = gpu.quartet_totals(False)
    assert (tot2 <= tot_all).all() and tot2.any()
    mask = np.array([[prob.should_calc_edge(u, w) for w in range(0, prob.n_nodes, 7)] for u in range(0, prob.n_nodes, 7)])
    assert not tot_all[::7, ::7][~mask].any()


def _n_devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_multi_gpu_handle_matches_single(gpu):
    """camus_b200_create(n_gpus > 1): one handle, every GPU of the process. Survivors, both edge
    tables, penalties, exported quartets and probes equal the single-GPU handle's, fused and
    table pipelines, also with a caller-side partition on top (2 x n_gpus shares)."""
    from camus_b200.cabi import CamusB200
    nd = min(_n_devices(), 4)
    multi = CamusB200(devices=list(range(nd)))
    for s, ms, qmode, thr in ((synth.make(72, 90, 12, support=True, drop_frac=0.1, nexus=True), 0.5, 1, 0.25),
                              (synth.make(64, 64, 13), 0.0, 2, 0.5)):
        prob = Problem(s.constraint, s.genes, s.fmt, ms)
        gpu.load(prob)
        multi.load(prob)
        for as_set in (False, True):
            gpu.set_score_hint(as_set)
            multi.set_score_hint(as_set)
            assert multi.count_filter(qmode, thr) == gpu.count_filter(qmode, thr)
            assert np.array_equal(multi.quartet_totals(as_set), gpu.quartet_totals(as_set))
        assert np.array_equal(multi.edge_penalties(), gpu.edge_penalties())
        for raw in (True, False):
            for a, b in zip(multi.export_quartets(raw=raw), gpu.export_quartets(raw=raw)):
                assert np.array_equal(a, b)
        quads = np.array([[0, 1, 2, 3], [5, 17, 40, 63], [9, 8, 7, 6]], np.int32)
        assert np.array_equal(multi.query_cells(quads), gpu.query_cells(quads))
        out, ns, sc = multi.run_resident(qmode, thr, False)
        assert (ns, sc) == gpu.count_filter(qmode, thr) and np.array_equal(out, gpu.quartet_totals(False))
        # the sum over the devices is formed once per scoring pass: repeated calls are idempotent
        assert np.array_equal(multi.quartet_totals(False), out)      # run_resident already reduced
        assert np.array_equal(multi.quartet_totals(False), out)
        assert np.array_equal(multi.quartet_totals_finish(), out)
        multi.set_score_hint(False)
        assert multi.count_filter(qmode, thr) == (ns, sc)
        assert np.array_equal(multi.quartet_totals_finish(), out)     # finish straight after a fused count_filter
        multi.quartet_totals_partial(False)
        multi.quartet_totals_partial(False)
        assert np.array_equal(multi.quartet_totals_finish(), out)
        # two such handles as two "processes": shares 0 and 1 of 2, summed by hand
        parts = []
        for r in range(2):
            multi.set_partition(r, 2)
            ns_r, sc_r = multi.count_filter(qmode, thr)
            multi.quartet_totals_partial(False)
            parts.append((ns_r, sc_r, multi.partial_tensor().clone()))
        assert (parts[0][0] + parts[1][0], parts[0][1] + parts[1][1]) == (ns, sc)
        multi.partial_tensor().copy_(parts[0][2] + parts[1][2])
        assert np.array_equal(multi.quartet_totals_finish(), out)
        multi.set_partition(0, 1)
        if s.fmt == "newick":  # ingest on every device of the handle (validated per device, appended everywhere)
            multi.load_newick(prob, s.genes, ms)
            assert multi.count_filter(qmode, thr) == (ns, sc)
            assert np.array_equal(multi.quartet_totals(False), out)
    gpu.set_score_hint(False)
    multi.close()


@pytest.mark.gpu
@pytest.mark.skipif(_n_devices() < 4, reason="needs four GPUs (gpurun --gpus 4)")
def test_four_gpu_handle_with_the_class_scheme(gpu, monkeypatch):
    """One handle over 4 GPUs with the class-scheme split forced: every device builds only its own
    tile classes; survivors, edge table and exported quartets equal the single-GPU handle's."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_PARTITION", "class")
    multi = CamusB200(devices=[0, 1, 2, 3])
    try:
        s = synth.make(88, 70, 31, drop_frac=0.05)
        prob = Problem(s.constraint, s.genes)
        gpu.load(prob)
        multi.load(prob)
        assert multi.count_filter(2, 0.4) == gpu.count_filter(2, 0.4)
        assert np.array_equal(multi.quartet_totals(False), gpu.quartet_totals(False))
        for a, b in zip(multi.export_quartets(), gpu.export_quartets()):
            assert np.array_equal(a, b)
        quads = np.array([[0, 1, 2, 3], [5, 17, 40, 63], [80, 8, 57, 6]], np.int32)
        assert np.array_equal(multi.query_cells(quads), gpu.query_cells(quads))
        out, ns, sc = multi.run_resident(2, 0.4, False)
        assert (ns, sc) == gpu.count_filter(2, 0.4) and np.array_equal(out, gpu.quartet_totals(False))
    finally:
        multi.close()


def _random_tree_with_chains(rng, n_taxa, chains):
    """Preorder (parent, taxon) arrays of a random binary tree on all taxa, with unary chains of
    the given lengths inserted above randomly chosen nodes (the first one above the root)."""
    kids = {}
    nodes = list(range(n_taxa))
    nxt = n_taxa
    while len(nodes) > 1:
        a = nodes.pop(rng.integers(len(nodes)))
        b = nodes.pop(rng.integers(len(nodes)))
        kids[nxt] = (a, b)
        nodes.append(nxt)
        nxt += 1
    root = nodes[0]
    above = {root: chains[0]}
    for length in chains[1:]:
        above[int(rng.integers(nxt))] = length
    parent, taxon = [], []
    stack = [(root, -1)]
    while stack:
        x, p = stack.pop()
        for _ in range(above.get(x, 0)):
            parent.append(p)
            taxon.append(-1)
            p = len(parent) - 1
        parent.append(p)
        taxon.append(x if x < n_taxa else -1)
        me = len(parent) - 1
        if x >= n_taxa:
            stack.append((kids[x][1], me))
            stack.append((kids[x][0], me))
    return np.array(parent, np.int32), np.array(taxon, np.int32)


def test_deep_trees_take_the_integer_plane_builder(gpu):
    """Unary chains make a gene tree deeper than 2047 edges although it has few taxa: fp16 depths
    are no longer exact and the library must switch to the integer plane builder by itself
    (camus_b200.cu, half_path). Raw counts, survivors and scores against the oracle; a tree deeper
    than the 16-bit depth tables can hold is refused."""
    rng = np.random.default_rng(3)
    s = synth.make(14, 1, 2)
    prob = Problem(s.constraint, s.genes)
    offs, pars, taxs = [0], [], []
    for g in range(45):
        chains = [3000, 40, 7] if g % 3 == 0 else ([5] if g % 3 == 1 else [0])
        p, t = _random_tree_with_chains(rng, 14, chains)
        pars.append(p)
        taxs.append(t)
        offs.append(offs[-1] + len(p))
    off, par, tax = np.array(offs, np.int64), np.concatenate(pars), np.concatenate(taxs)
    o = orc.Oracle()
    o.set_constraint(prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
    o.add_gene_trees(off, par, tax)
    gpu.set_constraint(prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
    gpu.add_gene_trees(off, par, tax)
    for qmode, thr in ((0, 0.0), (2, 0.3)):
        assert gpu.count_filter(qmode, thr) == o.count_filter(qmode, thr, False)
        assert np.array_equal(gpu.quartet_totals(False), o.quartet_totals(False, False))
    for a, b in zip(gpu.export_quartets(raw=True), o.export_raw_counts()):
        assert np.array_equal(a, b)
    from camus_b200.cabi import CamusB200Error
    p, t = _random_tree_with_chains(rng, 14, [70000])
    with pytest.raises(CamusB200Error, match="deeper|nodes"):
        gpu.add_gene_trees(np.array([0, len(p)], np.int64), p, t)


@pytest.mark.parametrize("seed", range(12))
def test_randomized_instances(gpu, seed):
    """Seeded sweep over shapes and options: taxa 9..75 (around the 8-taxon segment and 32-taxon
    box boundaries), 1..130 trees (around the 32-tree word), complete / missing taxa / collapsed
    edges / unrooted input, all filter modes, weighted and -asSet scores. Everything the C-ABI
    returns against the oracle, bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    n_taxa = int(rng.choice([9, 15, 16, 17, 31, 32, 33, 40, 47, 64, 65, 75]))
    n_trees = int(rng.choice([1, 2, 31, 32, 33, 50, 64, 65, 96, 130]))
    kind = seed % 4
    s = synth.make(n_taxa, n_trees, 500 + seed, support=(kind == 2), drop_frac=(0.15 if kind in (1, 2) else 0.0),
                   nni_rate=float(rng.choice([0.5, 3.0, 12.0])))
    genes = s.genes
    if kind == 3:
        genes = "\n".join(_unroot(l) for l in genes.strip().split("\n")) + "\n"
    prob = Problem(s.constraint, genes, s.fmt, 0.6 if kind == 2 else 0.0)
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    qmode = int(rng.integers(0, 3))
    thr = float(rng.choice([0.0, 0.1, 1.0 / 3.0, 0.5, 0.9, 1.0]))
    as_set = bool(rng.integers(0, 2))
    gpu.set_score_hint(as_set)
    try:
        assert gpu.count_filter(qmode, thr) == o.count_filter(qmode, thr, False)
        assert np.array_equal(gpu.quartet_totals(as_set), o.quartet_totals(as_set, False))
        assert np.array_equal(gpu.quartet_totals(not as_set), o.quartet_totals(not as_set, False))
        for a, b in zip(gpu.export_quartets(), o.export_quartets()):
            assert np.array_equal(a, b)
        for a, b in zip(gpu.export_quartets(raw=True), o.export_raw_counts()):
            assert np.array_equal(a, b)
        assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())
    finally:
        gpu.set_score_hint(False)


# ----------------------------------------------------------------- inputs beyond the bit-plane budget
def test_eight_sequential_passes_equal_one_pass(gpu, monkeypatch):
    """CAMUS_B200_PASSES=8: the 8 class-scheme shares run one after the other on one GPU (what the
    library does by itself when the bit-plane table does not fit, e.g. 1000 taxa x 4000 trees), each
    holding half of the bit-planes, junction table and scalars accumulating. Same survivors, both edge
    tables and a second quartet_totals with the other weight, against the oracle and the one-pass
    context: complete trees, incomplete trees (packed kernel), every filter mode."""
    from camus_b200.cabi import CamusB200
    monkeypatch.setenv("CAMUS_B200_PASSES", "8")
    multi = CamusB200(0)
    monkeypatch.delenv("CAMUS_B200_PASSES")
    cases = [(synth.make(70, 130, 61), 0.0), (synth.make(96, 64, 62, drop_frac=0.1), 0.0), (synth.make(23, 40, 63), 0.0),
             (synth.make(58, 90, 64, support=True, drop_frac=0.1, nexus=True), 0.7)]
    for s, ms in cases:
        prob = Problem(s.constraint, s.genes, s.fmt, ms)
        o = orc.Oracle().load(prob)
        for qmode, thr in ((0, 0.0), (2, 0.5), (1, 0.3)):
            want = o.count_filter(qmode, thr, False)
            for hint in (False, True):
                multi.load(prob)
                multi.set_score_hint(hint)
                assert multi.count_filter(qmode, thr) == want
                for as_set in (hint, not hint):
                    assert np.array_equal(multi.quartet_totals(as_set), o.quartet_totals(as_set, False))
                _, ns, sc = multi.run_resident(qmode, thr, hint, np.empty((prob.n_nodes, prob.n_nodes), np.uint64))
                assert (ns, sc) == want
            # the table pipeline after a passes run: the bit-planes in HBM are one share's, export rebuilds the whole table
            for a, b in zip(multi.export_quartets(), o.export_quartets()):
                assert np.array_equal(a, b)
            assert np.array_equal(multi.quartet_totals(False), o.quartet_totals(False, False))
        gpu.load(prob)
        assert gpu.count_filter(2, 0.5) == o.count_filter(2, 0.5, False)  # the one-pass context is untouched by the mode
    multi.close()
