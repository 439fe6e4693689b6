"""GPU parity at the sizes BASELINE.json names: the CUDA path (through the C-ABI) against the
oracle's dense mode (oracle/dense.hpp, pinned on the literal mode by tests/test_oracle_dense.py)
and against the LITERAL per-(u, w) QuartetScore loop at 1000 taxa x 1000 trees. Bit-exact."""
import json
import os
import zlib

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(__file__)
DIGESTS = os.path.join(HERE, "golden", "synthetic", "config_digests.json")


@pytest.fixture(scope="module")
def gpu():
    from camus_b200.cabi import CamusB200
    g = CamusB200(0)
    yield g
    g.close()


def table_digest(tab: np.ndarray) -> str:
    """crc32 + wrapped uint64 sum of an n x n uint64 table (committed in config_digests.json)."""
    return "%08x-%016x" % (zlib.crc32(np.ascontiguousarray(tab).tobytes()), int(tab.sum(dtype=np.uint64)))


def _config(cid, **kw):
    s, c = synth.make_config(cid, **kw)
    return Problem(s.constraint, s.genes, s.fmt, c["min_support"]), c


def _compare_tables(gpu, o, qmode, thr, as_sets, digests=True):
    """count_filter scalars + quartet_totals for every weight in as_sets: GPU vs dense oracle."""
    st = None
    for as_set in as_sets:
        gpu.set_score_hint(as_set)
        got = gpu.count_filter(qmode, thr)
        tab = gpu.quartet_totals(as_set)
        st, want = o.dense_run(qmode, thr, as_set, digests=digests and st is None)
        assert got == (st["n_survivors"], st["sum_counts"]), f"scalars as_set={as_set}"
        assert np.array_equal(tab, want), f"quartet_totals as_set={as_set}"
    gpu.set_score_hint(False)
    return st


def test_config2_full_size_vs_oracle(gpu):
    """BASELINE.json configs[1] in full (200 taxa x 1000 trees, -q 2): scalars, both edge tables,
    penalties, and every raw / surviving (quartet, count) through order-independent digests."""
    prob, c = _config(2)
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    st, _ = o.dense_run(c["qmode"], c["threshold"], False, table=False)
    _compare_tables(gpu, o, c["qmode"], c["threshold"], (False, True), digests=False)
    assert st["resolutions"] == gpu.workload()["resolutions"] == 1000 * (200 * 199 * 198 * 197 // 24)
    assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())
    gpu.count_filter(c["qmode"], c["threshold"])
    sk, sc = gpu.export_quartets(raw=False)
    assert len(sk) == st["n_survivors"] and int(sc.sum()) == st["sum_counts"] and orc.digest(sk, sc) == st["hash_surv"]
    del sk, sc
    rk, rc = gpu.export_quartets(raw=True)
    assert len(rk) == st["n_raw_keys"] and int(rc.sum(dtype=np.uint64)) == st["sum_raw"] and orc.digest(rk, rc) == st["hash_raw"]
    # the other filter modes at full size (scalars + table)
    for qmode, thr in ((0, 0.0), (1, 0.3)):
        _compare_tables(gpu, o, qmode, thr, (False,), digests=False)


def test_config5_full_size_sym_norm_vs_oracle(gpu):
    """BASELINE.json configs[4] in full (300 taxa x 2000 trees): both weights (sym always counts as
    a set, infer.go:84-85), penalties, and the DP's networks and float scores in sym and norm mode
    from the GPU tables vs from the oracle's tables (tolerance of the north star: 1e-9 relative)."""
    prob, c = _config(5)
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    pen = gpu.edge_penalties()
    assert np.array_equal(pen, o.edge_penalties())
    for mode, as_set in (("sym", True), ("norm", False)):
        gpu.set_score_hint(as_set)
        ns, sc = gpu.count_filter(c["qmode"], c["threshold"])
        tab = gpu.quartet_totals(as_set)
        st, want = o.dense_run(c["qmode"], c["threshold"], as_set, digests=False)
        assert (ns, sc) == (st["n_survivors"], st["sum_counts"])
        assert np.array_equal(tab, want)
        res = prob.run_dp(tab, pen, mode, as_set, ns, sc, prob.n_trees, 0.5)
        ref = prob.run_dp(want, o.edge_penalties(), mode, as_set, st["n_survivors"], st["sum_counts"], prob.n_trees, 0.5)
        assert res.newicks == ref.newicks and len(res.newicks) >= 1
        assert all(abs(a - b) <= 1e-9 * max(1.0, abs(b)) for a, b in zip(res.root_scores, ref.root_scores))
    gpu.set_score_hint(False)


def test_config4_500_taxa_nexus_collapse_asset_vs_oracle(gpu):
    """BASELINE.json configs[3] at its taxon count: 500 taxa, 256 NEXUS gene trees with support
    values collapsed at -s 0.7 (polytomies), 10 % of the taxa missing per tree, -asSet."""
    prob, c = _config(4, n_trees=256)
    assert prob.warned_missing
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    st = _compare_tables(gpu, o, c["qmode"], c["threshold"], (True, False), digests=False)
    assert st["resolutions"] == gpu.workload()["resolutions"]
    assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())
    sampled = np.random.default_rng(4).choice(500, (3000, 4))
    sampled = sampled[[len(set(q)) == 4 for q in sampled]].astype(np.int32)
    assert np.array_equal(gpu.query_cells(sampled), o.count_cells(sampled))


def test_1000_taxa_whole_table_vs_oracle(gpu):
    """1000 taxa (BASELINE.json configs[2] taxon count) x 32 gene trees: 10.7 M boxes, 125 segments,
    128 junction-table replicas, the regular-box epilogue, 64-bit indexing; the ENTIRE 1999 x 1999
    table and the scalars against the dense oracle (4.1e10 cells on the host cores)."""
    prob, c = _config(3, n_trees=32)
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    _compare_tables(gpu, o, c["qmode"], c["threshold"], (False,), digests=False)
    assert np.array_equal(gpu.edge_penalties(), o.edge_penalties())


def _sample_pairs_by_vertex(o, n, rng, n_vertices, max_leaves, pairs_per_vertex):
    lb = np.array([o.leaves_below(v) for v in range(n)])
    cands = [v for v in range(1, n) if 4 <= lb[v] <= max_leaves]
    picked = {}
    for v in rng.choice(cands, min(n_vertices, len(cands)), replace=False):
        v = int(v)
        sub = [x for x in range(v, n) if o.lca(v, x) == v]
        pairs = [(u, w) for u in sub for w in sub if o.lca(u, w) == v and o.should_calc_edge(u, w)]
        if len(pairs) > pairs_per_vertex:
            pairs = [pairs[i] for i in rng.choice(len(pairs), pairs_per_vertex, replace=False)]
        if pairs:
            picked[v] = pairs
    return picked


def test_config3_full_size_vs_literal_quartet_score(gpu):
    """BASELINE.json configs[2] in full (1000 taxa x 1000 trees, the bench workload): sampled
    entries of the edge table against the LITERAL algorithm -- td.Quartets(v) = surviving quartets
    with >= 3 taxa under v (treedata.go:197-222), then the per-quartet QuartetScore of
    quartettotals.go:78-157 for the pair -- for LCA vertices with at most 40 leaves below them.
    Plus sampled raw counts, and the committed digests of the result (also asserted by bench.py)."""
    prob, c = _config(3)
    n = prob.n_nodes
    o = orc.Oracle().load(prob)
    gpu.load(prob)
    ns, sc = gpu.count_filter(c["qmode"], c["threshold"])
    tab = gpu.quartet_totals(False)
    rng = np.random.default_rng(2026)
    picked = _sample_pairs_by_vertex(o, n, rng, 48, 40, 24)
    checked = nonzero = 0
    for v, pairs in picked.items():
        want, _ = o.dense_literal_at(v, pairs, c["qmode"], c["threshold"], False)
        for (u, w), x in zip(pairs, want):
            assert int(tab[u, w]) == int(x), (v, u, w)
            checked += 1
            nonzero += int(x) != 0
    assert checked >= 400 and nonzero >= 100, (checked, nonzero)
    quads = np.array([rng.choice(1000, 4, replace=False) for _ in range(2000)], np.int32)
    assert np.array_equal(gpu.query_cells(quads), o.count_cells(quads))
    # -asSet at full size: same sampled check on a smaller sample
    gpu.set_score_hint(True)
    assert gpu.count_filter(c["qmode"], c["threshold"]) == (ns, sc)
    tab_set = gpu.quartet_totals(True)
    gpu.set_score_hint(False)
    for v, pairs in list(picked.items())[:12]:
        want, _ = o.dense_literal_at(v, pairs, c["qmode"], c["threshold"], True)
        assert [int(tab_set[u, w]) for u, w in pairs] == [int(x) for x in want]
    got = {"n_survivors": ns, "sum_counts": sc, "quartet_totals": table_digest(tab), "quartet_totals_as_set": table_digest(tab_set)}
    want = json.load(open(DIGESTS))["config3"]
    assert got == want, got


def test_beyond_the_bit_plane_budget_1000_taxa_3000_trees(gpu):
    """1000 taxa x 3000 gene trees need a 197 GB bit-plane table: more than one B200 holds. The
    library then runs the 8 class-scheme shares one after the other (half of the table resident at
    a time, camus_b200.cu run_passes). Size-independent checks: with -q 0 the unfiltered edge table
    and both scalars are LINEAR in the gene trees, so the 3000-tree result must equal the sum over
    three 1000-tree runs (each of which fits and takes the ordinary one-pass path, itself compared
    with the oracle elsewhere in this file); and the filtered run must reproduce itself through
    run_resident."""
    s = synth.make(1000, 3000, 77)
    lines = s.genes.strip().split("\n")
    assert len(lines) == 3000
    parts = []
    for b in range(3):
        prob = Problem(s.constraint, "\n".join(lines[b * 1000:(b + 1) * 1000]) + "\n")
        gpu.load(prob)
        ns, sc = gpu.count_filter(0, 0.0)
        parts.append((ns, sc, gpu.quartet_totals(False)))
        assert gpu.stage_ms()[0]["count"] > 0
    prob = Problem(s.constraint, s.genes)
    gpu.load(prob)
    ns, sc = gpu.count_filter(0, 0.0)
    tab = gpu.quartet_totals(False)
    assert sc == sum(p[1] for p in parts)
    assert ns >= max(p[0] for p in parts)  # distinct surviving keys: a union, not a sum
    assert np.array_equal(tab, parts[0][2] + parts[1][2] + parts[2][2])
    want = gpu.count_filter(2, 0.5)
    t2 = gpu.quartet_totals(False)
    out, ns2, sc2 = gpu.run_resident(2, 0.5, False)
    assert (ns2, sc2) == want and np.array_equal(out, t2)
    gpu.clear_gene_trees()


def test_beyond_the_bit_plane_budget_1500_taxa_1000_trees(gpu):
    """1500 taxa x 1000 gene trees: 221 GB of bit-planes (188 segments), again more than one B200 holds; the 8 sequential
    passes keep half of it resident. Same size-independent check: the unfiltered (-q 0) table and the summed counts are
    linear in the gene trees, so they must equal the sum over the two 500-tree halves, each of which fits (110 GB) and
    takes the one-pass path."""
    s = synth.make(1500, 1000, 78)
    lines = s.genes.strip().split("\n")
    parts = []
    for b in range(2):
        prob = Problem(s.constraint, "\n".join(lines[b * 500:(b + 1) * 500]) + "\n")
        gpu.load(prob)
        ns, sc = gpu.count_filter(0, 0.0)
        parts.append((ns, sc, gpu.quartet_totals(False)))
    prob = Problem(s.constraint, s.genes)
    gpu.load(prob)
    ns, sc = gpu.count_filter(0, 0.0)
    tab = gpu.quartet_totals(False)
    assert sc == parts[0][1] + parts[1][1]
    assert ns >= max(parts[0][0], parts[1][0])
    assert np.array_equal(tab, parts[0][2] + parts[1][2])
    ms, _ = gpu.stage_ms()
    print(f"\n1500 x 1000 in 8 passes: prep {ms['prep']:.0f} ms, count {ms['count']:.0f} ms")
    gpu.clear_gene_trees()
