"""The oracle's dense mode (oracle/dense.hpp) against its literal mode: the dense mode is what the
GPU parity tests compare with at the sizes BASELINE.json names (config 2 in full, 1000 taxa), so it
is pinned here, on the CPU, on inputs the literal algorithms can still run."""
import itertools

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem
from oracle import oracle as orc


def _problem(n_taxa, n_trees, seed, **kw):
    ms = kw.pop("min_support", 0.0)
    s = synth.make(n_taxa, n_trees, seed, **kw)
    return Problem(s.constraint, s.genes, s.fmt, ms)


CASES = [
    dict(n_taxa=24, n_trees=40, seed=3),
    dict(n_taxa=33, n_trees=70, seed=5, drop_frac=0.15),
    dict(n_taxa=40, n_trees=50, seed=8, support=True, drop_frac=0.1, nexus=True, min_support=0.7),
    dict(n_taxa=17, n_trees=130, seed=9, nni_rate=6.0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("qmode,thr", [(0, 0.0), (1, 0.3), (2, 0.5), (2, 0.0)])
def test_dense_equals_literal(case, qmode, thr):
    kw = dict(CASES[case])
    prob = _problem(kw.pop("n_taxa"), kw.pop("n_trees"), kw.pop("seed"), **kw)
    o = orc.Oracle().load(prob)
    want = o.count_filter(qmode, thr, True)
    for as_set in (False, True):
        st, tab = o.dense_run(qmode, thr, as_set)
        assert (st["n_survivors"], st["sum_counts"]) == want
        assert np.array_equal(tab, o.quartet_totals(as_set, True))
    rk, rc = o.export_raw_counts()
    sk, sc = o.export_quartets()
    assert st["n_raw_keys"] == len(rk) and st["sum_raw"] == int(rc.sum())
    assert st["hash_raw"] == orc.digest(rk, rc) and st["hash_surv"] == orc.digest(sk, sc)
    assert st["cells"] == prob.n_taxa * (prob.n_taxa - 1) * (prob.n_taxa - 2) * (prob.n_taxa - 3) // 24
    # the digest is what it says: one mix per (key, count)
    assert orc.digest(sk[:3], sc[:3]) == sum(orc.lib().oracle_mix_key(int(k), int(c)) for k, c in zip(sk[:3], sc[:3])) % 2**64


def test_filter_cell_is_schedule_independent_and_equals_the_map_form():
    """filterQuartets visits the keys of a Go map in random order and re-reads counts after its own
    deletions; on one 4-taxon set every visiting order gives the same survivors, and they equal the
    map form (oracle filter_quartets, the literal mode's)."""
    rng = np.random.default_rng(11)
    base = orc.pack_quartet(3, 9, 14, 20) & ~(0xF << 60)
    topo = [0b1100, 0b1010, 0b0110]
    cases = [tuple(int(x) for x in rng.integers(0, 9, 3)) for _ in range(400)]
    cases += [(5, 5, 5), (0, 0, 7), (0, 4, 4), (4, 4, 0), (1, 2, 3), (0, 0, 0), (2, 2, 9), (9, 2, 2), (700, 200, 100)]
    for counts in cases:
        for qmode, thr in ((1, 0.0), (1, 0.3), (1, 0.5), (2, 0.0), (2, 0.1), (2, 1 / 3), (2, 0.5), (2, 0.9), (1, 1.0), (2, 1.0)):
            got = {orc.filter_cell(counts, qmode, thr, order) for order in itertools.permutations(range(3))}
            assert len(got) == 1, (counts, qmode, thr, got)
            keys = [base | (topo[i] << 60) for i in range(3) if counts[i]]
            ok, oc = orc.filter_quartets(keys, [c for c in counts if c], qmode, thr)
            want = [0, 0, 0]
            for k, c in zip(ok, oc):
                want[topo.index(int(k) >> 60)] = int(c)
            assert tuple(want) == got.pop(), (counts, qmode, thr)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_dense_literal_at_equals_literal_table(case):
    kw = dict(CASES[case])
    prob = _problem(kw.pop("n_taxa"), kw.pop("n_trees"), kw.pop("seed"), **kw)
    o = orc.Oracle().load(prob)
    o.count_filter(2, 0.5, True)
    n = prob.n_nodes
    for as_set in (False, True):
        want = o.quartet_totals(as_set, True)
        seen = 0
        for v in range(n):
            pairs = [(u, w) for u in range(n) for w in range(n) if o.lca(u, w) == v]
            got, nq = o.dense_literal_at(v, pairs, 2, 0.5, as_set)
            assert nq == o.num_quartets_at(v)
            for (u, w), x in zip(pairs, got):
                assert int(x) == int(want[u, w]), (v, u, w)
            seen += len(pairs)
        assert seen == n * n
