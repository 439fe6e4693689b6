"""Committed digests of the hot path's outputs on seeded synthetic inputs
(tests/golden/synthetic/vectors.json, made by tests/golden/make_synthetic.py with the oracle in
literal mode). CPU: the oracle's fast mode (four-point test + per-quartet scatter) reproduces them,
i.e. neither mode has drifted. GPU (-m gpu): the CUDA path through the C-ABI reproduces them without
the oracle in the loop."""
import json
import os

import pytest

from tests.golden import make_synthetic as M

VECTORS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synthetic", "vectors.json")))


def test_every_case_has_a_vector():
    assert sorted(VECTORS) == sorted(c[0] for c in M.CASES)


@pytest.mark.parametrize("name,kw,ms,qmode,thr,as_set", M.CASES, ids=[c[0] for c in M.CASES])
def test_oracle_fast_mode_reproduces_the_vectors(name, kw, ms, qmode, thr, as_set):
    prob = M.make_problem(kw, ms)
    assert M.digest(M.OracleBackend(prob, literal=False), prob, qmode, thr, as_set) == VECTORS[name]


class GpuBackend:
    def __init__(self, gpu, prob, as_set):
        self.g = gpu.load(prob)
        self.g.set_score_hint(as_set)

    def count_filter_plain(self, qmode, thr):
        return self.g.count_filter(qmode, thr)

    def export_quartets(self):
        return self.g.export_quartets()

    def totals_plain(self, as_set):
        return self.g.quartet_totals(as_set)

    def edge_penalties(self):
        return self.g.edge_penalties()


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from camus_b200.cabi import CamusB200
    g = CamusB200(0)
    yield g
    g.set_score_hint(False)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,ms,qmode,thr,as_set", M.CASES, ids=[c[0] for c in M.CASES])
def test_gpu_reproduces_the_vectors(gpu, name, kw, ms, qmode, thr, as_set):
    prob = M.make_problem(kw, ms)
    assert M.digest(GpuBackend(gpu, prob, as_set), prob, qmode, thr, as_set) == VECTORS[name]
