"""score.ReticulationScore (internal/score/score.go:25-72).

CPU part: the host mirror (ConvertToNetwork, Level1, getReticulationNodes) and the oracle's literal
restatement are pinned on the reference's own tests: TestConvertToNetwork (prep/io_test.go:115-211),
TestRecticulationScore (score/score_test.go:18-80) and the golden score_scores.csv of
TestCalculateRecticulationScore_Large (score_test.go:105-162; 1123 gene trees x 5 reticulations).
GPU part: camus_b200_reticulation_counts against the oracle (bit-exact integer counts) and against
the same goldens."""
import os

import numpy as np
import pytest

from camus_b200.hostlib import CamusError, NetworkProblem, read_text
from oracle import oracle as orc

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")

# prep/io_test.go:125-133, score/score_test.go:27-41
NET = "(((9,0),(7,(6,(#H1,8h0u)))),((#H3,(12,((3,(14h2w)#H3),10))h2u),((((5,(#H2,13h1u)),((2h1w)#H2,11))h0w)#H1,(1,4))));"
NET_RET = {"#H1": ("8h0u", "h0w"), "#H2": ("13h1u", "2h1w"), "#H3": ("h2u", "14h2w")}
KAT_GENES = ["((0,9),(5,7));", "((5,7),(9,6));", "((9,7),(5,6));", "((5,9),(7,6));"]
nan = float("nan")
KAT_EXPECTED = [[nan, nan, nan], [0.0, nan, nan], [1.0, nan, nan], [0.0, nan, nan]]
NOT_LEVEL1 = "(A,(B,(#H2,(C,(#H1,(D,(E,(F,((G,(H,((I,J))#H2)))#H1))))))));"  # score_test.go:44


def _same(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="module")
def large():
    return NetworkProblem.from_files(os.path.join(REF, "score_network.nwk"), os.path.join(REF, "infer_gene-trees.nwk.gz"))


def test_convert_to_network_maps_labels_to_nodes():
    p = NetworkProblem(NET, "")
    assert p.labels == ["#H1", "#H2", "#H3"]
    for r, lab in enumerate(p.labels):
        assert (p.node_names[p.uwvs[r, 0]], p.node_names[p.uwvs[r, 1]]) == NET_RET[lab]
    # the network tree keeps its #H tips (15 taxa + 3) and the unary #H nodes
    assert p.n_tips == 18 and p.n_nodes == 38


@pytest.mark.parametrize("text,err", [
    ("((a,b),(c,d));", "ErrNoReticulations"),                    # io_test.go:168 (a plain tree)
    ("((a,b),(c,d),e);", "ErrUnrooted"),                         # io_test.go:174
    ("((a,b,x),((c)#H1,(d,#H1)));", "ErrNonBinary"),             # io_test.go:137
    ("((a,#H1),((c)#H1,(d,#H1)));", "ErrInvalidFormat"),         # label matched twice (io.go:203)
    ("((a,b),((c,e),(d,#H1)));", "ErrInvalidFormat"),            # unmatched label (io.go:229)
    (NOT_LEVEL1, "ErrNotLevel1"),                                # score_test.go:43-48
])
def test_network_errors(text, err):
    with pytest.raises(CamusError) as e:
        NetworkProblem(text, "")
    assert e.value.name == err


def test_gene_tree_errors():
    with pytest.raises(CamusError) as e:
        NetworkProblem(NET, "((0,9),(5,zz));")
    assert e.value.name == "ErrTipNameMismatch"
    with pytest.raises(CamusError) as e:
        NetworkProblem(NET, "((0,9),(5,5));")
    assert e.value.name == "ErrMulTree"


def test_oracle_small_kat():
    p = NetworkProblem(NET, "\n".join(KAT_GENES))
    tot, sup = orc.reticulation_counts(p)
    assert _same(p.scores(tot, sup), KAT_EXPECTED)


def test_oracle_reproduces_the_golden_scores(large):
    tot, sup = orc.reticulation_counts(large)
    want = read_text(os.path.join(REF, "score_scores.csv")).strip()
    assert large.labels == ["#H1", "#H2", "#H3", "#H4", "#H5"] and large.n_trees == 1123
    assert large.csv(tot, sup) == want
    # and numerically, cell by cell
    rows = [l.split(",")[1:] for l in want.split("\n")[1:]]
    assert _same(large.scores(tot, sup), [[float(x) for x in r] for r in rows])


def test_records_are_what_quartet_score_reads(large):
    """Flags and attachment nodes of the flattened records against the tree itself."""
    par, rec = large.parent, large.records
    tip_node = {int(t): i for i, t in enumerate(large.tip_index) if t >= 0}

    def anc(x):
        out = [x]
        while par[out[-1]] >= 0:
            out.append(int(par[out[-1]]))
        return out

    for r, (u, w, v, ws) in enumerate(large.uwvs):
        for t in range(large.n_tips):
            a = anc(tip_node[t])
            lw = next(x for x in anc(int(w)) if x in a)
            lu = next(x for x in anc(int(u)) if x in a)
            assert rec[r, t, 0] == lw and rec[r, t, 1] == lu
            assert rec[r, t, 3] == (int(w) in a) | ((int(v) in a) << 1) | ((int(ws) in a) << 2) | ((int(u) in a) << 3)


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from camus_b200.cabi import CamusB200
    g = CamusB200(0)
    yield g
    g.close()


@pytest.mark.gpu
def test_gpu_small_kat(gpu):
    p = NetworkProblem(NET, "\n".join(KAT_GENES))
    tot, sup = gpu.reticulation_counts(p)
    assert _same(p.scores(tot, sup), KAT_EXPECTED)
    ot, os_ = orc.reticulation_counts(p)
    assert np.array_equal(tot, ot) and np.array_equal(sup, os_)


@pytest.mark.gpu
def test_gpu_golden_scores_and_oracle_counts(gpu, large):
    tot, sup = gpu.reticulation_counts(large)
    ot, os_ = orc.reticulation_counts(large)
    assert np.array_equal(tot, ot) and np.array_equal(sup, os_)
    assert large.csv(tot, sup) == read_text(os.path.join(REF, "score_scores.csv")).strip()


def _random_network_case(seed, n_taxa, n_genes):
    """A random level-1 network made by the product's own DP on a synthetic instance, then random
    polytomous gene trees with missing taxa and unary chains."""
    from camus_b200 import synth
    from camus_b200.hostlib import Problem
    s = synth.make(n_taxa, n_genes, seed, drop_frac=0.15)
    prob = Problem(s.constraint, s.genes)
    o = orc.Oracle().load(prob)
    ns, sc = o.count_filter(0, 0.0, False)
    res = prob.run_dp(o.quartet_totals(False, False), None, "max", False, ns, sc)
    return res.newicks[min(3, len(res.newicks)) - 1], s.genes


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_taxa,n_genes", [(5, 20, 40), (6, 31, 25), (8, 24, 30)])
def test_gpu_vs_oracle_on_inferred_networks(gpu, seed, n_taxa, n_genes):
    net, genes = _random_network_case(seed, n_taxa, n_genes)
    # unrooted / multifurcating / unary-chain variants of some gene trees (UnRoot and path-length rule)
    extra = ["((t0000,t0001,t0002),(t0003,t0004),t0005);", "(((((t0000,t0001)),t0002)),((t0003,t0004)));",
             "(t0000,(t0001,(t0002,(t0003,(t0004,t0005)))));", "((t0000,t0001),((t0002,t0003),(t0004,t0005)),(t0006,t0007));"]
    p = NetworkProblem(net, genes.strip() + "\n" + "\n".join(extra))
    assert p.n_ret >= 1
    tot, sup = gpu.reticulation_counts(p)
    ot, os_ = orc.reticulation_counts(p)
    assert np.array_equal(tot, ot) and np.array_equal(sup, os_)
    assert int(tot.sum()) > 0 and int(sup.sum()) > 0


@pytest.mark.gpu
def test_gpu_argument_checks(gpu, large):
    from camus_b200.cabi import CamusB200Error, lib, _p
    import ctypes as C
    rec = np.ascontiguousarray(large.records)
    off = np.array([0, 3], np.int64)
    par = np.array([-1, 0, 0], np.int32)
    tax = np.array([-1, 0, 0], np.int32)  # a taxon twice
    out = np.zeros(5, np.uint64)
    rc = lib().camus_b200_reticulation_counts(gpu._h, large.n_tips, large.n_ret, _p(rec), 1, _p(off), _p(par), _p(tax), _p(out), _p(out))
    assert rc == 4
    rc = lib().camus_b200_reticulation_counts(gpu._h, large.n_tips, 0, _p(rec), 1, _p(off), _p(par), _p(tax), _p(out), _p(out))
    assert rc == 1
    # more than 64 reticulations (score.ReticulationScore has no limit): batches of 64 inside the call.
    # The dataset's 5 reticulations repeated 14 times must give the 5 columns 14 times over.
    tot5, sup5 = gpu.reticulation_counts(large)
    reps = 14
    rec70 = np.ascontiguousarray(np.tile(rec.reshape(large.n_ret, -1), (reps, 1)))
    a = [np.ascontiguousarray(x) for x in (large.gene_node_off, large.gene_parent, large.gene_taxon)]
    tot = np.zeros((large.n_trees, large.n_ret * reps), np.uint64)
    sup = np.zeros_like(tot)
    rc = lib().camus_b200_reticulation_counts(gpu._h, large.n_tips, large.n_ret * reps, _p(rec70), large.n_trees, _p(a[0]), _p(a[1]), _p(a[2]),
                                              _p(tot), _p(sup))
    assert rc == 0
    assert np.array_equal(tot, np.tile(tot5, (1, reps))) and np.array_equal(sup, np.tile(sup5, (1, reps)))
