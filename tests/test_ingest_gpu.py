"""Gene-tree ingest on the GPU (camus_b200_add_gene_trees_newick): newick text -> flat
trees, against the host mirror of prep.readGeneTreesFile + the per-tree half of prep.processQuartets
(camus_b200/host/tree.hpp, treedata.hpp; pinned on TestReadInputFiles and the golden networks).
Bar: the SAME flat arrays (node offsets, parents, taxa), tree by tree."""
import gzip
import os

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem, read_text

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


@pytest.fixture(scope="module")
def gpu():
    from camus_b200.cabi import CamusB200
    g = CamusB200(0)
    yield g
    g.close()


def _same_flat(gpu, prob, text, min_support=0.0):
    n, missing = gpu.load_newick(prob, text, min_support)
    off, par, tax = gpu.get_gene_trees()
    assert n == prob.n_trees
    assert np.array_equal(off, prob.gene_node_off)
    assert np.array_equal(par, prob.gene_parent)
    assert np.array_equal(tax, prob.gene_taxon)
    nl = np.diff(np.concatenate([[0], np.cumsum(prob.gene_taxon >= 0)[prob.gene_node_off[1:] - 1]]))
    assert missing == bool((nl != prob.n_taxa).any())
    return n


def test_reference_dataset(gpu):
    """The reference's own gene-tree file (1123 trees, 4-22 of 22 taxa each, 482 polytomous, branch lengths)."""
    cons = read_text(os.path.join(REF, "infer_constraint.nwk"))
    genes = read_text(os.path.join(REF, "infer_gene-trees.nwk.gz"))
    prob = Problem(cons, genes)
    _same_flat(gpu, prob, genes)
    # and the results downstream of it
    a = gpu.count_filter(2, 0.5)
    ta = gpu.quartet_totals(False)
    gpu.load(prob)
    assert gpu.count_filter(2, 0.5) == a and np.array_equal(gpu.quartet_totals(False), ta)


@pytest.mark.parametrize("taxa,trees,seed,kw,ms", [
    (30, 40, 1, {}, 0.0), (64, 100, 2, {"drop_frac": 0.15}, 0.0), (50, 80, 3, {"support": True}, 0.7),
    (50, 80, 4, {"support": True, "drop_frac": 0.1}, 0.35), (50, 80, 5, {"support": True}, 0.0), (200, 300, 6, {}, 0.0),
])
def test_synthetic_inputs(gpu, taxa, trees, seed, kw, ms):
    s = synth.make(taxa, trees, seed, **kw)
    prob = Problem(s.constraint, s.genes, "newick", ms)
    _same_flat(gpu, prob, s.genes, ms)


def test_dialect_details(gpu):
    """Comments, blank lines, whitespace, internal node names, ':' lengths with exponents, a support on the
    root edge, unifurcations, a trifurcating root, appending to trees that are already there."""
    cons = "((a,b),((c,d),(e,f)));"
    genes = "\n".join([
        "  ((a:0.1,b:1e-3)0.9:0.2,((c,d)inner:3,(e,f)0.5)0.8)0.1;  ",
        "",
        "((a[&comment],b)[c2]0.2,(c,(d,(e,f)0.65)0.7));",
        "(a,b,(c,(d)));",
        "\t(((a , b ) 0.70 , c ) 0.69 , d );",
    ]) + "\n"
    for ms in (0.0, 0.7):
        prob = Problem(cons, genes, "newick", ms)
        _same_flat(gpu, prob, genes, ms)
    prob = Problem(cons, genes + genes, "newick", 0.7)
    gpu.load_newick(prob, genes, 0.7)
    gpu.add_gene_trees_newick(genes, 0.7)
    off, par, tax = gpu.get_gene_trees()
    assert np.array_equal(off, prob.gene_node_off) and np.array_equal(par, prob.gene_parent) and np.array_equal(tax, prob.gene_taxon)


@pytest.mark.parametrize("bad,ms,why", [
    ("((a,b),('c d',e));", 0.0, "quoted"),
    ("((a,b),(c,x));", 0.0, "not in the constraint"),
    ("((a,b),(c,a));", 0.0, "twice"),
    ("((a,b)1e-1,(c,d));", 0.5, "plain decimal"),
    ("((a,b),(c,d);", 0.0, "malformed"),
    ("((a,b),(c,d)); extra", 0.0, "malformed"),
    ("(a,b),(c,d));", 0.0, "malformed"),
])
def test_fallback_adds_nothing(gpu, bad, ms, why):
    from camus_b200.cabi import CamusB200Error
    cons = "((a,b),((c,d),e));"
    good = "((a,b),((c,d),e));\n((a,c),((b,d),e));\n"
    prob = Problem(cons, good)
    gpu.load_newick(prob, good)
    before = gpu.get_gene_trees()
    with pytest.raises(CamusB200Error) as ei:
        gpu.add_gene_trees_newick(good + bad + "\n" + good, ms)
    assert ei.value.code == 5 and ei.value.bad_tree == 3 and why in str(ei.value)
    after = gpu.get_gene_trees()
    assert all(np.array_equal(x, y) for x, y in zip(before, after))
    assert gpu.count_filter(0, 0.0)[1] > 0  # the context still works


def test_config3_size_and_time(gpu):
    """1000 taxa x 1000 trees (8 MB of text): identical flat arrays; prints the device and host times."""
    import time
    s, c = synth.make_config(3)
    t0 = time.perf_counter()
    prob = Problem(s.constraint, s.genes)
    t_host = time.perf_counter() - t0
    gpu.load_newick(prob, s.genes)  # warm
    t0 = time.perf_counter()
    gpu.clear_gene_trees()
    gpu.add_gene_trees_newick(s.genes.encode())
    t_dev = time.perf_counter() - t0
    off, par, tax = gpu.get_gene_trees()
    assert np.array_equal(off, prob.gene_node_off) and np.array_equal(par, prob.gene_parent) and np.array_equal(tax, prob.gene_taxon)
    print(f"\ningest 1000 x 1000: host parse + flatten {t_host * 1e3:.1f} ms (all cores, constraint included), GPU {t_dev * 1e3:.1f} ms")
