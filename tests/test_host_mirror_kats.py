"""Host-side mirror (camus_b200/host) against the reference tests that surround the hot path:
TestReadInputFiles and the file-level cases of TestConvertToNetwork (prep/io_test.go:11-211),
TestMakeNetwork (graphs/net_test.go:10-57), TestParseScorerMap / TestWithNGtrees / TestWithAlpha
(score/scorers_test.go:47-146). CPU only."""
import os

import pytest

from camus_b200.hostlib import CamusError, NetworkProblem, Problem, check_scorer_options, read_text

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")
CONSTRAINT = read_text(os.path.join(REF, "prep_constraint.nwk"))
QUARTETS_NWK = read_text(os.path.join(REF, "quartets.nwk"))
QUARTETS_NEX = read_text(os.path.join(REF, "quartets.nex"))
BADTREE, BADTREE_NOSEMI, EMPTY = "((a,b);\n", "((()()(\n", ""  # prep/testdata/{badtree,badtree-nosemi,empty}.nwk
NET = "(((9,0),(7,(6,(#H1,8h0u)))),((#H3,(12,((3,(14h2w)#H3),10))h2u),((((5,(#H2,13h1u)),((2h1w)#H2,11))h0w)#H1,(1,4))));"


# prep/io_test.go:20-93
@pytest.mark.parametrize("name,tree,genes,fmt,err", [
    ("basic", CONSTRAINT, QUARTETS_NWK, "newick", None),
    ("non-unique const tree", QUARTETS_NWK, QUARTETS_NWK, "newick", "ErrInvalidFile"),
    ("bad const tree", BADTREE, QUARTETS_NWK, "newick", "ErrInvalidFormat"),
    ("bad const tree (no ;)", BADTREE_NOSEMI, QUARTETS_NWK, "newick", "ErrInvalidFormat"),
    ("empty const tree", EMPTY, QUARTETS_NWK, "newick", "ErrInvalidFile"),
    ("bad gene tree trees", CONSTRAINT, BADTREE, "newick", "ErrInvalidFormat"),
    ("basic nexus", CONSTRAINT, QUARTETS_NEX, "nexus", None),
])
def test_read_input_files(name, tree, genes, fmt, err):
    if err is None:
        p = Problem(tree, genes, fmt)
        assert p.tip_names == list("ABCDEFGHIJ") and p.n_trees == 2
    else:
        with pytest.raises(CamusError) as e:
            Problem(tree, genes, fmt)
        assert e.value.name == err


# prep/io_test.go:136-176 (the cases that fail while reading or converting the network file)
@pytest.mark.parametrize("name,text,err", [
    ("unresolved", read_text(os.path.join(REF, "unresolved.nwk")), "ErrNonBinary"),
    ("non-unique network", NET + "\n" + NET + "\n", "ErrInvalidFile"),
    ("bad network", BADTREE, "ErrInvalidFormat"),
    ("bad network (no ;)", BADTREE_NOSEMI, "ErrInvalidFormat"),
    ("empty", EMPTY, "ErrInvalidFile"),
    ("no reticulations", CONSTRAINT, "ErrNoReticulations"),
    ("unrooted", "(((A,B)#H1,C),(D,(E,#H1)),(F,G));", "ErrUnrooted"),
])
def test_convert_to_network_file_errors(name, text, err):
    with pytest.raises(CamusError) as e:
        NetworkProblem(text, "")
    assert e.value.name == err


# graphs/net_test.go:17-22
def test_make_network_string():
    p = Problem("[&R]((A,(B,(C,F)a)b)c,(D,E)d)e;", "")
    got = p.make_network([(p.node_id("F"), p.node_id("E"))])
    assert got == "((A,(B,(C,(#H1,F))a)b)c,(D,(E)#H1)d)e;"


# score/scorers_test.go:47-146
def test_scorer_options():
    for mode in ("max", "norm", "sym"):
        check_scorer_options(mode, 3, 0.1)
    with pytest.raises(KeyError):
        check_scorer_options("bad")
    for n in (0, -2):
        with pytest.raises(CamusError) as e:
            check_scorer_options("norm", n)
        assert e.value.name == "ErrInvalidScorerOption"
    check_scorer_options("sym", alpha=1.0)
    for a in (1.1, 0.0, -1.0):
        with pytest.raises(CamusError) as e:
            check_scorer_options("sym", alpha=a)
        assert e.value.name == "ErrInvalidScorerOption"


# prep/io_test.go:66-73 ("empty gene trees" -> ErrInvalidFile), through the CLI, which reads files
def test_cli_empty_gene_tree_file(tmp_path):
    import subprocess
    from camus_b200 import build
    c, g = tmp_path / "c.nwk", tmp_path / "g.nwk"
    c.write_text(CONSTRAINT)
    g.write_text("")
    r = subprocess.run([build.build_cli(), "-o", str(tmp_path / "out"), str(c), str(g)], capture_output=True, text=True)
    assert r.returncode == 1 and "invalid file" in r.stderr and "empty gene tree file" in r.stderr
