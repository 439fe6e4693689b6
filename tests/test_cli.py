"""The `camus` command line (camus_b200/host/camus_main.cpp): same flags and CSV as camus.go."""
import gzip
import os
import shutil
import subprocess

import pytest

from camus_b200 import build
from tests.golden import ref_kats as K

REF = os.path.join(os.path.dirname(__file__), "golden", "ref")


def cli():
    return build.build_cli()


def test_cli_usage_and_argument_errors(tmp_path):
    exe = cli()
    r = subprocess.run([exe, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage: camus [flags]... <const_tree_file> <gene_tree_file>" in r.stderr
    r = subprocess.run([exe, "-hh"], capture_output=True, text=True)
    assert "-asSet" in r.stderr and "-sm mode" in r.stderr
    r = subprocess.run([exe, "only-one"], capture_output=True, text=True)
    assert r.returncode == 1 and "two positional arguments required" in r.stderr
    r = subprocess.run([exe, "-sm", "bogus", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "is not a valid score mode" in r.stderr
    r = subprocess.run([exe, "-q", "5", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "out of type range" in r.stderr
    r = subprocess.run([exe, "-o", str(tmp_path / "x"), "/nonexistent.nwk", "/nonexistent2.nwk"], capture_output=True, text=True)
    assert r.returncode == 1 and "camus encountered an error ::" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("qmode,thr,mode,alpha,nnets,golden", K.INFER_LARGE)
def test_cli_reproduces_golden_networks(tmp_path, qmode, thr, mode, alpha, nnets, golden):
    exe = cli()
    genes = tmp_path / "gene-trees.nwk"
    with gzip.open(os.path.join(REF, "infer_gene-trees.nwk.gz"), "rb") as f, open(genes, "wb") as g:
        shutil.copyfileobj(f, g)
    prefix = str(tmp_path / "out")
    cmd = [exe, "-o", prefix, "-q", str(qmode), "-t", str(thr), "-sm", mode]
    if mode == "sym":
        cmd += ["-a", str(alpha)]
    cmd += [os.path.join(REF, "infer_constraint.nwk"), str(genes)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = [l.strip() for l in open(os.path.join(REF, golden)) if l.strip()]
    import csv
    rows = list(csv.reader(r.stdout.splitlines()))
    assert rows[0] == ["Number of Branches", "Quartet Satisfied Percent", "Extended Newick"]
    assert rows[1][0] == "0" and rows[1][1] == "0"
    assert [row[2] for row in rows[2:]] == want
    assert [row[0] for row in rows[2:]] == [str(i + 1) for i in range(nnets)]
    assert open(prefix + ".csv").read() == r.stdout
    log = open(prefix + ".log").read()
    assert "1123 gene trees provided, containing %d quartets not in the constraint tree" % (14423 if qmode == 0 else 1150) in log
    assert "%d edges identified" % nnets in log


def test_parallel_newick_parse_reports_first_bad_line():
    """Gene-tree files above 256 KB are parsed on all host threads; results and the error for the
    FIRST malformed line (io.go:141-147 reports the line index) do not depend on that."""
    from camus_b200 import synth
    from camus_b200.hostlib import Problem, CamusError
    s = synth.make(300, 400, 5)
    assert len(s.genes) > (1 << 18)
    whole = Problem(s.constraint, s.genes)
    lines = s.genes.strip().split("\n")
    half = Problem(s.constraint, "\n".join(lines[:100]))  # below the threshold: serial parse
    n100 = int(half.gene_node_off[100])
    assert whole.n_trees == 400 and half.n_trees == 100
    assert (whole.gene_parent[:n100] == half.gene_parent).all() and (whole.gene_taxon[:n100] == half.gene_taxon).all()
    lines[123] = lines[123][:50]
    lines[300] = "((a,b"
    with pytest.raises(CamusError, match="line 123"):
        Problem(s.constraint, "\n".join(lines))
