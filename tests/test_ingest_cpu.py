"""The newick -> flat-tree automaton of the GPU ingest (camus_b200/csrc/kernels_ingest.cuh: the per-tree functions
k_nwk_pass1 / k_nwk_pass2 call) compiled for the host with g++ (tests/ingest_probe.cpp) and compared with the host
mirror of prep.readGeneTreesFile + the per-tree half of prep.processQuartets: the SAME flat arrays, and a fallback
status for everything outside the dialect. No GPU involved; tests/test_ingest_gpu.py runs the kernels themselves."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from camus_b200 import synth
from camus_b200.hostlib import Problem, read_text

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "ref")


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ingest") / "libingest_probe.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "ingest_probe.cpp")], check=True)
    L = C.CDLL(so)
    L.probe_ingest.restype = C.c_longlong
    L.probe_ingest.argtypes = [C.c_char_p, C.c_longlong, C.c_double, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_longlong, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


def ingest(L, tip_names, text, min_support=0.0):
    enc = [n.encode() for n in tip_names]
    off = np.zeros(len(enc) + 1, np.int64)
    off[1:] = np.cumsum([len(e) for e in enc])
    pool = b"".join(enc) or b"\0"
    raw = text.encode()
    cap = len(raw) + 8
    node_off = np.zeros(raw.count(b"\n") + 3, np.int64)
    parent = np.zeros(cap, np.int32)
    taxon = np.zeros(cap, np.int32)
    st, miss = C.c_int(), C.c_int()
    r = L.probe_ingest(raw, len(raw), float(min_support), len(enc), pool, off.ctypes.data, node_off.ctypes.data, parent.ctypes.data,
                       taxon.ctypes.data, cap, C.byref(st), C.byref(miss))
    if r < 0:
        return int(r), st.value, None
    n = int(node_off[r])
    return int(r), bool(miss.value), (node_off[:r + 1].copy(), parent[:n].copy(), taxon[:n].copy())


def same_flat(L, prob, text, ms=0.0):
    n, missing, flat = ingest(L, prob.tip_names, text, ms)
    assert n == prob.n_trees
    assert np.array_equal(flat[0], prob.gene_node_off) and np.array_equal(flat[1], prob.gene_parent) and np.array_equal(flat[2], prob.gene_taxon)
    nl = np.diff(np.concatenate([[0], np.cumsum(prob.gene_taxon >= 0)[prob.gene_node_off[1:] - 1]]))
    assert missing == bool((nl != prob.n_taxa).any())


def test_reference_dataset(probe):
    cons = read_text(os.path.join(REF, "infer_constraint.nwk"))
    genes = read_text(os.path.join(REF, "infer_gene-trees.nwk.gz"))
    same_flat(probe, Problem(cons, genes), genes)


@pytest.mark.parametrize("taxa,trees,seed,kw,ms", [
    (30, 40, 1, {}, 0.0), (64, 100, 2, {"drop_frac": 0.15}, 0.0), (50, 80, 3, {"support": True}, 0.7),
    (50, 80, 4, {"support": True, "drop_frac": 0.1}, 0.35), (50, 80, 5, {"support": True}, 0.0), (120, 60, 6, {}, 0.0),
])
def test_synthetic_inputs(probe, taxa, trees, seed, kw, ms):
    s = synth.make(taxa, trees, seed, **kw)
    same_flat(probe, Problem(s.constraint, s.genes, "newick", ms), s.genes, ms)


def test_dialect_details(probe):
    cons = "((a,b),((c,d),(e,f)));"
    genes = "\n".join([
        "  ((a:0.1,b:1e-3)0.9:0.2,((c,d)inner:3,(e,f)0.5)0.8)0.1;  ",
        "",
        "((a[&comment],b)[c2]0.2,(c,(d,(e,f)0.65)0.7));",
        "(a,b,(c,(d)));",
        "\t(((a , b ) 0.70 , c ) 0.69 , d );",
        "((a,b)0.69999999999999,(c,d)0.7000000000001);",
        "((a,b)000.5,(c,d).5,(e,f)5.);",
    ]) + "\n"
    for ms in (0.0, 0.7, 0.5):
        same_flat(probe, Problem(cons, genes, "newick", ms), genes, ms)


@pytest.mark.parametrize("bad,ms,code", [
    ("((a,b),('c d',e));", 0.0, 2), ("((a,b),(c,x));", 0.0, 4), ("((a,b),(c,a));", 0.0, 5), ("((a,b)1e-1,(c,d));", 0.5, 3),
    ("((a,b)inf,(c,d));", 0.5, 3), ("((a,b)0.1234567890123456,(c,d));", 0.5, 3), ("((a,b),(c,d);", 0.0, 1),
    ("((a,b),(c,d)); extra", 0.0, 1), ("(a,b),(c,d));", 0.0, 1), ("((a,b),(c,d))", 0.0, 1), ("((a,b),[(c,d));", 0.0, 1),
])
def test_everything_else_is_a_fallback(probe, bad, ms, code):
    good = "((a,b),((c,d),e));\n((a,c),((b,d),e));\n"
    n, status, flat = ingest(probe, ["a", "b", "c", "d", "e"], good + bad + "\n" + good, ms)
    assert n == -3 and status == code and flat is None


def test_support_values_match_strtod(probe):
    """Plain decimals of <= 15 significant digits: double(M) / 10^f is the correctly rounded value strtod returns,
    so `support < threshold` agrees with the host parser for thresholds right at, below and above the value."""
    rng = np.random.default_rng(5)
    cons = "((a,b),((c,d),e));"
    for _ in range(300):
        digits = int(rng.integers(1, 16))
        frac = int(rng.integers(0, digits + 1))
        m = int(rng.integers(0, 10 ** digits))
        txt = f"{m:0{digits}d}"
        lab = (txt[:digits - frac] or "0") + ("." + txt[digits - frac:] if frac else "")
        v = float(lab)
        for thr in (v, np.nextafter(v, np.inf), np.nextafter(v, -np.inf) if v > 0 else 1.0):
            if thr == 0:
                continue
            genes = f"((a,b){lab},((c,d){lab},e));\n"
            same_flat(probe, Problem(cons, genes, "newick", float(thr)), genes, float(thr))
