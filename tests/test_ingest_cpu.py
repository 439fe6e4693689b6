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


def _random_tree(rng, taxa, p_poly=0.2, p_unary=0.05, p_support=0.5, p_len=0.5, p_ws=0.2, p_comment=0.1):
    """A random newick string over a random subset of `taxa`, with the syntax the dialect allows."""
    leaves = [t for t in taxa if rng.random() > 0.15] or [taxa[0]]
    rng.shuffle(leaves)

    def ws():
        return rng.choice([" ", "\t", "  "]) if rng.random() < p_ws else ""

    def deco(internal):
        out = ""
        if internal and rng.random() < p_support:
            out += rng.choice(["0.%d" % rng.integers(0, 1000), "%d" % rng.integers(0, 101), "1", "0.7", "0.69", "inner%d" % rng.integers(9), "1.", "00.50"])
        if rng.random() < p_comment:
            out += "[c%d]" % rng.integers(9)
        if rng.random() < p_len:
            out += ":" + ws() + rng.choice(["0.1", "1e-3", "2", "1.5E+2", "0"])
        return out

    def build(items):
        if len(items) == 1:
            node = items[0] + ws() + deco(False)
            if rng.random() < p_unary:
                node = "(" + ws() + node + ws() + ")" + deco(True)
            return node
        k = 2 if rng.random() > p_poly or len(items) < 3 else int(rng.integers(3, min(5, len(items)) + 1))
        cuts = sorted(rng.choice(np.arange(1, len(items)), size=k - 1, replace=False).tolist())
        parts = [items[a:b] for a, b in zip([0] + cuts, cuts + [len(items)])]
        return "(" + ws() + ("," + ws()).join(build(p) for p in parts) + ws() + ")" + deco(True)

    body = build(leaves)
    if not body.startswith("("):
        body = "(" + body + ")"
    return ws() + body + ws() + ";" + ws()


def test_randomised_differential_against_the_host_parser(probe):
    """Whenever the automaton accepts a file it must produce exactly the host parser's flat trees (no silent divergence);
    it may be stricter (fallback), never different. Random syntax; then random single-character corruptions."""
    rng = np.random.default_rng(int(os.environ.get("CAMUS_FUZZ_SEED", "11")))
    taxa = ["t%02d" % i for i in range(14)]
    cons = synth.make(14, 1, 3).constraint
    names = Problem(cons, "").tip_names
    taxa = list(names)
    accepted = fallbacks = 0
    for it in range(int(os.environ.get("CAMUS_FUZZ_ITERS", "250"))):
        ms = float(rng.choice([0.0, 0.0, 0.5, 0.7, 0.695]))
        text = "\n".join(_random_tree(rng, taxa) for _ in range(int(rng.integers(1, 5)))) + "\n"
        if it % 3 == 2:  # corrupt one character
            pos = int(rng.integers(0, len(text) - 1))
            text = text[:pos] + rng.choice(list("(),;:[]' x")) + text[pos + 1:]
        try:
            prob = Problem(cons, text, "newick", ms)
        except Exception:
            prob = None
        n, aux, flat = ingest(probe, names, text, ms)
        if n >= 0:
            accepted += 1
            assert prob is not None, f"automaton accepted what the host parser rejects: {text!r}"
            assert n == prob.n_trees and np.array_equal(flat[0], prob.gene_node_off), text
            assert np.array_equal(flat[1], prob.gene_parent) and np.array_equal(flat[2], prob.gene_taxon), text
        else:
            fallbacks += 1
    assert accepted > 100 and fallbacks > 20
