"""ctypes binding of the CPU oracle (oracle/camus_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product (camus_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")

QNEQ, QEQ, QDIFF = 9, 10, 11  # graphs/quartet.go:27-31 (iota continues)

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "camus_oracle.cpp")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        vp, i32p, i64p, u64p, u32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.oracle_create.restype = vp
        L.oracle_destroy.argtypes = [vp]
        L.oracle_last_error.argtypes = [vp]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_constraint.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
        L.oracle_add_gene_trees.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.oracle_set_quartets.argtypes = [vp, C.c_uint64, vp, vp]
        L.oracle_tree_quartets.argtypes = [C.c_int32, vp, vp, C.c_int, vp, C.c_uint64, u64p]
        L.oracle_count_filter.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, u64p, u64p]
        L.oracle_literal_emissions.argtypes = [vp]
        L.oracle_literal_emissions.restype = C.c_uint64
        L.oracle_export_quartets.argtypes = [vp, vp, vp, C.c_uint64, u64p]
        L.oracle_export_raw_counts.argtypes = [vp, vp, vp, C.c_uint64, u64p]
        L.oracle_count_cells.argtypes = [vp, C.c_int32, vp, vp]
        L.oracle_quartet_totals.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
        L.oracle_edge_penalties.argtypes = [vp, vp]
        L.oracle_quartet_score.argtypes = [vp, C.c_uint64, C.c_int, C.c_int]
        L.oracle_calculate_penalty.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_calculate_penalty.restype = C.c_uint64
        L.oracle_should_calc_edge.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_cycle_length.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_lca.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_num_quartets_at.argtypes = [vp, C.c_int]
        L.oracle_num_quartets_at.restype = C.c_uint64
        L.oracle_pack_quartet.argtypes = [C.c_int] * 4
        L.oracle_pack_quartet.restype = C.c_uint64
        L.oracle_threshold_keep.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_dense_run.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp]
        L.oracle_dense_literal_at.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vp, u64p]
        L.oracle_filter_cell.argtypes = [vp, C.c_int, C.c_double, vp]
        L.oracle_filter_cell.restype = None
        L.oracle_filter_quartets.argtypes = [C.c_uint64, vp, vp, C.c_int, C.c_double, vp, vp, u64p]
        L.oracle_mix_key.argtypes = [C.c_uint64, C.c_uint32]
        L.oracle_mix_key.restype = C.c_uint64
        L.oracle_leaves_below.argtypes = [vp, C.c_int]
        L.oracle_depth.argtypes = [vp, C.c_int]
        L.oracle_reticulation_counts.argtypes = [C.c_int32, vp, vp, C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack_quartet(t0: int, t1: int, t2: int, t3: int) -> int:
    """graphs.makeQuartet(setTopology(..)) for the pairing (t0,t1 | t2,t3) of constraint tip indices."""
    return lib().oracle_pack_quartet(t0, t1, t2, t3)


def unpack_quartet(q: int):
    taxa = [(q >> (15 * i)) & 0x7FFF for i in range(4)]
    return taxa, (q >> 60) & 0xF


def threshold_keep(t: float, c0: int, c1: int, c2: int) -> bool:
    return bool(lib().oracle_threshold_keep(t, c0, c1, c2))


def filter_cell(counts, qmode: int, threshold: float, order=(0, 1, 2)):
    """prep.filterQuartets on the three topology counts of one 4-taxon set (0 = key absent), visiting
    the keys in `order`."""
    c = np.array(counts, np.uint32)
    o = np.array(order, np.int32)
    lib().oracle_filter_cell(_p(c), qmode, threshold, _p(o))
    return tuple(int(x) for x in c)


def filter_quartets(keys, counts, qmode: int, threshold: float):
    """prep.filterQuartets (the map form the literal mode uses) on an explicit (key, count) list."""
    keys = np.ascontiguousarray(keys, np.uint64)
    counts = np.ascontiguousarray(counts, np.uint32)
    ok, oc, n = np.zeros(len(keys), np.uint64), np.zeros(len(keys), np.uint32), C.c_uint64()
    lib().oracle_filter_quartets(len(keys), _p(keys), _p(counts), qmode, threshold, _p(ok), _p(oc), C.byref(n))
    return ok[:n.value], oc[:n.value]


def digest(keys: np.ndarray, counts: np.ndarray) -> int:
    """Order-independent 64-bit digest of a (packed quartet, count) list: the sum the oracle's dense
    mode accumulates over raw / surviving quartets (dense.hpp mix_key)."""
    with np.errstate(over="ignore"):
        z = keys.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * (counts.astype(np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))


def tree_quartets(parent: np.ndarray, taxon: np.ndarray, literal: bool = True) -> np.ndarray:
    """graphs.QuartetsFromTree for one flat tree: sorted array of packed keys."""
    parent = np.ascontiguousarray(parent, np.int32)
    taxon = np.ascontiguousarray(taxon, np.int32)
    n = C.c_uint64()
    lib().oracle_tree_quartets(len(parent), _p(parent), _p(taxon), int(literal), None, 0, C.byref(n))
    keys = np.zeros(n.value, np.uint64)
    lib().oracle_tree_quartets(len(parent), _p(parent), _p(taxon), int(literal), _p(keys), n.value, C.byref(n))
    return keys


def reticulation_counts(net):
    """score.ReticulationScore (score.go:25-72), literal, for a camus_b200.hostlib.NetworkProblem:
    (totals, supported) as uint64 [n_trees, n_ret]."""
    tot = np.zeros((net.n_trees, net.n_ret), np.uint64)
    sup = np.zeros((net.n_trees, net.n_ret), np.uint64)
    a = [np.ascontiguousarray(x) for x in (net.parent, net.tip_index, net.uwvs, net.gene_node_off, net.gene_parent, net.gene_taxon)]
    rc = lib().oracle_reticulation_counts(net.n_nodes, _p(a[0]), _p(a[1]), net.n_tips, net.n_ret, _p(a[2]), net.n_trees,
                                          _p(a[3]), _p(a[4]), _p(a[5]), _p(tot), _p(sup))
    assert rc == 0
    return tot, sup


class Oracle:
    """Same call sequence as the C-ABI (include/camus_b200.h), computed on the CPU."""

    def __init__(self, nthreads: int | None = None):
        self._h = C.c_void_p(lib().oracle_create())
        self.nthreads = nthreads or (os.cpu_count() or 1)
        self.n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def set_constraint(self, parent, child0, child1, tip_index, n_taxa: int):
        a = [np.ascontiguousarray(x, np.int32) for x in (parent, child0, child1, tip_index)]
        self.n = len(a[0])
        lib().oracle_set_constraint(self._h, self.n, n_taxa, *[_p(x) for x in a])

    def add_gene_trees(self, node_off, parent, taxon):
        node_off = np.ascontiguousarray(node_off, np.int64)
        parent = np.ascontiguousarray(parent, np.int32)
        taxon = np.ascontiguousarray(taxon, np.int32)
        lib().oracle_add_gene_trees(self._h, len(node_off) - 1, _p(node_off), _p(parent), _p(taxon))

    def load(self, prob):
        """prob: camus_b200.hostlib.Problem (flat arrays only are read)."""
        self.set_constraint(prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
        if prob.n_trees:
            self.add_gene_trees(prob.gene_node_off, prob.gene_parent, prob.gene_taxon)
        return self

    def set_quartets(self, keys, counts):
        keys = np.ascontiguousarray(keys, np.uint64)
        counts = np.ascontiguousarray(counts, np.uint32)
        lib().oracle_set_quartets(self._h, len(keys), _p(keys), _p(counts))

    def count_filter(self, qmode: int = 0, threshold: float = 0.0, literal: bool = True):
        ns, sc = C.c_uint64(), C.c_uint64()
        rc = lib().oracle_count_filter(self._h, qmode, threshold, int(literal), self.nthreads, C.byref(ns), C.byref(sc))
        assert rc == 0
        return ns.value, sc.value

    def literal_emissions(self) -> int:
        return lib().oracle_literal_emissions(self._h)

    def _export(self, fn):
        n = C.c_uint64()
        fn(self._h, None, None, 0, C.byref(n))
        keys = np.zeros(n.value, np.uint64)
        counts = np.zeros(n.value, np.uint32)
        fn(self._h, _p(keys), _p(counts), n.value, C.byref(n))
        return keys, counts

    def export_quartets(self):
        return self._export(lib().oracle_export_quartets)

    def export_raw_counts(self):
        return self._export(lib().oracle_export_raw_counts)

    def count_cells(self, quads) -> np.ndarray:
        """Raw topology counts of the given 4-taxon sets (tip indices), reference topology order."""
        quads = np.ascontiguousarray(quads, np.int32).reshape(-1, 4)
        out = np.zeros((len(quads), 3), np.uint32)
        lib().oracle_count_cells(self._h, len(quads), _p(quads), _p(out))
        return out

    def dense_run(self, qmode: int = 0, threshold: float = 0.0, as_set: bool = False, table: bool = True, digests: bool = True):
        """The whole path in dense mode (oracle/dense.hpp): no quartet map, any size that fits the
        LCA-depth tables. Returns (stats dict, n x n totals or None)."""
        out = np.zeros((self.n, self.n), np.uint64) if table else None
        st = np.zeros(8, np.uint64)
        rc = lib().oracle_dense_run(self._h, qmode, threshold, int(as_set), self.nthreads, int(digests),
                                    _p(out) if table else None, _p(st))
        if rc != 0:
            raise RuntimeError(lib().oracle_last_error(self._h).decode())
        names = ("n_survivors", "sum_counts", "n_raw_keys", "sum_raw", "hash_raw", "hash_surv", "resolutions", "cells")
        return {k: int(v) for k, v in zip(names, st)}, out

    def dense_literal_at(self, v: int, pairs, qmode: int = 0, threshold: float = 0.0, as_set: bool = False):
        """Literal quartetsTotal (quartettotals.go:78-93) for (u, w) pairs whose LCA is v, at any
        input size: (values, len(td.Quartets(v)))."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        out = np.zeros(len(pairs), np.uint64)
        nq = C.c_uint64()
        rc = lib().oracle_dense_literal_at(self._h, v, qmode, threshold, int(as_set), self.nthreads, len(pairs), _p(pairs),
                                           _p(out), C.byref(nq))
        if rc != 0:
            raise RuntimeError(lib().oracle_last_error(self._h).decode())
        return out, nq.value

    def leaves_below(self, v: int) -> int:
        return lib().oracle_leaves_below(self._h, v)

    def depth(self, v: int) -> int:
        return lib().oracle_depth(self._h, v)

    def quartet_totals(self, as_set: bool = False, literal: bool = True) -> np.ndarray:
        out = np.zeros((self.n, self.n), np.uint64)
        rc = lib().oracle_quartet_totals(self._h, int(as_set), int(literal), self.nthreads, _p(out))
        if rc != 0:
            raise RuntimeError(lib().oracle_last_error(self._h).decode())
        return out

    def edge_penalties(self) -> np.ndarray:
        out = np.zeros((self.n, self.n), np.uint64)
        lib().oracle_edge_penalties(self._h, _p(out))
        return out

    def calculate_penalty(self, u: int, w: int) -> int:
        return lib().oracle_calculate_penalty(self._h, u, w)

    def quartet_score(self, q: int, u: int, w: int) -> int:
        return lib().oracle_quartet_score(self._h, q, u, w)

    def should_calc_edge(self, u: int, w: int) -> bool:
        return bool(lib().oracle_should_calc_edge(self._h, u, w))

    def cycle_length(self, u: int, w: int) -> int:
        return lib().oracle_cycle_length(self._h, u, w)

    def lca(self, a: int, b: int) -> int:
        return lib().oracle_lca(self._h, a, b)

    def num_quartets_at(self, v: int) -> int:
        return lib().oracle_num_quartets_at(self._h, v)
