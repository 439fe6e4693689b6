"""ctypes binding of the CUDA C-ABI (include/camus_b200.h, libcamus_b200.so).

This is the same boundary a cgo shim binds. There is no fallback: if the library
is missing or no sm_100 device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build

STAGES = ["upload", "prep", "count", "filter", "score", "final", "penalty", "_"]

_lib = None


class CamusB200Error(RuntimeError):
    pass


_dp_table = None


def dp_accel_table():
    """The ten camus_b200_dp_* entry points as an array of function pointers, in the order
    camus_host_run_dp_accel expects (camus_b200/host/dp.hpp DpAccelApi)."""
    global _dp_table
    if _dp_table is None:
        L = lib()
        names = ("create", "destroy", "last_error", "set_node", "set_vertex", "set_path_columns", "across", "launches",
                 "set_conv_columns", "across_maxplus")
        _dp_table = (C.c_void_p * 10)(*[C.cast(getattr(L, "camus_b200_dp_" + n), C.c_void_p).value for n in names])
    return _dp_table


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = os.environ.get("CAMUS_B200_LIB", build.CUDA_SO)  # override: A/B builds of the same library
        if not os.path.exists(so):
            raise CamusB200Error(
                f"{so} is missing: build it with `python -m camus_b200.build cuda` "
                "(there is no CPU fallback for the hot path)")
        L = C.CDLL(so)
        vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
        L.camus_b200_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
        L.camus_b200_destroy.argtypes = [vp]
        L.camus_b200_last_error.argtypes = [vp]
        L.camus_b200_last_error.restype = C.c_char_p
        L.camus_b200_set_partition.argtypes = [vp, C.c_int, C.c_int]
        L.camus_b200_set_constraint.argtypes = [vp, C.c_int32, C.c_int32, vp, vp, vp, vp]
        L.camus_b200_add_gene_trees.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.camus_b200_clear_gene_trees.argtypes = [vp]
        L.camus_b200_get_gene_trees.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(C.c_int32), u64p]
        L.camus_b200_set_tip_labels.argtypes = [vp, C.c_int32, C.c_char_p, vp]
        L.camus_b200_add_gene_trees_newick.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.camus_b200_set_score_hint.argtypes = [vp, C.c_int]
        L.camus_b200_count_filter.argtypes = [vp, C.c_int, C.c_double, u64p, u64p]
        L.camus_b200_quartet_totals.argtypes = [vp, C.c_int, vp]
        L.camus_b200_edge_penalties.argtypes = [vp, vp]
        L.camus_b200_export_quartets.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, u64p]
        L.camus_b200_query_cells.argtypes = [vp, C.c_int32, vp, vp]
        L.camus_b200_quartet_totals_partial.argtypes = [vp, C.c_int]
        L.camus_b200_partial_buffer.argtypes = [vp, C.POINTER(vp), u64p]
        L.camus_b200_quartet_totals_finish.argtypes = [vp, vp]
        L.camus_b200_run_resident.argtypes = [vp, C.c_int, C.c_double, C.c_int, vp, u64p, u64p]
        L.camus_b200_stage_ms.argtypes = [vp, C.POINTER(C.c_float), u64p]
        L.camus_b200_workload.argtypes = [vp, u64p, u64p, u64p]
        L.camus_b200_constraint_tables.argtypes = [vp, vp, vp, vp]
        L.camus_b200_comm_unique_id.argtypes = [vp]
        L.camus_b200_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.camus_b200_reticulation_counts.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


EXPORTED = [
    "camus_b200_create", "camus_b200_destroy", "camus_b200_last_error", "camus_b200_set_partition",
    "camus_b200_set_constraint", "camus_b200_add_gene_trees", "camus_b200_clear_gene_trees",
    "camus_b200_set_score_hint", "camus_b200_count_filter", "camus_b200_quartet_totals", "camus_b200_edge_penalties",
    "camus_b200_export_quartets", "camus_b200_query_cells", "camus_b200_quartet_totals_partial", "camus_b200_partial_buffer",
    "camus_b200_quartet_totals_finish", "camus_b200_run_resident", "camus_b200_stage_ms", "camus_b200_workload",
    "camus_b200_reticulation_counts", "camus_b200_constraint_tables", "camus_b200_comm_unique_id", "camus_b200_comm_init",
    "camus_b200_set_tip_labels", "camus_b200_add_gene_trees_newick", "camus_b200_get_gene_trees",
    "camus_b200_dp_create", "camus_b200_dp_destroy", "camus_b200_dp_last_error", "camus_b200_dp_set_node",
    "camus_b200_dp_set_vertex", "camus_b200_dp_set_path_columns", "camus_b200_dp_across", "camus_b200_dp_launches",
    "camus_b200_dp_set_conv_columns", "camus_b200_dp_across_maxplus",
]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class CamusB200:
    """One context on one GPU (or, with `devices`, on several GPUs of this process). Method names
    and argument meaning follow include/camus_b200.h."""

    def __init__(self, device: int = 0, devices=None):
        """devices: list of GPU ordinals for one handle that drives several GPUs of this process
        (camus_b200_create with n_gpus > 1); default is the single GPU `device`."""
        L = lib()
        h = C.c_void_p()
        devs = list(devices) if devices is not None else [device]
        dev = (C.c_int * len(devs))(*devs)
        rc = L.camus_b200_create(C.byref(h), len(devs), dev)
        if rc != 0:
            raise CamusB200Error(L.camus_b200_last_error(None).decode())
        self._h = h
        self.n = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().camus_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _ck(self, rc: int):
        if rc != 0:
            raise CamusB200Error(f"[{rc}] " + lib().camus_b200_last_error(self._h).decode())

    def set_partition(self, rank: int, world: int):
        self._ck(lib().camus_b200_set_partition(self._h, rank, world))

    def set_constraint(self, parent, child0, child1, tip_index, n_taxa: int):
        a = [np.ascontiguousarray(x, np.int32) for x in (parent, child0, child1, tip_index)]
        self.n = len(a[0])
        self._ck(lib().camus_b200_set_constraint(self._h, self.n, n_taxa, *[_p(x) for x in a]))

    def add_gene_trees(self, node_off, parent, taxon):
        node_off = np.ascontiguousarray(node_off, np.int64)
        parent = np.ascontiguousarray(parent, np.int32)
        taxon = np.ascontiguousarray(taxon, np.int32)
        self._ck(lib().camus_b200_add_gene_trees(self._h, len(node_off) - 1, _p(node_off), _p(parent), _p(taxon)))

    def set_tip_labels(self, names):
        """The constraint's tip names in tip-index order (camus_b200_set_tip_labels)."""
        enc = [n.encode() for n in names]
        off = np.zeros(len(enc) + 1, np.int64)
        off[1:] = np.cumsum([len(e) for e in enc])
        pool = b"".join(enc) or b"\0"
        self._ck(lib().camus_b200_set_tip_labels(self._h, len(enc), pool, _p(off)))

    def add_gene_trees_newick(self, text, min_support: float = 0.0):
        """Newick text -> flat trees on the GPU (camus_b200_add_gene_trees_newick). Returns
        (n_trees, missing_taxa); raises CamusB200Error with code 5 (fallback) and .bad_tree set when
        the caller has to parse the text itself."""
        if isinstance(text, str):
            text = text.encode()
        nt, mt, bt = C.c_int32(), C.c_int32(), C.c_int32()
        rc = lib().camus_b200_add_gene_trees_newick(self._h, text, len(text), float(min_support), C.byref(nt), C.byref(mt), C.byref(bt))
        if rc != 0:
            e = CamusB200Error(f"[{rc}] " + lib().camus_b200_last_error(self._h).decode())
            e.code, e.bad_tree = rc, bt.value
            raise e
        return nt.value, bool(mt.value)

    def get_gene_trees(self):
        """(node_off, parent, taxon) of the flat gene trees the context holds."""
        nt, nn = C.c_int32(), C.c_uint64()
        self._ck(lib().camus_b200_get_gene_trees(self._h, None, None, None, 0, C.byref(nt), C.byref(nn)))
        off = np.zeros(nt.value + 1, np.int64)
        par = np.zeros(max(nn.value, 1), np.int32)
        tax = np.zeros(max(nn.value, 1), np.int32)
        self._ck(lib().camus_b200_get_gene_trees(self._h, _p(off), _p(par), _p(tax), nn.value, C.byref(nt), C.byref(nn)))
        return off, par[:nn.value], tax[:nn.value]

    def load_newick(self, prob, gene_text, min_support: float = 0.0):
        """As load(), the gene trees parsed and flattened on the GPU from their newick text."""
        self.set_constraint(prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
        self.set_tip_labels(prob.tip_names)
        return self.add_gene_trees_newick(gene_text, min_support)

    def clear_gene_trees(self):
        self._ck(lib().camus_b200_clear_gene_trees(self._h))

    def load(self, prob):
        """prob: camus_b200.hostlib.Problem (its flat arrays are what the Go side would pass)."""
        self.set_constraint(prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
        if prob.n_trees:
            self.add_gene_trees(prob.gene_node_off, prob.gene_parent, prob.gene_taxon)
        return self

    def set_score_hint(self, as_set: bool):
        self._ck(lib().camus_b200_set_score_hint(self._h, int(as_set)))

    def count_filter(self, qmode: int = 0, threshold: float = 0.0):
        ns, sc = C.c_uint64(), C.c_uint64()
        self._ck(lib().camus_b200_count_filter(self._h, qmode, threshold, C.byref(ns), C.byref(sc)))
        return ns.value, sc.value

    def quartet_totals(self, as_set: bool = False, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n, self.n), np.uint64)
        self._ck(lib().camus_b200_quartet_totals(self._h, int(as_set), _p(out)))
        return out

    def edge_penalties(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n, self.n), np.uint64)
        self._ck(lib().camus_b200_edge_penalties(self._h, _p(out)))
        return out

    def export_quartets(self, raw: bool = False):
        n = C.c_uint64()
        self._ck(lib().camus_b200_export_quartets(self._h, int(raw), None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, np.uint64)
        counts = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(lib().camus_b200_export_quartets(self._h, int(raw), _p(keys), _p(counts), n.value, C.byref(n)))
        return keys, counts

    def constraint_tables(self):
        """(lca [n, n] int32, depth [n] int32, leaves_below [n] uint64): graphs.MakeTreeData's tables
        (treedata.go:130-194) in the caller's node ids."""
        lca = np.empty((self.n, self.n), np.int32)
        depth = np.empty(self.n, np.int32)
        below = np.empty(self.n, np.uint64)
        self._ck(lib().camus_b200_constraint_tables(self._h, _p(lca), _p(depth), _p(below)))
        return lca, depth, below

    @staticmethod
    def comm_unique_id() -> bytes:
        """Rank 0: the 128-byte NCCL id every rank passes to comm_init."""
        buf = (C.c_uint8 * 128)()
        if lib().camus_b200_comm_unique_id(buf) != 0:
            raise CamusB200Error(lib().camus_b200_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        """Attach this context to the library-owned NCCL communicator (collective call); sets the
        partition (rank, world). Afterwards count_filter / quartet_totals / run_resident return
        global results on every rank."""
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(lib().camus_b200_comm_init(self._h, buf, rank, world))

    def query_cells(self, quads) -> np.ndarray:
        """Raw counts of arbitrary 4-taxon sets (tip indices), reference topology order."""
        quads = np.ascontiguousarray(quads, np.int32).reshape(-1, 4)
        out = np.zeros((len(quads), 3), np.uint32)
        self._ck(lib().camus_b200_query_cells(self._h, len(quads), _p(quads), _p(out)))
        return out

    def reticulation_counts(self, net):
        """score.ReticulationScore's per-gene-tree loop (score.go:31-60) for a hostlib.NetworkProblem:
        (totals, supported), uint64 [n_trees, n_ret]."""
        tot = np.zeros((net.n_trees, net.n_ret), np.uint64)
        sup = np.zeros((net.n_trees, net.n_ret), np.uint64)
        a = [np.ascontiguousarray(x) for x in (net.records, net.gene_node_off, net.gene_parent, net.gene_taxon)]
        self._ck(lib().camus_b200_reticulation_counts(self._h, net.n_tips, net.n_ret, _p(a[0]), net.n_trees, _p(a[1]), _p(a[2]),
                                                      _p(a[3]), _p(tot), _p(sup)))
        return tot, sup

    # ---- split form for multi-process runs ----
    def quartet_totals_partial(self, as_set: bool = False):
        self._ck(lib().camus_b200_quartet_totals_partial(self._h, int(as_set)))

    def partial_buffer(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().camus_b200_partial_buffer(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def quartet_totals_finish(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n, self.n), np.uint64)
        self._ck(lib().camus_b200_quartet_totals_finish(self._h, _p(out)))
        return out

    # ---- instrumentation ----
    def run_resident(self, qmode: int, threshold: float, as_set: bool, out: np.ndarray | None = None,
                     partial_only: bool = False):
        ns, sc = C.c_uint64(), C.c_uint64()
        if partial_only:
            self._ck(lib().camus_b200_run_resident(self._h, qmode, threshold, int(as_set), None, C.byref(ns), C.byref(sc)))
            return None, ns.value, sc.value
        if out is None:
            out = np.empty((self.n, self.n), np.uint64)
        self._ck(lib().camus_b200_run_resident(self._h, qmode, threshold, int(as_set), _p(out), C.byref(ns), C.byref(sc)))
        return out, ns.value, sc.value

    def partial_tensor(self):
        """The partial junction table as a zero-copy torch int64 tensor (for torch.distributed)."""
        import torch
        ptr, n = self.partial_buffer()

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

        return torch.as_tensor(_Holder(), device="cuda")

    def stage_ms(self):
        ms = (C.c_float * 8)()
        nl = C.c_uint64()
        self._ck(lib().camus_b200_stage_ms(self._h, ms, C.byref(nl)))
        return {STAGES[i]: ms[i] for i in range(7)}, nl.value

    def workload(self):
        r, c, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(lib().camus_b200_workload(self._h, C.byref(r), C.byref(c), C.byref(b)))
        return {"resolutions": r.value, "cells": c.value, "boxes": b.value}
