"""ctypes binding of libcamus_host.so: the host-side mirror of the reference's Go packages
(parsing + flattening before the hot path, scorers + DP + network writer after it).

Nothing here counts or scores quartets; that is `camus_b200.cabi` (CUDA, include/camus_b200.h).
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
from dataclasses import dataclass

import numpy as np

from . import build

# camus::Err (host/tree.hpp) <-> the reference's sentinel errors
ERR_NAMES = {
    1: "ErrUnrooted", 2: "ErrNonBinary", 3: "ErrMulTree", 4: "ErrTipNameMismatch",
    5: "ErrInvalidFile", 6: "ErrInvalidFormat", 7: "ErrTypeOutRange", 8: "ErrInvalidScorerOption",
    9: "ErrDevice", 10: "ErrInternal", 11: "ErrNoReticulations", 12: "ErrNotLevel1",
}


class CamusError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.name = ERR_NAMES.get(code, str(code))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(build.HOST_SO):
            build.build_host()
        L = C.CDLL(build.HOST_SO)
        L.camus_host_last_error.restype = C.c_char_p
        L.camus_host_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.POINTER(C.c_void_p)]
        L.camus_host_free.argtypes = [C.c_void_p]
        L.camus_host_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3 + [C.POINTER(C.c_int64)]
        for nm in ("constraint_parent", "constraint_child0", "constraint_child1", "constraint_tip_index",
                   "gene_parent", "gene_taxon"):
            f = getattr(L, "camus_host_" + nm)
            f.argtypes = [C.c_void_p]
            f.restype = C.POINTER(C.c_int32)
        L.camus_host_gene_node_off.argtypes = [C.c_void_p]
        L.camus_host_gene_node_off.restype = C.POINTER(C.c_int64)
        L.camus_host_warned_missing.argtypes = [C.c_void_p]
        L.camus_host_percent_no_support.argtypes = [C.c_void_p]
        L.camus_host_percent_no_support.restype = C.c_double
        L.camus_host_constraint_newick.argtypes = [C.c_void_p]
        L.camus_host_constraint_newick.restype = C.c_char_p
        L.camus_host_tip_name.argtypes = [C.c_void_p, C.c_int]
        L.camus_host_tip_name.restype = C.c_char_p
        L.camus_host_node_name.argtypes = [C.c_void_p, C.c_int]
        L.camus_host_node_name.restype = C.c_char_p
        L.camus_host_should_calc_edge.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.camus_host_cycle_length.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.camus_host_run_dp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64,
                                        C.c_uint64, C.c_int, C.c_double, C.POINTER(C.c_void_p)]
        L.camus_host_run_dp_accel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64,
                                              C.c_uint64, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_int64,
                                              C.POINTER(C.c_void_p)]
        L.camus_host_dp_mock_api.restype = C.c_void_p
        L.camus_host_results_accel_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.camus_host_results_accel_stats.restype = None
        L.camus_host_results_free.argtypes = [C.c_void_p]
        L.camus_host_results_count.argtypes = [C.c_void_p]
        L.camus_host_results_branches.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        L.camus_host_results_qsat.argtypes = [C.c_void_p, C.c_int]
        L.camus_host_results_qsat.restype = C.c_double
        L.camus_host_results_root_score.argtypes = [C.c_void_p, C.c_int]
        L.camus_host_results_root_score.restype = C.c_double
        L.camus_host_results_newick.argtypes = [C.c_void_p, C.c_int]
        L.camus_host_results_newick.restype = C.c_char_p
        L.camus_host_make_network.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        L.camus_host_make_network.restype = C.c_char_p
        L.camus_host_check_scorer_options.argtypes = [C.c_char_p, C.c_int, C.c_double]
        vp = C.c_void_p
        L.camus_host_network_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
        L.camus_host_network_free.argtypes = [vp]
        L.camus_host_network_dims.argtypes = [vp] + [C.POINTER(C.c_int32)] * 4 + [C.POINTER(C.c_int64)]
        L.camus_host_network_records.argtypes = [vp]
        L.camus_host_network_records.restype = C.POINTER(C.c_uint32)
        L.camus_host_network_gene_node_off.argtypes = [vp]
        L.camus_host_network_gene_node_off.restype = C.POINTER(C.c_int64)
        for nm in ("gene_parent", "gene_taxon", "parent", "tip_index", "uwvs"):
            f = getattr(L, "camus_host_network_" + nm)
            f.argtypes = [vp]
            f.restype = C.POINTER(C.c_int32)
        L.camus_host_network_label.argtypes = [vp, C.c_int]
        L.camus_host_network_label.restype = C.c_char_p
        L.camus_host_network_node_name.argtypes = [vp, C.c_int]
        L.camus_host_network_node_name.restype = C.c_char_p
        L.camus_host_format_ratio.argtypes = [vp, C.c_uint64, C.c_uint64]
        L.camus_host_format_ratio.restype = C.c_char_p
        _lib = L
    return _lib


def read_text(path: str) -> str:
    if path.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return f.read()
    with open(path) as f:
        return f.read()


@dataclass
class DPResult:
    branches: list[list[tuple[int, int]]]  # per k
    qsat: list[float]
    root_scores: list[float]               # DP[root][k], k = 0..m
    newicks: list[str]
    accel: dict | None = None              # DP offload statistics (run_dp(gpu=...))


class Problem:
    """Parsed + validated + flattened input: what the Go side hands to the C-ABI.

    Mirrors prep.ReadInputFiles (io.go:81) and the host half of prep.Preprocess
    (preprocess.go:26-44, 87-100)."""

    def __init__(self, constraint_text: str, gene_text: str, fmt: str = "newick", min_support: float = 0.0):
        L = lib()
        h = C.c_void_p()
        rc = L.camus_host_load(constraint_text.encode(), gene_text.encode(), 1 if fmt == "nexus" else 0,
                               float(min_support), C.byref(h))
        if rc != 0:
            raise CamusError(rc, L.camus_host_last_error().decode())
        self._h = h
        nn, nt, ng = C.c_int32(), C.c_int32(), C.c_int32()
        tot = C.c_int64()
        L.camus_host_dims(h, C.byref(nn), C.byref(nt), C.byref(ng), C.byref(tot))
        self.n_nodes, self.n_taxa, self.n_trees, self.total_gene_nodes = nn.value, nt.value, ng.value, tot.value

        def arr(fn, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(fn(h), shape=(n,)).astype(dt, copy=True)

        self.parent = arr(L.camus_host_constraint_parent, self.n_nodes, np.int32)
        self.child0 = arr(L.camus_host_constraint_child0, self.n_nodes, np.int32)
        self.child1 = arr(L.camus_host_constraint_child1, self.n_nodes, np.int32)
        self.tip_index = arr(L.camus_host_constraint_tip_index, self.n_nodes, np.int32)
        self.gene_node_off = arr(L.camus_host_gene_node_off, self.n_trees + 1, np.int64)
        self.gene_parent = arr(L.camus_host_gene_parent, self.total_gene_nodes, np.int32)
        self.gene_taxon = arr(L.camus_host_gene_taxon, self.total_gene_nodes, np.int32)
        self.warned_missing = bool(L.camus_host_warned_missing(h))
        self.constraint_newick = L.camus_host_constraint_newick(h).decode()
        self.tip_names = [L.camus_host_tip_name(h, i).decode() for i in range(self.n_taxa)]
        self.node_names = [L.camus_host_node_name(h, i).decode() for i in range(self.n_nodes)]

    @classmethod
    def from_files(cls, constraint_file: str, gene_file: str, fmt: str = "newick", min_support: float = 0.0):
        return cls(read_text(constraint_file), read_text(gene_file), fmt, min_support)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().camus_host_free(self._h)
            self._h = None

    def node_id(self, label: str) -> int:
        return self.node_names.index(label)

    def tip_id(self, label: str) -> int:
        return self.tip_names.index(label)

    def should_calc_edge(self, u: int, w: int) -> bool:
        return bool(lib().camus_host_should_calc_edge(self._h, u, w))

    def cycle_length(self, u: int, w: int) -> int:
        return lib().camus_host_cycle_length(self._h, u, w)

    def make_network(self, branches) -> str:
        """graphs.MakeNetwork(td, branches).Newick() (graphs/net.go:38-101); branches = [(u, w), ...] node ids."""
        flat = (C.c_int32 * (2 * len(branches)))(*[x for b in branches for x in b])
        r = lib().camus_host_make_network(self._h, len(branches), flat)
        if r is None:
            raise CamusError(10, lib().camus_host_last_error().decode())
        return r.decode()

    def run_dp(self, totals: np.ndarray, penalties: np.ndarray | None = None, mode: str = "max", as_set: bool = False,
               n_survivors: int = 0, sum_counts: int = 0, n_gtrees: int | None = None, alpha: float = 0.1,
               gpu: int | None = None, min_work: int = -1) -> DPResult:
        """infer.Infer after Scorer.Init (infer.go:79-95): DP + traceback + MakeNetwork. gpu = device
        index: the DP's across-scan (main_dp.go:247-287) runs there through the camus_b200_dp_* calls
        of libcamus_b200.so; None = the host scan."""
        L = lib()
        accel = None
        if gpu == "mock":  # CPU stand-in of the device scan (host/dp_mock.hpp): tests without a GPU
            accel, gpu = C.c_void_p(L.camus_host_dp_mock_api()), 0
        elif gpu is not None:
            from . import cabi
            accel = cabi.dp_accel_table()
        totals = np.ascontiguousarray(totals, dtype=np.uint64)
        assert totals.size == self.n_nodes * self.n_nodes
        pen_p = None
        if penalties is not None:
            penalties = np.ascontiguousarray(penalties, dtype=np.uint64)
            pen_p = penalties.ctypes.data_as(C.c_void_p)
        r = C.c_void_p()
        rc = L.camus_host_run_dp_accel(self._h, {"max": 0, "norm": 1, "sym": 2}[mode], totals.ctypes.data_as(C.c_void_p),
                                       pen_p, int(as_set), int(n_survivors), int(sum_counts),
                                       int(self.n_trees if n_gtrees is None else n_gtrees), float(alpha),
                                       C.cast(accel, C.c_void_p) if accel is not None else None, int(gpu or 0), int(min_work),
                                       C.byref(r))
        if rc != 0:
            raise CamusError(rc, L.camus_host_last_error().decode())
        try:
            m = L.camus_host_results_count(r)
            branches, qsat, newicks = [], [], []
            roots = [L.camus_host_results_root_score(r, 0)]
            for k in range(1, m + 1):
                buf = (C.c_int32 * (2 * k))()
                nb = L.camus_host_results_branches(r, k, buf)
                branches.append([(buf[2 * i], buf[2 * i + 1]) for i in range(nb)])
                qsat.append(L.camus_host_results_qsat(r, k))
                roots.append(L.camus_host_results_root_score(r, k))
                newicks.append(L.camus_host_results_newick(r, k).decode())
            res = DPResult(branches, qsat, roots, newicks)
            st = (C.c_uint64 * 8)()
            L.camus_host_results_accel_stats(r, st)
            res.accel = {"device_steps": st[0], "host_steps": st[1], "nan_steps": st[2], "launches": st[3],
                         "upload_s": st[4] * 1e-6, "across_s": st[5] * 1e-6,
                         "create_s": st[6] * 1e-6, "destroy_s": st[7] * 1e-6}
            return res
        finally:
            L.camus_host_results_free(r)


def check_scorer_options(mode: str, n_gtrees: int = 1, alpha: float = 0.1) -> None:
    """score.ParseScorer + WithNGtrees / WithAlpha as infer.Infer applies them (infer.go:79-88,
    scorers.go:12-38). Raises KeyError for an unknown mode, CamusError(ErrInvalidScorerOption) else."""
    rc = lib().camus_host_check_scorer_options(mode.encode(), int(n_gtrees), float(alpha))
    if rc == -1:
        raise KeyError(mode)
    if rc != 0:
        raise CamusError(rc, lib().camus_host_last_error().decode())


class NetworkProblem:
    """Inputs of score.ReticulationScore (internal/score/score.go:25-72), parsed and flattened: the
    level-1 network (prep.ConvertToNetwork, io.go:173-236) reduced to one 16-byte record per
    (reticulation, network tip), and the gene trees as CSR with NETWORK tip indices. What the Go
    side would hand to camus_b200_reticulation_counts."""

    def __init__(self, network_text: str, gene_text: str, fmt: str = "newick"):
        L = lib()
        h = C.c_void_p()
        rc = L.camus_host_network_load(network_text.encode(), gene_text.encode(), 1 if fmt == "nexus" else 0, C.byref(h))
        if rc != 0:
            raise CamusError(rc, L.camus_host_last_error().decode())
        self._h = h
        nn, nt, nr, ng = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        tot = C.c_int64()
        L.camus_host_network_dims(h, C.byref(nn), C.byref(nt), C.byref(nr), C.byref(ng), C.byref(tot))
        self.n_nodes, self.n_tips, self.n_ret, self.n_trees, self.total_gene_nodes = nn.value, nt.value, nr.value, ng.value, tot.value

        def arr(fn, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(fn(h), shape=(n,)).astype(dt, copy=True)

        self.records = arr(L.camus_host_network_records, self.n_ret * self.n_tips * 4, np.uint32).reshape(self.n_ret, self.n_tips, 4)
        self.gene_node_off = arr(L.camus_host_network_gene_node_off, self.n_trees + 1, np.int64)
        self.gene_parent = arr(L.camus_host_network_gene_parent, self.total_gene_nodes, np.int32)
        self.gene_taxon = arr(L.camus_host_network_gene_taxon, self.total_gene_nodes, np.int32)
        self.parent = arr(L.camus_host_network_parent, self.n_nodes, np.int32)
        self.tip_index = arr(L.camus_host_network_tip_index, self.n_nodes, np.int32)
        self.uwvs = arr(L.camus_host_network_uwvs, 4 * self.n_ret, np.int32).reshape(self.n_ret, 4)
        self.labels = [L.camus_host_network_label(h, r).decode() for r in range(self.n_ret)]
        self.node_names = [L.camus_host_network_node_name(h, i).decode() for i in range(self.n_nodes)]

    @classmethod
    def from_files(cls, network_file: str, gene_file: str, fmt: str = "newick"):
        return cls(read_text(network_file), read_text(gene_file), fmt)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().camus_host_network_free(self._h)
            self._h = None

    def scores(self, totals: np.ndarray, supported: np.ndarray) -> np.ndarray:
        """float64(supported) / float64(totals), NaN where nothing was comparable (score.go:62-68)."""
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(totals == 0, np.nan, supported.astype(np.float64) / totals.astype(np.float64))

    def csv(self, totals: np.ndarray, supported: np.ndarray, names=None) -> str:
        """prep.WriteRetScoresToCSV (io.go:315-340)."""
        L = lib()
        rows = ["gene," + ",".join(self.labels)]
        for g in range(self.n_trees):
            cells = [L.camus_host_format_ratio(self._h, int(supported[g, r]), int(totals[g, r])).decode() for r in range(self.n_ret)]
            rows.append(",".join([str(names[g] if names else g + 1)] + cells))
        return "\n".join(rows)
