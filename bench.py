#!/usr/bin/env python
"""bench.py -- throughput of the CAMUS quartet hot path on B200 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)

A step = one pass of the whole hot path (LCA-depth tables -> rooted-triple bit-planes -> quartet
count -> -q/-t filter + constraint removal -> edge-score scatter -> n x n table) over one batch of
synthetic gene trees. `value` = gene-tree quartets resolved per second with the flat trees already
in HBM; `e2e` = the same through the C-ABI with HOST buffers (H2D of the trees and D2H of the
table inside the timed region). With N ranks the 4-taxon-set index space is sharded N ways, every
rank holds all trees, and the partial junction tables are summed with one NCCL all-reduce.

`--impl reference` times the reference's own algorithm (the oracle's LITERAL mode: bipartition
enumerator, per-tree set, map merge, filterQuartets, constraint removal, mapQuartetsToVertices and
the per-(u, w) QuartetScore loop) on the host cores: config 1 in full; for the larger configurations
a bounded sample (all gene trees induced on 50 of the taxa = the size the reference runs as-is), and
says so. Both arms print the same `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "gene_tree_quartets_resolved_per_sec"
UNIT = "quartets/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        period = 0.1
        while not self.stop_flag:
            t0 = time.perf_counter()
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            # NVML queries take a driver lock that CUDA calls wait on: 50 per second delayed the launches of small
            # configurations by milliseconds, and on some boxes one query pair costs tens of milliseconds (config 5:
            # 17 -> 34 ms per step with 10 samples per second). Back off when the queries are slow.
            if time.perf_counter() - t0 > 0.005:
                period = 1.0
            time.sleep(period)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_problem(cfg_id: int, n_taxa, n_trees):
    from camus_b200 import synth
    from camus_b200.hostlib import Problem
    s, c = synth.make_config(cfg_id, n_taxa, n_trees)
    t0 = time.perf_counter()
    prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
    prob.host_parse_s = time.perf_counter() - t0  # parse + validate + flatten on all host cores (constraint included)
    prob.gene_text, prob.gene_fmt = s.genes, s.fmt
    return prob, c


def workload_name(cfg_id, c):
    return (f"config{cfg_id}: {c['n_taxa']}-taxon constraint + {c['n_trees']} gene trees, -q {c['qmode']} -t {c['threshold']}"
            f"{' -asSet' if c['as_set'] else ''} {c['mode']} mode")


def config_dict(cfg_id, c):
    """The `config` object of the JSON line: identical for both arms (`--impl ours|reference`)."""
    from math import comb
    n, g = c["n_taxa"], c["n_trees"]
    m = -(-n // 8)
    tiles = (m + 2) * (m + 1) * m // 6
    boxes = comb(m + 3, 4)
    groups = -(-g // 32)
    table = groups * 6144 * tiles
    l2 = ("explicit flush: 512 MB written before every timed step" if table < (1 << 30) else
          f"no flush needed: every step streams a {table / 1e9:.2f} GB rooted-triple bit-plane table "
          f"({boxes * groups * 4 * 4096 / 1e9:.0f} to {boxes * groups * 4 * 6144 / 1e9:.0f} GB of tile reads), far beyond the 126 MB L2")
    return {"workload": workload_name(cfg_id, c), "taxa": n, "gene_trees": g, "cells": comb(n, 4), "boxes": boxes, "l2": l2}


def load_digests(cfg_id, c):
    p = os.path.join(ROOT, "tests", "golden", "synthetic", "config_digests.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    from camus_b200 import synth
    base = synth.CONFIGS[cfg_id]
    if c["n_taxa"] != base["n_taxa"] or c["n_trees"] != base["n_trees"]:
        return None
    return d.get("config%d" % cfg_id)


def table_digest(tab):
    import zlib
    return "%08x-%016x" % (zlib.crc32(np.ascontiguousarray(tab).tobytes()), int(tab.sum(dtype=np.uint64)))


# ------------------------------------------------------------------------------------------
def restrict_to_taxa(prob, keep_tips, n_trees):
    """Induced sub-problem on a subset of the taxa (flat arrays only): the constraint tree with
    unifurcations suppressed, the first n_trees gene trees with the other leaves pruned."""
    keep = np.zeros(prob.n_taxa, bool)
    keep[list(keep_tips)] = True
    new_tip = np.cumsum(keep) - 1
    n = prob.n_nodes
    kept_below = np.zeros(n, np.int64)
    for i in range(n - 1, -1, -1):  # ids are preorder: children after parents
        if prob.child0[i] < 0:
            kept_below[i] = int(keep[prob.tip_index[i]])
        else:
            kept_below[i] = kept_below[prob.child0[i]] + kept_below[prob.child1[i]]
    retained = [i for i in range(n) if kept_below[i] > 0 and (prob.child0[i] < 0 or (kept_below[prob.child0[i]] > 0 and kept_below[prob.child1[i]] > 0))]
    new_id = {x: j for j, x in enumerate(retained)}
    m = len(retained)
    par, c0, c1, tip = (np.full(m, -1, np.int32) for _ in range(4))
    for j, x in enumerate(retained):
        p = prob.parent[x]
        while p >= 0 and p not in new_id:
            p = prob.parent[p]
        par[j] = new_id[p] if p >= 0 else -1
        if par[j] >= 0:
            if c0[par[j]] < 0:
                c0[par[j]] = j
            else:
                c1[par[j]] = j
        if prob.child0[x] < 0:
            tip[j] = new_tip[prob.tip_index[x]]
    off, gp, gt = [0], [], []
    for g in range(min(n_trees, prob.n_trees)):
        lo, hi = int(prob.gene_node_off[g]), int(prob.gene_node_off[g + 1])
        p, t = prob.gene_parent[lo:hi].tolist(), prob.gene_taxon[lo:hi].tolist()
        m = hi - lo
        kb = [0] * m   # kept leaves below
        nk = [0] * m   # children with kept leaves below
        for i in range(m - 1, -1, -1):
            if t[i] >= 0 and keep[t[i]]:
                kb[i] += 1
            if p[i] >= 0 and kb[i] > 0:
                kb[p[i]] += kb[i]
                nk[p[i]] += 1
        # a node stays if it is a kept leaf or still branches (unifurcations are suppressed, as
        # gotree does when the reference reads a pruned tree)
        idx = [-1] * m
        anc = [-1] * m  # nearest retained ancestor (new index)
        k = 0
        for i in range(m):
            up = anc[p[i]] if p[i] >= 0 else -1
            if kb[i] > 0 and ((t[i] >= 0 and keep[t[i]]) or nk[i] >= 2):
                idx[i] = k
                k += 1
                gp.append(up)
                gt.append(int(new_tip[t[i]]) if t[i] >= 0 else -1)
                anc[i] = idx[i]
            else:
                anc[i] = up
        off.append(len(gp))
    return (par, c0, c1, tip, int(keep.sum())), (np.array(off, np.int64), np.array(gp, np.int32), np.array(gt, np.int32))


REF_TAXA = 50  # the reference runs BASELINE.json configs[0] (50 taxa) as-is; larger inputs are sampled at that size


class CpuSample:
    """The reference's own algorithm (oracle LITERAL mode) on a bounded sample of the workload:
    the whole problem when it has at most REF_TAXA taxa (config 1: in full, nothing extrapolated),
    else ALL gene trees induced on REF_TAXA of the taxa (every k-th tip). One run = processQuartets
    + filterQuartets + constraint removal (preprocess.go:44-57), then mapQuartetsToVertices +
    CalculateQuartetTotals (treedata.go:197-222, quartettotals.go:43-61)."""

    def __init__(self, prob, c, nthreads):
        from math import comb

        from oracle import oracle as orc
        self.c, self.nthreads, self.orc = c, nthreads, orc
        self.full = prob.n_taxa <= REF_TAXA
        if self.full:
            self.cons = (prob.parent, prob.child0, prob.child1, prob.tip_index, prob.n_taxa)
            self.genes = (prob.gene_node_off, prob.gene_parent, prob.gene_taxon)
            self.desc = f"the whole input ({prob.n_taxa} taxa x {prob.n_trees} gene trees), nothing sampled or extrapolated"
        else:
            step = -(-prob.n_taxa // REF_TAXA)
            self.cons, self.genes = restrict_to_taxa(prob, range(0, prob.n_taxa, step), prob.n_trees)
            self.desc = (f"all {prob.n_trees} gene trees induced on {self.cons[4]} of the {prob.n_taxa} taxa (every {step}th tip); the "
                         "literal algorithm gets slower per resolution as taxa grow (duplicate emissions, map size), so this "
                         "OVERSTATES the reference at the full size, which it cannot run at all above a few hundred taxa")
        off, tax = self.genes[0], self.genes[2]
        self.resolutions = 0
        for g in range(len(off) - 1):
            self.resolutions += comb(int((tax[off[g]:off[g + 1]] >= 0).sum()), 4)

    def run(self, literal=True):
        """(seconds count+filter+removal, seconds edge totals)"""
        o = self.orc.Oracle(self.nthreads)
        o.set_constraint(*self.cons[:4], self.cons[4])
        o.add_gene_trees(*self.genes)
        t0 = time.perf_counter()
        if literal:
            o.count_filter(self.c["qmode"], self.c["threshold"], True)
            t1 = time.perf_counter()
            o.quartet_totals(self.c["as_set"], True)
        else:
            o.dense_run(self.c["qmode"], self.c["threshold"], self.c["as_set"], digests=False)
            t1 = time.perf_counter()
        return t1 - t0, time.perf_counter() - t1

    def record(self, n_runs=1):
        ts = [self.run(True) for _ in range(n_runs)]
        lit = sum(a + b for a, b in ts) / n_runs
        fast = sum(self.run(False))
        return {
            "value": self.resolutions / lit, "unit": UNIT, "cores": self.nthreads, "kind": "port",
            "sample": "oracle literal mode (the Go algorithms, function by function) end to end -- processQuartets + filterQuartets + "
                      "constraint removal + mapQuartetsToVertices + per-(u,w) QuartetScore loop -- on " + self.desc,
            "sample_resolutions": self.resolutions, "seconds": lit,
            "seconds_count_filter": sum(a for a, _ in ts) / n_runs, "seconds_edge_totals": sum(b for _, b in ts) / n_runs,
            "full_size": self.full,
            "fast_mode": {"value": self.resolutions / fast, "seconds": fast,
                          "what": "same sample, oracle dense mode (vectorised four-point counting + per-quartet scatter: a "
                                  "different, faster CPU algorithm than the reference's)"},
        }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, c = make_problem(args.config, args.taxa, args.trees)
    nthreads = min(os.cpu_count() or 1, args.ref_threads)
    smp = CpuSample(prob, c, nthreads)
    times = []
    for i in range(args.warmup + args.steps):
        a, b = smp.run(True)
        if i >= args.warmup:
            times.append(a + b)
    dt = sum(times) / len(times)
    val = smp.resolutions / dt
    fast = sum(smp.run(False))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": config_dict(args.config, c),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "oracle literal mode (the Go algorithms, function by function) end to end -- processQuartets + "
                                   "filterQuartets + constraint removal + mapQuartetsToVertices + per-(u,w) QuartetScore loop -- on " + smp.desc,
                         "sample_resolutions": smp.resolutions, "full_size": smp.full,
                         "fast_mode": {"value": smp.resolutions / fast, "seconds": fast,
                                       "what": "same sample, oracle dense mode (a different, faster CPU algorithm than the reference's)"}},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from camus_b200.cabi import CamusB200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob, c = make_problem(args.config, args.taxa, args.trees)
    # Launched under torchrun: one process per GPU. Launched as a plain process with --gpus N > 1:
    # ONE handle drives the N GPUs (camus_b200_create n_gpus = N), no torch.distributed involved.
    n_dev = args.gpus if world == 1 and args.gpus > 1 else 1
    gpu = CamusB200(devices=list(range(n_dev))) if n_dev > 1 else CamusB200(local)
    shares = world * n_dev
    # One process per GPU: the collective lives in the library (camus_b200_comm_init: rank 0's NCCL id
    # is handed round once, then every call returns global results). --collective torch keeps the
    # round-1 form (set_partition + torch.distributed.all_reduce on the partial table).
    lib_comm = world > 1 and args.collective == "library"
    if lib_comm:
        obj = [CamusB200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        gpu.comm_init(obj[0], rank, world)
    else:
        gpu.set_partition(rank, world)
    gpu.load(prob)
    n = prob.n_nodes
    out = np.empty((n, n), np.uint64)
    qmode, thr, as_set = c["qmode"], c["threshold"], c["as_set"]
    gpu.set_score_hint(as_set)  # as the host mirror's Preprocess does: one fused pass serves both calls of a step
    scal = torch.zeros(2, dtype=torch.int64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_and_finish(ns, sc):
        if world > 1 and not lib_comm:
            z = gpu.partial_tensor()
            dist.all_reduce(z)
            scal[0], scal[1] = ns, sc
            dist.all_reduce(scal)
            torch.cuda.synchronize()
            ns, sc = int(scal[0]), int(scal[1])
        gpu.quartet_totals_finish(out)
        return ns, sc

    def step_e2e():
        gpu.clear_gene_trees()
        gpu.add_gene_trees(prob.gene_node_off, prob.gene_parent, prob.gene_taxon)
        ns, sc = gpu.count_filter(qmode, thr)
        ms1, l1 = gpu.stage_ms()
        gpu.quartet_totals_partial(as_set)
        r = reduce_and_finish(ns, sc)
        return r, ms1, l1

    def step_resident():
        if lib_comm:
            _, ns, sc = gpu.run_resident(qmode, thr, as_set, out)
            return ns, sc
        _, ns, sc = gpu.run_resident(qmode, thr, as_set, None, partial_only=True)
        return reduce_and_finish(ns, sc)

    # first call uploads the trees and builds everything once (also a correctness anchor)
    first, _, _ = step_e2e()
    R = gpu.workload()["resolutions"]

    m_seg = -(-prob.n_taxa // 8)
    num_tiles = (m_seg + 2) * (m_seg + 1) * m_seg // 6
    groups = -(-prob.n_trees // 32)
    table_bytes = groups * 6144 * num_tiles
    # Timing rule: inputs larger than L2, or an explicit flush between steps. The bit-plane table a
    # step streams is far larger than the 126 MB L2 from config 2 up; for small inputs (config 1)
    # every step is timed on its own after a 512 MB write.
    flush = table_bytes < (1 << 30)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if flush else None

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        if flush:
            dt = 0.0
            for _ in range(steps):
                flush_buf.zero_()
                barrier()
                t0 = time.perf_counter()
                r = fn()
                torch.cuda.synchronize()
                dt += (time.perf_counter() - t0) / steps
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                r = fn()
            barrier()
            dt = (time.perf_counter() - t0) / steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        return dt, r

    sampler = ClockSampler(local)
    sampler.start()
    dt_res, r_res = timed(step_resident, args.steps, args.warmup)
    stage_ms, launches = gpu.stage_ms()
    clocks = sampler.result()
    dt_e2e, r_e2e = timed(lambda: step_e2e()[0], args.steps, max(1, args.warmup // 2))
    assert r_res == first and r_e2e == first, (first, r_res, r_e2e)
    # the result of the timed steps against the committed digests of this configuration
    digest_want = load_digests(args.config, c)
    digest_got = {"n_survivors": first[0], "sum_counts": first[1],
                  "quartet_totals_as_set" if as_set else "quartet_totals": table_digest(out)}
    if digest_want is not None and rank == 0:
        for k, v in digest_got.items():
            assert digest_want[k] == v, f"result differs from tests/golden/synthetic/config_digests.json: {k} = {v}, expected {digest_want[k]}"

    if rank == 0:
        peak, peak_src = load_peaks()
        wl = gpu.workload()
        cells, num_boxes = wl["cells"], wl["boxes"]
        # ---- roofline of the dominant kernel: the fused count kernel (count + filter + score) ----
        # Algorithmic bytes per launch, SURVEY.md 8(d) convention: 8 B per resolution (one count
        # cell read + written per (gene tree, 4-set)) + 16 B per cell (one 128-bit cell read by
        # filter / score), for the share of the 4-set space this rank owns.
        alg_bytes_kernel = (8 * R + 16 * cells) / shares
        kern_ms = stage_ms["count"]
        achieved = alg_bytes_kernel / (kern_ms * 1e-3) / 1e9
        b_alg = 8 * R + 16 * cells + 24 * n * n
        # Bytes the kernel's own design has to move per launch: four triple tiles per (box, tree
        # group), read once each (HBM or L2); 8 of their 12 planes for complete binary trees.
        nl = np.diff(np.concatenate([[0], np.cumsum(prob.gene_taxon >= 0)[prob.gene_node_off[1:] - 1]]))
        complete = bool((nl == prob.n_taxa).all()) and not c.get("support", False)
        design_bytes = (num_boxes / shares) * groups * 4 * 6144 * (2.0 / 3.0 if complete else 1.0)
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf) and shares == 1:  # the captures are single-GPU launches
            traffic = json.load(open(tf)).get(f"config{args.config}_k_count_dram_bytes_per_launch")
        h2d = int(prob.gene_node_off.nbytes + prob.gene_parent.nbytes + prob.gene_taxon.nbytes)
        sm_mhz = clocks.get("sm_mhz") or 1965
        # The kernel's floor: one POPC per (cell, counted topology, 32 trees) on the XU pipe, 16 lanes per clock
        # and SM (one warp-POPC per 8 clocks and SM sub-partition). The rate was verified on this GPU with a
        # synthetic POPC loop under ncu (tools/lds_probe.cu, profiles/r02_pipe_probe.txt: 8.1 - 8.3 clocks per
        # warp-POPC and sub-partition at 8 .. 32 resident warps).
        packed = (not complete) and prob.n_trees <= 1023 and os.environ.get("CAMUS_B200_NO_PACKED", "0") != "1"
        popcs = (num_boxes / shares) * groups * 4096 * (2 if complete else 3)
        xu_floor_ms = popcs / (16 * 148 * sm_mhz * 1e6) * 1e3
        # shared-memory side of the same loop: wavefronts per thread and tree group (LDS.128 = 4, or 2 when the
        # lanes of a quad share addresses; same probe), 8 warps per box, one wavefront per clock and SM
        wavefronts = 48 if complete else 72
        lsu_floor_ms = (num_boxes / shares) * groups * 8 * wavefronts / (148 * sm_mhz * 1e6) * 1e3
        from camus_b200 import partition as part
        scheme = "one share" if shares == 1 else (
            "class scheme: a rank owns boxes by the classes of their first two segments and builds only the "
            f"{'1/2' if shares == 8 else '1/2 or 3/4'} of the bit-planes it reads" if part.use_class_scheme(prob.n_taxa, shares)
            else "chunk scheme: chunks of boxes round-robin, every rank builds every bit-plane")
        line = {
            "metric": METRIC, "value": R / dt_res, "unit": UNIT, "n_gpus": shares, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_res * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": config_dict(args.config, c),
            "sharding": f"4-taxon-set index space / {shares} ({scheme}), all trees on every GPU, " +
                        (("one ncclAllReduce of the junction table inside the library" if lib_comm else
                          "one NCCL all-reduce of the junction table (torch.distributed)") if n_dev == 1 else
                         "one process and one handle for all GPUs, junction tables summed by NVLink peer reads"),
            "l2_flush": "512 MB write before every timed step (steps timed one by one)" if flush else "none needed (see config.l2)",
            "resolutions_per_step": R,
            "clocks": clocks,
            "e2e": {"value": R / dt_e2e, "unit": UNIT, "ms_per_step": dt_e2e * 1e3, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(out.nbytes + 16)},
            "gpu_launches": int(launches) * args.steps,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "roofline": {"bound": "xu", "binding_pipe": "XU (POPC) with the shared-memory load pipe close behind; not HBM, not tensor",
                         "kernel": "k_count<fused, %s> (count + filter + score in one launch, one CTA per 8x8x8x8 box)" % (
                             "complete" if complete else ("packed" if packed else "general")),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_kernel,
                         "launch_ms": kern_ms,
                         "xu_floor_ms": xu_floor_ms, "lsu_floor_ms": lsu_floor_ms,
                         "frac_binding": max(xu_floor_ms, lsu_floor_ms) / kern_ms,
                         "design_bytes_per_launch": design_bytes,
                         "design_achieved": design_bytes / (kern_ms * 1e-3) / 1e9,
                         "hbm_frac_measured": (traffic / (kern_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "note": "achieved / frac follow the survey's accounting (8 B per resolution), as the contract asks; the "
                                 "kernel keeps counts in registers and resolves 32 gene trees per 32-bit word, so it moves ~60x fewer "
                                 "bytes than that model and frac exceeds 1: it says nothing about efficiency. frac_binding = the larger of "
                                 "two floors over launch_ms: POPCs at 16 lanes per clock and SM (xu_floor_ms; rate verified with a "
                                 "synthetic loop under ncu, profiles/r02_pipe_probe.txt) and shared-memory wavefronts at one per clock "
                                 "and SM (lsu_floor_ms); design_* = "
                                 "bytes of triple tiles streamed (L2 + HBM); traffic = DRAM bytes per launch from ncu "
                                 "(profiles/traffic.json), hbm_frac_measured = that over the measured HBM peak"},
            "roofline_whole_path": {"algorithmic_bytes": b_alg, "achieved": b_alg / dt_res / 1e9, "frac": b_alg / dt_res / 1e9 / peak},
            "result": dict(digest_got, digests_checked=digest_want is not None),
        }
        extras = not args.no_extras and shares == 1
        if extras:
            # ---- the metric's second half: end-to-end infer seconds (Preprocess + Scorer.Init + RunDP,
            # infer.go:70-96) from the parsed trees, hot path on the GPU, DP on the host mirror ----
            t0 = time.perf_counter()
            gpu.clear_gene_trees()
            gpu.add_gene_trees(prob.gene_node_off, prob.gene_parent, prob.gene_taxon)
            ns, sc = gpu.count_filter(qmode, thr)
            tab = gpu.quartet_totals(as_set)
            pen = gpu.edge_penalties() if c["mode"] != "max" else None
            t1 = time.perf_counter()
            prob.run_dp(tab, pen, c["mode"], as_set, ns, sc, prob.n_trees, 0.5, gpu=local)  # untimed warm-up: loads the DP kernels
            t1b = time.perf_counter()
            res = prob.run_dp(tab, pen, c["mode"], as_set, ns, sc, prob.n_trees, 0.5, gpu=local)
            t2 = time.perf_counter()
            host = prob.run_dp(tab, pen, c["mode"], as_set, ns, sc, prob.n_trees, 0.5) if c["mode"] == "max" or prob.n_taxa <= 400 else None
            t3 = time.perf_counter()
            if host is not None:
                assert host.branches == res.branches and host.newicks == res.newicks and host.root_scores == res.root_scores
            line["infer"] = {"infer_s": (t1 - t0) + (t2 - t1b), "hot_path_s": t1 - t0, "dp_s": t2 - t1b, "dp_first_call_s": t1b - t1, "score_mode": c["mode"],
                             "networks": len(res.newicks), "dp_offload": res.accel,
                             "dp_host_only_s": (t3 - t2) if host is not None else None,
                             "what": "flat trees -> count/filter -> edge table (GPU, C-ABI) -> DP: post-order driver, cycleDP.update and "
                                     "traceback on the host (C++ mirror of infer/main_dp.go), the scoreEdgesAcross / FourWayBestSplit scan "
                                     "of every large step on the GPU (camus_b200_dp_across); dp_host_only_s = the same DP with the host "
                                     "scan (identical networks asserted; not run for float modes above 400 taxa, where the literal scan "
                                     "takes minutes); parsing excluded"}
            # ---- gene-tree ingest on the GPU: newick text -> flat trees, against the host parser ----
            if prob.gene_fmt == "newick":
                text = prob.gene_text.encode()
                gpu.set_tip_labels(prob.tip_names)
                ti = []
                for _ in range(3):
                    gpu.clear_gene_trees()
                    t4 = time.perf_counter()
                    gpu.add_gene_trees_newick(text, c["min_support"])
                    ti.append(time.perf_counter() - t4)
                off_g, par_g, tax_g = gpu.get_gene_trees()
                assert np.array_equal(off_g, prob.gene_node_off) and np.array_equal(par_g, prob.gene_parent) and np.array_equal(tax_g, prob.gene_taxon)
                line["ingest"] = {"newick_bytes": len(text), "gpu_s": min(ti), "host_parse_flatten_s": prob.host_parse_s,
                                  "what": "camus_b200_add_gene_trees_newick (text H2D + two kernels + flat trees back into the host staging) "
                                          "against the host mirror's threaded parser + flatten; identical flat arrays asserted"}
                ns, sc = gpu.count_filter(qmode, thr)  # the trees were re-added: bring the context back
            # the DP in the other score modes on the same tables
            other = {}
            pen_all = gpu.edge_penalties()
            for m in ("max", "norm", "sym"):
                if m == c["mode"]:
                    continue
                a_s = as_set or m == "sym"
                gpu.set_score_hint(a_s)
                ns_m, sc_m = gpu.count_filter(qmode, thr)
                tab_m = gpu.quartet_totals(a_s)
                t4 = time.perf_counter()
                r_m = prob.run_dp(tab_m, pen_all if m != "max" else None, m, a_s, ns_m, sc_m, prob.n_trees, 0.5, gpu=local)
                other[m] = {"dp_s": time.perf_counter() - t4, "networks": len(r_m.newicks), "dp_offload": r_m.accel}
            gpu.set_score_hint(as_set)
            line["infer"]["dp_other_modes"] = other
        gpu.close()
        if extras:
            # ---- the kernel real (incomplete / polytomous) inputs take, on the same workload ----
            if complete:
                os.environ["CAMUS_B200_NO_COMPLETE"] = "1"
                try:
                    gen = CamusB200(local)
                finally:
                    del os.environ["CAMUS_B200_NO_COMPLETE"]
                gen.load(prob)
                gen.set_score_hint(as_set)
                assert gen.count_filter(qmode, thr) == first
                tg = []
                for _ in range(2):
                    t0 = time.perf_counter()
                    gen.run_resident(qmode, thr, as_set, out)
                    tg.append(time.perf_counter() - t0)
                gms, _ = gen.stage_ms()
                assert digest_got["quartet_totals_as_set" if as_set else "quartet_totals"] == table_digest(out)
                gen.close()
                gfloor = num_boxes * groups * 4096 * 3 / (16 * 148 * sm_mhz * 1e6) * 1e3
                line["general_kernel"] = {"ms_per_step": min(tg) * 1e3, "value": R / min(tg), "count_ms": gms["count"], "prep_ms": gms["prep"],
                                          "xu_floor_ms": gfloor, "frac_binding": gfloor / gms["count"],
                                          "what": "same workload with CAMUS_B200_NO_COMPLETE=1: the 12-plane, 3-POPC kernel that gene "
                                                  "trees with missing taxa or polytomies take; results identical"}
            # ---- same-size CPU comparison where one exists: BASELINE.json configs[0] in full ----
            if not args.no_cpu:
                nthr = min(os.cpu_count() or 1, args.ref_threads)
                line["cpu_baseline"] = CpuSample(prob, c, nthr).record()
                if args.config != 1:
                    p1, c1 = make_problem(1, None, None)
                    rec1 = CpuSample(p1, c1, nthr).record()
                    g1 = CamusB200(local)
                    o1 = np.empty((p1.n_nodes, p1.n_nodes), np.uint64)
                    ts = []
                    for _ in range(5):
                        t0 = time.perf_counter()
                        g1.load(p1)
                        g1.set_score_hint(c1["as_set"])
                        g1.count_filter(c1["qmode"], c1["threshold"])
                        g1.quartet_totals(c1["as_set"], o1)
                        ts.append(time.perf_counter() - t0)
                    g1.close()
                    line["config1_same_size"] = {
                        "workload": workload_name(1, c1), "resolutions": rec1["sample_resolutions"],
                        "gpu_e2e_s": min(ts), "cpu_literal_s": rec1["seconds"], "cpu_fast_mode_s": rec1["fast_mode"]["seconds"], "cores": nthr,
                        "what": "the one configuration the reference's algorithm runs in full: whole hot path, host buffers in, table out"}
        elif not args.no_cpu:
            nthr = min(os.cpu_count() or 1, args.ref_threads)
            line["cpu_baseline"] = CpuSample(prob, c, nthr).record()
        print(json.dumps(line))
    else:
        gpu.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json configs[] index + 1 (3 = 1000 taxa x 1000 trees)")
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--trees", type=int, default=None)
    ap.add_argument("--ref-threads", type=int, default=64)
    ap.add_argument("--collective", default="library", choices=["library", "torch"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the infer / general-kernel / config-1 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
