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
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

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
    prob = Problem(s.constraint, s.genes, s.fmt, c["min_support"])
    return prob, c


def workload_name(cfg_id, c):
    return (f"config{cfg_id}: {c['n_taxa']}-taxon constraint + {c['n_trees']} gene trees, -q {c['qmode']} -t {c['threshold']}"
            f"{' -asSet' if c['as_set'] else ''} {c['mode']} mode")


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
        p, t = prob.gene_parent[lo:hi], prob.gene_taxon[lo:hi]
        kb = np.zeros(hi - lo, np.int64)
        for i in range(hi - lo - 1, -1, -1):
            if t[i] >= 0:
                kb[i] += int(keep[t[i]])
            if p[i] >= 0:
                kb[p[i]] += kb[i]
        idx = np.full(hi - lo, -1, np.int64)
        k = 0
        for i in range(hi - lo):
            if kb[i] > 0:
                idx[i] = k
                k += 1
                gp.append(idx[p[i]] if p[i] >= 0 else -1)
                gt.append(new_tip[t[i]] if t[i] >= 0 else -1)
        off.append(len(gp))
    return (par, c0, c1, tip, int(keep.sum())), (np.array(off, np.int64), np.array(gp, np.int32), np.array(gt, np.int32))


def cpu_reference_sample(prob, c, n_sample_trees: int, nthreads: int, max_taxa: int = 100):
    """The reference's own algorithm (oracle literal mode: bipartition enumerator + per-tree set +
    map merge + filter + constraint removal) on a bounded sample: the first n_sample_trees gene
    trees, induced on every other taxon when the problem has more than max_taxa taxa (one
    200-taxon tree alone costs the hash-map algorithm over a minute per core)."""
    from math import comb

    from oracle import oracle as orc
    step = max(1, -(-prob.n_taxa // max_taxa))
    keep = range(0, prob.n_taxa, step)
    cons, genes = restrict_to_taxa(prob, keep, n_sample_trees)
    o = orc.Oracle(nthreads)
    o.set_constraint(*cons[:4], cons[4])
    o.add_gene_trees(*genes)
    k = len(genes[0]) - 1
    res = 0
    for g in range(k):
        nl = int((genes[2][genes[0][g]:genes[0][g + 1]] >= 0).sum())
        res += comb(nl, 4)
    t0 = time.perf_counter()
    o.count_filter(c["qmode"], c["threshold"], True)
    dt = time.perf_counter() - t0
    desc = (f"first {k} of {prob.n_trees} gene trees induced on {cons[4]} of {prob.n_taxa} taxa (every {step}th tip), "
            f"{nthreads} threads")
    return res, dt, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob, c = make_problem(args.config, args.taxa, args.trees)
    cores = os.cpu_count() or 1
    nthreads = min(cores, args.ref_threads)
    per_step = max(1, args.ref_trees)
    times, res = [], 0
    for i in range(args.warmup + args.steps):
        res, dt, desc = cpu_reference_sample(prob, c, per_step, nthreads)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    val = res / dt
    sample = ("literal processQuartets+filterQuartets+constraint removal (oracle port of the Go algorithm) on the " + desc +
              "; mapQuartetsToVertices and the per-(u,w) QuartetScore loop excluded (~1e12 calls at this size)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": {"workload": workload_name(args.config, c)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
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
        if world > 1:
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
        _, ns, sc = gpu.run_resident(qmode, thr, as_set, None, partial_only=True)
        return reduce_and_finish(ns, sc)

    # first call uploads the trees and builds everything once (also a correctness anchor)
    first, _, _ = step_e2e()
    R = gpu.workload()["resolutions"]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
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

    if rank == 0:
        peak, peak_src = load_peaks()
        wl = gpu.workload()
        cells, num_boxes = wl["cells"], wl["boxes"]
        m_seg = -(-prob.n_taxa // 8)
        num_tiles = (m_seg + 2) * (m_seg + 1) * m_seg // 6
        # ---- roofline of the dominant kernel: the fused k_count<true> (count + filter + score) ----
        # Algorithmic bytes per launch, SURVEY.md 8(d) convention: 8 B per resolution (one count
        # cell read + written per (gene tree, 4-set)) + 16 B per cell (one 128-bit cell read by
        # filter / score), for the share of the 4-set space this rank owns.
        alg_bytes_kernel = (8 * R + 16 * cells) / shares
        kern_ms = stage_ms["count"]
        n_launch = max(1, -(-num_boxes // shares // (1 << 30)))
        achieved = alg_bytes_kernel / (kern_ms * 1e-3) / 1e9
        b_alg = 8 * R + 16 * cells + 24 * n * n
        # Bytes the kernel's own design has to move per launch: four 6 KB triple tiles per
        # (box, tree group), read once each (HBM or L2), and one pass over the junction tables.
        groups = -(-prob.n_trees // 32)
        design_bytes = (num_boxes / shares) * groups * 4 * 6144
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(f"config{args.config}_k_count_dram_bytes_per_launch")
        if shares > 1:
            traffic = None  # the captures are single-GPU launches
        h2d = int(prob.gene_node_off.nbytes + prob.gene_parent.nbytes + prob.gene_taxon.nbytes)
        # complete binary gene trees take the 8-plane kernel variant (2/3 of the tile bytes)
        nl = np.diff(np.concatenate([[0], np.cumsum(prob.gene_taxon >= 0)[prob.gene_node_off[1:] - 1]]))
        complete = bool((nl == prob.n_taxa).all()) and not c.get("support", False)
        if complete:
            design_bytes *= 2.0 / 3.0
        from camus_b200 import partition as part
        scheme = "one share" if shares == 1 else (
            "class scheme: a rank owns boxes by the classes of their first two segments and builds only the "
            f"{'1/2' if shares == 8 else '1/2 or 3/4'} of the bit-planes it reads" if part.use_class_scheme(prob.n_taxa, shares)
            else "chunk scheme: chunks of boxes round-robin, every rank builds every bit-plane")
        line = {
            "metric": METRIC, "value": R / dt_res, "unit": UNIT, "n_gpus": shares, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_res * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, c), "resolutions_per_step": R, "cells": cells, "boxes": num_boxes,
                       "sharding": f"4-taxon-set index space / {shares} ({scheme}), all trees on every GPU, " +
                                   ("one NCCL all-reduce of the junction table" if n_dev == 1 else
                                    "one process and one handle for all GPUs, junction tables summed by NVLink peer reads"),
                       "l2": "no explicit flush: the rooted-triple bit-planes streamed by every step "
                             f"({design_bytes * shares / 1e9:.1f} GB of tile reads over a {groups * 6144 * num_tiles / 1e9:.2f} GB table) exceed the 126 MB L2 many times over"},
            "clocks": clocks,
            "e2e": {"value": R / dt_e2e, "unit": UNIT, "ms_per_step": dt_e2e * 1e3, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(out.nbytes + 16)},
            "gpu_launches": int(launches) * args.steps,
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "roofline": {"bound": "hbm", "kernel": "k_count<fused, %s> (count + filter + score in one launch)" % ("complete" if complete else "general"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_kernel / n_launch,
                         "launch_ms": kern_ms / n_launch,
                         "design_bytes_per_launch": design_bytes / n_launch,
                         "design_achieved": design_bytes / (kern_ms * 1e-3) / 1e9,
                         # the pipe that actually bounds the kernel: one POPC per (cell, counted topology, 32 trees)
                         # on the XU pipe, 16 lanes per clock and SM (B300_MICROARCH.md), 148 SMs
                         "xu_floor_ms": (num_boxes / shares) * groups * 4096 * (2 if complete else 3)
                                        / (16 * 148 * (clocks.get("sm_mhz") or 1965) * 1e6) * 1e3,
                         "note": "achieved uses the survey's accounting (8 B per resolution); the kernel resolves 32 gene trees per "
                                 "32-bit word in registers, so it moves ~60x fewer bytes than that model and frac exceeds 1; "
                                 "design_* = bytes of triple tiles the kernel actually streams (L2 + HBM); traffic = its DRAM bytes "
                                 "(ncu); the binding resource is the POPC (XU) pipe: xu_floor_ms / launch_ms is the fraction of "
                                 "that floor reached, see DESIGN.md"},
            "roofline_whole_path": {"algorithmic_bytes": b_alg, "achieved": b_alg / dt_res / 1e9, "frac": b_alg / dt_res / 1e9 / peak},
            "result": {"n_survivors": first[0], "sum_counts": first[1]},
        }
        if not args.no_cpu:
            nthr = min(os.cpu_count() or 1, args.ref_threads)
            res, dtc, desc = cpu_reference_sample(prob, c, args.ref_trees, nthr)
            line["cpu_baseline"] = {
                "value": res / dtc, "unit": UNIT, "cores": nthr, "kind": "port",
                "sample": "literal count+filter+removal (oracle port of the Go algorithm) on the " + desc + "; edge scoring excluded",
            }
        print(json.dumps(line))
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
    ap.add_argument("--ref-trees", type=int, default=8, help="gene trees per CPU-reference sample")
    ap.add_argument("--ref-threads", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
