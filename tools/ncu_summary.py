#!/usr/bin/env python
"""Developer tool: condense an .ncu-rep (ncu --set full) into the few metrics DESIGN.md quotes.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALL = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print("kernel:", row[hdr.index("Kernel Name")])
    d = {h: (v, u) for h, v, u in zip(hdr, row, units)}
    for w in WANT:
        if w in d:
            print(f"  {w:84s} {d[w][0]:>18s} {d[w][1]}")
    print("  warp stall reasons (cycles per issued instruction):")
    st = sorted(((float(v[0].replace(',', '')), h) for h, v in d.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)
    for val, h in st[:8]:
        print(f"    {h[len(STALL):-len('_per_issue_active.ratio')]:28s} {val:6.2f}")
