#!/bin/bash
set -u
tag=${1:-s7}
out=gpurun_out
P=build_variants/lds_probe
timeout 120 $P lds > $out/${tag}_lds.log 2>&1
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $out/${tag}_lds_ncu.csv $P lds > /dev/null 2>&1
timeout 120 $P popc > $out/${tag}_popc.log 2>&1
for d in 0 1 2 3; do
  CAMUS_B200_DEBUG=$d timeout 300 python tools/stage_times.py --config 3 --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c3_debug$d.log
done
timeout 1500 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1
