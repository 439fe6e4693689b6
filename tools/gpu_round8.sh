#!/bin/bash
set -u
tag=${1:-s8}
out=gpurun_out
P=build_variants/lds_probe
timeout 120 $P lds > $out/${tag}_lds.log 2>&1
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $out/${tag}_lds_ncu.csv $P lds > /dev/null 2>&1
timeout 120 $P popc > $out/${tag}_popc.log 2>&1
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
CFG=3; run default A=1; run nodesc CAMUS_B200_NO_BOXDESC=1; run debug1 CAMUS_B200_DEBUG=1; run debug2 CAMUS_B200_DEBUG=2; run debug3 CAMUS_B200_DEBUG=3
CFG=2; run default A=1; run nodesc CAMUS_B200_NO_BOXDESC=1; run debug2 CAMUS_B200_DEBUG=2
CFG=4; run default A=1; run debug2 CAMUS_B200_DEBUG=2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1
