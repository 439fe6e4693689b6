#!/bin/bash
set -u
tag=${1:-s24}
out=gpurun_out
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
for CFG in 3 2 4 5; do run new A=1; run prev CAMUS_B200_LIB=build_variants/libcamus_prev.so; done
CFG=3; run new2 A=1; run prev2 CAMUS_B200_LIB=build_variants/libcamus_prev.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1
