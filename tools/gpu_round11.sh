#!/bin/bash
set -u
tag=${1:-s11}
out=gpurun_out
timeout 900 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
timeout 600 python tools/dp_times.py --config 5 > $out/${tag}_dp_c5.log 2>&1
timeout 900 python tools/dp_times.py --config 3 --host max > $out/${tag}_dp_c3.log 2>&1
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
CFG=3; run nopopc CAMUS_B200_DEBUG=2 CAMUS_B200_LIB=build_variants/libcamus_nopopc.so; run nolds CAMUS_B200_DEBUG=2 CAMUS_B200_LIB=build_variants/libcamus_nolds.so
