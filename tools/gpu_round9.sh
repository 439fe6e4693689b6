#!/bin/bash
set -u
tag=${1:-s9}
out=gpurun_out
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
CFG=3; run default A=1; run debug2 CAMUS_B200_DEBUG=2
CFG=2; run default A=1
CFG=4; run default A=1
CFG=5; run default A=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1
