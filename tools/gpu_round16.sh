#!/bin/bash
set -u
tag=${1:-s16}
out=gpurun_out
for mw in 30000 8000 2000; do
  CAMUS_DP_PROFILE=1 CAMUS_DP_ACCEL_MIN_WORK=$mw timeout 600 python tools/dp_times.py --config 3 --host none > $out/${tag}_dp_c3_mw$mw.log 2>&1
done
