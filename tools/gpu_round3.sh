#!/bin/bash
set -u
tag=${1:-s3}
out=gpurun_out
for v in stream v1; do
  if [ $v = v1 ]; then export CAMUS_B200_COUNT_V1=1; else unset CAMUS_B200_COUNT_V1; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_${v}_c2 \
    python tools/stage_times.py --config 2 --reps 1 > $out/${tag}_ncu_${v}_c2.log 2>&1
done
unset CAMUS_B200_COUNT_V1
