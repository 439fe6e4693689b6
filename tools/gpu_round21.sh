#!/bin/bash
out=gpurun_out; tag=${1:-s21}
for c in 2 5; do
  timeout 300 python tools/host_overhead.py $c > $out/${tag}_ovh_c$c.log 2>&1
  CAMUS_B200_PASSES=1 timeout 300 python tools/host_overhead.py $c > $out/${tag}_ovh_c${c}_p1.log 2>&1
done
