#!/bin/bash
set -u
tag=${1:-s29}
out=gpurun_out
for i in 1 2 3; do timeout 900 python tools/dp_times.py --config 3 --host none > $out/${tag}_dp_c3_$i.log 2>&1; done
