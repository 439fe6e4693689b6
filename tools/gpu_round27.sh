#!/bin/bash
set -u
tag=${1:-s27}
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dp_across_tiled --launch-skip 810 -c 1 -f -o $out/${tag}_dp_across_root python tools/dp_times.py --config 3 --modes norm --host none > $out/${tag}_ncu_dp.log 2>&1
