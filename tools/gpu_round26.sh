#!/bin/bash
set -u
tag=${1:-s26}
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dp_across_tiled --launch-skip 560 -c 3 -f -o $out/${tag}_dp_across python tools/dp_times.py --config 3 --modes norm --host none > $out/${tag}_ncu_dp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dp --csv --log-file $out/${tag}_dp_launches.csv python tools/dp_times.py --config 3 --modes norm --host none > $out/${tag}_ncu_dp2.log 2>&1
