#!/bin/bash
set -u
tag=${1:-s25}
out=gpurun_out
timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
CAMUS_DP_PROFILE=1 timeout 900 python tools/dp_times.py --config 3 --host none > $out/${tag}_dp_c3.log 2>&1
CAMUS_DP_PROFILE=1 timeout 900 python tools/dp_times.py --config 3 --host none > $out/${tag}_dp_c3b.log 2>&1
timeout 600 python tools/dp_times.py --config 5 --host none > $out/${tag}_dp_c5.log 2>&1
