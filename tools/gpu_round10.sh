#!/bin/bash
set -u
tag=${1:-s10}
out=gpurun_out
timeout 900 python -m pytest tests/test_dp_offload.py -m gpu -q -x > $out/${tag}_pytest_dp.log 2>&1
timeout 600 python tools/dp_times.py --config 5 > $out/${tag}_dp_c5.log 2>&1
timeout 900 python tools/dp_times.py --config 3 --host max > $out/${tag}_dp_c3.log 2>&1
