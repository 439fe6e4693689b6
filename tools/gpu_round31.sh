#!/bin/bash
set -u
tag=${1:-s31}
out=gpurun_out
timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
