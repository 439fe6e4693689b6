#!/bin/bash
set -u
tag=${1:-s35}
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multirank_gpu.py -m gpu -q -x -k "multi or library_owned" > $out/${tag}_pytest_multi.log 2>&1
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > $out/${tag}_bench_onehandle_n2.json 2> $out/${tag}_bench_onehandle_n2.err
