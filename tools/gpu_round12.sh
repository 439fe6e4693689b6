#!/bin/bash
set -u
tag=${1:-s12}
out=gpurun_out
timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
timeout 600 python tools/dp_times.py --config 5 > $out/${tag}_dp_c5.log 2>&1
timeout 900 python tools/dp_times.py --config 3 --host max > $out/${tag}_dp_c3.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "passes or packed or config1" > $out/${tag}_pytest_passes.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity_large.py -m gpu -q -x -k "beyond" > $out/${tag}_pytest_beyond.log 2>&1
