#!/bin/bash
set -u
tag=${1:-s33}
out=gpurun_out
timeout 900 python -m pytest tests/test_ingest_gpu.py tests/test_dp_offload.py -m gpu -q -s > $out/${tag}_pytest.log 2>&1
