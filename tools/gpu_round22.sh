#!/bin/bash
set -u
out=gpurun_out; tag=${1:-s22}
CAMUS_B200_PASSES=8 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not partition and not multi and not export and not query and not table_pipelines and not unfused" > $out/${tag}_pytest_passes_env.log 2>&1
CAMUS_B200_NO_BOXDESC=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or small_shapes or regular_box or randomized or mid_size or packed" > $out/${tag}_pytest_nodesc.log 2>&1
CAMUS_DP_ACCEL_MIN_WORK=0 timeout 900 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp_mw0.log 2>&1
