#!/bin/bash
set -u
tag=${1:-s14}
out=gpurun_out
timeout 900 python -m pytest tests/test_ingest_gpu.py -m gpu -q -s > $out/${tag}_pytest_ingest.log 2>&1
timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
CAMUS_DP_NO_TILED=1 timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp_notiled.log 2>&1
timeout 900 python tools/dp_times.py --config 3 --host max > $out/${tag}_dp_c3.log 2>&1
timeout 600 python tools/dp_times.py --config 5 > $out/${tag}_dp_c5.log 2>&1
timeout 300 ncu --metrics sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $out/${tag}_popc_ncu.csv build_variants/lds_probe popc > $out/${tag}_popc_under_ncu.log 2>&1
