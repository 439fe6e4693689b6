#!/bin/bash
set -u
tag=${1:-s6}
out=gpurun_out
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
CFG=3; run default A=1; run nobar CAMUS_B200_LIB=build_variants/libcamus_nobar.so
run gen_packed CAMUS_B200_NO_COMPLETE=1; run gen_packed_nobar CAMUS_B200_NO_COMPLETE=1 CAMUS_B200_LIB=build_variants/libcamus_nobar.so
CFG=4; run packed A=1; run packed_nobar CAMUS_B200_LIB=build_variants/libcamus_nobar.so; run wide CAMUS_B200_NO_PACKED=1
CFG=2; run default A=1; run nobar CAMUS_B200_LIB=build_variants/libcamus_nobar.so
CFG=5; run default A=1; run nobar CAMUS_B200_LIB=build_variants/libcamus_nobar.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1
CAMUS_B200_LIB=build_variants/libcamus_nobar.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or small_shapes or regular_box or complete_tree or randomized or mid_size or process_quartets or packed" > $out/${tag}_pytest_nobar.log 2>&1
