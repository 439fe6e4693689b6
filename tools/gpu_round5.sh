#!/bin/bash
set -u
tag=${1:-s5}
out=gpurun_out
for cfg in 3 2; do
  CAMUS_B200_LIB=build_variants/libcamus_occ4.so timeout 300 python tools/stage_times.py --config $cfg --reps 3 > $out/${tag}_c${cfg}_occ4.log 2>&1
  timeout 300 python tools/stage_times.py --config $cfg --reps 3 > $out/${tag}_c${cfg}_box.log 2>&1
done
CAMUS_B200_LIB=build_variants/libcamus_occ4.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or small_shapes or regular_box or complete_tree or randomized or mid_size or process_quartets" > $out/${tag}_pytest_occ4.log 2>&1
CAMUS_B200_LIB=build_variants/libcamus_occ4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_occ4_c2 \
    python tools/stage_times.py --config 2 --reps 1 > $out/${tag}_ncu_occ4_c2.log 2>&1
