#!/bin/bash
set -u
tag=${1:-s32}
out=gpurun_out
for z in 128 64 32 16; do CAMUS_B200_ZREP=$z timeout 600 python tools/share_times.py --config 3 --world 8 --reps 2 --skip-whole 2>&1 | head -3 > $out/${tag}_share8_z$z.log; done
