#!/bin/bash
set -u
tag=${1:-s20}
out=gpurun_out
for c in 5 2 1 4; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-extras > $out/${tag}_bench_c$c.json 2> $out/${tag}_bench_c$c.err; done
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
