#!/bin/bash
set -u
tag=${1:-s19}
out=gpurun_out
N=${2:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
timeout 600 python bench.py --gpus $N --steps 5 --warmup 3 --no-extras > $out/${tag}_bench_onehandle_n$N.json 2> $out/${tag}_bench_onehandle_n$N.err
