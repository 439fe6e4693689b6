#!/bin/bash
set -u
tag=${1:-s17}
out=gpurun_out
timeout 900 python -m pytest tests/test_multirank_gpu.py -m gpu -q -x > $out/${tag}_pytest_multi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $out/${tag}_bench_ref_n2.json 2> $out/${tag}_bench_ref_n2.err
timeout 600 python bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > $out/${tag}_bench_onehandle_n2.json 2> $out/${tag}_bench_onehandle_n2.err
