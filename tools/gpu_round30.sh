#!/bin/bash
set -u
tag=${1:-s30}
out=gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
for c in 1 2 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c$c.json 2> $out/${tag}_bench_c$c.err; done
timeout 900 python -m pytest tests/test_ingest_gpu.py tests/test_reticulation_score.py tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest.log 2>&1
