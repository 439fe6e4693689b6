#!/bin/bash
set -u
tag=${1:-s4}
out=gpurun_out
for cfg in 3 2 4; do
  timeout 300 python tools/stage_times.py --config $cfg --reps 3 > $out/${tag}_c${cfg}_box.log 2>&1
done
timeout 1700 python -m pytest tests -m gpu -q -x --durations=8 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
