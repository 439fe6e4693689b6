#!/bin/bash
set -u
tag=${1:-s18}
out=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_all.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
