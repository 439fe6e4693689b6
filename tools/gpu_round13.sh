#!/bin/bash
# round-2 evidence: bench lines, ncu launch list + full captures, sanitizer logs, pipe experiments
set -u
tag=${1:-s13}
out=gpurun_out
timeout 600 python -m pytest tests/test_dp_offload.py -m gpu -q > $out/${tag}_pytest_dp.log 2>&1
timeout 900 python tools/dp_times.py --config 3 --host max > $out/${tag}_dp_c3.log 2>&1
run() { local name=$1; shift; env "$@" timeout 300 python tools/stage_times.py --config $CFG --reps 3 2>&1 | grep -v "^N=" > $out/${tag}_c${CFG}_${name}.log; }
CFG=3; run bare CAMUS_B200_DEBUG=2 CAMUS_B200_LIB=build_variants/libcamus_bare.so
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
for c in 1 2 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c$c.json 2> $out/${tag}_bench_c$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-extras > $out/${tag}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_k_count_c3 python tools/stage_times.py --config 3 --reps 1 > $out/${tag}_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_k_count_c2 python tools/stage_times.py --config 2 --reps 1 > $out/${tag}_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_k_count_c4 python tools/stage_times.py --config 4 --reps 1 > $out/${tag}_ncu_c4.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_racecheck.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "passes" > $out/${tag}_memcheck_passes.log 2>&1
