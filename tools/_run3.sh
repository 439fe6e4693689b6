set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r3_pytest.log 2>&1
tail -4 gpurun_out/r3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; tail -1 gpurun_out/r3_smoke.log
timeout 600 python tools/retic_times.py --config 2 --cpu-trees 0 > gpurun_out/r3_retic_config2.log 2>&1; tail -1 gpurun_out/r3_retic_config2.log
timeout 600 python bench.py > gpurun_out/r3_bench_n1.json 2> gpurun_out/r3_bench_n1.err; tail -c 600 gpurun_out/r3_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches_config3.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r3_ncu.log 2>&1
tail -2 gpurun_out/r3_ncu.log | cut -c1-300
