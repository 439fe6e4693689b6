set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r8_pytest.log 2>&1
tail -4 gpurun_out/r8_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/r8_bench_n1.json 2> gpurun_out/r8_bench_n1.err; tail -c 300 gpurun_out/r8_bench_n1.json
timeout 300 python bench.py --config 2 --no-cpu > gpurun_out/r8_bench_c2.json 2> gpurun_out/r8_bench_c2.err; tail -c 300 gpurun_out/r8_bench_c2.json
