set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "four_gpu or multi_gpu_handle" ) > gpurun_out/r4_pytest_multi.log 2>&1
tail -4 gpurun_out/r4_pytest_multi.log
timeout 200 python bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r4_bench_onehandle_n4.json 2> gpurun_out/r4_bench_onehandle_n4.err; tail -c 400 gpurun_out/r4_bench_onehandle_n4.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r4_bench_n2.json 2> gpurun_out/r4_bench_n2.err; tail -c 300 gpurun_out/r4_bench_n2.json
