set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_reticulation_score.py -m gpu -x -q ) > gpurun_out/r7_pytest_retic.log 2>&1
tail -5 gpurun_out/r7_pytest_retic.log
timeout 600 python tools/retic_times.py --config 2 --cpu-trees 1 > gpurun_out/r7_retic_config2.log 2>&1; tail -1 gpurun_out/r7_retic_config2.log
timeout 600 python tools/retic_times.py --config 1 --taxa 120 --trees 400 --cpu-trees 2 > gpurun_out/r7_retic_120x400.log 2>&1; tail -1 gpurun_out/r7_retic_120x400.log
