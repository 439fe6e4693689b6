set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/share_times.py --world 4 --skip-whole > gpurun_out/r5_share4_class.log 2>&1; tail -5 gpurun_out/r5_share4_class.log
for m in 3 5 6; do CAMUS_B200_LIB=$GRAFT_REPO_ROOT/build_variants/libcamus_minb$m.so timeout 200 python tools/stage_times.py --config 3 --reps 3 > gpurun_out/r5_minb$m.log 2>&1; tail -1 gpurun_out/r5_minb$m.log; done
timeout 200 python tools/stage_times.py --config 3 --reps 3 > gpurun_out/r5_minb4.log 2>&1; tail -1 gpurun_out/r5_minb4.log
