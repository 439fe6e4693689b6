set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ret -c 2 -f -o gpurun_out/r6_retic python tools/retic_times.py --config 1 --taxa 120 --trees 400 --cpu-trees 0 --reps 1 > gpurun_out/r6_retic_ncu.log 2>&1
tail -2 gpurun_out/r6_retic_ncu.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tree_prepare_warp -c 1 -f -o gpurun_out/r6_prepare python tools/stage_times.py --config 3 --reps 1 > gpurun_out/r6_prepare_ncu.log 2>&1
tail -2 gpurun_out/r6_prepare_ncu.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
