mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "plane_builders or sliced or small_shapes or dataset or class_partitions or complete_tree or config4_like" ) > gpurun_out/s7_pytest_new.log 2>&1
tail -4 gpurun_out/s7_pytest_new.log
for v in default nodirect; do
  for c in 3 4 2; do
    case $v in
      default) env= ;;
      nodirect) env="CAMUS_B200_NO_DIRECT=1" ;;
    esac
    env $env timeout 300 python tools/stage_times.py --config $c --reps 3 > gpurun_out/s7_c${c}_$v.log 2>&1
    echo "== $v c$c"; tail -1 gpurun_out/s7_c${c}_$v.log | cut -c1-200
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s7_launches.csv python tools/stage_times.py --config 3 --reps 1 > gpurun_out/s7_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_triple_planes_bs|k_depth_slices_direct" -s 2 -c 2 -f -o gpurun_out/s7_prep python tools/stage_times.py --config 3 --reps 1 > gpurun_out/s7_ncu_prep.log 2>&1
