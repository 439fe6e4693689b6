#!/bin/bash
# Developer tool: one gpurun call that regenerates the evidence under profiles/ (copy what you want judged from
# gpurun_out/ afterwards): tests, bench lines, ncu launch list + full captures, pipe probe.
#   gpurun --timeout 3600 -- 'bash tools/gpu_evidence.sh r03'
set -u
tag=${1:-ev}
out=gpurun_out
mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest_all.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
for c in 1 2 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $out/${tag}_bench_c$c.json 2> $out/${tag}_bench_c$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --steps 1 --warmup 1 --no-extras > $out/${tag}_ncu_launches.log 2>&1
for c in 3 2 4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_count -c 1 -f -o $out/${tag}_k_count_c$c python tools/stage_times.py --config $c --reps 1 > $out/${tag}_ncu_c$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_triple_planes_bs|k_depth_slices_direct" -s 2 -c 2 -f -o $out/${tag}_prep python tools/stage_times.py --config 3 --reps 1 > $out/${tag}_ncu_prep.log 2>&1
if [ -x build_variants/lds_probe ]; then  # nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build_variants/lds_probe tools/lds_probe.cu
  timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none --csv --log-file $out/${tag}_lds_ncu.csv build_variants/lds_probe lds > /dev/null 2>&1
  timeout 300 ncu --metrics sm__inst_executed_pipe_xu.sum,sm__cycles_elapsed.max --clock-control none --csv --log-file $out/${tag}_popc_ncu.csv build_variants/lds_probe popc > /dev/null 2>&1
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
