#!/bin/bash
set -u
tag=${1:-s2}
out=gpurun_out
st() { # name, env..., -- args
  local name=$1; shift
  env "$@" CAMUS_B200_VERBOSE=1 timeout 300 python tools/stage_times.py --config $CFG --reps $REPS > $out/${tag}_c${CFG}_${name}.log 2>&1
}
CFG=3 REPS=2
st dyn A=1
st static CAMUS_B200_STREAM_STATIC=1
st ctas2 CAMUS_B200_STREAM_CTAS=2
st relaxed CAMUS_B200_LIB=build_variants/libcamus_relaxed.so
st v1 CAMUS_B200_COUNT_V1=1
CFG=2 REPS=5
st dyn A=1
st relaxed CAMUS_B200_LIB=build_variants/libcamus_relaxed.so
st v1 CAMUS_B200_COUNT_V1=1
CFG=4 REPS=3
st dyn A=1
st relaxed CAMUS_B200_LIB=build_variants/libcamus_relaxed.so
st v1 CAMUS_B200_COUNT_V1=1
timeout 300 python tools/make_config_digests.py --gpu > $out/${tag}_digests.log 2>&1
cp tests/golden/synthetic/config_digests.json $out/${tag}_config_digests.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_count_stream -c 1 -f -o $out/${tag}_stream_c3 \
  python tools/stage_times.py --config 3 --reps 1 > $out/${tag}_ncu_c3.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_racecheck.log 2>&1
