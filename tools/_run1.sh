set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest.log 2>&1
tail -5 gpurun_out/r1_pytest.log
timeout 300 python tools/share_times.py --world 8 > gpurun_out/r1_share8_class.log 2>&1
CAMUS_B200_PARTITION=chunk timeout 300 python tools/share_times.py --world 8 --skip-whole > gpurun_out/r1_share8_chunk.log 2>&1
timeout 300 python tools/share_times.py --world 4 --skip-whole > gpurun_out/r1_share4_class.log 2>&1
for b in 4 8 16 32; do CAMUS_B200_PREP_BATCH=$b timeout 200 python tools/stage_times.py --config 3 --reps 3 > gpurun_out/r1_batch$b.log 2>&1; done
CAMUS_B200_SERIAL_PREPARE=1 timeout 200 python tools/stage_times.py --config 3 --reps 3 > gpurun_out/r1_serialprep.log 2>&1
tail -3 gpurun_out/r1_share8_class.log gpurun_out/r1_batch*.log gpurun_out/r1_serialprep.log
