#!/bin/bash
# One gpurun call of the round-2 loop: GPU tests, A/B stage times, sanitizer. Logs under gpurun_out/.
set -u
tag=${1:-s1}
out=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_smi.txt 2>&1
nproc > $out/${tag}_host.txt; free -g >> $out/${tag}_host.txt
python - > $out/${tag}_occ.txt 2>&1 <<'PY'
import ctypes, sys
sys.path.insert(0, '.')
from camus_b200.cabi import CamusB200
g = CamusB200(0); print("create ok"); g.close()
PY
timeout 1700 python -m pytest tests -m gpu -q -x --durations=15 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
for cfg in 3 4 2; do
  timeout 300 python tools/stage_times.py --config $cfg --reps 3 > $out/${tag}_stage_c${cfg}_stream.log 2>&1
  CAMUS_B200_COUNT_V1=1 timeout 300 python tools/stage_times.py --config $cfg --reps 3 > $out/${tag}_stage_c${cfg}_v1.log 2>&1
done
