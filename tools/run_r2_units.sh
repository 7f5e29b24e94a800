#!/bin/bash
# unit-list launch grid + component-split reduce: parity, N = 1e4 / 1e5 / 1e6 timing, 8-rank emulation at 1e5
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidevice.py -m gpu -q -x 2>&1 | tail -3
python tools/rank_emulation.py 100000 8 2>&1 | tail -5
python bench.py --particles 10000 --steps 30 --no-cpu --no-sweep 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1e4', l['ms_per_step'], 'median', sorted(l['ms_steps_rank0'])[15], 'kernel', l['roofline']['kernel_ms'], l['roofline']['launch'])"
python bench.py --particles 100000 --steps 10 --no-cpu --no-sweep 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1e5', l['ms_per_step'], 'kernel', l['roofline']['kernel_ms'], l['roofline']['launch'])"
python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1e6', l['ms_per_step'], 'kernel', l['roofline']['kernel_ms'], l['roofline']['launch'], l['acc_checksum']['sha256_16'])"
RB2_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_1e4_nograph.csv python tools/prof_step.py --n 10000 --steps 3 > /dev/null 2>&1; python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_1e4_nograph.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-10:]: print(r[4][:50], r[7], r[8], r[-1])
PY
