#!/bin/bash
for o in 1 0 1 0; do
echo "== RB2_UNIT_ORDER=$o"
RB2_UNIT_ORDER=$o python tools/rank_emulation.py 100000 8 2>&1 | grep "waves=64"
RB2_UNIT_ORDER=$o python bench.py --steps 2 --warmup 2 --no-cpu --no-sweep 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1e6', l['ms_per_step'], 'kernel', l['roofline']['kernel_ms'], l['clocks'])"
done
