#!/bin/bash
# A/B of the cheaper far-partner inverse cube in the pair-symmetric kernel (option "sym_far"), parity then timing
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python tools/far_probe.py 20000
for f in 1 0; do
  echo "== RB2_SYM_FAR=$f"
  RB2_SYM_FAR=$f python bench.py --no-sweep --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'], l['roofline']['frac'], l['acc_checksum'])"
  RB2_SYM_FAR=$f python bench.py --no-sweep --particles 100000 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'], l.get('parity'))"
done
