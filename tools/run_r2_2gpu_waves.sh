#!/bin/bash
mkdir -p gpurun_out
for w in 16 32 64 128; do
  RB2_SYM_WAVES=$w python bench.py --gpus 1 --particles 100000 --steps 10 --no-cpu --no-sweep > gpurun_out/w1_$w.json 2>> gpurun_out/r2_waves.err
  RB2_SYM_WAVES=$w python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --particles 100000 --steps 10 --no-cpu --no-sweep > gpurun_out/w2_$w.json 2>> gpurun_out/r2_waves.err
  python - $w <<'PY'
import json,sys
w=sys.argv[1]
a=json.loads(open(f'gpurun_out/w1_{w}.json').read().strip().splitlines()[-1]); b=json.loads(open(f'gpurun_out/w2_{w}.json').read().strip().splitlines()[-1])
print('waves',w,'1e5: 1 GPU', round(a['ms_per_step'],3), ' 2 GPUs', round(b['ms_per_step'],3), ' eff', round(a['ms_per_step']/b['ms_per_step']/2,4), 'split', a['roofline']['launch']['j_chunk'], b['roofline']['launch']['j_chunk'])
PY
done | tee gpurun_out/r2_waves.log
