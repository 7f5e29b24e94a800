#!/bin/bash
# Eight GPUs: the contract line (one process per GPU, torchrun) with its N = 1e4 / 1e5 sweep, and one process driving all eight.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --no-cpu > gpurun_out/r2_8gpu_torchrun_1e6.json 2> gpurun_out/r2_8gpu.err
python bench.py --gpus 8 --single-process --steps 5 --no-cpu > gpurun_out/r2_8gpu_single_process_1e6.json 2>> gpurun_out/r2_8gpu.err
python - <<'PY'
import json
for f in ('r2_8gpu_torchrun_1e6','r2_8gpu_single_process_1e6'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'n_gpus',d['n_gpus'],'processes',d.get('processes'),'ms/step',round(d['ms_per_step'],2),'value',d['value'],'e2e ms',round(d['e2e']['ms_per_step'],2),'sum_abs',d['acc_checksum']['sum_abs'],d['acc_checksum']['replicas_identical'])
        for s in d.get('sweep') or []: print('   sweep', s['n_particles'], round(s['ms_per_step_median'],4), round(s['kernel_ms_median'],4), round(s['frac'],3))
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/r2_8gpu.err
