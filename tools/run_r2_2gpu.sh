#!/bin/bash
# Two GPUs: single-process multi-device mode against the one-process-per-GPU mode.
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_p2p.py -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -40 | tee gpurun_out/r2_2gpu_tests.log
python bench.py --gpus 2 --single-process --steps 3 --no-cpu > gpurun_out/r2_2gpu_single_process_1e6.json 2> gpurun_out/r2_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --no-cpu > gpurun_out/r2_2gpu_torchrun_1e6.json 2>> gpurun_out/r2_2gpu.err
python - <<'PY'
import json
for f in ('r2_2gpu_single_process_1e6','r2_2gpu_torchrun_1e6'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'n_gpus',d['n_gpus'],'processes',d.get('processes'),'ms/step',round(d['ms_per_step'],2),'value',d['value'],'e2e ms',round(d['e2e']['ms_per_step'],2),'sum_abs',d['acc_checksum']['sum_abs'],d['acc_checksum']['replicas_identical'])
        for s in d.get('sweep') or []: print('   sweep', s['n_particles'], round(s['ms_per_step_median'],4), round(s['frac'],3))
    except Exception as e: print(f, 'failed', e)
PY
tail -5 gpurun_out/r2_2gpu.err
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "split_over_two_ranks or symmetric_kernel" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --particles 100000 --steps 10 --no-cpu --no-sweep > gpurun_out/r2_2gpu_torchrun_1e5.json 2>> gpurun_out/r2_2gpu.err
python bench.py --gpus 1 --particles 100000 --steps 10 --no-cpu --no-sweep > gpurun_out/r2_1gpu_1e5.json 2>> gpurun_out/r2_2gpu.err
python - <<'PY'
import json
a=json.loads(open('gpurun_out/r2_1gpu_1e5.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/r2_2gpu_torchrun_1e5.json').read().strip().splitlines()[-1])
print('1e5: 1 GPU', a['ms_per_step'], ' 2 GPUs', b['ms_per_step'], ' efficiency', a['ms_per_step']/b['ms_per_step']/2, 'kernel', a['roofline']['kernel_ms'], b['roofline']['kernel_ms'])
PY
