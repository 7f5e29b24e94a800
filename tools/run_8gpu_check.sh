#!/bin/bash
# 8-GPU check: bench.py at N = 1e6 and 1e5 with the peer-memory exchange (default).
G=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521"
for n in 1000000 100000; do
  timeout 400 $TR bench.py --gpus $G --particles $n --steps 5 --warmup 3 --no-cpu --no-sweep 2>gpurun_out/b${G}_${n}.err | tail -1 > gpurun_out/b${G}_${n}.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b${G}_${n}.json").read())
    print("$n", "%.4e" % d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], "%.4e" % d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["ms_steps_rank0"])
except Exception as e:
    print("$n failed", e); print(open("gpurun_out/b${G}_${n}.err").read()[-2500:])
PY
done
