#!/bin/bash
# 2-GPU check: the peer-memory exchange test, then bench.py at 3 sizes with both exchanges.
timeout 600 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for n in 1000000 100000 10000; do
  for ex in p2p nccl; do
    timeout 300 $TR bench.py --gpus 2 --particles $n --steps 6 --warmup 3 --exchange $ex 2>gpurun_out/b2_${n}_${ex}.err | tail -1 > gpurun_out/b2_${n}_${ex}.json
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b2_${n}_${ex}.json").read())
    print("$n $ex", "%.4e" % d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], "%.4e" % d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["ms_steps_rank0"])
except Exception as e:
    print("$n $ex failed", e); print(open("gpurun_out/b2_${n}_${ex}.err").read()[-1500:])
PY
  done
done
