#!/bin/bash
# 2-GPU check: the peer-memory / multi-device tests, then bench.py at N = 1e6 with the SAME step counts as the 1-GPU line
# (the checksum of the accelerations must agree to rounding), from two processes and from one.
timeout 900 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_multidevice.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], "%.4e" % d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["acc_checksum"]["sum_abs"], d["acc_checksum"]["replicas_identical"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
timeout 400 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-sweep 2>gpurun_out/b2_torchrun.err | tail -1 > gpurun_out/b2_torchrun.json; show gpurun_out/b2_torchrun.json
timeout 400 python bench.py --gpus 2 --single-process --steps 3 --warmup 3 --no-cpu --no-sweep 2>gpurun_out/b2_single.err | tail -1 > gpurun_out/b2_single.json; show gpurun_out/b2_single.json
