#!/usr/bin/env python
"""Two-rank diagnostic (torchrun): host wall-clock of the step phases with the peer-memory exchange attached."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
mode = sys.argv[2] if len(sys.argv) > 2 else "p2p"
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n, device=local)
hp = rb.HotPath(cfg)
hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
hp.set_option("pair_mode", 2)
hp.set_pair_rank(rank, world)
if mode == "p2p":
    handles = [None] * world
    dist.all_gather_object(handles, hp.p2p_export(n))
    hp.p2p_attach(world, rank, handles)
def t(f):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); f(); return (time.perf_counter() - t0) * 1e3
for it in range(4):
    a = t(lambda: hp.Update_Particle_Position(it + 1))
    if mode == "p2p":
        b = t(lambda: hp.Calculate_Acceleration_Particles())
    else:
        b = t(lambda: hp.accel_partial())
    c = t(lambda: hp.Update_Particle_Velocity())
    d = t(lambda: hp.Update_Position(10 + it)) if mode == "p2p" else 0.0
    r = hp.Update_Position(20 + it) if mode == "p2p" else None
    print(f"rank {rank} it {it}: position {a:.3f} ms  accel {b:.3f} ms  velocity {c:.3f} ms  full step {d:.3f} ms"
          + (f"  (device: accel {r.accel_ms:.3f} step {r.step_ms:.3f})" if r else ""), flush=True)
dist.barrier(); torch.cuda.synchronize()
if mode == "p2p": hp.p2p_detach()
hp.close(); dist.destroy_process_group()
