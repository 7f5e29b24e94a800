#!/usr/bin/env python
"""One GPU playing each of `world` ranks in turn (rb2_set_pair_rank + rb2_accel_partial): the slowest rank's share of the
pair-symmetric evaluation against 1/world of the undivided one, per sym_waves setting -- the kernel-side part of the
strong-scaling efficiency, without the exchange.   usage: python tools/rank_emulation.py [n] [world]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000*NM, (1000*NM,)*3, 1e-16, True, 1, capacity=n)

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
    hp.set_option("pair_mode", 2)
    for waves in (8, 16, 32, 64, 128):
        hp.set_option("sym_waves", waves)
        hp.set_pair_rank(0, 1)
        hp.accel_partial()
        whole = timed(hp.accel_partial)
        info1 = hp.last_accel_info()
        per = []
        for r in range(world):
            hp.set_pair_rank(r, world)
            hp.accel_partial()
            per.append(timed(hp.accel_partial))
        info = hp.last_accel_info()
        hp.set_pair_rank(0, 1)
        print(f"n={n} world={world} waves={waves}: whole {whole:.3f} ms (G*1000+K {info1['j_chunk']}), ranks min {min(per):.3f} max {max(per):.3f} "
              f"(G*1000+K {info['j_chunk']}, bands {info['grid_y']}); kernel-side efficiency {whole / world / max(per):.3f}", flush=True)
