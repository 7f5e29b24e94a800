#!/usr/bin/env python
"""Sweep the pair-symmetric kernel's scheduling knobs (sym_waves, sym_budget_mb) at a given N."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
    hp.set_option("pair_mode", 2)
    for budget in (256, 1024, 2048, 8192):
        for waves in (4, 8, 16, 32, 64):
            hp.set_option("sym_budget_mb", budget)
            hp.set_option("sym_waves", waves)
            ts = []
            for k in range(reps):
                hp.Calculate_Acceleration_Particles()
                ts.append(hp.last_accel_info()["ms"])
            info = hp.last_accel_info()
            print(f"n={n} budget={budget:5d} MB waves={waves:3d}: best {min(ts):9.3f} ms  pair-int/s {n*(n-1)/min(ts)*1e3:.4e}  bands {info['grid_y']} G {info['j_chunk']}", flush=True)
