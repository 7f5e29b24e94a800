#!/usr/bin/env python
"""Acceleration evaluation time against the shape of the pair-symmetric work units (K target superblocks x G source
tiles per CTA), the scratch budget and the number of waves per launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000*NM, (1000*NM,)*3, 1e-16, True, 1, capacity=n)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
    for budget in (2048, 4096):
        for waves in (16, 32):
            for kmax, gmax in ((64, 24), (24, 24), (12, 24), (12, 12), (6, 12)):
                hp.set_option("sym_budget_mb", budget); hp.set_option("sym_waves", waves)
                hp.set_option("sym_kmax", kmax); hp.set_option("sym_gmax", gmax)
                ts = []
                for k in range(3):
                    hp.Calculate_Acceleration_Particles()
                    ts.append(hp.last_accel_info()["ms"])
                info = hp.last_accel_info()
                print(f"n={n} budget={budget} waves={waves} kmax={kmax} gmax={gmax}: best {min(ts):.2f} ms  bands {info['grid_y']} G*1000+K {info['j_chunk']}", flush=True)
