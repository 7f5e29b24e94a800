#!/usr/bin/env python
"""Pair-symmetric kernel: one vs two targets per lane (sym_tpl) x sym_waves at several N; prints the checksum too."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

sizes = [int(float(a)) for a in sys.argv[1:]] or [16384, 100000]
for n in sizes:
    pos = make_cloud(n)
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
        hp.set_option("pair_mode", 2)
        for tpl in (1, 2):
            for waves in (16, 64):
                hp.set_option("sym_tpl", tpl)
                hp.set_option("sym_waves", waves)
                ts = []
                for k in range(3 if n < 500000 else 2):
                    hp.Calculate_Acceleration_Particles()
                    ts.append(hp.last_accel_info()["ms"])
                info = hp.last_accel_info()
                acc = hp.download(("acc",))["acc"]
                print(f"n={n} tpl={tpl} waves={waves:3d}: best {min(ts):9.3f} ms  pair-int/s {n*(n-1)/min(ts)*1e3:.4e}  "
                      f"bands {info['grid_y']} G {info['j_chunk']}  checksum {float(np.abs(acc).sum()):.17g}", flush=True)
