#!/usr/bin/env python
"""Where the pair-symmetric kernel overtakes the gather kernel: acceleration evaluation time vs N."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

for n in (500, 1000, 2000, 3000, 4000, 6000, 8000, 10000, 16384, 32768):
    pos = make_cloud(n)
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
        out = []
        for mode, tpl in ((1, 1), (2, 1), (2, 2)):
            hp.set_option("pair_mode", mode)
            hp.set_option("sym_tpl", tpl)
            ts = []
            for k in range(8):
                hp.Calculate_Acceleration_Particles()
                ts.append(hp.last_accel_info()["ms"])
            out.append(min(ts))
        print(f"n={n:6d}  gather {out[0]*1e3:9.1f} us   sym tpl=1 {out[1]*1e3:9.1f} us   sym tpl=2 {out[2]*1e3:9.1f} us", flush=True)
