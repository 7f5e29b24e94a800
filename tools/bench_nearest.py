#!/usr/bin/env python
"""Times rb2_nearest_electron (Sample_Elec_Position sweep) on the synthetic cloud; host arrays out."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

for n in [int(float(a)) for a in sys.argv[1:]] or [10000, 100000, 1000000]:
    pos = make_cloud(n)
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
        hp.Sample_Elec_Position()
        ts = []
        for k in range(3):
            t0 = time.perf_counter(); d, i = hp.Sample_Elec_Position(); ts.append(time.perf_counter() - t0)
        t = min(ts)
        print(f"n={n}: {t*1e3:.3f} ms  {n*(n-1)/t:.3e} ordered pairs/s  mean nearest distance {d.mean()/NM:.3f} nm", flush=True)
