#!/usr/bin/env python
"""Tip geometry: acceleration evaluation (gather kernel, sphere image on) and a field batch at several N --
ordered pair evaluations per second and the FP64 pipe share at ~36 instructions per pair (rb2_tip_math.cuh)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0

NM = 1e-9
for n in (1000, 10000, 100000):
    rng = np.random.default_rng(n)
    pos = np.stack([rng.uniform(-400, 400, n), rng.uniform(-400, 400, n), rng.uniform(105, 900, n)], axis=1) * NM
    cfg = rb.tip_config(2.0e3, 900 * NM, 100 * NM, 100 * NM, (100 * NM, 100 * NM, 1000 * NM), 0.25e-15, True, capacity=n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
        peak, _ = hp.fp64_peak(30.0)
        ts = []
        for k in range(4):
            hp.Calculate_Acceleration_Particles()
            ts.append(hp.last_accel_info()["ms"])
        t = min(ts[1:])
        M = 10000
        pts = np.stack([rng.uniform(-300, 300, M), rng.uniform(-300, 300, M), rng.uniform(101, 400, M)], axis=1) * NM
        hp.Calc_Field_at_Batch(pts)
        t0 = time.perf_counter(); hp.Calc_Field_at_Batch(pts); tf = time.perf_counter() - t0
        print(f"tip N={n}: acceleration {t:.3f} ms = {n*(n-1)/t*1e3:.3e} ordered pair evaluations/s "
              f"(~{36*2*n*(n-1)/t*1e3/1e12/peak:.2f} of the FP64 peak at 36 instructions per pair); "
              f"field batch M={M}: {tf*1e3:.3f} ms = {M*n/tf:.3e} point-interactions/s", flush=True)
