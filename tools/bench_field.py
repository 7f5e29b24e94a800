#!/usr/bin/env python
"""Batched surface-field evaluation sweep: M points on z = 0 against N electrons, through the C ABI (host
buffers in / out).  Two entry points: rb2_field_batch (Calc_Field_at_Batch, all three components, general
kernel) and rb2_field_surface_z (E_z only, mirror-antisymmetric form).  Reports call latency and
point-interactions/s; BASELINE.md section 4."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM


def timed(fn, pts, work):
    reps = 200 if work < 1e8 else (20 if work < 1e10 else 3)
    fn(pts)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(pts)
    return (time.perf_counter() - t0) / reps


for n in (1000, 10000, 100000, 1000000):
    pos = make_cloud(n)
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
        peak, _ = hp.fp64_peak(30.0)
        rng = np.random.default_rng(1)
        for M in (1, 32, 256, 4096, 65536):
            if M * n > 7e10:
                continue
            pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), np.zeros(M)], axis=1) * NM
            t = timed(hp.Calc_Field_at_Batch, pts, M * n)
            ts = timed(hp.field_surface_z, pts, M * n)
            # FP64 instructions per (point, particle) at N_ic_max = 1: 74 general, 31 surface (FMA = 2 flops)
            rec = dict(N=n, M=M, batch_call_us=round(t * 1e6, 2), batch_rate=M * n / t,
                       batch_fp64_pipe_frac=round(74 * 2 * M * n / t / 1e12 / peak, 4),
                       surface_call_us=round(ts * 1e6, 2), surface_rate=M * n / ts,
                       surface_fp64_pipe_frac=round(31 * 2 * M * n / ts / 1e12 / peak, 4), speedup=round(t / ts, 2))
            print(json.dumps(rec), flush=True)
