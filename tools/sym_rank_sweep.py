#!/usr/bin/env python
"""Per-rank time of the pair-symmetric kernel when its work units are dealt to `world` ranks, emulated on ONE
GPU (rb2_set_pair_rank + rb2_accel_partial): shows how evenly the split divides the single-GPU time and how
the CTA-group size (sym_waves) should follow the rank count."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
    hp.set_option("pair_mode", 2)
    t1 = None
    for world in (1, 2, 4, 8):
        for waves in (16, 4, 64):
            hp.set_option("sym_waves", waves)
            ts = []
            for rank in sorted({0, world // 2, world - 1}):
                hp.set_pair_rank(rank, world)
                hp.accel_partial()
                t0 = time.perf_counter()
                hp.accel_partial()
                ts.append((time.perf_counter() - t0) * 1e3)
            if world == 1 and waves == 16:
                t1 = ts[0]
            print(f"n={n} world={world} sym_waves={waves:4d}: per-rank ms {['%.1f' % t for t in ts]}  ideal {t1 / world:.1f}  "
                  f"efficiency {t1 / world / max(ts):.3f}", flush=True)
