"""Two rb2_mh_tip calls (214 chains x 80 jumps against ~1e3 electrons above the Tip-FE deck's tip) for an ncu launch list,
and their wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, Q_0
NM = 1e-9
N, M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 214
cfg = rb.tip_config(2000.0, 1000 * NM, 250 * NM, 500 * NM, (0.0, 0.0, 1500 * NM), 1e-16, True, capacity=N + 1024)
with rb.HotPath(cfg) as hp:
    rng = np.random.default_rng(1)
    pos = np.stack([rng.uniform(-60, 60, N), rng.uniform(-60, 60, N), rng.uniform(520, 1400, N)], axis=1) * NM
    hp.upload(pos, np.full(N, -Q_0), np.full(N, M_0))
    hp.mh_tip(M, seed=1)
    t = []
    for k in range(3):
        t0 = time.perf_counter(); r = hp.mh_tip(M, seed=2 + k); t.append(time.perf_counter() - t0)
    print(f"rb2_mh_tip N={N} M={M}: {min(t)*1e3:.3f} ms per call ({min(t)*1e6/80:.1f} us per jump), mean normal field {r[0].mean():.4e}")
