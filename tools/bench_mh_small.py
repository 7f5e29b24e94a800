"""ms per rb2_mh_planar call (M chains x 200 jumps against N electrons): the single-barrier kernel for <= 32 chains
(k_mh_small) against the two-barrier cooperative kernel (k_mh_persistent, option mh_small = 0)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, Q_0

NM = 1e-9
emit, d = 1000 * NM, 1000 * NM
for N in [int(float(a)) for a in sys.argv[1:]] or (0, 1000, 5000, 10000, 19000, 40000, 75000):
    cfg = rb.planar_config(2000.0, d, (emit, emit, d), 1e-16, True, 1, capacity=max(N, 1) + 1024)
    with rb.HotPath(cfg) as hp:
        if N:
            rng = np.random.default_rng(1)
            pos = np.stack([rng.uniform(-0.5 * emit, 0.5 * emit, N), rng.uniform(-0.5 * emit, 0.5 * emit, N),
                            rng.uniform(1 * NM, 0.9 * d, N)], axis=1)
            hp.upload(pos, np.full(N, -Q_0), np.full(N, M_0))
        args = dict(emit_pos=(-0.5 * emit, -0.5 * emit), emit_dim=(emit, emit), w_theta=((4.7,),))
        for M in [int(m) for m in os.environ.get("MH_M", "1,10,32").split(",")]:
            out = {"N": N, "M": M}
            for small in (1, 0):
                hp.set_option("mh_small", small)
                hp.mh_planar(M, seed=1, **args)
                t = []
                for k in range(5):
                    t0 = time.perf_counter(); r = hp.mh_planar(M, seed=2 + k, **args); t.append(time.perf_counter() - t0)
                out["small_ms" if small else "persistent_ms"] = round(float(np.median(t)) * 1e3, 3)
                out["F_small" if small else "F_persistent"] = float(np.mean(r[1]))
            out["speedup"] = round(out["persistent_ms"] / out["small_ms"], 2)
            print(json.dumps(out), flush=True)
