"""One rb2_mh_planar call per kernel variant for ncu: N electrons, M chains (args), mh_small 1 then 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, Q_0
NM = 1e-9
N, M = int(sys.argv[1]), int(sys.argv[2])
emit, d = 1000 * NM, 1000 * NM
cfg = rb.planar_config(2000.0, d, (emit, emit, d), 1e-16, True, 1, capacity=N + 1024)
with rb.HotPath(cfg) as hp:
    rng = np.random.default_rng(1)
    pos = np.stack([rng.uniform(-0.5 * emit, 0.5 * emit, N), rng.uniform(-0.5 * emit, 0.5 * emit, N), rng.uniform(1 * NM, 0.9 * d, N)], axis=1)
    hp.upload(pos, np.full(N, -Q_0), np.full(N, M_0))
    args = dict(emit_pos=(-0.5 * emit, -0.5 * emit), emit_dim=(emit, emit), w_theta=((4.7,),))
    for small in (1, 0):
        hp.set_option("mh_small", small)
        r = hp.mh_planar(M, seed=3, **args)
        print("small", small, "F mean", r[1].mean(), "MH_std", r[4])
