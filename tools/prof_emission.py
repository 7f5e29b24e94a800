#!/usr/bin/env python
"""One call of each emission-side kernel for ncu: rb2_field_surface_z (M points) and rb2_mh_planar (M chains,
a few jumps) against N electrons.  Usage: prof_emission.py [N] [M] [ndim]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ndim = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), np.zeros(M)], axis=1) * NM
    ez = hp.field_surface_z(pts)
    full = hp.Calc_Field_at_Batch(pts)
    print("max rel diff surface vs general:", float(np.max(np.abs(ez - full[:, 2]) / np.abs(ez))))
    df, F, p, a_rate, sd = hp.mh_planar(M, (-500 * NM, -500 * NM), (1000 * NM, 1000 * NM), ((4.7,),), 1234, ndim=ndim)
    print("chains:", M, "mean F", float(F.mean()), "a_rate", a_rate, "MH_std", sd)
