#!/usr/bin/env python
"""Small driver for ncu captures: uploads the synthetic cloud and runs a few hot-path steps
(and optionally field batches) without any of bench.py's extra legs."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rumdeed_b200 as rb  # noqa: E402
from bench import NM, make_cloud  # noqa: E402
from rumdeed_b200.api import M_0, Q_0  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--nic", type=int, default=1)
ap.add_argument("--field", type=int, default=0, help="also run this many field batches of --m points")
ap.add_argument("--m", type=int, default=256)
a = ap.parse_args()
pos = make_cloud(a.n)
cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1.0e-16, True, a.nic, capacity=a.n)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, np.full(a.n, -Q_0), np.full(a.n, M_0))
    for s in range(a.steps):
        r = hp.Update_Position(s + 1)
        print("step", s + 1, "accel_ms", r.accel_ms, "step_ms", r.step_ms, flush=True)
    rng = np.random.default_rng(1)
    for k in range(a.field):
        pts = np.stack([rng.uniform(-500, 500, a.m), rng.uniform(-500, 500, a.m), np.zeros(a.m)], axis=1) * NM
        hp.Calc_Field_at_Batch(pts)
