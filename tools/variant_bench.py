#!/usr/bin/env python
"""Times the acceleration evaluation of every kernel variant under tools/variants/ (one process each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
code = r'''
import sys, os, numpy as np
sys.path.insert(0, %r)
import rumdeed_b200 as rb
from rumdeed_b200.api import Q_0, M_0
from bench import make_cloud, NM
n = %d
lib = sys.argv[1]
pos = make_cloud(n)
cfg = rb.planar_config(2000.0, 1000*NM, (1000*NM,)*3, 1e-16, True, 1, capacity=n)
hp = rb.HotPath(cfg, lib_path=lib)
hp.upload(pos, np.full(n, -Q_0), np.full(n, M_0))
ts = []
for k in range(6):
    hp.Calculate_Acceleration_Particles()
    ts.append(hp.last_accel_info()["ms"])
info = hp.last_accel_info()
acc = hp.download(("acc",))["acc"]
print("%%-10s best %%.3f ms  median %%.3f ms  grid %%dx%%d  checksum %%.17g" %% (os.path.basename(os.path.dirname(lib)), min(ts), sorted(ts)[len(ts)//2], info["grid_x"], info["grid_y"], float(np.abs(acc).sum())))
hp.close()
''' % (ROOT, n)
vdir = os.path.join(ROOT, "tools", "variants")
for name in sorted(os.listdir(vdir)):
    lib = os.path.join(vdir, name, "librumdeed_b200.so")
    if os.path.exists(lib):
        r = subprocess.run([sys.executable, "-c", code, lib], capture_output=True, text=True)
        print((r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
