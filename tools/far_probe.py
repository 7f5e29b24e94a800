"""Largest relative error of the two pair kernels against the long-double CPU checker with and without the cheaper
far-partner inverse cube (option sym_far), headline-shaped cloud (d = 1000 nm).  usage: python tools/far_probe.py [n]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import rumdeed_b200 as rb
from oracle.oracle import Oracle
orc = Oracle()
from test_gpu_parity import cloud, planar, relerr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfg, p = planar(orc)
pos, q, m, sp = cloud(n, 99, ions=False)
truth = orc.accel_gather_ld(p, pos, q, m)
with rb.HotPath(cfg) as hp:
    hp.upload(pos, q, m, species=sp)
    for far in (0, 1):
        hp.set_option("sym_far", far)
        for mode in (1, 2):
            hp.set_option("pair_mode", mode)
            hp.Calculate_Acceleration_Particles()
            acc = hp.download(("acc",))["acc"]
            print(f"n={n} sym_far={far} pair_mode={mode} max rel err vs long double: {relerr(acc, truth):.3e}")
