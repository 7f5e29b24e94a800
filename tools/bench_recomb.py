#!/usr/bin/env python
"""Times the collision step (rb2_discrete_recombination / rb2_continuous_ionization) on a synthetic plasma:
n electrons with 10 - 2000 eV, n/10 ions, uniform in the 1 um^3 gap (positions of bench.make_cloud).

Prints, per size: device time of the recombination call (CUDA events, median of 5), (ion, electron) pairs/s,
candidates that reached the quartic, and -- for the smaller sizes -- the CPU oracle's restatement of
Do_Discrete_Recombination_ots (every pair through the quartic, like the reference) on a bounded sample of ions.
usage: tools/bench_recomb.py [n_electrons ...] [--cpu]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, M_N2P, Q_0, SPECIES_ELEC, SPECIES_ION
from bench import NM, make_cloud
from test_oracle_collisions import synthetic_tables

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cpu = "--cpu" in sys.argv
N_D = 101325.0 / (1.380649e-23 * 293.15)
DT = 1.0e-16

for ne in [int(float(a)) for a in args] or [10000, 100000, 1000000]:
    ni = max(ne // 10, 1)
    n = ne + ni
    rng = np.random.default_rng(ne)
    pos = make_cloud(n)
    species = np.full(n, SPECIES_ELEC, np.int32)
    species[rng.choice(n, ni, replace=False)] = SPECIES_ION
    ion = species == SPECIES_ION
    E = 10.0 ** rng.uniform(1.0, 3.3, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    vel = d * np.sqrt(2 * Q_0 * E / M_0)[:, None]; vel[ion] = 0.0
    acc = np.tile([0.0, 0.0, 3.5e20], (n, 1)); acc[ion] = 0.0
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, DT, True, 1, capacity=n + 1024)
    with rb.HotPath(cfg) as hp:
        hp.Init_Collisions(2, *synthetic_tables(), n_d=N_D, cyl_radius=1000 * NM)
        hp.upload(pos, np.where(ion, Q_0, -Q_0), np.where(ion, M_N2P, M_0), vel=vel, acc=acc, prev_pos=pos - vel * DT,
                  species=species, life=np.where(ion, 10 ** 8, -1).astype(np.int32))
        ms, wall = [], []
        for k in range(6):
            t0 = time.perf_counter(); r = hp.Do_Discrete_Recombination(1); wall.append(time.perf_counter() - t0)
            ms.append(r.ms)
        t = float(np.median(ms[1:])) * 1e-3
        print(f"recombination  n_elec={ne} n_ion={ni}: device {t*1e3:.3f} ms (wall {np.median(wall[1:])*1e3:.3f} ms)  "
              f"{ne*ni/t:.3e} pairs/s  candidates {r.n_candidates}  recombinations {r.nrRecombinations}", flush=True)
        ms = []
        for k in range(4):
            r = hp.Do_Continuous_Ionization(2 + k, 1234 + k); ms.append(r.ms)
        print(f"ionisation     n_elec={ne}: device {np.median(ms[1:]):.3f} ms  collisions {r.nrCollisions} ionisations {r.nrIonizations}", flush=True)
    if cpu and ne <= 200000:
        from oracle.collisions import Collisions
        col = Collisions(tables=synthetic_tables())
        rr = col.collision_data(vel)[:, 3]
        sample = max(1, min(ni, int(2.0e7 // ne)))   # ~2e7 quartic solves
        keep = np.ones(n, bool); keep[np.nonzero(ion)[0][sample:]] = False
        t0 = time.perf_counter()
        nr = col.discrete_recombination(pos[keep], vel[keep], acc[keep], species[keep], np.ones(keep.sum(), np.int32),
                                        np.where(ion, 10 ** 8, -1).astype(np.int32)[keep], np.zeros(keep.sum(), np.int32),
                                        np.ones(keep.sum(), np.int32), rr[keep], 1, DT)[0]
        tc = time.perf_counter() - t0
        print(f"cpu oracle     {sample} ions x {ne} electrons: {tc:.2f} s  {sample*ne/tc:.3e} pairs/s (1 core)  -> GPU/CPU {ne*ni/t/(sample*ne/tc):.0f}x", flush=True)
