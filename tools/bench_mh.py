"""Wall time of one emission-sampler call (M lock-step chains x 200 jumps against N particles):
host loop (mh_batch=1, one rb2_field_batch round trip per jump) vs device-resident chains (mh_batch=2,
rb2_mh_planar).  Prints one JSON line per (N, M)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rumdeed_b200 as rb
from rumdeed_b200.host_api import Simulation

NM = 1e-9


def run(mh_batch, N, M, reps=3):
    emit, d = 1000 * NM, 1000 * NM
    sim = Simulation(seed=5, emission_mode=10, V_s=2000.0, box_dim=(emit, emit, d), time_step=0.25e-15, image_charge=True,
                     N_ic_max=1, emitters_pos=(-0.5 * emit, -0.5 * emit, 0.0), emitters_dim=(emit, emit, 0.0), emitters_type=2,
                     mh_batch=mh_batch, w_theta=((4.7,),), max_particles=max(N + 1000, 20000))
    with sim:
        if N:
            rng = np.random.default_rng(1)
            pos = np.stack([rng.uniform(-0.5 * emit, 0.5 * emit, N), rng.uniform(-0.5 * emit, 0.5 * emit, N),
                            rng.uniform(1 * NM, 0.9 * d, N)], axis=1)
            rb.HotPath.attach().Add_Particles(pos, np.zeros((N, 3)), np.ones(N, dtype=np.int32), 0)
        sim.Metropolis_Hastings_rectangle_J_batch(M)
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            df, F, pos = sim.Metropolis_Hastings_rectangle_J_batch(M)
            t.append(time.perf_counter() - t0)
    return min(t), float(np.mean(F))


if __name__ == "__main__":
    for N in (0, 1000, 10000, 100000):
        for M in (10, 100, 1000, 10000):
            if N * M > 2e9:
                continue
            th, Fh = run(1, N, M)
            td, Fd = run(2, N, M)
            print(json.dumps({"N": N, "M": M, "host_loop_ms": round(th * 1e3, 3), "device_ms": round(td * 1e3, 3),
                              "speedup": round(th / td, 2), "F_mean_host": Fh, "F_mean_device": Fd}), flush=True)
