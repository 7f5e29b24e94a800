"""Generates tests/golden/deck_oracle_runs.json: the CPU oracle's own runs of the Photo and Checkerboard-TFE
configurations (tools/decks/photo, tools/decks/checkerboard_tfe = the parameter values of the reference's
Examples/Photo and Examples/Checkerboard-TFE), recorded per step, for the GPU deck tests to compare with.

The oracle restates the reference's serial loops (mod_photo_emission.f90:603-686,
mod_field_thermo_emission.F90:136-364, mod_verlet.F90 Beeman step); the full-size Photo deck puts ~45 000
electrons into the gap in its first step, which takes the CPU ~10 minutes for 200 steps -- hence a fixture.

    python tests/golden/make_deck_fixtures.py            (about 15 minutes on 8 cores)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import SUPPLY_GTF, Emission, Oracle  # noqa: E402

NM = 1.0e-9


def run_photo(orc, emit_nm, steps, seed):
    box = (emit_nm * NM, emit_nm * NM, 1000 * NM)
    p = orc.params_planar(2000.0, 1000 * NM, box, 1.0e-16, True, 1)
    st = orc.store(200000)
    em = Emission(orc, p, st, (-0.5 * emit_nm * NM, -0.5 * emit_nm * NM, 0.0), (emit_nm * NM, emit_nm * NM, 0.0), ((2.5,),), seed=seed)
    emitted, nrpart, ramo = [], [], []
    for i in range(1, steps + 1):
        e = em.get_laser_energy(4.7, 0.02)  # laser file: 2 2 2 / 4.7 0.02
        emitted.append(int(em.do_photo_emission_rectangle(i, e, 2, -1)))
        st.step(p)
        ramo.append(float(st.s.ramo_current[1]))
        nrpart.append(int(st.s.nrPart))
        st.remove(i)
    return dict(emit_nm=emit_nm, steps=steps, seed=seed, emitted=emitted, nrPart=nrpart, ramo=ramo,
                absorbed_bot=int(sum(emitted)) - int(st.s.nrPart))


def run_tfe(orc, steps, seed, grid=16):
    w = ((2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0), (2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0))
    box = (100 * NM, 100 * NM, 1000 * NM)
    p = orc.params_planar(2000.0, 1000 * NM, box, 1.0e-16, True, 1)
    st = orc.store(200000)
    em = Emission(orc, p, st, (-50 * NM, -50 * NM, 0.0), (100 * NM, 100 * NM, 0.0), w, T_temp=1000.0, seed=seed)
    emitted, nrpart, ramo, nsup = [], [], [], []
    sec_ramo = np.zeros(16)
    for i in range(1, steps + 1):
        N_sup, _ = em.supply_grid(SUPPLY_GTF, grid)
        nsup.append(float(N_sup))
        emitted.append(int(em.do_field_thermo_emission_planar(i, N_sup)))
        st.step(p)
        ramo.append(float(st.s.ramo_current[1]))
        sec_ramo += st.ramo_current_emit(16)
        nrpart.append(int(st.s.nrPart))
        st.remove(i)
    # sections of every electron still in the gap + of those absorbed is not tracked by the store: count at emission
    return dict(steps=steps, seed=seed, emitted=emitted, nrPart=nrpart, ramo=ramo, N_sup=nsup,
                sec_ramo_sum=[float(x) for x in sec_ramo], sec_in_gap=np.bincount(st.section, minlength=17)[1:17].tolist())


if __name__ == "__main__":
    orc = Oracle()
    out = {}
    t0 = time.time()
    out["tfe"] = [run_tfe(orc, 400, 100 + k) for k in range(3)]
    print("tfe", time.time() - t0, [sum(r["emitted"]) for r in out["tfe"]], flush=True)
    out["photo_small"] = [run_photo(orc, 100.0, 200, 200 + k) for k in range(3)]
    print("photo 100 nm", time.time() - t0, [sum(r["emitted"]) for r in out["photo_small"]], flush=True)
    if "--no-full" not in sys.argv:
        out["photo_full"] = [run_photo(orc, 500.0, 200, 300)]
        print("photo 500 nm", time.time() - t0, sum(out["photo_full"][0]["emitted"]), flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "deck_oracle_runs.json"), "w") as f:
        json.dump(out, f)
    print("done", time.time() - t0)
