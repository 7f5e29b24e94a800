"""The CPU oracle against tests/golden/reference_vectors.json (values transcribed from the reference's src/mod_tests.F90,
see tests/golden/README.md).  Tolerances: the reference's own 2 % where it compares against externally computed numbers
(the probe point of the field vectors is built from single-precision literals there), 1e-9 .. 1e-12 for closed forms."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import SPECIES_ELEC, SPECIES_ION

NM = 1.0e-9
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def _three(orc, g):
    inp = g["field_three_particles_no_image_charge"]["inputs"]
    R = np.array(inp["particles_nm"]) * NM
    k = orc.k
    q = np.where(np.array(inp["species"]) == 2, k.q_0, -k.q_0)
    sp = np.array(inp["species"], dtype=np.int32)
    d = inp["d_nm"] * NM
    probe = np.array(inp["probe_nm_single_precision"], dtype=np.float32).astype(np.float64) * NM
    return R, q, sp, d, inp["V"], probe


def test_field_vectors(orc, gold):
    R, q, sp, d, V, probe = _three(orc, gold)
    p = orc.params_planar(V, d, (100 * NM, 100 * NM, d), 0.25e-15, False, 0)
    got = orc.calc_field_at(p, R, q, probe, sp)
    want = np.array(gold["field_three_particles_no_image_charge"]["expected_V_per_m"])
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-6
    # the disabled variant was generated for an older two-partner model (partners at -z and at 2d - z, mod_tests.F90:740-765):
    # adding those two partners of every particle as explicit opposite charges reproduces it
    probe_d = np.array(gold["field_three_particles_no_image_charge"]["inputs"]["probe_nm_single_precision"]) * NM
    Ra = R * np.array([1.0, 1.0, -1.0])
    Rb = R * np.array([1.0, 1.0, -1.0]) + np.array([0.0, 0.0, 2.0 * d])
    got2 = orc.calc_field_at(p, np.concatenate([R, Ra, Rb]), np.concatenate([q, -q, -q]), probe_d,
                             np.concatenate([sp, sp, sp]))
    want2 = np.array(gold["field_three_particles_image_charge_n0"]["expected_V_per_m"])
    assert np.linalg.norm(got2 - want2) / np.linalg.norm(want2) < gold["tolerance_rel_reference"]


def test_fowler_nordheim_values(orc, gold):
    g = gold["fowler_nordheim_4p7eV"]
    p = orc.params_planar(2000.0, 1000 * NM, (100 * NM, 100 * NM, 1000 * NM), 1e-16, True, 0)
    w = g["inputs"]["w_theta_eV"]
    for c in g["cases"]:
        F = c["F_V_per_m"]
        assert orc.fn_v_y(p, F, w) == pytest.approx(c["v_y"], rel=1e-12)
        assert orc.fn_t_y(p, F, w) == pytest.approx(c["t_y"], rel=1e-12)
        assert orc.fn_escape_prob_log(p, F, w) == pytest.approx(c["ln_D"], rel=1e-10)


def test_collision_math_values(orc, gold):
    from oracle.collisions import Collisions
    col = Collisions(orc)
    g = gold["collision_math"]
    for c in g["normal_dist"]:
        assert col.normal_dist(c["mu"], c["sigma"], c["x"]) == pytest.approx(c["expected"], rel=1e-12)
    for c in g["folded_normal_dist"]:
        assert col.folded_normal_dist(c["mu"], c["sigma"], c["x"]) == pytest.approx(c["expected"], rel=1e-12)
    for c in g["kramers_cross_section"]:
        assert col.kramers(c["energy_eV"]) * c["scale"] == pytest.approx(c["expected"], rel=1e-9)


def test_particle_removal_fixture(orc, gold):
    g = gold["particle_removal"]
    p = orc.params_planar(2000.0, 1000 * NM, (100 * NM, 100 * NM, 1000 * NM), 1e-16, True, 0)
    st = orc.store(16)
    n = g["inputs"]["particles"]
    for i in range(n):
        st.add(p, [i * NM, 0.0, (10 + i) * NM], [0, 0, 0], SPECIES_ELEC if i % 2 == 0 else SPECIES_ION, 0, 1)
    for s in g["inputs"]["remove_1_based"]:
        st.mark(s - 1, 1)
        assert st.charge[s - 1] == 0.0
    st.remove(g["inputs"]["life_time_steps"])
    assert st.n == len(g["expected"]["survivor_ids"])
    assert list(st.ids) == g["expected"]["survivor_ids"]
    assert [round(x / NM) + 1 for x in st.pos[:, 0]] == g["expected"]["survivor_slots_1_based"]
