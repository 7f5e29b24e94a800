"""Emission rows: the host mirror of the reference's emission plugins running on
the device field evaluation, checked against the CPU oracle.

The reference's RNG (compiler RANDOM_NUMBER) and its quadrature (Cuba) are unpinned, so parity here
is statistical: distributions of sampled positions / fields, supply integrals within the reference's
own tolerance contract (epsabs 0.5, epsrel 1e-3), and whole-system observables (emitted / absorbed
counts, steady-state current band of mod_tests.F90:2020-2024).  CPU tests pin the deterministic
helpers (work function, GTF, FN, namelist reader) to golden values.
"""
import math
import os

import ctypes as C

import numpy as np
import pytest

import rumdeed_b200 as rb
from rumdeed_b200.host_api import SUPPLY_FE, SUPPLY_GTF, Simulation, kevin_jgtf_v2

NM = 1.0e-9
REF_EXAMPLES = "/root/reference/Examples"


# ---------------------------------------------------------------------------------------------------------
# CPU: deterministic helpers
def test_checkerboard_work_function_golden(orc):
    """mod_tests.F90:1918-1978, both the host mirror and the oracle."""
    from oracle.oracle import Emission
    w = ((4.10, 4.20), (4.30, 4.40))
    sim = Simulation(init=False, V_s=1000.0, box_dim=(100 * NM, 100 * NM, 500 * NM), time_step=0.25e-15,
                     emitters_pos=(0, 0, 0), emitters_dim=(100 * NM, 100 * NM, 0), w_theta=w)
    p = orc.params_planar(1000.0, 500 * NM, (100 * NM, 100 * NM, 500 * NM), 0.25e-15, True, 0)
    em = Emission(orc, p, orc.store(4), (0, 0, 0), (100 * NM, 100 * NM, 0), w)
    cases = [((25, 25), 4.30, 1), ((75, 25), 4.40, 2), ((25, 75), 4.10, 3), ((75, 75), 4.20, 4),
             ((150, 25), 4.40, 2), ((-10, 130), 4.10, 3)]
    for (x, y), val, sec in cases:
        pos = np.array([x, y, 0.0]) * NM
        assert sim.w_theta_xy(pos) == (val, sec)
        assert em.w_theta_xy(pos) == (val, sec)
    sim.close()


def test_gtf_function_host_equals_oracle(orc):
    from oracle.oracle import Emission
    p = orc.params_planar(1000.0, 500 * NM, (100 * NM, 100 * NM, 500 * NM), 0.25e-15, True, 0)
    em = Emission(orc, p, orc.store(4), (0, 0, 0), (100 * NM, 100 * NM, 0))
    for F in (-1e8, -5e8, -2e9, -4e9, -8e9):
        for T in (300.0, 1000.0, 1500.0, 2500.0):
            for w in (2.0, 2.5, 4.7):
                a, b = kevin_jgtf_v2(F, T, w), em.kevin_jgtf_v2(F, T, w)
                assert a == pytest.approx(b, rel=1e-13)
                assert a > 0.0
    assert kevin_jgtf_v2(-0.1, 1000.0, 4.7) == 0.0  # negligible-field guard (mod_kevin_rjgtf_v2.f90:66-69)
    # Richardson limit: at low field the GTF current approaches A T^2 exp(-(phi - sqrt(4 Q F))/kT)
    T, w, F = 2000.0, 4.7, -1.0e7
    kb, Q = 1.0 / 11604.50635, (1.0 / 137.035999084) * 0.6582119571 * 299.7924580 / 4.0
    phix = w - math.sqrt(4.0 * Q * abs(F) * 1e-9)
    rld = 120.173 * T * T * math.exp(-phix / (kb * T)) * 1e4
    assert kevin_jgtf_v2(F, T, w) == pytest.approx(rld, rel=0.05)


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference decks not present on this box")
def test_namelist_reader_on_reference_decks():
    """The reference's own example decks parse (units scaled like Read_Input_Variables, main.F90:348-383)."""
    s = Simulation(os.path.join(REF_EXAMPLES, "Checkerboard-TFE"), init=False)
    assert s.steps_in_input == 2000
    assert s.w_theta_xy(np.array([-40.0, -40.0, 0.0]) * NM) == (2.5, 1)   # bottom-left cell of the 4x4 board
    assert s.w_theta_xy(np.array([-40.0, 40.0, 0.0]) * NM) == (2.0, 13)   # top-left
    s.close()
    s = Simulation(os.path.join(REF_EXAMPLES, "Planar-FE", "2.0eV"), init=False)
    assert s.steps_in_input == 2000 and s.w_theta_xy(np.zeros(3))[0] == 2.0
    s.close()
    s = Simulation(os.path.join(REF_EXAMPLES, "Tip-FE"), init=False)
    assert s.steps_in_input == 50000
    s.close()


def test_namelist_reader_fixture(tmp_path):
    (tmp_path / "input").write_text(
        "&INPUT\n  V_S = 1.0d3,   ! volts\n  BOX_DIM = 0.0d0, 0.0d0, 500.0d0,\n  TIME_STEP = 0.25d-3,\n  STEPS = 17,\n"
        "  EMISSION_MODE = 10,\n  NREMIT = 1,\n  IMAGE_CHARGE = .True.,\n  N_IC_MAX = 1,\n  MH_BATCH = .true.,\n"
        "  EMITTERS_DIM(1:3, 1) = 100.0d0, 100.0d0, 0.0d0,\n  EMITTERS_POS(1:3, 1) = -50.0d0, -50.0d0, 0.0d0,\n"
        "  EMITTERS_TYPE(1) = 2,\n  EMITTERS_DELAY(1) = 0,\n  PLANES_N = 2,\n  PLANES_Z = 10.0d0, 250.0d0,\n"
        "  WRITE_POSITION_FILE = .true.,\n  SAMPLE_ELEC_FILE = .true.,\n  SAMPLE_ELEC_RATE = 200,\n/\n")
    (tmp_path / "work").write_text("1\n2 2\n4.1 4.2\n4.3 4.4\n")
    s = Simulation(str(tmp_path), init=False)
    assert s.steps_in_input == 17
    assert s.w_theta_xy(np.array([-25.0, -25.0, 0.0]) * NM) == (4.3, 1)
    assert s.w_theta_xy(np.array([25.0, 25.0, 0.0]) * NM) == (4.2, 4)
    s.close()


# ---------------------------------------------------------------------------------------------------------
# GPU: samplers against the oracle
gpu = pytest.mark.gpu


def _planar_pair(orc, seed, nic=1, mh_batch=False, w=((2.0,),), mode=10, T=293.15, V=1000.0, d=500 * NM, emit=100 * NM,
                 dt=0.25e-15, cap=20000, laser=None):
    """A host Simulation on the GPU and an oracle Emission with the same (empty) system."""
    from oracle.oracle import Emission
    box = (emit, emit, d)
    pos, dim = (-0.5 * emit, -0.5 * emit, 0.0), (emit, emit, 0.0)
    sim = Simulation(seed=seed, emission_mode=mode, V_s=V, box_dim=box, time_step=dt, image_charge=True, N_ic_max=nic,
                     emitters_pos=pos, emitters_dim=dim, emitters_type=2, T_temp=T, mh_batch=mh_batch, w_theta=w,
                     max_particles=cap, laser=laser)
    p = orc.params_planar(V, d, box, dt, True, nic)
    st = orc.store(cap)
    em = Emission(orc, p, st, pos, dim, w, T_temp=T, seed=seed + 1000)
    return sim, p, st, em


def _preload(sim, st, p, n, seed, emit=100 * NM, d=500 * NM):
    """Put the same space charge into both systems (n electrons above the emitter)."""
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(-0.5 * emit, 0.5 * emit, n), rng.uniform(-0.5 * emit, 0.5 * emit, n),
                    rng.uniform(1 * NM, 0.6 * d, n)], axis=1)
    hp = rb.HotPath.attach()
    hp.Add_Particles(pos, np.zeros((n, 3)), np.ones(n, dtype=np.int32), 0)
    for r in pos:
        st.add(p, r, [0, 0, 0], 1, 0, 1)
    return pos


@gpu
@pytest.mark.parametrize("kind,mode,T", [(SUPPLY_FE, 10, 293.15), (SUPPLY_GTF, 9, 1400.0)])
def test_supply_integral_matches_oracle_grid(orc, kind, mode, T):
    """The Cuba stand-in honours the reference's tolerance contract against the oracle's midpoint grid."""
    w = ((2.0, 2.5), (2.5, 2.0))
    sim, p, st, em = _planar_pair(orc, 11, w=w, mode=mode, T=T)
    with sim:
        for n_pre in (0, 150):
            if n_pre:
                _preload(sim, st, p, n_pre, 3)
            truth, _ = em.supply_grid(kind, 96)
            val, err, neval, fail = sim.Cuba_Integrate(kind)
            tol = max(0.5, 1e-3 * abs(truth))
            assert fail == 0 and neval >= 1000
            assert err <= tol
            assert abs(val - truth) <= 4.0 * tol, (val, truth, err)


@gpu
@pytest.mark.parametrize("kind,mode,T", [(SUPPLY_FE, 10, 293.15), (SUPPLY_GTF, 9, 1400.0)])
def test_supply_quadrature_on_the_device(orc, kind, mode, T):
    """rb2_planar_supply_level (MH_DEVICE runs): every level of the lattice rule -- nodes, cathode-plane field, integrand,
    per-shift sums -- on the device.  Same seed, same shifts: the same levels, evaluation count and (to the summation
    order) the same integral and error as the host loop over rb2_field_surface_z; and the oracle's grid as before."""
    w = ((2.0, 2.5), (2.5, 2.0))
    res = {}
    for mh_batch in (False, 2):
        sim, p, st, em = _planar_pair(orc, 11, w=w, mode=mode, T=T, mh_batch=mh_batch)
        with sim:
            out = []
            for n_pre in (0, 150):
                if n_pre:
                    _preload(sim, st, p, n_pre, 3)
                truth, _ = em.supply_grid(kind, 96)
                val, err, neval, fail = sim.Cuba_Integrate(kind)
                tol = max(0.5, 1e-3 * abs(truth))
                assert fail == 0 and err <= tol and abs(val - truth) <= 4.0 * tol, (mh_batch, val, truth, err)
                out.append((val, err, neval))
            res[mh_batch] = out
    for (v0, e0, n0), (v1, e1, n1) in zip(res[False], res[2]):
        assert n0 == n1
        assert v1 == pytest.approx(v0, rel=1e-12) and e1 == pytest.approx(e0, rel=1e-6, abs=1e-12 * abs(v0))


def _ks(a, b):
    from scipy.stats import ks_2samp
    return ks_2samp(a, b).pvalue


@gpu
def test_fe_samplers_match_oracle_distribution(orc):
    """Serial and lock-step chains (mod_field_emission_v2.F90:1122, :1284) sample the supply density:
    positions, surface fields and escape exponents follow the oracle's distributions."""
    w = ((2.0, 2.4), (2.4, 2.0))
    sim, p, st, em = _planar_pair(orc, 21, w=w)
    with sim:
        _preload(sim, st, p, 120, 5)
        M = 600
        df_b, F_b, pos_b = sim.Metropolis_Hastings_rectangle_J_batch(M)
        df_o, F_o, pos_o = em.mh_rectangle_J_batch(M)
        ser = [sim.Metropolis_Hastings_rectangle_J() for _ in range(150)]
        ser_o = [em.mh_rectangle_J() for _ in range(150)]
    assert np.all(F_b < 0) and np.all(F_o < 0)
    for k in (0, 1):
        assert _ks(pos_b[:, k], pos_o[:, k]) > 1e-3
        assert _ks(np.array([s_[3][k] for s_ in ser]), pos_o[:, k]) > 1e-3
        assert _ks(np.array([s_[3][k] for s_ in ser_o]), pos_o[:, k]) > 1e-3
    assert _ks(F_b, F_o) > 1e-3 and _ks(df_b, df_o) > 1e-3
    assert _ks(np.array([s_[2] for s_ in ser]), F_o) > 1e-3
    # the low work function cells (2.0 eV) must hold most of the supply-weighted samples in both
    in_low = lambda ps: np.mean(((ps[:, 0] < 0) & (ps[:, 1] > 0)) | ((ps[:, 0] > 0) & (ps[:, 1] < 0)))
    assert abs(in_low(pos_b) - in_low(pos_o)) < 0.1


@gpu
def test_device_resident_fe_sampler(orc):
    """mh_batch = 2 (rb2_mh_planar): the lock-step chains of mod_field_emission_v2.F90:1284-1458 run on the
    device; same distributions as the oracle's chains, reproducible for a given seed, adaptive step in range."""
    w = ((2.0, 2.4), (2.4, 2.0))
    sim, p, st, em = _planar_pair(orc, 22, w=w, mh_batch=2)
    emit = 100 * NM
    with sim:
        _preload(sim, st, p, 120, 5)
        M = 800
        df_d, F_d, pos_d = sim.Metropolis_Hastings_rectangle_J_batch(M)
        df_o, F_o, pos_o = em.mh_rectangle_J_batch(M)
        hp = rb.HotPath.attach()
        args = dict(emit_pos=(-0.5 * emit, -0.5 * emit), emit_dim=(emit, emit), w_theta=w, seed=987654321)
        r1 = hp.mh_planar(257, **args)
        r2 = hp.mh_planar(257, **args)
        r3 = hp.mh_planar(257, **{**args, "seed": 5})
    assert np.all(F_d < 0) and np.all(np.abs(pos_d[:, :2]) <= 0.5 * emit) and np.all(pos_d[:, 2] == 0)
    for k in (0, 1):
        assert _ks(pos_d[:, k], pos_o[:, k]) > 1e-3
    assert _ks(F_d, F_o) > 1e-3 and _ks(df_d, df_o) > 1e-3
    in_low = lambda ps: np.mean(((ps[:, 0] < 0) & (ps[:, 1] > 0)) | ((ps[:, 0] > 0) & (ps[:, 1] < 0)))
    assert abs(in_low(pos_d) - in_low(pos_o)) < 0.1
    for a, b in zip(r1[:3], r2[:3]):
        assert np.array_equal(a, b)                      # counter-based RNG: same seed, same chains
    assert r1[3:] == r2[3:] and not np.array_equal(r1[2], r3[2])
    assert 0.0 < r1[3] <= 1.0 and 0.00005 <= r1[4] <= 0.125
    # escape exponent consistent with the returned field and position (Escape_Prob_log, :568)
    k = 7
    col, row = int((r1[2][k, 0] / emit + 0.5) * 2), 1 - int((r1[2][k, 1] / emit + 0.5) * 2)
    assert r1[0][k] == pytest.approx(orc.fn_escape_prob_log(p, r1[1][k], w[row][col]), rel=1e-12)


@gpu
@pytest.mark.parametrize("n_pre,nic,kind,M", [(0, 1, 1, 24), (120, 1, 1, 24), (700, 0, 1, 24), (2500, 2, 1, 24), (300, 1, 2, 24),
                                              (0, 1, 1, 33), (120, 1, 1, 100), (2500, 1, 1, 128), (300, 1, 2, 70),
                                              (0, 1, 1, 129), (2500, 1, 1, 324), (700, 0, 1, 512), (300, 1, 2, 400), (9000, 2, 1, 200)])
def test_device_sampler_single_barrier_kernel(orc, n_pre, nic, kind, M):
    """Up to 512 chains run in k_mh_small (records resident in shared memory, chain state in registers, one barrier per
    jump; 4 warps per CTA up to 128 chains, 16 beyond).  Same generator keys and proposals as k_mh_persistent: for the same seed the chains coincide except where
    an accept decision sits on a rounding knife edge (the partial sums are joined in a different order)."""
    w = ((2.0, 2.4), (2.4, 2.0))
    mode, T = (10, 293.15) if kind == 1 else (9, 1200.0)
    sim, p, st, em = _planar_pair(orc, 41, nic=nic, w=w, mh_batch=2, mode=mode, T=T)
    emit = 100 * NM
    calls = 25 if M <= 32 else 8
    with sim:
        if n_pre:
            _preload(sim, st, p, n_pre, 9)
        hp = rb.HotPath.attach()
        args = dict(emit_pos=(-0.5 * emit, -0.5 * emit), emit_dim=(emit, emit), w_theta=w, kind=kind, T_temp=T)
        small = [hp.mh_planar(M, seed=1000 + k, **args) for k in range(calls)]
        again = hp.mh_planar(M, seed=1000, **args)
        one = hp.mh_planar(1, seed=77, **args)
        full = hp.mh_planar(32, seed=78, **args)
        hp.set_option("mh_small", 0)
        big = [hp.mh_planar(M, seed=1000 + k, **args) for k in range(calls)]
        hp.set_option("mh_small", 1)
    for a, b in zip(small[0][:3], again[:3]):
        assert np.array_equal(a, b)                       # reproducible
    assert small[0][3:] == again[3:]
    ps = np.concatenate([r[2] for r in small]); pb = np.concatenate([r[2] for r in big])
    Fs = np.concatenate([r[1] for r in small]); Fb = np.concatenate([r[1] for r in big])
    assert np.all(Fs < 0) and np.all(np.abs(ps[:, :2]) <= 0.5 * emit) and np.all(ps[:, 2] == 0)
    assert one[1].shape == (1,) and one[1][0] < 0 and full[1].shape == (32,) and np.all(full[1] < 0)
    same = np.all(np.abs(ps - pb) <= 1e-9 * emit, axis=1)
    assert same.mean() > 0.9, same.mean()
    assert np.allclose(Fs[same], Fb[same], rtol=1e-9)
    # adaptive step: same trajectory of MH_std when the chains coincide
    n_same_std = sum(abs(a[4] - b[4]) <= 1e-9 * b[4] for a, b in zip(small, big))
    assert n_same_std >= 0.6 * calls
    for k in (0, 1):
        assert _ks(ps[:, k], pb[:, k]) > 1e-3
    if kind == 1:
        assert 0.00005 <= small[0][4] <= 0.125
        k = 3
        col, row = int((small[0][2][k, 0] / emit + 0.5) * 2), 1 - int((small[0][2][k, 1] / emit + 0.5) * 2)
        assert small[0][0][k] == pytest.approx(orc.fn_escape_prob_log(p, small[0][1][k], w[row][col]), rel=1e-12)


@gpu
def test_device_resident_thermo_sampler(orc):
    w = ((2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0), (2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0))
    sim, p, st, em = _planar_pair(orc, 32, w=w, mode=9, T=1000.0, V=2000.0, d=1000 * NM, dt=1e-16, mh_batch=2)
    with sim:
        _preload(sim, st, p, 80, 6, d=1000 * NM)
        # MH_std adapts once per chain in the serial routine and once per jump iteration in lock-step: let
        # both settle (it grows from 1.25 % of the emitter side towards its cap) before comparing
        for _ in range(30):
            sim.Metropolis_Hastings_rectangle_J_thermo_batch(40)
        a, ok = sim.Metropolis_Hastings_rectangle_J_thermo_batch(600)
        b = np.array([em.mh_rectangle_J_thermo()[1] for _ in range(900)])[500:]
        assert sim.state().MH_std > 0.05
    assert np.all(ok == 1)
    for k in (0, 1):
        assert _ks(a[:, k], b[:, k]) > 1e-3
    cell = lambda ps: ((np.floor((ps[:, 0] / (100 * NM) + 0.5) * 4) + np.floor((ps[:, 1] / (100 * NM) + 0.5) * 4)) % 2)
    assert abs(np.mean(cell(a)) - np.mean(cell(b))) < 0.1


@gpu
def test_thermo_sampler_matches_oracle_distribution(orc):
    w = ((2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0), (2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0))
    sim, p, st, em = _planar_pair(orc, 31, w=w, mode=9, T=1000.0, V=2000.0, d=1000 * NM, dt=1e-16)
    with sim:
        _preload(sim, st, p, 80, 6, d=1000 * NM)
        a = np.array([sim.Metropolis_Hastings_rectangle_J_thermo()[1] for _ in range(400)])
        b = np.array([em.mh_rectangle_J_thermo()[1] for _ in range(400)])
    for k in (0, 1):
        assert _ks(a[:, k], b[:, k]) > 1e-3
    cell = lambda ps: ((np.floor((ps[:, 0] / (100 * NM) + 0.5) * 4) + np.floor((ps[:, 1] / (100 * NM) + 0.5) * 4)) % 2)
    assert abs(np.mean(cell(a)) - np.mean(cell(b))) < 0.1


def _tip_pair(orc, seed, V=1000.0, cap=20000, dt=0.25e-16, dims=(1000 * NM, 250 * NM, 500 * NM), box_z=1000 * NM, mh_batch=False):
    from oracle.oracle import Emission
    d_tip, R_base, h_tip = dims
    box = (0.0, 0.0, box_z)
    sim = Simulation(seed=seed, emission_mode=3, V_s=V, box_dim=box, time_step=dt, image_charge=True, N_ic_max=1,
                     emitters_pos=(0, 0, 0), emitters_dim=(d_tip, R_base, h_tip), emitters_type=1, max_particles=cap,
                     mh_batch=mh_batch)
    p = orc.params_tip(V, d_tip, R_base, h_tip, box, dt, True)
    st = orc.store(cap)
    em = Emission(orc, p, st, (0, 0, 0), (d_tip, R_base, h_tip), seed=seed + 1000)
    return sim, p, st, em


@gpu
def test_tip_supply_and_sampler_match_oracle(orc):
    sim, p, st, em = _tip_pair(orc, 41)
    with sim:
        n_s, F_avg = sim.Tip_Supply_Grid(100, 100)
        n_o, F_o = em.tip_supply_grid(100, 100)
        assert n_s == pytest.approx(n_o, rel=1e-9) and F_avg == pytest.approx(F_o, rel=1e-9)
        # space charge above the apex, then compare again and sample
        rng = np.random.default_rng(2)
        pos = np.stack([rng.uniform(-30, 30, 60), rng.uniform(-30, 30, 60), rng.uniform(503, 700, 60)], axis=1) * NM
        rb.HotPath.attach().Add_Particles(pos, np.zeros((60, 3)), np.ones(60, dtype=np.int32), 0)
        for r in pos:
            st.add(p, r, [0, 0, 0], 1, 0, 1)
        n_s, _ = sim.Tip_Supply_Grid(100, 100)
        n_o, _ = em.tip_supply_grid(100, 100)
        assert n_s == pytest.approx(n_o, rel=1e-9)
        a = np.array([sim.Metro_algo_tip_v3(80)[1:5] for _ in range(300)])
        b = np.array([em.metro_algo_tip_v3(80)[1:5] for _ in range(300)])
        eta_f_b, df_b, _ = sim.Metro_algo_tip_v3_batch(600, 80)   # lock-step variant (mh_batch)
    assert _ks(a[:, 0], b[:, 0]) > 1e-3          # xi
    assert _ks(a[:, 2], b[:, 2]) > 1e-3          # normal field
    assert _ks(a[:, 3], b[:, 3]) > 1e-3          # escape probability
    assert _ks(eta_f_b, b[:, 2]) > 1e-3 and _ks(df_b, b[:, 3]) > 1e-3


@gpu
def test_tip_supply_grid_on_the_device(orc):
    """rb2_tip_supply (MH_DEVICE runs): the 100 x 100 supply sum of Do_Field_Emission_Tip_OLDCODE with the grid resident on
    the device -- against the oracle's serial double loop, against the host loop over an rb2_field_batch of the same
    nodes, without and with space charge, and again after the particle set changed (the grid stays, the field does not)."""
    sim_h, p, st, em = _tip_pair(orc, 41)
    with sim_h:
        host0 = sim_h.Tip_Supply_Grid(100, 100)
    sim, p, st, em = _tip_pair(orc, 41, mh_batch=2)
    with sim:
        dev0 = sim.Tip_Supply_Grid(100, 100)
        n_o, F_o = em.tip_supply_grid(100, 100)
        assert dev0[0] == pytest.approx(n_o, rel=1e-9) and dev0[1] == pytest.approx(F_o, rel=1e-9)
        assert dev0[0] == pytest.approx(host0[0], rel=1e-12) and dev0[1] == pytest.approx(host0[1], rel=1e-12)
        rng = np.random.default_rng(2)
        pos = np.stack([rng.uniform(-30, 30, 60), rng.uniform(-30, 30, 60), rng.uniform(503, 700, 60)], axis=1) * NM
        hp = rb.HotPath.attach()
        hp.Add_Particles(pos, np.zeros((60, 3)), np.ones(60, dtype=np.int32), 0)
        for r in pos:
            st.add(p, r, [0, 0, 0], 1, 0, 1)
        dev1 = sim.Tip_Supply_Grid(100, 100)
        n_o1, F_o1 = em.tip_supply_grid(100, 100)
        assert dev1[0] == pytest.approx(n_o1, rel=1e-9) and dev1[1] == pytest.approx(F_o1, rel=1e-9)
        assert dev1[0] < dev0[0]                      # the space charge screens the apex
        assert sim.Tip_Supply_Grid(100, 100) == dev1  # deterministic
        # the raw entry points with a grid of our own: 7 nodes, unit weights, normals along z
        nodes = pos[:7] * [0.2, 0.2, 0.0] + [0, 0, 600 * NM]
        nrm = np.tile([0.0, 0.0, 1.0], (7, 1))
        hp.tip_supply_set_grid(nodes, nrm, np.ones(7))
        n_s, F_sum = hp.tip_supply()
        fz = hp.Calc_Field_at_Batch(nodes)[:, 2]
        assert F_sum == pytest.approx(fz.sum(), rel=1e-13)
        assert (n_s > 0.0) == bool(np.any(fz < 0.0))


@gpu
def test_device_resident_tip_sampler(orc):
    """rb2_mh_tip: the lock-step tip chains with every jump queued on the device (proposal kernel, the tip field kernel
    of rb2_field_batch, accept kernel).  Same distributions as the oracle's serial chains and as the host lock-step
    variant, reproducible for a seed, positions on the tip surface, escape probability consistent with the field."""
    sim, p, st, em = _tip_pair(orc, 43, mh_batch=True)
    with sim:
        rng = np.random.default_rng(2)
        pos = np.stack([rng.uniform(-30, 30, 60), rng.uniform(-30, 30, 60), rng.uniform(503, 700, 60)], axis=1) * NM
        hp = rb.HotPath.attach()
        hp.Add_Particles(pos, np.zeros((60, 3)), np.ones(60, dtype=np.int32), 0)
        for r in pos:
            st.add(p, r, [0, 0, 0], 1, 0, 1)
        b = np.array([em.metro_algo_tip_v3(80)[1:5] for _ in range(300)])         # xi, phi, eta_f, df
        eta_h, df_h, pos_h = sim.Metro_algo_tip_v3_batch(600, 80)                 # host lock-step loop
        eta_d, df_d, pos_d, a_rate, mh_std = hp.mh_tip(600, seed=12345)
        again = hp.mh_tip(600, seed=12345)
        other = hp.mh_tip(600, seed=6)
        few = hp.mh_tip(3, seed=1)
        # a sharper look at the means than KS on 600 samples gives: 4000 chains each way
        eta_h4, df_h4, _ = sim.Metro_algo_tip_v3_batch(4000, 80)
        eta_d4, df_d4, _, _, _ = hp.mh_tip(4000, seed=99)
    assert np.all(eta_d < 0) and np.all((df_d > 0) & (df_d <= 1))
    assert _ks(eta_d, b[:, 2]) > 1e-3 and _ks(df_d, b[:, 3]) > 1e-3              # against the oracle's serial chains
    assert _ks(eta_d, eta_h) > 1e-3 and _ks(df_d, df_h) > 1e-3                    # against the host lock-step loop
    assert _ks(np.hypot(pos_d[:, 0], pos_d[:, 1]), np.hypot(pos_h[:, 0], pos_h[:, 1])) > 1e-3
    for x, y in zip(again[:3], (eta_d, df_d, pos_d)):
        assert np.array_equal(x, y)
    assert again[3:] == (a_rate, mh_std) and not np.array_equal(other[2], pos_d)
    assert 0.0 < a_rate <= 1.0 and 0.0005 <= mh_std <= 0.125
    assert few[0].shape == (3,) and np.all(few[0] < 0)
    for a4, b4 in ((eta_d4, eta_h4), (df_d4, df_h4)):
        se = math.sqrt(a4.var() / a4.size + b4.var() / b4.size)
        assert abs(a4.mean() - b4.mean()) < 4.5 * se, (a4.mean(), b4.mean(), se)
    # positions lie on the hyperboloid eta = eta_1 (src/mod_hyperboloid_tip.f90:36-76) and the escape probability is
    # Escape_Prob_Tip of the returned field
    for k in (0, 17, 599):
        eta = orc.lib.orc_eta_coor(C.byref(p), pos_d[k, 0], pos_d[k, 1], pos_d[k, 2])
        assert eta == pytest.approx(p.eta_1, rel=1e-9)
        assert df_d[k] == pytest.approx(orc.tip_escape_prob(p, eta_d[k], 4.7), rel=1e-10)


M32 = 0xFFFFFFFF


def _philox(ctr, key):
    """Philox4x32-10, the generator of the device samplers (rb2_mh.cu)."""
    ctr, key = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = 0xD2511F53 * ctr[0], 0xCD9E8D57 * ctr[2]
        ctr = [((p1 >> 32) ^ ctr[1] ^ key[0]) & M32, p1 & M32, ((p0 >> 32) ^ ctr[3] ^ key[1]) & M32, p0 & M32]
        key = [(key[0] + 0x9E3779B9) & M32, (key[1] + 0xBB67AE85) & M32]
    return ctr


def _rand2(seed, chain, it, purpose, attempt):
    key = [(seed & M32) ^ ((chain * 0x9E3779B1) & M32), ((seed >> 32) + chain) & M32]
    r = _philox([it & M32, purpose, attempt, chain], key)
    u53 = lambda hi, lo: ((hi >> 5) * 67108864.0 + (lo >> 6)) * (1.0 / 9007199254740992.0)
    return u53(r[0], r[1]), u53(r[2], r[3])


@gpu
@pytest.mark.parametrize("n_pre,w", [(0, ((2.0,),)), (300, ((2.0, 3.2), (3.2, 2.0)))])
def test_serial_sampler_kernel_follows_the_serial_algorithm(orc, n_pre, w):
    """rb2_mh_planar_serial (the reference's default mh_batch = .false. semantics in one kernel) against a line-by-line
    host replay of Metropolis_Hastings_rectangle_J inside the insert loop (mod_field_emission_v2.F90:1122-1265,
    :322-380) that uses the SAME counter-based random numbers and one rb2_field_surface_z call per jump, with every
    emitted electron inserted into the store before the next chain starts: same emission decisions, same positions
    (to rounding: libm vs CUDA log / sqrt), same step adaptation.  Chain s therefore sees the electrons of chains < s."""
    emit, d, V = 60 * NM, 500 * NM, 3000.0
    sim, p, st, em = _planar_pair(orc, 3, w=w, V=V, d=d, emit=emit, dt=1e-16)
    epos, edim = (-0.5 * emit, -0.5 * emit), (emit, emit)
    M, ndim, nfirst, seed = 16, 40, 10, 987654321
    warr = np.atleast_2d(np.asarray(w, dtype=float))
    h_bar = 6.62607015e-34 / (2.0 * math.pi)
    b_FN = -4.0 / (3.0 * h_bar) * math.sqrt(2.0 * rb.api.M_0 * rb.api.Q_0)
    eps0 = 1.0 / (1.25663706212e-6 * 299792458.0 ** 2)
    l_const = rb.api.Q_0 / (4.0 * math.pi * eps0)

    def w_at(x, y):
        return sim.w_theta_xy(np.array([x, y, 0.0]))[0]

    def fn_l(F, wv):
        return min(1.0, l_const * (-F) / (wv * wv))

    def target(F, wv):
        l = fn_l(F, wv)
        return 2.0 * math.log(-F) - 2.0 * math.log(1.0 + l * (1.0 / 9.0 - math.log(l) / 18.0)) - math.log(wv)

    with sim:
        hp = rb.HotPath.attach()
        if n_pre:
            _preload(sim, st, p, n_pre, 8, emit=emit, d=d)
        df_d, F_d, pos_d, em_d, ar_d, sd_d = hp.mh_planar_serial(M, epos, edim, warr, seed, ndim=ndim, ndim_first=nfirst)
        again = hp.mh_planar_serial(M, epos, edim, warr, seed, ndim=ndim, ndim_first=nfirst)
        # host replay
        field = lambda x, y: float(hp.field_surface_z(np.array([[x, y, 0.0]]))[0])
        mh_std, a_rate = 0.0125, 1.0
        out = []
        for s in range(M):
            ok, rnd = False, 0
            while not ok and rnd < 10000:
                u, v = _rand2(seed, s, -(rnd + 1), 0, 0)
                cx, cy = u * edim[0] + epos[0], v * edim[1] + epos[1]
                Fc = field(cx, cy)
                ok = Fc < 0.0
                rnd += 1
            assert ok
            sup = target(Fc, w_at(cx, cy))
            ja = jr = 0
            for it in range(1, ndim + 1):
                for attempt in range(64):
                    u, v = _rand2(seed, s, it, 1, attempt)
                    a, b = 2.0 * u - 1.0, 2.0 * v - 1.0
                    ww = a * a + b * b
                    if 0.0 < ww < 1.0:
                        f = math.sqrt((-2.0 * math.log(ww)) / ww)
                        g0, g1 = a * f, b * f
                        break
                frac = mh_std if it > nfirst else 0.10
                qx, qy = cx + g0 * (edim[0] * frac), cy + g1 * (edim[1] * frac)
                x_min, x_max, y_min, y_max = epos[0], epos[0] + edim[0], epos[1], epos[1] + edim[1]
                if qx > x_max: qx = x_max - (qx - x_max)
                elif qx < x_min: qx = (x_min - qx) + x_min
                if qy > y_max: qy = y_max - (qy - y_max)
                elif qy < y_min: qy = (y_min - qy) + y_min
                Fz = field(qx, qy)
                acc = False
                if Fz < 0.0:
                    sup_new = target(Fz, w_at(qx, qy))
                    acc = sup_new >= sup or math.log(_rand2(seed, s, it, 2, 0)[0]) <= sup_new - sup
                    if acc:
                        cx, cy, sup, Fc = qx, qy, sup_new, Fz
                if it > nfirst:
                    ja, jr = ja + acc, jr + (not acc)
            if ja + jr > 0:
                a_rate = ja / (ja + jr)
                mh_std = min(max(mh_std * math.exp(0.025 * (a_rate - 0.35)), 0.00005), 0.125)
            wv = w_at(cx, cy)
            l = fn_l(Fc, wv)
            D_f = b_FN * math.sqrt(wv) ** 3 * (1.0 - l + l * math.log(l) / 6.0) / (-Fc)
            emitted = math.log(_rand2(seed, s, ndim + 1, 3, 0)[0]) <= D_f
            out.append((cx, cy, Fc, D_f, emitted))
            if emitted:  # Add_Particle at z = 1 nm: the next chains see it
                hp.Add_Particle([cx, cy, 1.0 * NM], [0.0, 0.0, 0.0], 1, 1, 1)
    for x, y in zip(again[:4], (df_d, F_d, pos_d, em_d)):
        assert np.array_equal(x, y)
    ref = np.array([(o[0], o[1], o[2], o[3]) for o in out])
    assert [bool(e) for e in em_d] == [o[4] for o in out]
    assert int(em_d.sum()) >= 1                          # the chains behind an emitted one ran against a same-step electron
    assert np.allclose(pos_d[:, 0], ref[:, 0], rtol=0, atol=1e-9 * emit) and np.allclose(pos_d[:, 1], ref[:, 1], rtol=0, atol=1e-9 * emit)
    assert np.allclose(F_d, ref[:, 2], rtol=1e-9) and np.allclose(df_d, ref[:, 3], rtol=1e-9)
    assert sd_d == pytest.approx(mh_std, rel=1e-12) and ar_d == pytest.approx(a_rate, rel=1e-12)


@gpu
@pytest.mark.parametrize("M,n_pre", [(100, 60), (214, 60), (214, 1500), (512, 700), (33, 0)])
def test_tip_sampler_persistent_kernel(orc, M, n_pre):
    """k_mh_tip_small: the tip chains of a time step as ONE persistent kernel (<= 512 chains): the same proposals,
    targets and generator keys as the three-launches-per-jump path (option mh_small = 0), so for one seed almost every
    chain ends in the same spot (the field sums differ in rounding only: a knife-edge accept decision may flip); same
    shared-step adaptation; distributions equal to the oracle's serial chains."""
    sim, p, st, em = _tip_pair(orc, 47, mh_batch=True)
    with sim:
        hp = rb.HotPath.attach()
        if n_pre:
            rng = np.random.default_rng(3)
            pos = np.stack([rng.uniform(-40, 40, n_pre), rng.uniform(-40, 40, n_pre), rng.uniform(503, 900, n_pre)], axis=1) * NM
            hp.Add_Particles(pos, np.zeros((n_pre, 3)), np.ones(n_pre, dtype=np.int32), 0)
            if n_pre <= 100:
                for r in pos:
                    st.add(p, r, [0, 0, 0], 1, 0, 1)
        l0 = hp.launch_count()
        eta_s, df_s, pos_s, a_s, std_s = hp.mh_tip(M, seed=777)
        l1 = hp.launch_count()
        again = hp.mh_tip(M, seed=777)
        hp.set_option("mh_small", 0)
        eta_l, df_l, pos_l, a_l, std_l = hp.mh_tip(M, seed=777)
        l2 = hp.launch_count()
        hp.set_option("mh_small", 1)
        if n_pre <= 100:
            b = np.array([em.metro_algo_tip_v3(80)[1:5] for _ in range(300)])
    assert l1 - l0 == 1 and l2 - l1 > 200          # one kernel against three launches per jump
    for x, y in zip(again[:3], (eta_s, df_s, pos_s)):
        assert np.array_equal(x, y)
    same = np.all(pos_s == pos_l, axis=1) | (np.linalg.norm(pos_s - pos_l, axis=1) < 1e-15)
    assert same.mean() > 0.9, same.mean()
    assert std_s == pytest.approx(std_l, rel=0.15) and a_s == pytest.approx(a_l, abs=0.1)
    assert np.all(eta_s < 0) and np.all((df_s > 0) & (df_s <= 1))
    assert np.allclose(eta_s[same], eta_l[same], rtol=1e-9)
    if n_pre <= 100 and M >= 100:
        assert _ks(eta_s, b[:, 2]) > 1e-3 and _ks(df_s, b[:, 3]) > 1e-3


# ---------------------------------------------------------------------------------------------------------
# GPU: whole-system runs
@gpu
@pytest.mark.parametrize("mh_batch", [-1, False, True, 2])
def test_planar_system_reference_test(orc, mh_batch):
    """mod_tests.F90:2013-2125 (Test_Planar_System): 250 steps, 1 kV over 500 nm, 100 x 100 nm emitter,
    2.0 eV, N_ic_max = 0.  Same assertions as the reference, plus agreement with an oracle run."""
    n_steps, d, V, dt = 250, 500 * NM, 1000.0, 0.25e-15
    sim, p, st, em = _planar_pair(orc, 12345, nic=0, mh_batch=mh_batch)
    q0 = rb.api.Q_0
    with sim:
        Q_ramo = Q_ss = 0.0
        n_peak = 0
        for i in range(1, n_steps + 1):
            s = sim.step(i)
            Q_ramo += s.ramo_current[1] * dt
            if i > n_steps // 2:
                Q_ss += s.ramo_current[1] * dt
            n_peak = max(n_peak, s.nrElec)
        s = sim.state()
        pos = rb.HotPath.attach().download(("pos",))["pos"]
    n_emit, n_top, n_bot = s.nrEmitted_total, s.nrAbsorbed_top, s.nrAbsorbed_bot
    assert n_emit > 100 and n_top > 50
    assert s.nrElec == n_emit - n_top - n_bot
    assert np.all(pos[:, 2] >= 0.0) and np.all(pos[:, 2] <= d) and np.all(np.abs(pos) < 1.0)
    z_emit = 1.0 * NM
    Q_exp = q0 * (n_top * (d - z_emit) - n_bot * z_emit + np.sum(pos[:, 2] - z_emit)) / d
    assert abs(Q_ramo - Q_exp) < 0.10 * abs(Q_exp)
    I_ss = Q_ss / (dt * (n_steps - n_steps // 2))
    assert 0.5e-3 < I_ss < 4.0e-3
    # the oracle's run of the same system (CPU, its own RNG): same steady-state current within noise
    Qo = 0.0
    for i in range(1, n_steps + 1):
        N_sup, _ = em.supply_grid(SUPPLY_FE, 16)
        em.do_field_emission_planar(i, N_sup, mh_batch in (True, 2))
        st.step(p)
        if i > n_steps // 2:
            Qo += st.s.ramo_current[1] * dt
        st.remove(i)
    I_o = Qo / (dt * (n_steps - n_steps // 2))
    assert 0.5e-3 < I_o < 4.0e-3
    assert abs(I_ss - I_o) < 0.15 * I_o, (I_ss, I_o)


@gpu
@pytest.mark.parametrize("mh_batch", [True, 2])
def test_tip_system_reference_test(orc, mh_batch):
    """mod_tests.F90:2131-2219 (Test_Tip_System) shape: emission from the tip reaches the anode and the
    bookkeeping closes; the emitted count tracks the oracle's."""
    # d = 1000 nm gap, tip 900/100/100 nm, 800 V, dt = 0.25 fs, 350 steps -- the reference's own numbers
    # (the lock-step sampler is used for the 350 steps: the serial one costs ~500 x 81 single-point field
    # calls per step here; their equivalence is covered by test_tip_supply_and_sampler_match_oracle)
    sim, p, st, em = _tip_pair(orc, 777, V=800.0, dt=0.25e-15, dims=(900 * NM, 100 * NM, 100 * NM), box_z=1000 * NM,
                               mh_batch=mh_batch)
    n_steps = 350
    q0 = rb.api.Q_0
    with sim:
        Q_ramo = 0.0
        emitted_100 = 0
        for i in range(1, n_steps + 1):
            s = sim.step(i)
            Q_ramo += s.ramo_current[1] * 0.25e-15
            if i == 100:
                emitted_100 = s.nrEmitted_total
        s = sim.state()
        pos = rb.HotPath.attach().download(("pos",))["pos"]
    n_top = s.nrAbsorbed_top
    assert s.nrEmitted_total > 20 and n_top > 5
    assert s.nrElec == s.nrEmitted_total - s.nrAbsorbed_top - s.nrAbsorbed_bot
    assert np.all(pos[:, 2] >= 0.0) and np.all(pos[:, 2] <= 1000 * NM) and np.all(np.abs(pos) < 1.0)
    transits = Q_ramo / q0
    assert transits > 0.8 * n_top - 2.0          # Shockley-Ramo lower bound
    assert transits < 1.0 * (n_top + s.nrElec) + 2.0
    for i in range(1, 101):  # the oracle's serial CPU run of the first 100 steps
        n_s, _ = em.tip_supply_grid(40, 40)
        em.do_field_emission_tip(i, n_s)
        st.step(p)
        st.remove(i)
    emitted_o = st.s.nrID
    assert abs(emitted_100 - emitted_o) < 5.0 * math.sqrt(emitted_o) + 0.15 * emitted_o, (emitted_100, emitted_o)


@gpu
@pytest.mark.parametrize("mh_batch", [-1, False, 2])
def test_thermo_field_system(orc, mh_batch):
    w = ((2.0, 2.5), (2.5, 2.0))
    sim, p, st, em = _planar_pair(orc, 99, w=w, mode=9, T=1000.0, V=2000.0, d=1000 * NM, dt=1e-16, mh_batch=mh_batch)
    n_steps = 60
    with sim:
        for i in range(1, n_steps + 1):
            s = sim.step(i)
        s = sim.state()
        vel = rb.HotPath.attach().download(("vel", "section"))
    assert s.nrEmitted_total > 50
    assert s.nrElec == s.nrEmitted_total - s.nrAbsorbed_top - s.nrAbsorbed_bot
    for i in range(1, n_steps + 1):
        N_sup, _ = em.supply_grid(SUPPLY_GTF, 16)
        em.do_field_thermo_emission_planar(i, N_sup)
        st.step(p)
        st.remove(i)
    assert abs(s.nrEmitted_total - st.s.nrID) < 5.0 * math.sqrt(st.s.nrID) + 0.1 * st.s.nrID
    # every electron carries a valid section of the 2 x 2 board and a Maxwell-Boltzmann launch velocity
    # with v_z >= 0 (mod_velocity.f90:53-68).  (With MH_std starting at 1.25 % of the emitter and 25 jumps
    # the reference's chains stay close to their uniform start, so the sections are NOT supply weighted.)
    sec = vel["section"]
    assert set(np.unique(sec)) <= {1, 2, 3, 4}
    v = vel["vel"]
    sd = math.sqrt(1.380649e-23 * 1000.0 / rb.api.M_0)
    assert np.all(np.abs(v[:, :2]) < 8 * sd + 1e5)


@gpu
def test_photo_emission_space_charge_limit(orc):
    """mod_photo_emission.f90:603-686: emission per step stops at the space-charge limit; the batched
    speculative evaluation reproduces the serial decisions, so counts track the oracle's serial loop."""
    laser = dict(gauss_mode=2, laser_mode=1, photon_mode=2, energy=4.7)
    sim, p, st, em = _planar_pair(orc, 5, w=((4.5,),), mode=1, V=2.0, d=1000 * NM, emit=200 * NM, dt=1e-16, laser=laser)
    counts, counts_o = [], []
    with sim:
        for i in range(1, 6):
            counts.append(sim.Do_Emission(i))
        s = sim.state()
    for i in range(1, 6):
        counts_o.append(em.do_photo_emission_rectangle(i, 4.7, 2, -1))
    assert counts[0] > 5 and counts_o[0] > 5
    assert abs(counts[0] - counts_o[0]) < 0.35 * counts_o[0] + 5
    assert sum(counts[1:]) < counts[0]  # later steps only top up what the first one left
    assert abs(sum(counts) - sum(counts_o)) < 0.3 * sum(counts_o) + 5


@gpu
@pytest.mark.parametrize("w,n_pre", [(((2.5,),), 0), (((2.5, 5.0), (5.0, 2.5)), 400)])
def test_photo_loop_speculative_batches_equal_serial_loop(orc, w, n_pre):
    """mod_photo_emission.f90:603-686 is an accept-and-insert loop in which every accepted electron changes the field
    the next attempt sees.  The device path evaluates batches of upcoming attempts speculatively (rb2_field_batch_delta
    against store + electrons accepted so far) and re-evaluates what an acceptance made stale; this must make EXACTLY
    the accept / insert sequence of the literal loop (one Calc_Field_at per probe, immediate Add_Particle) on the same
    random stream: same counts per step and bit-identical particle arrays after several full time steps."""
    laser = dict(gauss_mode=2, laser_mode=2, photon_mode=2, energy=4.7, variation=0.02)
    runs = []
    for serial in (1, 0):
        sim, p, st, em = _planar_pair(orc, 77, w=w, mode=1, V=2000.0, d=1000 * NM, emit=100 * NM, dt=1e-16, laser=laser)
        sim.set_option("photo_serial", serial)
        with sim:
            if n_pre:
                _preload(sim, st, p, n_pre, 5, d=1000 * NM)
            counts = []
            for i in range(1, 7):
                counts.append(sim.step(i).nrElecEmit)
            s = sim.state()
            d = rb.HotPath.attach().download(("pos", "vel", "acc", "id", "step", "section"))
        runs.append((counts, s.nrElec, d))
    (c1, n1, d1), (c0, n0, d0) = runs
    assert c1 == c0 and n1 == n0 and c1[0] > 500
    for key in d1:
        assert np.array_equal(d1[key], d0[key]), key
    if len(w) == 2:  # no electron starts above a 5.0 eV cell (photon energy 4.7 eV; the first row of `work` is the top row)
        new = d1["pos"][d1["step"] > 0]
        assert np.all((new[:, 0] < 0) != (new[:, 1] < 0))


@gpu
def test_full_store_drops_particles_without_records(orc, tmp_path):
    """Add_Particle at MAX_PARTICLES (src/mod_pair.F90:36-43): the particle is dropped and counted, nothing is written
    and nrID does not advance -- density_emit*.bin hold exactly the accepted particles."""
    d = _deck(tmp_path, "gpu_planar_fe")
    cap = 400
    with Simulation(d, write_files=True, seed=3, max_particles=cap) as sim:
        for i in range(1, 16):
            s = sim.step(i)
        assert s.nrPart == cap and s.nrID == cap and s.nrElec == cap
        k = rb.HotPath.attach().counts()
        assert k.nrPart_dropped > 0 and k.nrID == cap
    out = tmp_path / "out"
    assert (out / "density_emit.bin").stat().st_size == cap * 40
    assert (out / "density_emit_elec.bin").stat().st_size == cap * 32
    assert (out / "density_emit_ion.bin").stat().st_size == 0
    init = np.fromfile(out / "init.bin", dtype=np.dtype([("d", "<f8", 7), ("i", "<i4", 4)]))
    assert init.nbytes == 72 and list(init["i"][0]) == [5000000, 1, 9216, 1000]
    assert list(init["d"][0]) == [1.0, 1.0, 1.0, 1e-9, 1e-12, 1e-9 / 1e-12, 1.0]


def _deck(tmp_path, name, edits=(), extra=""):
    """Copy tools/decks/<name> to a scratch dir; edits = (old, new) replacements in `input`, extra = namelist lines."""
    import shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tools", "decks", name)
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path / f)
    txt = (tmp_path / "input").read_text()
    for old, new in edits:
        assert old in txt
        txt = txt.replace(old, new)
    if extra:
        txt = txt.replace("\n/", "\n  " + extra + "\n/")
    (tmp_path / "input").write_text(txt)
    return str(tmp_path)


def _golden_runs():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deck_oracle_runs.json")) as f:
        return json.load(f)


@gpu
def test_photo_deck_small_emitter_vs_oracle_run(orc, tmp_path):
    """tools/decks/photo (= Examples/Photo) with the emitter shrunk to 100 x 100 nm so that the CPU oracle can run the
    same 200 steps live (3 seeds): the first-step burst (space-charge limit of the bare cathode), the emitted total and
    the current of the last 50 steps agree within a few per cent; exact electron bookkeeping."""
    from oracle.oracle import Emission
    d = _deck(tmp_path, "photo", edits=[("-250.0d0, -250.0d0", "-50.0d0, -50.0d0"), ("500.0d0, 500.0d0, 0.0d0", "100.0d0, 100.0d0, 0.0d0")])
    steps = 200
    got = []
    for seed in (11, 12):
        with Simulation(d, seed=seed, max_particles=50000) as sim:
            em0 = sim.step(1).nrElecEmit
            cur = []
            for i in range(2, steps + 1):
                cur.append(sim.step(i).ramo_total)
            s = sim.state()
            assert s.nrElec == s.nrEmitted_total - s.nrAbsorbed_top - s.nrAbsorbed_bot
            got.append((em0, s.nrEmitted_total, float(np.mean(cur[-50:]))))
    ref = []
    for seed in (21, 22, 23):
        box = (100 * NM, 100 * NM, 1000 * NM)
        p = orc.params_planar(2000.0, 1000 * NM, box, 1e-16, True, 1)
        st = orc.store(50000)
        em = Emission(orc, p, st, (-50 * NM, -50 * NM, 0), (100 * NM, 100 * NM, 0), ((2.5,),), seed=seed)
        tot, cur, first = 0, [], None
        for i in range(1, steps + 1):
            n = em.do_photo_emission_rectangle(i, em.get_laser_energy(4.7, 0.02), 2, -1)
            first = n if first is None else first
            tot += n
            st.step(p)
            cur.append(st.s.ramo_current[1])
            st.remove(i)
        ref.append((first, tot, float(np.mean(cur[-50:]))))
    g, r = np.mean(got, axis=0), np.mean(ref, axis=0)
    assert abs(g[0] - r[0]) < 0.02 * r[0], (got, ref)
    assert abs(g[1] - r[1]) < 0.02 * r[1], (got, ref)
    assert abs(g[2] - r[2]) < 0.05 * abs(r[2]), (got, ref)


@gpu
def test_photo_deck_full_size_vs_oracle_fixture(tmp_path):
    """tools/decks/photo as shipped (500 x 500 nm emitter: ~45 000 electrons leave in the first step) for 200 steps
    against the CPU oracle's own run of the same deck (tests/golden/deck_oracle_runs.json, made by
    tests/golden/make_deck_fixtures.py: ~10 CPU-minutes)."""
    fx = _golden_runs()["photo_full"][0]
    d = _deck(tmp_path, "photo")
    with Simulation(d, seed=31, max_particles=200000) as sim:
        em, cur = [], []
        for i in range(1, fx["steps"] + 1):
            s = sim.step(i)
            em.append(s.nrElecEmit)
            cur.append(s.ramo_total)
        s = sim.state()
        assert s.nrElec == s.nrEmitted_total - s.nrAbsorbed_top - s.nrAbsorbed_bot
    # the burst stops at 100 consecutive failures -- a random stopping rule: two oracle seeds differ by 1 % here
    assert abs(em[0] - fx["emitted"][0]) < 0.03 * fx["emitted"][0], (em[0], fx["emitted"][0])
    assert abs(sum(em) - sum(fx["emitted"])) < 0.03 * sum(fx["emitted"])
    for a, b in ((100, 120), (180, 200)):
        assert np.mean(cur[a:b]) == pytest.approx(np.mean(fx["ramo"][a:b]), rel=0.04)
    assert s.nrElec == pytest.approx(fx["nrPart"][-1], rel=0.03)


@gpu
def test_checkerboard_tfe_deck_vs_oracle_run(orc, tmp_path):
    """tools/decks/checkerboard_tfe (= Examples/Checkerboard-TFE) as shipped + WRITE_RAMO_SEC, 700 steps (no electron
    has crossed the gap yet), 6 seeds, against the CPU oracle's live runs of the same deck (6 seeds): emitted count, current, share of the 16 sections
    among the electrons (chi-square) and in the per-section Ramo current; per step the sections add up to the total
    and ramo_current.bin holds (MAX_SECTIONS, MAX_EMITTERS) doubles per step (src/mod_pair.F90:822-826)."""
    from oracle.oracle import Emission
    from scipy import stats
    steps, nseed = 700, 6
    d = _deck(tmp_path, "checkerboard_tfe", extra="WRITE_RAMO_SEC = .True.,")
    g_emit, g_cur, g_sec, g_share = [], [], np.zeros(16), np.zeros(16)
    for seed in range(41, 41 + nseed):
        with Simulation(d, write_files=(seed == 41), seed=seed, max_particles=50000) as sim:
            cur = []
            for i in range(1, steps + 1):
                s = sim.step(i)
                cur.append(s.ramo_total)
                sec_now = sim.ramo_current_emit(16)
                assert sec_now.sum() == pytest.approx(s.ramo_total, rel=1e-10, abs=1e-30)
                g_share += sec_now
            s = sim.state()
            st_g = rb.HotPath.attach().download(("section", "vel", "charge"))
            # the table of the last step against the particle arrays (planar: E_zunit = -1/d)
            want = np.bincount(st_g["section"], weights=st_g["charge"] * st_g["vel"][:, 2] * (-1.0 / (1000 * NM)), minlength=17)[1:17]
            assert np.allclose(sec_now, want, rtol=1e-11, atol=1e-25)
            assert s.nrAbsorbed_top == 0
            g_emit.append(s.nrEmitted_total)
            g_cur.append(float(np.mean(cur[-200:])))
            g_sec += np.bincount(st_g["section"], minlength=17)[1:17]
        if seed == 41:
            out = tmp_path / "out"
            assert (out / "ramo_current.bin").stat().st_size == steps * 96 * 96 * 1 * 8
            last = np.fromfile(out / "ramo_current.bin", dtype="<f8").reshape(steps, 96 * 96)[-1]
            assert np.array_equal(last[:16], sec_now) and np.all(last[16:] == 0.0)
    w = ((2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0), (2.0, 2.5, 2.0, 2.5), (2.5, 2.0, 2.5, 2.0))
    o_emit, o_cur, o_sec, o_share = [], [], np.zeros(16), np.zeros(16)
    for seed in range(51, 51 + nseed):
        box = (100 * NM, 100 * NM, 1000 * NM)
        p = orc.params_planar(2000.0, 1000 * NM, box, 1e-16, True, 1)
        st = orc.store(50000)
        em = Emission(orc, p, st, (-50 * NM, -50 * NM, 0), (100 * NM, 100 * NM, 0), w, T_temp=1000.0, seed=seed)
        cur = []
        for i in range(1, steps + 1):
            N_sup, _ = em.supply_grid(SUPPLY_GTF, 16)
            em.do_field_thermo_emission_planar(i, N_sup)
            st.step(p)
            cur.append(st.s.ramo_current[1])
            o_share += st.ramo_current_emit(16)
            st.remove(i)
        o_emit.append(st.s.nrID)
        o_cur.append(float(np.mean(cur[-200:])))
        o_sec += np.bincount(st.section, minlength=17)[1:17]
    ge, oe = sum(g_emit), sum(o_emit)
    assert abs(ge - oe) < 4.0 * math.sqrt(ge + oe), (g_emit, o_emit)           # Poisson candidates: 4 sigma of the difference
    assert np.mean(g_cur) == pytest.approx(np.mean(o_cur), rel=0.10), (g_cur, o_cur)
    chi2, pval, _, _ = stats.chi2_contingency(np.stack([g_sec, o_sec]))
    assert pval > 1e-3, (g_sec, o_sec, pval)
    assert np.all(g_sec > 0) and g_sec.sum() == ge
    gs, os_ = g_share / g_share.sum(), o_share / o_share.sum()
    assert np.max(np.abs(gs - os_)) < 0.35 * np.max(os_), (gs, os_)


@gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rumdeed_b200", "rumdeed_b200_run")),
                    reason="driver executable not built")
def test_driver_executable_writes_reference_format_files(tmp_path):
    """program RUMDEED on a small deck: the output files parse with the record layouts of the reference's
    scripts/python_package/rumdeed_io.py (density_emit: 3 f64 + 4 i32; absorb_top / planes: 5 f64 + 3 i32)."""
    import subprocess
    (tmp_path / "input").write_text(
        "&INPUT\n  V_S = 1.0d3,\n  BOX_DIM = 0.0d0, 0.0d0, 500.0d0,\n  TIME_STEP = 0.25d-3,\n  STEPS = 400,\n  EMISSION_MODE = 10,\n"
        "  NREMIT = 1,\n  IMAGE_CHARGE = .True.,\n  N_IC_MAX = 1,\n  MH_BATCH = .true.,\n"
        "  EMITTERS_DIM(1:3, 1) = 100.0d0, 100.0d0, 0.0d0,\n  EMITTERS_POS(1:3, 1) = -50.0d0, -50.0d0, 0.0d0,\n"
        "  EMITTERS_TYPE(1) = 2,\n  EMITTERS_DELAY(1) = 0,\n  PLANES_N = 2,\n  PLANES_Z = 10.0d0, 250.0d0,\n"
        "  WRITE_POSITION_FILE = .true.,\n  SAMPLE_ELEC_FILE = .true.,\n  SAMPLE_ELEC_RATE = 200,\n/\n")
    (tmp_path / "work").write_text("1\n1 1\n2.00\n")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rumdeed_b200", "rumdeed_b200_run")
    r = subprocess.run([exe, str(tmp_path), "4242", "0", "50000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "out"
    ramo = np.loadtxt(out / "ramo_current.dt")
    assert ramo.shape == (400, 14) and np.all(ramo[:, 1] == np.arange(1, 401))
    emit_dt = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("emit", "<i4"), ("sec", "<i4"), ("id", "<i4"), ("species", "<i4")])
    de = np.fromfile(out / "density_emit.bin", dtype=emit_dt)
    assert len(de) > 50 and np.all(de["z"] == 1.0) and np.array_equal(de["id"], np.arange(len(de)))
    assert np.all(np.abs(de["x"]) <= 50.0) and np.all(de["species"] == 1)
    abs_dt = np.dtype([("x", "<f8"), ("y", "<f8"), ("vx", "<f8"), ("vy", "<f8"), ("vz", "<f8"), ("emit", "<i4"), ("sec", "<i4"), ("id", "<i4")])
    top = np.fromfile(out / "density_absorb_top.bin", dtype=abs_dt)
    pl2 = np.fromfile(out / "planes-2.bin", dtype=abs_dt)
    absorbed = np.loadtxt(out / "absorbed_top.dt")
    assert len(top) == int(absorbed[:, 3].sum()) and len(top) > 10
    assert np.all(top["vz"] > 0) and len(pl2) >= len(top)
    emitted = np.loadtxt(out / "emitted.dt")
    assert int(emitted[:, 2].sum()) == len(de)
    # position.bin (Write_Position) read the way scripts/python_package/rumdeed_io.py does; elec-<step>.bin
    # (Sample_Elec_Position): x, y, z, nearest distance of every electron
    pdt = np.dtype([("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("emit", "<i4"), ("sec", "<i4"), ("id", "<i4")])
    with open(out / "position.bin", "rb") as f:
        assert list(np.fromfile(f, count=2, dtype=np.int32)) == [400, 1]
        last = None
        for k in range(1, 401):
            step_k, nr = np.fromfile(f, count=2, dtype=np.int32)
            assert step_k == k and nr == int(ramo[k - 1, 4])
            last = np.fromfile(f, count=nr, dtype=pdt)
        assert f.read() == b""
    el = np.fromfile(out / "elec-400.bin", dtype="<f8").reshape(-1, 4)
    assert len(el) == len(last) and np.array_equal(el[:, 0], last["x"]) and np.array_equal(el[:, 2], last["z"])
    d = el[:, None, :3] - el[None, :, :3]
    dd = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    np.fill_diagonal(dd, np.inf)
    assert np.array_equal(el[:, 3], dd.min(axis=1))
    assert (out / "elec-200.bin").exists() and not (out / "elec-100.bin").exists()
    # steady-state current of the reference's test system stays in its band
    I_ss = ramo[250:, 2].mean()
    assert 0.5e-3 < I_ss < 4.0e-3


@gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rumdeed_b200", "rumdeed_b200_run")),
                    reason="driver executable not built")
def test_driver_ion_deck_writes_collision_files(tmp_path):
    """program RUMDEED on the Ion configuration (tools/decks/ion: planar field emission into N2 at NTP, COLLISION_MODE = 2):
    collisions.dt, ionization_data.bin and the particle counts are consistent with each other; record layouts of
    Write_Ionization_Data (src/mod_pair.F90:930-935: i32, 8 f64, 4 i32) and of the collisions.dt line (mod_collisions.F90:74)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in ("input", "work", "N2-tot-cross.txt", "N2-ion-cross.txt"):
        shutil.copy(os.path.join(root, "tools", "decks", "ion", f), tmp_path / f)
    txt = (tmp_path / "input").read_text().replace("\n/", "\n  MH_DEVICE = .True.,\n/")
    (tmp_path / "input").write_text(txt)
    exe = os.path.join(root, "rumdeed_b200", "rumdeed_b200_run")
    steps = 600
    r = subprocess.run([exe, str(tmp_path), "777", str(steps), "200000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "out"
    coll = np.loadtxt(out / "collisions.dt", dtype=np.int64)
    assert coll.shape == (steps, 4) and np.all(coll[:, 0] == np.arange(1, steps + 1))
    assert np.all(coll[:, 1] >= coll[:, 2] + coll[:, 3])          # collisions >= ionisations + recombinations
    n_ion = int(coll[:, 2].sum())
    assert n_ion > 100
    idt = np.dtype([("step", "<i4"), ("pos", "<f8", 3), ("in_speed", "<f8"), ("out_speed", "<f8"), ("new_speed", "<f8"),
                    ("ion_dist", "<f8"), ("ion_rad", "<f8"), ("in_id", "<i4"), ("new_id", "<i4"), ("ion_id", "<i4"), ("emit", "<i4")])
    assert idt.itemsize == 84
    ev = np.fromfile(out / "ionization_data.bin", dtype=idt)
    assert len(ev) == n_ion
    assert np.all(np.diff(ev["step"]) >= 0) and np.all(ev["ion_id"] == ev["new_id"] + 1)
    assert np.all(ev["in_id"] >= 1) and np.all((ev["emit"] == 1) | (ev["emit"] == 2))
    # energy balance per event: the two outgoing electrons share E_in - 15.581 eV (src/mod_collisions.F90:626-629)
    half_m = 0.5 * 9.1093837015e-31 / 1.602176634e-19
    e_in, e_out, e_new = half_m * ev["in_speed"] ** 2, half_m * ev["out_speed"] ** 2, half_m * ev["new_speed"] ** 2
    assert np.all(e_in > 15.581) and np.allclose(e_out + e_new, e_in - 15.581, rtol=1e-9, atol=1e-9)
    assert np.all(ev["pos"][:, 2] > 0) and np.all(np.hypot(ev["pos"][:, 0], ev["pos"][:, 1]) <= 500e-9)
    # every ion created before the last step is in the gap when the last ramo_current.dt line is written (none leaves or
    # expires within 600 steps; the line of step k precedes the collisions of step k)
    ramo = np.loadtxt(out / "ramo_current.dt")
    if int(coll[:, 3].sum()) == 0:
        assert int(ramo[-1, 6]) == int(coll[:-1, 2].sum())
    rec_size = (out / "recombination_data.bin").stat().st_size
    assert rec_size == 68 * int(coll[:, 3].sum())                  # i32 + 6 f64 + 4 i32 per recombination
    assert (out / "density_absorb_recom.bin").stat().st_size == 2 * 48 * int(coll[:, 3].sum())
    assert (out / "absorbed_recom.dt").exists()
