"""GPU parity tests: the CUDA hot path, called through the C ABI, against the CPU oracle on
identical seeded inputs (oracle = tests' checker only).

Tolerances (BASELINE.json north_star): fields and forces within 1e-11 relative in FP64,
measured like the reference does (mod_tests.F90:1664-1675): ||a - a_ref|| / max(||a_ref||, 1)
per particle; particle indexing, removal and the Beeman update are bit-exact.
"""
import math

import numpy as np
import pytest

import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, M_N2P, Q_0, REMOVE_BOT, REMOVE_TOP, SPECIES_ELEC, SPECIES_ION

pytestmark = pytest.mark.gpu

NM = 1.0e-9
TOL = 1.0e-11


def cloud(n, seed, box=(1000.0, 1000.0, 1000.0), ions=True, zmin=1.0):
    rng = np.random.default_rng(np.random.PCG64(seed))
    pos = np.stack([rng.uniform(-0.5 * box[0], 0.5 * box[0], n),
                    rng.uniform(-0.5 * box[1], 0.5 * box[1], n),
                    rng.uniform(zmin, box[2] - zmin, n)], axis=1) * NM
    ion = ((np.arange(n) % 10) == 9) if ions else np.zeros(n, dtype=bool)
    q = np.where(ion, Q_0, -Q_0)
    m = np.where(ion, M_N2P, M_0)
    sp = np.where(ion, SPECIES_ION, SPECIES_ELEC).astype(np.int32)
    return pos, q, m, sp


def relerr(a, ref):
    scale = np.maximum(np.linalg.norm(ref, axis=-1), 1.0)
    return float(np.max(np.linalg.norm(a - ref, axis=-1) / scale))


def planar(orc, V=2000.0, d=1000.0 * NM, ic=True, nic=1, dt=1.0e-16, cap=1 << 16, planes=()):
    box = (1000 * NM, 1000 * NM, d)
    cfg = rb.planar_config(V, d, box, dt, ic, nic, capacity=cap, planes_z=planes)
    p = orc.params_planar(V, d, box, dt, ic, nic)
    p.set_planes(planes)
    return cfg, p


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ic,nic", [(False, 0), (True, 0), (True, 1), (True, 2), (True, 3)])
@pytest.mark.parametrize("n", [1, 2, 3, 127, 128, 129, 1000, 2500])
def test_planar_acceleration_vs_oracle(orc, ic, nic, n):
    cfg, p = planar(orc, ic=ic, nic=nic)
    pos, q, m, sp = cloud(n, 20261017 + n)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
    truth = orc.accel_gather_ld(p, pos, q, m)
    ref = orc.accel_gather(p, pos, q, m)
    assert relerr(acc, truth) < TOL
    assert relerr(acc, ref) < TOL
    if n <= 1000:
        scat = orc.accel_planar(p, pos, q, m, sp)  # the CPU pair loop (i<j scatter)
        assert relerr(acc, scat) < TOL


@pytest.mark.parametrize("ic,nic", [(False, 0), (True, 0), (True, 1), (True, 2)])
@pytest.mark.parametrize("tpl", [1, 2])
@pytest.mark.parametrize("n,budget_mb", [(1, 2048), (2, 2048), (33, 2048), (127, 2048), (128, 2048), (129, 0.01), (255, 2048),
                                          (256, 0.01), (257, 2048), (300, 0.01), (385, 0.01), (1000, 0.02), (2500, 0.05),
                                          (2500, 2048), (6000, 0.2)])
def test_pair_symmetric_kernel_vs_oracle(orc, ic, nic, n, budget_mb, tpl):
    """The pair-symmetric kernel (each unordered pair once, like the reference's CPU loop
    mod_verlet.F90:763-884) against the long-double oracle, with small scratch budgets to force
    several bands / groups, sizes around the 32-, 128- and 256-particle tile edges, ions mixed in, and one or
    two targets per lane (sym_tpl)."""
    cfg, p = planar(orc, ic=ic, nic=nic)
    pos, q, m, sp = cloud(n, 777 + n)
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", 2)
        hp.set_option("sym_tpl", tpl)
        hp.set_option("sym_budget_mb", budget_mb)
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        hp.Calculate_Acceleration_Particles()
        again = hp.download(("acc",))["acc"]
        hp.set_option("pair_mode", 1)
        hp.Calculate_Acceleration_Particles()
        gather = hp.download(("acc",))["acc"]
    assert np.array_equal(acc, again)  # fixed summation order: bit-identical from run to run
    if n <= 2500:
        assert relerr(acc, orc.accel_gather_ld(p, pos, q, m)) < TOL
    assert relerr(acc, gather) < 1e-12


@pytest.mark.parametrize("tpl", [1, 2])
def test_pair_symmetric_split_over_two_ranks(orc, tpl):
    """Work units dealt to two 'processes' (run one after the other here) and summed give the full result."""
    cfg, p = planar(orc, ic=True, nic=1)
    pos, q, m, sp = cloud(3000, 4242)
    raws = []
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", 2)
        hp.set_option("sym_tpl", tpl)
        hp.set_option("sym_budget_mb", 0.1)
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        full = hp.download(("acc",))["acc"]
        import torch
        for r in range(2):
            hp.set_pair_rank(r, 2)
            hp.accel_partial()
            ptr, nbytes = hp.device_buffer("raw")
            raws.append(_dev_to_numpy(ptr, nbytes // 8))
        total = raws[0] + raws[1]
        _numpy_to_dev(total, ptr)
        hp.accel_finalize()
        hp.set_pair_rank(0, 1)
        split = hp.download(("acc",))["acc"]
    assert relerr(split, full) < 1e-13
    assert relerr(split, orc.accel_gather(p, pos, q, m)) < TOL


class _Alias:
    """Expose a raw device pointer of the library to torch without a copy."""

    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _dev_to_numpy(ptr, nelem):
    import torch
    t = torch.as_tensor(_Alias(ptr, nelem), device="cuda")
    out = t.cpu().numpy().copy()
    torch.cuda.synchronize()
    return out


def _numpy_to_dev(a, ptr):
    import torch
    t = torch.as_tensor(_Alias(ptr, a.size), device="cuda")
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    torch.cuda.synchronize()


def test_reference_golden_three_particles(orc):
    """mod_tests.F90:405-518: closed-form Coulomb accelerations and the Python field vector."""
    d, V = 100 * NM, 2.0
    box = (100 * NM, 100 * NM, d)
    R = np.array([[3.0, -10.0, 2.0], [-9.0, 26.0, 80.0], [6.0, -24.0, 56.53]]) * NM
    q = np.array([-Q_0, -Q_0, Q_0]); m = np.array([M_0, M_0, M_N2P])
    sp = np.array([1, 1, 2], dtype=np.int32)
    cfg = rb.planar_config(V, d, box, 0.25e-15, False, 0, capacity=64)
    pre = Q_0 ** 2 * rb.api.DIV_FAC_C
    E = np.array([0, 0, -V / d])
    coul = lambda a, b: (a - b) / np.linalg.norm(a - b) ** 3
    want = np.stack([(pre * coul(R[0], R[1]) - pre * coul(R[0], R[2])) / M_0 - Q_0 / M_0 * E,
                     (pre * coul(R[1], R[0]) - pre * coul(R[1], R[2])) / M_0 - Q_0 / M_0 * E,
                     (-pre * coul(R[2], R[0]) - pre * coul(R[2], R[1])) / M_N2P + Q_0 / M_N2P * E])
    with rb.HotPath(cfg) as hp:
        probe = np.array([-4.55, -2.34, 96.44]) * NM
        assert np.allclose(hp.Calc_Field_at(probe), E, rtol=0, atol=0)  # empty system: vacuum field
        hp.Add_Particle(R[0], [0, 0, 0], SPECIES_ELEC, 1, 0)
        hp.Add_Particle(R[1], [0, 0, 0], SPECIES_ELEC, 1, 0)
        hp.Add_Particle(R[2], [0, 0, 0], SPECIES_ION, 1, 0)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        assert relerr(acc, want) < 1e-8  # closed form differs by the 1e-18 m softening only
        E_python = np.array([-314559.29097098, 1423979.07058996, -20246038.87978313])
        got = hp.Calc_Field_at(probe)
        assert np.all(np.abs(got - E_python) / np.abs(E_python) < 1e-7)
        p = orc.params_planar(V, d, box, 0.25e-15, False, 0)
        assert relerr(got[None], orc.calc_field_at(p, R, q, probe, ld=True)[None]) < TOL


def test_forty_particle_reference_case(orc):
    """mod_tests.F90:1610-1676: 40 sin/cos-placed particles, every 5th an ion, N_ic_max = 2 and IC off."""
    d, V = 1000 * NM, 2.0e3
    box = (100 * NM, 100 * NM, d)
    i = np.arange(1, 41, dtype=np.float64)
    pos = np.stack([(0.5 + 0.45 * np.sin(1.7 * i)) * box[0], (0.5 + 0.45 * np.cos(2.3 * i)) * box[1],
                    (0.5 + 0.45 * np.sin(3.1 * i + 0.5)) * box[2]], axis=1)
    ion = (np.arange(1, 41) % 5) == 0
    q = np.where(ion, Q_0, -Q_0); m = np.where(ion, M_N2P, M_0)
    sp = np.where(ion, 2, 1).astype(np.int32)
    for ic, nic in ((True, 2), (False, 0)):
        cfg = rb.planar_config(V, d, box, 0.25e-15, ic, nic, capacity=64)
        p = orc.params_planar(V, d, box, 0.25e-15, ic, nic)
        with rb.HotPath(cfg) as hp:
            hp.upload(pos, q, m, species=sp)
            hp.Calculate_Acceleration_Particles()
            acc = hp.download(("acc",))["acc"]
        assert relerr(acc, orc.accel_generic(p, pos, q, m, sp)) < 1e-12
        assert relerr(acc, orc.accel_planar(p, pos, q, m, sp)) < 1e-12


def test_coincident_and_marked_particles(orc):
    """Exact coincidence gives a zero Coulomb term (softened 1/r^3 times a zero offset), and a
    marked particle (q = 0) neither exerts nor feels pair forces (mod_pair.F90:207)."""
    cfg, p = planar(orc, ic=True, nic=1)
    pos, q, m, sp = cloud(300, 7)
    pos[17] = pos[3]
    q = q.copy(); q[40] = 0.0; q[41] = 0.0
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
    assert np.all(np.isfinite(acc))
    assert relerr(acc, orc.accel_gather(p, pos, q, m)) < TOL
    assert np.all(acc[40] == 0.0) and np.all(acc[41] == 0.0)


@pytest.mark.parametrize("mode,tpl", [(1, 0), (2, 1), (2, 2)])
@pytest.mark.parametrize("d_nm", [1000.0, 2500.0, 500.0])
def test_far_partner_shortcut_stays_inside_the_bar(orc, d_nm, mode, tpl):
    """Option sym_far (default on): with a gap d >= 1 um the three image partners that are at least d away are evaluated
    without the softening term -- a change of at most 3e-18 / d of those terms.  With and without it the result is
    within 1e-11 of the long-double oracle and the two differ by less than 4e-12; below 1 um the option changes nothing
    (bit for bit).  Particles right at both electrodes included (the near partners keep the full form)."""
    d = d_nm * NM
    n = 3000
    box = (1000 * NM, 1000 * NM, d)
    cfg = rb.planar_config(2000.0, d, box, 1.0e-16, True, 1, capacity=n)
    p = orc.params_planar(2000.0, d, box, 1.0e-16, True, 1)
    pos, q, m, sp = cloud(n, 5150, box=(1000.0, 1000.0, d_nm), ions=True, zmin=0.0)
    pos[:40, 2] = np.linspace(0.0, 2.0, 40) * NM            # grazing the cathode
    pos[40:80, 2] = d - np.linspace(0.0, 2.0, 40) * NM      # ... and the anode
    res = {}
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", mode)
        if tpl:
            hp.set_option("sym_tpl", tpl)
        hp.upload(pos, q, m, species=sp)
        for far in (1, 0):
            hp.set_option("sym_far", far)
            hp.Calculate_Acceleration_Particles()
            res[far] = hp.download(("acc",))["acc"]
    truth = orc.accel_gather_ld(p, pos, q, m)
    assert relerr(res[0], truth) < TOL and relerr(res[1], truth) < TOL
    if d_nm >= 1000.0:
        assert relerr(res[1], res[0]) < 4e-12
        assert not np.array_equal(res[1], res[0])  # the shortcut is actually taken
    else:
        assert np.array_equal(res[1], res[0])
    # a charged particle OUTSIDE the plates (only a caller can put it there: the time step strips such a particle of its
    # charge) brings "far" partners close: the shortcut must switch itself off -- bit-identical with and without the
    # option, through the resident state and through rb2_accel_host
    pos2 = pos.copy()
    pos2[5, 2] = 1.99 * d
    pos2[6, 2] = -0.3 * d
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", mode)
        if tpl:
            hp.set_option("sym_tpl", tpl)
        out = {}
        for far in (1, 0):
            hp.set_option("sym_far", far)
            hp.upload(pos2, q, m, species=sp)
            hp.Calculate_Acceleration_Particles()
            out[far] = (hp.download(("acc",))["acc"], hp.accel_host(pos2, q, m))
    assert np.array_equal(out[1][0], out[0][0]) and np.array_equal(out[1][1], out[0][1])
    assert relerr(out[1][0], orc.accel_gather_ld(p, pos2, q, m)) < TOL


@pytest.mark.parametrize("mode,tpl", [(1, 0), (2, 1), (2, 2)])
@pytest.mark.parametrize("n", [300, 4000])
def test_close_pairs_use_reference_arithmetic(orc, n, mode, tpl):
    """The fast inverse cube is first order in the 1e-18 m softening: good to 6 (eps/r)^2, i.e. not to 1e-11 below
    r ~ 1e-12 m.  Pairs whose lateral offset is below 1e-11 m are re-evaluated with the reference's own sqrt / divide
    (src/mod_verlet.F90:1302-1303): separations from 1e-15 to 1e-11 m, along every axis, in the same tile, across
    tiles and superblocks, against an image partner at either electrode, both kernels -- all within 1e-11 of the
    long-double oracle."""
    d = 1000 * NM
    cfg, p = planar(orc, ic=True, nic=1, cap=n + 64)
    pos, q, m, sp = cloud(n, 99 + n, ions=True)
    rng = np.random.default_rng(n)
    seps = [1e-15, 3e-15, 1e-14, 1e-13, 3e-13, 1e-12, 5e-12, 1e-11, 3e-11]
    partners = []
    k = 0
    for r in seps:
        for axis in range(3):
            # i: a particle somewhere in the array, j: another one far away in index (other tile / superblock) or adjacent
            i = int(rng.integers(0, n))
            j = (i + 1) % n if k % 2 == 0 else (i + n // 2 + 17) % n
            off = np.zeros(3); off[axis] = r
            if axis < 2:
                off[2] = 0.3 * r
            pos[j] = pos[i] + off
            partners.append((i, j))
            k += 1
    # both particles grazing the cathode / the anode, laterally 1e-13 m apart: the n = 0 partner (distance z_i + z_j) and
    # the 2d - z partner come as close as the direct pair
    base = n - 8
    pos[base] = [10 * NM, -20 * NM, 2e-13]; pos[base + 1] = [10 * NM + 1e-13, -20 * NM, 3e-13]
    pos[base + 2] = [-30 * NM, 5 * NM, d - 2e-13]; pos[base + 3] = [-30 * NM, 5 * NM + 1e-13, d - 1e-13]
    pos[base + 4] = pos[3]  # exactly coincident: contributes exactly 0 (softened weight times zero offset)
    touched = sorted({x for ij in partners for x in ij} | set(range(base, base + 5)) | {3})
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", mode)
        if tpl:
            hp.set_option("sym_tpl", tpl)
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        pts = np.concatenate([pos[[i for i, _ in partners]] + [0, 0, 2e-14], [[10 * NM, -20 * NM, 0.0], [-30 * NM, 5 * NM, d]]])
        fld = hp.Calc_Field_at_Batch(pts)
        ez = hp.field_surface_z(np.array([[10 * NM, -20 * NM, 0.0], [10 * NM + 5e-14, -20 * NM, 0.0]]))
    assert np.all(np.isfinite(acc))
    truth = orc.accel_gather_ld(p, pos, q, m)
    assert relerr(acc, truth) < TOL
    assert relerr(acc[touched], truth[touched]) < TOL
    want = np.array([orc.calc_field_at(p, pos, q, pt, sp, ld=True) for pt in pts])
    assert relerr(fld, want) < TOL
    wz = np.array([orc.calc_field_at(p, pos, q, pt, sp, ld=True)[2] for pt in ([10 * NM, -20 * NM, 0.0], [10 * NM + 5e-14, -20 * NM, 0.0])])
    assert np.max(np.abs(ez - wz) / np.abs(wz)) < TOL


def test_fused_step_refuses_an_i_partition_and_remove_reasons(orc):
    """The fused step has no exchange: with an i-partition only the owned rows would get an acceleration and the others
    would be integrated with a = 0 (silently wrong) -- it must fail instead.  Mark_Particles_Remove: reason 4
    (remove_ion, src/mod_pair.F90:273-279, :327-333) has its own counters, an unknown reason is an error."""
    cfg, p = planar(orc, cap=256)
    pos, q, m, sp = cloud(100, 4, ions=True)
    sp[7] = 3; q[7] = 0.0
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.set_partition(10, 60)
        with pytest.raises(rb.api.Rb2Error) as e:
            hp.Update_Position(1)
        assert "partition" in str(e.value)
        assert np.array_equal(hp.download(("pos",))["pos"], pos)  # nothing was integrated
        hp.Update_Particle_Position(1)                            # the three phases keep working with a partition
        hp.Calculate_Acceleration_Particles()
        hp.set_partition(0, -1)
        hp.Update_Position(2)
        with pytest.raises(rb.api.Rb2Error):
            hp.Mark_Particles_Remove([1], 9)
        assert hp.counts().nrPart_remove == 0
        hp.Mark_Particles_Remove([0, 7], rb.api.REMOVE_ION)       # an electron and the atom
        k = hp.counts()
        assert (k.nrPart_remove, k.nrPart_remove_ion, k.nrElec_remove_ion, k.nrAtom_remove_ion) == (2, 2, 1, 1)
        assert (k.nrPart_remove_top, k.nrPart_remove_bot) == (0, 0)
        k = hp.Remove_Particles(3)
        assert (k.nrPart, k.nrAtom, k.nrPart_remove_ion) == (98, 0, 0)


def test_partition_and_host_buffer_paths(orc):
    cfg, p = planar(orc, ic=True, nic=1)
    pos, q, m, sp = cloud(1500, 11)
    ref = orc.accel_gather(p, pos, q, m)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        full = hp.download(("acc",))["acc"]
        # i-partition: two halves written in place give the same bits as the full evaluation
        hp.upload(pos, q, m, species=sp)
        hp.set_partition(0, 700)
        hp.Calculate_Acceleration_Particles()
        hp.set_partition(700, 1500)
        hp.Calculate_Acceleration_Particles()
        hp.set_partition(0, -1)
        halves = hp.download(("acc",))["acc"]
        host = hp.accel_host(pos, q, m)
    assert relerr(full, ref) < TOL
    assert relerr(halves, ref) < TOL
    assert relerr(host, ref) < TOL
    assert np.array_equal(host, full)


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ic,nic", [(False, 0), (True, 0), (True, 1), (True, 2)])
@pytest.mark.parametrize("M", [1, 4, 33, 300])
def test_planar_field_batch_vs_oracle(orc, ic, nic, M):
    cfg, p = planar(orc, ic=ic, nic=nic)
    pos, q, m, sp = cloud(2000, 5)
    rng = np.random.default_rng(99 + M)
    pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), np.zeros(M)], axis=1) * NM
    if M >= 4:
        pts[1, 2] = 1.0 * NM    # the photo-emission probe height
        pts[2, 2] = 500.0 * NM
    with rb.HotPath(cfg) as hp:
        vac = hp.Calc_Field_at_Batch(pts)
        assert np.all(vac[:, :2] == 0.0) and np.all(vac[:, 2] == cfg.E_z)
        hp.upload(pos, q, m, species=sp)
        hp.Particles_To_Device()
        fld = hp.Calc_Field_at_Batch(pts)
        one = hp.Calc_Field_at(pts[0])
        hp.Release_Device_Particles()
    want = np.stack([orc.calc_field_at(p, pos, q, pts[k], sp, ld=True) for k in range(M)])
    assert relerr(fld, want) < TOL
    assert relerr(fld, orc.calc_field_at_batch(p, pos, q, pts, sp)) < TOL
    assert relerr(one[None], want[:1]) < TOL


@pytest.mark.parametrize("ic,nic", [(False, 0), (True, 0), (True, 1), (True, 2), (True, 4)])
@pytest.mark.parametrize("M,n", [(1, 0), (1, 1), (5, 127), (33, 2000), (300, 129), (2000, 5000)])
def test_surface_field_z_vs_oracle(orc, ic, nic, M, n):
    """rb2_field_surface_z: E_z on the cathode plane from the mirror-antisymmetric form of the image series
    equals the z component of Calc_Field_at (mod_verlet.F90:1466) evaluated there, to the FP64 gate."""
    cfg, p = planar(orc, ic=ic, nic=nic)
    pos, q, m, sp = cloud(max(n, 1), 7)
    pos, q, m, sp = pos[:n], q[:n], m[:n], sp[:n]
    rng = np.random.default_rng(17 + M)
    pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), np.zeros(M)], axis=1) * NM
    with rb.HotPath(cfg) as hp:
        if n:
            hp.upload(pos, q, m, species=sp)
        ez = hp.field_surface_z(pts)
        full = hp.Calc_Field_at_Batch(pts)
        bad = pts.copy()
        bad[-1, 2] = 1.0 * NM
        with pytest.raises(rb.Rb2Error):
            hp.field_surface_z(bad)
    if n == 0:
        assert np.all(ez == cfg.E_z)
        return
    want = np.array([orc.calc_field_at(p, pos, q, pts[k], sp, ld=True)[2] for k in range(min(M, 300))])
    assert np.max(np.abs(ez[:len(want)] - want) / np.abs(want)) < TOL
    assert np.max(np.abs(ez - full[:, 2]) / np.abs(ez)) < TOL
    if ic:  # the lateral components the general kernel returns there are rounding noise
        assert np.max(np.abs(full[:, :2])) < 1e-9 * np.max(np.abs(full[:, 2]))


def test_reference_planar_batch_case(orc):
    """mod_tests.F90:1554-1603 (N_ic_max = 1, 3 particles, 4 points; vacuum check)."""
    d, V = 100 * NM, 2.0
    box = (100 * NM, 100 * NM, d)
    cfg = rb.planar_config(V, d, box, 0.25e-15, True, 1, capacity=64)
    p = orc.params_planar(V, d, box, 0.25e-15, True, 1)
    pts = np.array([[0, 0, 10.0], [12, -4, 21], [-20, 30, 80], [40, 40, 95]]) * NM
    R = np.array([[10, -5, 20.0], [-15, 8, 60], [5, 25, 40]]) * NM
    q = np.array([-Q_0, -Q_0, Q_0])
    with rb.HotPath(cfg) as hp:
        vac = hp.Calc_Field_at_Batch(pts)
        assert np.all(vac == np.array([0.0, 0.0, -V / d]))
        for r, s in zip(R, (1, 1, 2)):
            hp.Add_Particle(r, [0, 0, 0], s, 1, 1)
        batch = hp.Calc_Field_at_Batch(pts)
        singles = np.stack([hp.Calc_Field_at(pt) for pt in pts])
    assert relerr(batch, orc.calc_field_at_batch(p, R, q, pts)) < TOL
    assert relerr(batch, singles) < 1e-13


def test_field_delta_is_linear(orc):
    """snapshot + delta == field of the enlarged system (serial sampler semantics)."""
    cfg, p = planar(orc, ic=True, nic=1)
    pos, q, m, sp = cloud(900, 21)
    new_pos, new_q, _, _ = cloud(37, 22, ions=False)
    new_pos[:, 2] = 1.0 * NM
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(-500, 500, 50), rng.uniform(-500, 500, 50), np.zeros(50)], axis=1) * NM
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        delta = hp.Calc_Field_at_Batch_delta(pts, new_pos, new_q)
    want = orc.calc_field_at_batch(p, np.concatenate([pos, new_pos]), np.concatenate([q, new_q]), pts)
    assert relerr(delta, want) < TOL


# --------------------------------------------------------------------------------------------------
def tip_setup(orc, V=2.0e3, cap=4096, dt=0.25e-3 * 1e-12, boxz=1000 * NM, ic=True):
    box = (100 * NM, 100 * NM, boxz)
    cfg = rb.tip_config(V, 900 * NM, 100 * NM, 100 * NM, box, dt, ic, capacity=cap)
    p = orc.params_tip(V, 900 * NM, 100 * NM, 100 * NM, box, dt, ic)
    return cfg, p


def test_tip_config_matches_oracle(orc):
    cfg, p = tip_setup(orc)
    for name in ("a_foci", "eta_1", "shift_z", "pre_fac_E_tip", "pre_fac_E_tip_unit_voltage", "h_tip", "r_tip", "max_xi", "d"):
        assert getattr(cfg, name) == pytest.approx(getattr(p, name), rel=1e-15), name


def test_tip_reference_case(orc):
    """mod_tests.F90:988-1084: three particles at the apex; batch == point."""
    cfg, p = tip_setup(orc)
    R = np.array([[2.0, 1.0, 103.0], [-2.0, 2.5, 106.0], [1.5, -3.0, 110.0]]) * NM
    q = np.array([-Q_0, -Q_0, Q_0]); m = np.array([M_0, M_0, M_N2P])
    sp = np.array([1, 1, 2], dtype=np.int32)
    pts = np.array([[0, 0, 102.0], [30, -10, 130], [-50, 40, 300], [15, 8, 500]]) * NM
    with rb.HotPath(cfg) as hp:
        for r, s in zip(R, sp):
            hp.Add_Particle(r, [0, 0, 0], int(s), 1, 0)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        batch = hp.Calc_Field_at_Batch(pts)
        singles = np.stack([hp.Calc_Field_at(pt) for pt in pts])
    assert relerr(acc, orc.accel_generic(p, R, q, m, sp)) < TOL
    assert relerr(acc, orc.accel_gather_ld(p, R, q, m)) < TOL
    assert relerr(batch, orc.calc_field_at_batch(p, R, q, pts, sp)) < TOL
    assert relerr(batch, singles) < 1e-13


@pytest.mark.parametrize("ic", [True, False])
def test_tip_close_pairs_and_large_cloud(orc, ic):
    """The tip kernels run on MUFU-seeded inverse cubes (rb2_tip_math.cuh: Coulomb and sphere-image term rewritten on
    shared distances); a pair closer than 1e-11 m goes back to the literal sqrt / divide arithmetic.  2600 particles (21
    source tiles, pairs with lower- and higher-indexed sources in every row) with separations from 1e-15 to 3e-11 m planted
    across tiles, field points 2e-14 m from a particle, image charge on and off; 260 rows against the long-double oracle."""
    n = 2600
    cfg, p = tip_setup(orc, ic=ic)
    rng = np.random.default_rng(77)
    pos = np.stack([rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(105, 900, n)], axis=1) * NM
    ion = (np.arange(n) % 7) == 6
    q = np.where(ion, Q_0, -Q_0); m = np.where(ion, M_N2P, M_0)
    seps = [1e-15, 1e-14, 1e-13, 1e-12, 5e-12, 1e-11, 3e-11]
    touched = []
    for k, r in enumerate(seps):
        for axis in range(3):
            i = int(rng.integers(0, n)); j = (i + 1) % n if (k + axis) % 2 == 0 else (i + n // 2 + 31) % n
            off = np.zeros(3); off[axis] = r
            pos[j] = pos[i] + off
            touched += [i, j]
    rows = np.unique(np.concatenate([touched, rng.choice(n, 220, replace=False), [0, 127, 128, n - 1]]))
    pts = np.concatenate([pos[touched[:6]] + [0, 0, 2e-14], np.stack([rng.uniform(-30, 30, 30), rng.uniform(-30, 30, 30), rng.uniform(101, 400, 30)], axis=1) * NM])
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=np.where(ion, 2, 1).astype(np.int32))
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        fld = hp.Calc_Field_at_Batch(pts)
        hp.set_option("tip_field_small", 0)
        fld_tiled = hp.Calc_Field_at_Batch(pts)
    assert np.all(np.isfinite(acc))
    truth = orc.accel_gather_ld_rows(p, pos, q, m, rows)
    assert relerr(acc[rows], truth) < TOL
    want = np.stack([orc.calc_field_at(p, pos, q, pt, ld=True) for pt in pts])
    assert relerr(fld, want) < TOL and relerr(fld_tiled, want) < TOL


@pytest.mark.parametrize("n", [5, 130, 700])
def test_tip_cloud_vs_oracle(orc, n):
    cfg, p = tip_setup(orc)
    rng = np.random.default_rng(n)
    pos = np.stack([rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(105, 900, n)], axis=1) * NM
    ion = (np.arange(n) % 7) == 6
    q = np.where(ion, Q_0, -Q_0); m = np.where(ion, M_N2P, M_0)
    pts = np.stack([rng.uniform(-30, 30, 40), rng.uniform(-30, 30, 40), rng.uniform(101, 400, 40)], axis=1) * NM
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=np.where(ion, 2, 1).astype(np.int32))
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        fld = hp.Calc_Field_at_Batch(pts)               # small batch: the CTA-per-point tip kernel
        hp.set_option("tip_field_small", 0)
        fld_tiled = hp.Calc_Field_at_Batch(pts)         # the tiled pair kernel
        hp.set_option("tip_field_small", 1)
    assert relerr(acc, orc.accel_gather_ld(p, pos, q, m)) < TOL
    assert relerr(acc, orc.accel_gather(p, pos, q, m)) < TOL
    want = np.stack([orc.calc_field_at(p, pos, q, pts[k], ld=True) for k in range(40)])
    assert relerr(fld, want) < TOL and relerr(fld_tiled, want) < TOL
    assert relerr(fld, fld_tiled) < 1e-12


# --------------------------------------------------------------------------------------------------
def test_beeman_kinematics_reference_case(orc):
    """mod_tests.F90:1403-1452."""
    d, V, dt, vx0 = 1000 * NM, 2.0e3, 0.25e-15, 1.0e3
    cfg = rb.planar_config(V, d, (100 * NM, 100 * NM, d), dt, False, 0, capacity=16)
    with rb.HotPath(cfg) as hp:
        hp.Add_Particle([0.0, 0.0, 500 * NM], [vx0, 0.0, 0.0], SPECIES_ELEC, 1, 1)
        for i in range(1, 4):
            r = hp.Update_Position(i)
        s = hp.download(("pos", "vel"))
    a_z = Q_0 * V / (M_0 * d)
    assert abs((s["pos"][0, 2] - 500 * NM) - 4.5 * a_z * dt ** 2) / (4.5 * a_z * dt ** 2) < 1e-9
    assert abs(s["vel"][0, 2] - 3.0 * a_z * dt) / (3.0 * a_z * dt) < 1e-12
    assert abs(s["pos"][0, 0] - 3.0 * vx0 * dt) / (3.0 * vx0 * dt) < 1e-12
    assert s["vel"][0, 0] == vx0 and s["pos"][0, 1] == 0.0
    ramo = Q_0 * 3.0 * a_z * dt / d
    assert abs(r.ramo_current[SPECIES_ELEC] - ramo) / ramo < 1e-12


def test_particle_removal_reference_case(orc):
    """mod_tests.F90:1193-1328: survivors 1,3,5,7 in order, ids 0,2,4,6, life-time bin 10."""
    d = 100 * NM
    cfg = rb.planar_config(2.0, d, (100 * NM, 100 * NM, d), 0.25e-15, False, 0, capacity=16)
    spc = [1, 1, 2, 1, 1, 2, 1]
    R = np.array([[1.0 * i, -2.0 * i, 10.0 * i] for i in range(1, 8)]) * NM
    Vel = np.array([[100.0 * i, -50.0 * i, 25.0 * i] for i in range(1, 8)])
    with rb.HotPath(cfg) as hp:
        for i in range(7):
            hp.Add_Particle(R[i], Vel[i], spc[i], 1, 1)
        k = hp.counts()
        assert (k.nrPart, k.nrElec, k.nrIon, k.nrID) == (7, 5, 2, 7)
        hp.Mark_Particles_Remove(1, REMOVE_TOP)
        hp.Mark_Particles_Remove(1, REMOVE_TOP)  # no-op
        hp.Mark_Particles_Remove(3, REMOVE_BOT)
        hp.Mark_Particles_Remove(5, REMOVE_TOP)
        k = hp.counts()
        assert (k.nrPart_remove, k.nrElec_remove, k.nrIon_remove) == (3, 2, 1)
        assert (k.nrElec_remove_top, k.nrElec_remove_bot, k.nrIon_remove_top) == (1, 1, 1)
        assert hp.download(("charge",))["charge"][1] == 0.0
        k = hp.Remove_Particles(11)
        assert (k.nrPart, k.nrElec, k.nrIon) == (4, 3, 1)
        s = hp.download(("pos", "prev_pos", "vel", "charge", "mass", "species", "id", "emitter", "section", "mask"))
        keep = [0, 2, 4, 6]
        assert np.array_equal(s["pos"], R[keep]) and np.array_equal(s["vel"], Vel[keep])
        assert np.all(s["prev_pos"] == -1.0 * NM)
        assert list(s["species"]) == [1, 2, 1, 1] and list(s["id"]) == [0, 2, 4, 6]
        assert list(s["emitter"]) == [1] * 4 and list(s["section"]) == [1] * 4
        assert s["charge"][0] == -Q_0 and s["charge"][1] == Q_0 and s["mass"][0] == M_0 and s["mass"][1] == M_N2P
        lt = hp.life_time()
        assert lt[10, SPECIES_ELEC] == 2 and lt[10, SPECIES_ION] == 1
        assert np.all(s["mask"] == 1)
        k = hp.counts()
        assert (k.nrPart_remove, k.nrElec_remove, k.nrIon_remove) == (0, 0, 0)
        # remove everything (the branch that skips the compaction), then reuse the store
        hp.Mark_Particles_Remove(np.arange(4), REMOVE_TOP)
        k = hp.Remove_Particles(21)
        assert k.nrPart == 0 and k.nrElec == 0
        hp.Add_Particle(np.array([5, 6, 50.0]) * NM, [0, 0, 0], SPECIES_ELEC, 22, 1)
        s = hp.download(("pos", "id"))
        assert np.array_equal(s["pos"][0], np.array([5, 6, 50.0]) * NM) and s["id"][0] == 7


@pytest.mark.parametrize("geom,event_buffer", [("planar", None), ("tip", None), ("planar", 1)])
def test_stepped_trajectory_bit_exact_bookkeeping(orc, geom, event_buffer):
    """Many Beeman steps with absorption at both electrodes, plane crossings, mid-run additions:
    particle order, ids, records and counters must match the oracle exactly; positions and
    velocities agree to rounding (the accelerations feeding them agree to ~1e-13).  event_buffer = 1: the
    record list never fits its device buffer at first, so it is rebuilt after the step from the saved
    pre-update velocities."""
    rng = np.random.default_rng(5)
    if geom == "planar":
        d = 200 * NM
        planes = (5 * NM, 50 * NM, 150 * NM)
        box = (100 * NM, 100 * NM, d)
        dt = 2.0e-16
        cfg = rb.planar_config(40.0, d, box, dt, True, 1, capacity=2048, planes_z=planes)
        p = orc.params_planar(40.0, d, box, dt, True, 1)
        n0 = 300
        pos = np.stack([rng.uniform(-50, 50, n0), rng.uniform(-50, 50, n0), rng.uniform(1, 199, n0)], axis=1) * NM
        vel = np.stack([rng.normal(0, 2e4, n0), rng.normal(0, 2e4, n0), rng.normal(0, 4e5, n0)], axis=1)
    else:
        planes = (150 * NM, 400 * NM)
        box = (100 * NM, 100 * NM, 600 * NM)
        dt = 2.0e-16
        cfg = rb.tip_config(500.0, 900 * NM, 100 * NM, 100 * NM, box, dt, True, capacity=2048, planes_z=planes)
        p = orc.params_tip(500.0, 900 * NM, 100 * NM, 100 * NM, box, dt, True)
        n0 = 200
        pos = np.stack([rng.uniform(-20, 20, n0), rng.uniform(-20, 20, n0), rng.uniform(102, 590, n0)], axis=1) * NM
        vel = np.stack([rng.normal(0, 2e4, n0), rng.normal(0, 2e4, n0), rng.normal(0, 6e5, n0)], axis=1)
    p.set_planes(planes)
    species = np.where((np.arange(n0) % 9) == 8, 2, 1).astype(np.int32)
    st = orc.store(2048)
    with rb.HotPath(cfg) as hp:
        if event_buffer is not None:
            hp.set_option("event_buffer", event_buffer)
        hp.set_option("ramo_sections", 5)
        for i in range(n0):
            st.add(p, pos[i], vel[i], int(species[i]), 0, 1, -1, 1 + (i % 5))
        hp.Add_Particles(pos, vel, species, 0, emit=np.ones(n0, dtype=np.int32), sec=(1 + (np.arange(n0) % 5)).astype(np.int32))
        total_events = 0
        for step in range(1, 121):
            st.clear_events()
            st.step(p)
            if event_buffer is not None and step % 2 == 0:
                hp.set_option("event_buffer", event_buffer)  # shrink again: every other step overflows
            r = hp.Update_Position(step)
            ev_o, ev_g = st.events(), hp.events()
            assert r.n_events == len(ev_o) == len(ev_g)
            for a, b in zip(ev_o, ev_g):
                assert (a["kind"], a["plane"], a["index"], a["emit"], a["sec"], a["id"]) == \
                       (b["kind"], b["plane"], b["index"], b["emit"], b["sec"], b["id"])
                for key in ("x", "y", "vx", "vy", "vz"):
                    assert b[key] == pytest.approx(a[key], rel=1e-9, abs=1e-30)
            total_events += len(ev_o)
            k = hp.counts()
            assert (k.nrPart_remove, k.nrElec_remove, k.nrIon_remove) == (st.s.nrPart_remove, st.s.nrElec_remove, st.s.nrIon_remove)
            assert (k.nrElec_remove_top, k.nrElec_remove_bot) == (st.s.nrElec_remove_top, st.s.nrElec_remove_bot)
            for sp_ in (1, 2):
                assert r.ramo_current[sp_] == pytest.approx(st.s.ramo_current[sp_], rel=1e-9, abs=1e-30)
            sec_g, sec_o = hp.ramo_current_emit(5), st.ramo_current_emit(5)
            assert np.allclose(sec_g, sec_o, rtol=1e-9, atol=1e-9 * np.max(np.abs(sec_o)) + 1e-300)  # src/mod_verlet.F90:489-492
            st.remove(step)
            k = hp.Remove_Particles(step)
            assert (k.nrPart, k.nrElec, k.nrIon) == (st.s.nrPart, st.s.nrElec, st.s.nrIon)
            if step % 10 == 0:  # emission: new electrons 1 nm above the cathode / tip
                kadd = 7
                if geom == "planar":
                    npos = np.stack([rng.uniform(-50, 50, kadd), rng.uniform(-50, 50, kadd), np.full(kadd, 1.0)], axis=1) * NM
                else:
                    npos = np.stack([rng.uniform(-5, 5, kadd), rng.uniform(-5, 5, kadd), np.full(kadd, 102.0)], axis=1) * NM
                for i in range(kadd):
                    st.add(p, npos[i], [0, 0, 0], 1, step, 1, -1, 3)
                hp.Add_Particles(npos, np.zeros((kadd, 3)), np.ones(kadd, dtype=np.int32), step,
                                 emit=np.ones(kadd, dtype=np.int32), sec=np.full(kadd, 3, dtype=np.int32))
            if step % 20 == 0:
                s = hp.download(("pos", "vel", "id", "species", "step", "section"))
                assert np.array_equal(s["id"], st.ids) and np.array_equal(s["species"], st.species)
                assert np.array_equal(s["step"], st.step_born) and np.array_equal(s["section"], st.section)
                assert np.allclose(s["pos"], st.pos, rtol=1e-10, atol=1e-22)
                assert np.allclose(s["vel"], st.vel, rtol=1e-8, atol=1e-6)
        assert total_events > 20 and st.s.nrPart < n0 + 12 * 7
        lt = hp.life_time()
        for sp_ in (1, 2):
            ref = np.array([st.life_time(t, sp_) for t in range(rb.api.MAX_LIFE_TIME + 1)])
            assert np.array_equal(lt[:, sp_], ref)


@pytest.mark.parametrize("geom", ["planar", "tip"])
@pytest.mark.parametrize("board,n", [(2, 3), (2, 1000), (4, 300), (4, 70001), (96, 70001)])
def test_ramo_current_per_section_vs_oracle(orc, geom, board, n):
    """ramo_current_emit(sec, emit) (src/mod_verlet.F90:489-492, written by Write_Ramo_Current src/mod_pair.F90:822-826)
    for a board x board checkerboard of sections: every section against the oracle's serial sum, atoms skipped,
    sections beyond the table ignored, bit-identical from run to run, and the sections add up to the species totals."""
    rng = np.random.default_rng(1000 * board + n)
    nsec = board * board
    if geom == "planar":
        d, dt = 1000 * NM, 1.0e-16
        box = (1000 * NM, 1000 * NM, d)
        cfg = rb.planar_config(2000.0, d, box, dt, True, 1, capacity=n + 8)
        p = orc.params_planar(2000.0, d, box, dt, True, 1)
        pos = np.stack([rng.uniform(-500, 500, n), rng.uniform(-500, 500, n), rng.uniform(1, 999, n)], axis=1) * NM
    else:
        box, dt = (100 * NM, 100 * NM, 900 * NM), 1.0e-16
        cfg = rb.tip_config(500.0, 900 * NM, 100 * NM, 100 * NM, box, dt, True, capacity=n + 8)
        p = orc.params_tip(500.0, 900 * NM, 100 * NM, 100 * NM, box, dt, True)
        pos = np.stack([rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(110, 890, n)], axis=1) * NM
    vel = np.stack([rng.normal(0, 3e4, n), rng.normal(0, 3e4, n), rng.normal(2e5, 4e5, n)], axis=1)
    species = np.where((np.arange(n) % 7) == 6, 2, 1).astype(np.int32)
    if n > 10:
        species[5] = 3  # an atom: no contribution (src/mod_verlet.F90:463)
    sec = rng.integers(1, nsec + 1, n).astype(np.int32)
    acc = np.zeros((n, 3))
    q = np.where(species == 2, Q_0, np.where(species == 1, -Q_0, 0.0))
    m = np.where(species == 2, M_N2P, M_0)
    st = orc.store(n + 8)
    for i in range(n):
        st.add(p, pos[i], vel[i], int(species[i]), 0, 1, -1, int(sec[i]))
    st.acc[:] = 0.0; st.acc_prev[:] = 0.0; st.acc_prev2[:] = 0.0
    st.update_velocity(p)
    want = st.ramo_current_emit(nsec)
    got = []
    for rep in range(2):
        with rb.HotPath(cfg) as hp:
            hp.set_option("ramo_sections", nsec)
            hp.upload(pos, q, m, vel=vel, acc=acc, acc_prev=acc, acc_prev2=acc, species=species, section=sec)
            r = hp.Update_Particle_Velocity()
            got.append(hp.ramo_current_emit(nsec))
            wide = hp.ramo_current_emit(nsec + 3)
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(wide[:nsec], got[0]) and np.all(wide[nsec:] == 0.0)
    scale = np.max(np.abs(want))
    assert scale > 0.0
    assert np.max(np.abs(got[0] - want)) <= 1e-12 * scale * max(1.0, math.sqrt(n / nsec))
    tot = r.ramo_current[1] + r.ramo_current[2]
    assert got[0].sum() == pytest.approx(tot, rel=1e-10)
    with rb.HotPath(cfg) as hp:  # off by default: asking for the table is an error, the step itself is unchanged
        hp.upload(pos, q, m, vel=vel, acc=acc, acc_prev=acc, acc_prev2=acc, species=species, section=sec)
        r2 = hp.Update_Particle_Velocity()
        assert r2.ramo_current[1] == r.ramo_current[1]
        with pytest.raises(rb.api.Rb2Error):
            hp.ramo_current_emit(nsec)


@pytest.mark.parametrize("n", [700, 5000, 20000])
def test_step_graph_replay_is_bit_identical(orc, n):
    """rb2_step replayed as a CUDA graph (captured once two consecutive steps queue identical work) against plain
    launches: bit-identical trajectories, records and Ramo currents; a change of the particle count or of a scalar
    drops the graph and the next pair of identical steps captures a new one."""
    out = []
    for graph in (1, 0):
        cfg, p = planar(orc, cap=n + 64, planes=(300 * NM, 600 * NM))
        pos, q, m, sp = cloud(n, 31 + n, ions=True, zmin=5.0)
        vel = np.zeros((n, 3)); vel[:, 2] = np.linspace(-3e6, 3e6, n)
        with rb.HotPath(cfg) as hp:
            hp.set_option("step_graph", graph)
            hp.set_option("ramo_sections", 3)
            hp.upload(pos, q, m, vel=vel, species=sp, section=(1 + np.arange(n) % 3).astype(np.int32))
            log = []
            for step in range(1, 13):
                r = hp.Update_Position(step)
                log.append((r.n_events, tuple(r.ramo_current), tuple(hp.ramo_current_emit(3)), tuple(sorted((e["index"], e["kind"], e["plane"]) for e in hp.events()))))
                if step == 6:   # absorbed particles leave: another particle count, the graph of steps 3-6 is dropped
                    hp.Remove_Particles(step)
                if step == 9:   # a scalar passed by value changes
                    cfg.E_z *= 1.5
                    hp.update_config(cfg)
            out.append((log, hp.download(("pos", "vel", "acc", "acc_prev", "acc_prev2", "charge")), hp.stat("graph_replays")))
    (lg, dg, rg), (l0, d0, r0) = out
    assert rg >= 3 and r0 == 0
    assert lg == l0
    for key in dg:
        assert np.array_equal(dg[key], d0[key]), key
    assert sum(x[0] for x in lg) > 0  # plane crossings / absorptions did happen


def test_beeman_update_is_bit_exact_without_pair_forces(orc):
    """With the pair forces off (single species far apart is not needed: use q = 0 atoms-free
    trick: one particle per run) the position/velocity arithmetic itself is bit-identical."""
    d, V, dt = 300 * NM, 30.0, 1.0e-16
    cfg = rb.planar_config(V, d, (100 * NM, 100 * NM, d), dt, False, 0, capacity=8)
    p = orc.params_planar(V, d, (100 * NM, 100 * NM, d), dt, False, 0)
    rng = np.random.default_rng(1)
    for trial in range(5):
        pos = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(1, 299)]) * NM
        vel = rng.normal(0, 1e5, 3)
        st = orc.store(8)
        st.add(p, pos, vel, 1, 0, 1)
        with rb.HotPath(cfg) as hp:
            hp.Add_Particle(pos, vel, 1, 0, 1)
            for step in range(1, 40):
                st.step(p)
                hp.Update_Position(step)
            s = hp.download(("pos", "vel", "acc", "prev_pos"))
        assert np.array_equal(s["pos"], st.pos) and np.array_equal(s["vel"], st.vel)
        assert np.array_equal(s["acc"], st.acc) and np.array_equal(s["prev_pos"], st.prev_pos)


# --------------------------------------------------------------------------------------------------
def test_large_cloud_properties(orc):
    """BASELINE sizes: rows sampled against the long-double oracle, Newton's third law on the
    Coulomb-only variant, and linearity of the field in the sources."""
    n = 100_000
    cfg, p = planar(orc, ic=True, nic=1, cap=n)
    pos, q, m, sp = cloud(n, 20261017, ions=True)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        info = hp.last_accel_info()
        assert info["ms"] > 0 and info["grid_x"] * info["grid_y"] >= 148
        # 1200 rows against the long-double oracle: fixed ones (tile edges, both target sub-sets of a superblock, the
        # last partial tile, ions) and a seeded random draw over the whole index range
        rng = np.random.default_rng(7)
        rows = np.unique(np.concatenate([[0, 1, 127, 128, 129, 255, 256, 383, 384, 5000, 49999, 50000, 77777, n - 33, n - 2, n - 1],
                                         np.arange(9, n, 1009), rng.integers(0, n, 1100)])).astype(np.int32)
        assert rows.size >= 1000 and np.any(sp[rows] == SPECIES_ION)
        truth = orc.accel_gather_ld_rows(p, pos, q, m, rows)
        assert relerr(acc[rows], truth) < TOL
        # every row: the pair-symmetric kernel (default at this size) against the independent gather kernel
        hp.set_option("pair_mode", 1)
        hp.Calculate_Acceleration_Particles()
        acc_g = hp.download(("acc",))["acc"]
        assert not np.array_equal(acc, acc_g)  # two different kernels (summation orders) did run
        assert relerr(acc, acc_g) < 1e-12
        assert relerr(acc_g[rows], truth) < TOL
        hp.set_option("pair_mode", 0)
        # Coulomb only, no vacuum field: sum_i m_i a_i = 0
        cfg0 = rb.planar_config(0.0, 1000 * NM, (1000 * NM, 1000 * NM, 1000 * NM), 1e-16, False, 0, capacity=n)
        hp.update_config(cfg0)
        hp.Calculate_Acceleration_Particles()
        a0 = hp.download(("acc",))["acc"]
        f = a0 * m[:, None]
        assert np.linalg.norm(f.sum(axis=0)) < 1e-9 * np.abs(f).sum()


# --------------------------------------------------------------------------------------------------
def test_full_size_cloud_properties(orc):
    """BASELINE.json's largest configuration (N = 1e6, planar image charges, N_ic_max = 1), where the oracle cannot
    evaluate every row: rows sampled against the long-double oracle (1e-11), bit-identical repetition (fixed summation
    order, no atomics), Newton's third law on the Coulomb-only variant, and Add / Mark / Remove bookkeeping at scale."""
    n = 1_000_000
    cfg, p = planar(orc, ic=True, nic=1, cap=n + 16)
    pos, q, m, sp = cloud(n, 20261017, ions=True)
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        hp.Calculate_Acceleration_Particles()
        acc = hp.download(("acc",))["acc"]
        assert np.all(np.isfinite(acc))
        # 300 rows against the long-double oracle, spread over the whole index range (every band of source tiles, both
        # target sub-sets of a superblock, the last partial tile of 64 particles, ions)
        rng = np.random.default_rng(11)
        rows = np.unique(np.concatenate([[0, 127, 128, 255, 256, 16383, 16384, 500_000, 777_777, n - 129, n - 65, n - 64, n - 1],
                                         np.arange(9, n, 5003), rng.integers(0, n, 100)])).astype(np.int32)
        assert rows.size >= 256 and np.any(sp[rows] == SPECIES_ION)
        truth = orc.accel_gather_ld_rows(p, pos, q, m, rows)
        assert relerr(acc[rows], truth) < TOL
        hp.Calculate_Acceleration_Particles()
        assert np.array_equal(hp.download(("acc",))["acc"], acc)
        # stable compaction at scale: remove every 7th particle, the survivors keep their order and ids
        kill = np.arange(3, n, 7, dtype=np.int32)
        hp.Mark_Particles_Remove(kill, REMOVE_TOP)
        k = hp.Remove_Particles(1)
        keep = np.ones(n, bool); keep[kill] = False
        assert k.nrPart == int(keep.sum())
        st = hp.download(("pos", "id", "charge"))
        assert np.array_equal(st["id"], np.arange(n, dtype=np.int32)[keep])
        assert np.array_equal(st["pos"], pos[keep]) and np.array_equal(st["charge"], q[keep])
        # Coulomb only, no vacuum field: sum_i m_i a_i = 0
        cfg0 = rb.planar_config(0.0, 1000 * NM, (1000 * NM, 1000 * NM, 1000 * NM), 1e-16, False, 0, capacity=n + 16)
        hp.update_config(cfg0)
        hp.Calculate_Acceleration_Particles()
        f = hp.download(("acc",))["acc"] * m[keep][:, None]
        assert np.linalg.norm(f.sum(axis=0)) < 1e-9 * np.abs(f).sum()


# --------------------------------------------------------------------------------------------------
def test_empty_and_overflowing_store(orc):
    """Edge cases the reference guards: an empty system steps to nothing (mod_verlet.F90:123-162 with
    nrPart = 0), zero field points are a no-op (:1658), Add_Particle beyond MAX_PARTICLES drops the particle
    and counts it (mod_pair.F90:37-43), and bad indices are refused instead of touching memory."""
    d = 100 * NM
    cfg = rb.planar_config(2.0, d, (100 * NM, 100 * NM, d), 0.25e-15, True, 1, capacity=4)
    with rb.HotPath(cfg) as hp:
        r = hp.Update_Position(1)
        assert r.counts.nrPart == 0 and r.n_events == 0 and list(r.ramo_current) == [0.0] * 4
        assert hp.Remove_Particles(1).nrPart == 0
        hp.Calculate_Acceleration_Particles()
        assert hp.Calc_Field_at_Batch(np.zeros((0, 3))).shape == (0, 3)
        assert hp.field_surface_z(np.zeros((0, 3))).shape == (0,)
        assert hp.download(("pos",))["pos"].shape == (0, 3)
        pos = np.array([[0.0, 0.0, 10.0 * (k + 1)] for k in range(6)]) * NM
        hp.Add_Particles(pos, np.zeros((6, 3)), np.ones(6, dtype=np.int32), 1)
        k = hp.counts()
        assert (k.nrPart, k.nrElec, k.nrPart_dropped) == (4, 4, 2)
        got = hp.download(("pos",))["pos"]
        assert np.array_equal(got, pos[:4])
        for bad in (-1, 4, 1000):
            with pytest.raises(rb.Rb2Error):
                hp.Mark_Particles_Remove(bad, REMOVE_TOP)
        hp.Update_Position(2)          # the full store still steps
        assert hp.counts().nrPart == 4
        hp.Mark_Particles_Remove(0, REMOVE_TOP)
        assert hp.Remove_Particles(2).nrPart == 3
        hp.Add_Particle(pos[5], [0, 0, 0], SPECIES_ELEC, 3, 1)   # room again
        k = hp.counts()
        assert (k.nrPart, k.nrPart_dropped) == (4, 2)


def test_many_field_points_and_determinism(orc):
    """M far above one wave of CTAs; repeated evaluation is bit-identical (fixed summation order)."""
    cfg, p = planar(orc, ic=True, nic=1)
    pos, q, m, sp = cloud(3000, 11)
    rng = np.random.default_rng(3)
    M = 70000
    pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), rng.uniform(0, 900, M)], axis=1) * NM
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        a = hp.Calc_Field_at_Batch(pts)
        b = hp.Calc_Field_at_Batch(pts)
        hp.Calculate_Acceleration_Particles()
        acc1 = hp.download(("acc",))["acc"]
        hp.Calculate_Acceleration_Particles()
        acc2 = hp.download(("acc",))["acc"]
    assert np.array_equal(a, b) and np.array_equal(acc1, acc2)
    idx = rng.choice(M, 64, replace=False)
    want = np.stack([orc.calc_field_at(p, pos, q, pts[k], sp, ld=True) for k in idx])
    assert relerr(a[idx], want) < TOL


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 5000, 20011])
def test_nearest_electron_sweep_is_bit_exact(orc, n):
    """Sample_Elec_Position (mod_pair.F90:975-1037): nearest other electron of every electron -- distance and
    index identical to the serial scan (ions do not count, electrons marked for removal still do, ties go to the
    lowest index: duplicated positions and a lattice patch provoke them)."""
    cfg, p = planar(orc, cap=max(n, 16))
    pos, q, m, sp = cloud(n, 31 + n)
    if n >= 1000:
        pos[10] = pos[3]                      # coincident electrons: distance exactly 0
        pos[500:520] = pos[400:420]           # more exact ties
        g = np.arange(64)
        pos[600:664] = np.stack([(g % 4) * 2.0, ((g // 4) % 4) * 2.0, 300.0 + (g // 16) * 2.0], axis=1) * NM  # lattice: equal distances
    with rb.HotPath(cfg) as hp:
        hp.upload(pos, q, m, species=sp)
        if n >= 1000:
            hp.Mark_Particles_Remove(7, REMOVE_TOP)   # still an electron until Remove_Particles
        dist, idx = hp.Sample_Elec_Position()
    dist_o, idx_o = orc.nearest_elec(pos, sp)
    assert np.array_equal(idx, idx_o)
    assert np.array_equal(dist, dist_o)
    elec = sp == SPECIES_ELEC
    assert np.all(dist[~elec] == 1000.0) and np.all(idx[~elec] == -1)
    if n >= 1000:
        assert dist[10] == 0.0 and idx[10] == 3 and idx[3] == 10
