"""One process driving several GPUs (rb2_set_devices): the single-process Fortran host's way to more than one device
(src/main.F90:5-29).  The two-device cases need a box with two GPUs (gpurun --gpus 2) and are skipped otherwise."""
import numpy as np
import pytest

import rumdeed_b200 as rb
from rumdeed_b200.api import M_0, M_N2P, Q_0

pytestmark = pytest.mark.gpu
NM = 1.0e-9


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(-500, 500, n), rng.uniform(-500, 500, n), rng.uniform(1, 999, n)], axis=1) * NM
    ion = (np.arange(n) % 10) == 9
    return pos, np.where(ion, Q_0, -Q_0), np.where(ion, M_N2P, M_0), np.where(ion, 2, 1).astype(np.int32)


def test_device_list_is_validated():
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=1024)
    with rb.HotPath(cfg) as hp:
        hp.set_devices([0])  # one device: nothing to do
        with pytest.raises(rb.api.Rb2Error):
            hp.set_devices([1, 0])          # the first entry must be the device of rb2_init
        with pytest.raises(rb.api.Rb2Error):
            hp.set_devices([0, 0])
        with pytest.raises(rb.api.Rb2Error):
            hp.set_devices([0, 99])
        pos, q, m, sp = _cloud(100, 1)
        hp.upload(pos, q, m, species=sp)
        if _ngpu() >= 2:
            with pytest.raises(rb.api.Rb2Error):
                hp.set_devices([0, 1])      # too late: particles exist
        hp.Calculate_Acceleration_Particles()  # the context is still usable


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n", [5000, 20001])
def test_two_devices_from_one_process(n):
    """The same run on one device and on two devices of one process: accelerations within 1e-13 (the two partial sums
    are added in device order), identical bookkeeping, records and Ramo currents to rounding over several steps with
    additions and removals."""
    pos, q, m, sp = _cloud(n, 7 + n)
    vel = np.zeros((n, 3)); vel[:, 2] = np.linspace(-2e6, 2e6, n)
    runs = []
    for devices in ([0], [0, 1]):
        cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n + 64, planes_z=(400 * NM,))
        with rb.HotPath(cfg) as hp:
            hp.set_option("pair_mode", 2)
            hp.set_devices(devices)
            hp.upload(pos, q, m, vel=vel, species=sp)
            hp.Calculate_Acceleration_Particles()
            acc0 = hp.download(("acc",))["acc"]
            host = hp.accel_host(pos, q, m)
            log = []
            for step in range(1, 6):
                r = hp.Update_Position(step)
                log.append((r.n_events, r.counts.nrPart_remove, r.ramo_current[1], r.ramo_current[2], [e["index"] for e in hp.events()]))
                k = hp.Remove_Particles(step)
                if step == 2:
                    hp.Add_Particles(pos[:5] + [0, 0, 1e-9], np.zeros((5, 3)), np.ones(5, dtype=np.int32), step)
            st = hp.download(("pos", "vel", "id", "charge"))
            runs.append((acc0, host, log, st, k.nrPart))
    (a1, h1, l1, s1, n1), (a2, h2, l2, s2, n2) = runs
    scale = np.maximum(np.linalg.norm(a1, axis=1), 1.0)
    assert np.max(np.linalg.norm(a2 - a1, axis=1) / scale) < 1e-13
    assert np.max(np.linalg.norm(h2 - a2, axis=1) / scale) < 1e-13 and np.array_equal(h1, a1)
    assert n1 == n2 and np.array_equal(s1["id"], s2["id"]) and np.array_equal(s1["charge"], s2["charge"])
    assert np.allclose(s1["pos"], s2["pos"], rtol=1e-12, atol=1e-22) and np.allclose(s1["vel"], s2["vel"], rtol=1e-10, atol=1e-6)
    for x, y in zip(l1, l2):
        assert x[0] == y[0] and x[1] == y[1] and x[4] == y[4]
        assert x[2] == pytest.approx(y[2], rel=1e-10) and x[3] == pytest.approx(y[3], rel=1e-10, abs=1e-30)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_field_batches_are_split_over_the_devices():
    """With several devices in the process a large batch of field points is dealt out in slices (every device holds the
    whole store): equal to the one-device result to rounding (the split of the particle range into chunks, i.e. the
    summation order, depends on the number of points of a launch), for rb2_field_batch and rb2_field_surface_z."""
    n, M = 30000, 5000
    pos, q, m, sp = _cloud(n, 3)
    rng = np.random.default_rng(9)
    pts = np.stack([rng.uniform(-500, 500, M), rng.uniform(-500, 500, M), np.zeros(M)], axis=1) * NM
    out = []
    for devices in ([0], [0, 1]):
        cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM,) * 3, 1e-16, True, 1, capacity=n + 64)
        with rb.HotPath(cfg) as hp:
            hp.set_devices(devices)
            hp.upload(pos, q, m, species=sp)
            out.append((hp.Calc_Field_at_Batch(pts), hp.field_surface_z(pts), hp.Calc_Field_at_Batch(pts[:100])))
    for a, b in zip(out[0], out[1]):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-13 * np.max(np.abs(a)))
    assert np.array_equal(out[0][2], out[1][2])  # a small batch stays on the first device
