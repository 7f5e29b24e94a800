"""Multi-process test of the peer-memory exchange (rb2_p2p_*): the pair work of the pair-symmetric kernel split
over two processes, partial sums read from the peer's exchange block inside the finalise kernel (CUDA IPC; NVLink
when the box has two GPUs, the same device otherwise).  Checked against the unsplit single-process run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, n, steps, outdir):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "p2p_worker.py"), str(r), str(world), str(port), str(n),
                               str(steps), str(outdir)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=240)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
    return [np.load(os.path.join(outdir, f"rank{r}of{world}.npz")) for r in range(world)]


def _rel(a, ref):
    return float(np.max(np.linalg.norm(a - ref, axis=-1) / np.maximum(np.linalg.norm(ref, axis=-1), 1.0)))


@pytest.mark.parametrize("n", [4000, 20001])
def test_p2p_exchange_two_ranks_vs_single(tmp_path, n):
    steps = 3
    single = _run(1, n, steps, tmp_path)[0]
    r0, r1 = _run(2, n, steps, tmp_path)
    for k in ("acc0", "pos", "vel", "acc", "ramo"):
        assert np.array_equal(r0[k], r1[k]), f"{k}: the replicas diverged"  # same rank-order sum everywhere
    assert _rel(r0["acc0"], single["acc0"]) < 1e-13
    assert _rel(r0["acc"], single["acc"]) < 1e-12
    assert np.allclose(r0["pos"], single["pos"], rtol=1e-14, atol=0)
    assert np.allclose(r0["ramo"], single["ramo"], rtol=1e-12, atol=0)
