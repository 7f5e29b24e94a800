"""World-size-2 gloo test (CPU) of the multi-GPU host logic: equal-chunk i-partition, in-place
all-gather of the (3,N) acceleration slices, identical result on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from rumdeed_b200.partition import row_partition
    orc = Oracle()
    nm = 1e-9
    rng = np.random.default_rng(7)
    pos = np.stack([rng.uniform(-50, 50, n), rng.uniform(-50, 50, n), rng.uniform(1, 999, n)], axis=1) * nm
    q = np.where(np.arange(n) % 4 == 3, orc.k.q_0, -orc.k.q_0)
    m = np.where(np.arange(n) % 4 == 3, orc.k.m_N2p, orc.k.m_0)
    p = orc.params_planar(2000.0, 1000 * nm, (100 * nm, 100 * nm, 1000 * nm), 1e-16, True, 1)
    chunk, i0, i1, cap = row_partition(n, world, rank)
    acc = torch.zeros(3 * cap, dtype=torch.float64)            # the (3, capacity) buffer
    rows = orc.accel_gather_ld(p, pos, q, m, i0, i1)           # this rank's rows only (global indices)
    acc[3 * i0: 3 * i1] = torch.from_numpy(rows.reshape(-1))
    dist.all_gather_into_tensor(acc, acc[3 * rank * chunk: 3 * (rank + 1) * chunk].clone())
    full = orc.accel_gather_ld(p, pos, q, m)
    ok = np.array_equal(acc[: 3 * n].numpy().reshape(n, 3), full)
    np.save(os.path.join(out_dir, f"acc_{rank}.npy"), acc[: 3 * n].numpy())
    assert ok
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 101])
def test_partition_allgather_world2(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    a0 = np.load(tmp_path / "acc_0.npy")
    a1 = np.load(tmp_path / "acc_1.npy")
    assert np.array_equal(a0, a1)


def test_row_partition_covers_everything():
    from rumdeed_b200.partition import row_partition
    for n in (0, 1, 7, 8, 1000, 1_000_000):
        for world in (1, 2, 4, 8):
            seen = 0
            for r in range(world):
                chunk, i0, i1, cap = row_partition(n, world, r)
                assert i0 == min(n, r * chunk) and i0 <= i1 <= n and cap >= n and cap % world == 0 or n == 0
                seen += i1 - i0
            assert seen == n
