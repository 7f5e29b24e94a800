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


# ---- the pair-symmetric deal of work units (the multi-GPU path bench.py runs): host-only planning, CPU test ----------
def _plan_worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rumdeed_b200.api import sym_plan_probe
    p = sym_plan_probe(n, world=world, rank=rank)
    # every rank computes the same owner table and the same cost per rank ...
    h = torch.tensor([p["table_hash"] & 0x7FFFFFFFFFFFFFFF] + p["rank_cost"], dtype=torch.int64)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    assert all(torch.equal(hs[0], x) for x in hs)
    # ... and the ranks' unit lists are disjoint and cover every non-empty unit of the triangle
    cnt = torch.tensor([len(p["units"])], dtype=torch.int64)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    m = int(max(c.item() for c in cnts))
    mine = torch.full((m, 3), -1, dtype=torch.int64)
    mine[: len(p["units"])] = torch.from_numpy(p["units"])
    alls = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(alls, mine)
    if rank == 0:
        units = np.concatenate([a.numpy()[: int(c.item())] for a, c in zip(alls, cnts)])
        np.save(os.path.join(out_dir, "units.npy"), units)
        np.save(os.path.join(out_dir, "shape.npy"), np.array([p[k] for k in ("T", "K", "G", "Wb", "nsb", "nIb")] + p["rank_cost"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(3500, 2), (100_000, 2), (100_000, 8), (1_000_000, 8)])
def test_pair_symmetric_units_dealt_over_ranks(tmp_path, n, world):
    """rb2_sym_plan_probe on every rank of a gloo group: same table everywhere, disjoint unit lists whose union is the
    whole triangle of (target superblock set, source tile group) units, cost per rank within 1 % (plus one unit)."""
    mp.spawn(_plan_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    units = np.load(tmp_path / "units.npy")
    T, K, G, Wb, nsb, nIb, *cost = np.load(tmp_path / "shape.npy").tolist()
    assert len({tuple(u) for u in units}) == len(units)            # no unit twice
    want = set()
    total = 0
    for b, b0 in enumerate(range(0, nsb, Wb)):
        blen = min(Wb, nsb - b0)
        nI = (b0 + blen - 1) // T + 1
        for s in range((nI + K - 1) // K):
            for g in range((blen + G - 1) // G):
                J0, J1 = b0 + g * G, min(b0 + g * G + G, b0 + blen)
                c = 0
                for I in range(s * K, min(s * K + K, nIb)):
                    Jb = max(J0, T * I)
                    if Jb >= J1:
                        break
                    c += T * (J1 - Jb) - sum(T - 1 - (J - T * I) for J in range(Jb, min(J1, T * I + T)))
                if c > 0:
                    want.add((b, s, g))
                    total += c
    assert {tuple(u) for u in units} == want
    assert sum(cost) == total == nsb * (nsb + 1) // 2                # every (tile, tile) pair of the triangle once
    biggest_unit = K * G * T
    assert max(cost) - min(cost) <= max(0.01 * total / world, biggest_unit * (nsb + Wb - 1) // Wb)
