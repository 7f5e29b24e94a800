"""Worker of tests/test_gpu_p2p.py: one process per rank (one GPU each when the box has enough, else sharing
cuda:0 -- CUDA IPC and the flag protocol work the same, the kernels of the two processes are time-sliced).
Exchanges the IPC handles over gloo, runs MD steps with the pair work split over the ranks and the partial sums
exchanged over peer memory, and saves the final state."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, port, n, steps, outdir = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    import torch
    import torch.distributed as dist

    import rumdeed_b200 as rb
    from rumdeed_b200.api import M_0, M_N2P, Q_0

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    nm = 1.0e-9
    rng = np.random.default_rng(np.random.PCG64(99 + n))
    pos = np.stack([rng.uniform(-500, 500, n), rng.uniform(-500, 500, n), rng.uniform(1, 999, n)], axis=1) * nm
    ion = (np.arange(n) % 10) == 9
    q = np.where(ion, Q_0, -Q_0)
    m = np.where(ion, M_N2P, M_0)
    cfg = rb.planar_config(2000.0, 1000 * nm, (1000 * nm,) * 3, 1.0e-16, True, 1, capacity=n, device=dev)
    with rb.HotPath(cfg) as hp:
        hp.set_option("pair_mode", 2)
        hp.set_option("sym_budget_mb", 4.0)  # several bands
        hp.upload(pos, q, m)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, hp.p2p_export(n))
            hp.p2p_attach(world, rank, handles)
        hp.Calculate_Acceleration_Particles()
        acc0 = hp.download(("acc",))["acc"]
        for s in range(steps):
            r = hp.Update_Position(s + 1)
        out = hp.download(("pos", "vel", "acc"))
        np.savez(os.path.join(outdir, f"rank{rank}of{world}.npz"), acc0=acc0, ramo=np.array(r.ramo_current), **out)
        dist.barrier()  # nobody unmaps while a peer may still read
        if world > 1:
            hp.p2p_detach()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
