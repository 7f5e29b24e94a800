#!/usr/bin/env python
"""bench.py -- RUMDEED hot-path benchmark (contract: see the task statement / DESIGN.md section 6).

Workload (BASELINE.json configs[4]): synthetic planar-diode electron cloud,
d = 1000 nm, V = 2000 V, image_charge on, N_ic_max = 1, dt = 1e-4 ps, N electrons i.i.d.
uniform in x,y in [-500,500] nm, z in [1,999] nm, NumPy PCG64(20261017).  Default N = 1e6.

A step is one MD step of the hot path: Beeman position update + boundary/plane checks,
all-pairs Coulomb + image-charge acceleration, velocity update + Ramo current, with the
particle state resident in HBM.  value = N(N-1) ordered pair interactions per step / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n PARTICLES] [--impl ours|reference]

N > 1: one process per GPU (torchrun); i-particles are partitioned, every rank integrates
a replica of the O(N) state and the acceleration slices are all-gathered over NCCL.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NM = 1.0e-9
FLOPS_PER_PAIR = {(-1): 21, 0: 39, 1: 99, 2: 159}  # 21 + ic*(18 + 60*N_ic_max)
# What the kernels EXECUTE at N_ic_max = 1 (ncu opcode counts, profiles/ncu_opmix_r02_n1e5.txt): FP64-pipe instructions and
# flops (DFMA = 2, DMUL / DADD = 1) per pair evaluation.  The pair-symmetric kernel evaluates an UNORDERED pair
# once (both ordered interactions), the gather kernel an ordered pair.
EXECUTED = {"sym": {"fp64_instr": 75.57, "flops": 116.8, "per": "unordered pair"},
            "gather": {"fp64_instr": 71.04, "flops": 108.0, "per": "ordered pair"}}
SURFACE_FLOPS_PER_POINT_PARTICLE = 54  # k_surface_field at N_ic_max = 1: 31 FP64 instructions, 23 of them FMAs (+ 1 compare)
METRIC = "pair_interactions_per_s"
UNIT = "pair-interactions/s"


def make_cloud(n, seed=20261017):
    rng = np.random.default_rng(np.random.PCG64(seed))
    pos = np.empty((n, 3))
    pos[:, 0] = rng.uniform(-500.0, 500.0, n) * NM
    pos[:, 1] = rng.uniform(-500.0, 500.0, n) * NM
    pos[:, 2] = rng.uniform(1.0, 999.0, n) * NM
    return pos


def flops_per_pair(image_charge, nic):
    return 21 + (18 + 60 * nic if image_charge else 0)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), read
    in-process through NVML from a sampling thread.  (Spawning `nvidia-smi -lms` instead stalls the GPUs for tens
    of milliseconds at start-up and per query once peer mappings exist -- measured: 13.3 -> 18-22 ms per step at
    N = 1e5 on 2 GPUs -- so it is only the fallback when NVML cannot be loaded.)"""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0, period_s=0.25):
        self.gpu = gpu_index
        self.period = period_s
        self.samples = []
        self.reasons = set()
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None
        self.handle = None
        self.mx = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: match by PCI bus id
            try:
                import torch
                bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id
                dom = getattr(torch.cuda.get_device_properties(gpu_index), "pci_domain_id", 0)
                dev = torch.cuda.get_device_properties(gpu_index).pci_device_id
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self._sample()  # first use of each query is slow (tens of ms, and it blocks kernel launches): do it now
            self.samples.clear()
            self.reasons.clear()
        except Exception:
            self.nvml = None

    def _sample(self):
        nv = self.nvml
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.samples.append((sm, pw))
        for name, bit in self.REASONS:
            if bits & bit:
                self.reasons.add(name)

    def _loop(self):
        # a query holds a driver lock that kernel launches also take (a launch-bound step at N = 1e4 stalls for
        # milliseconds per query), so sample sparsely: the first one half a period in, then every period
        if self.stop_flag.wait(0.5 * self.period):
            return
        while True:
            try:
                self._sample()
            except Exception:
                pass
            if self.stop_flag.wait(self.period):
                return

    def start(self):
        if self.nvml is None:
            return
        self.samples.clear()
        self.reasons.clear()
        self.stop_flag.clear()
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": "NVML, in-process sampling thread"}
        if self.thread is None:
            out["how"] = "NVML unavailable"
            return out
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            # region shorter than the sampling period: one sample right behind the last timed step
            try:
                self._sample()
                out["how"] += " (timed region shorter than the sampling period: one sample taken right behind it)"
            except Exception:
                pass
        if self.samples:
            sm = [x[0] for x in self.samples]
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=self.mx, power_w_max=float(max(x[1] for x in self.samples)),
                       reasons=sorted(self.reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
def load_fast_oracle():
    """The CPU restatement built with the reference's flags (-O3 -march=native -fopenmp).
    Rebuilt natively on this box when gcc is here, else the prebuilt x86-64-v3 file."""
    from oracle import oracle as om
    path = None
    try:
        td = tempfile.mkdtemp(prefix="rb2_oracle_")
        om.build(force=True, march="native", out_dir=td)
        path = os.path.join(td, "liboracle_fast.so")
        if not os.path.exists(path):
            path = None
    except Exception:
        path = None
    if path is None:
        om.build()
        return om.Oracle(fast=True), "prebuilt -O3 -march=x86-64-v3 -fopenmp"
    return om.Oracle(path=path), "-O3 -march=native -fopenmp (built on this host)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_stride(n):
    """The SAME bounded sample in both CPU legs (cpu_baseline and --impl reference): every row up to N = 2e5 (5.5 s at
    1e5 on 16 cores), every 64th row of the i<j loop at N = 1e6 (a fixed 1/64 slice, ~9 s on 16 cores)."""
    return 1 if n <= 200_000 else max(1, int(round(64 * (n / 1.0e6) ** 2)))


def cpu_sample(orc, p, pos, q, m):
    """Time a bounded, strided sample of the rows of the reference's CPU pair loop
    (Calculate_Acceleration_Particles_Planar, i<j scatter) on all host cores.  Returns ordered pair-interactions/s."""
    n = pos.shape[0]
    orc.set_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1: say what we use
    stride = cpu_stride(n)
    total_pairs = n * (n - 1) / 2
    t0 = time.perf_counter()
    _, pairs = orc.accel_planar_rows(p, pos, q, m, 0, n, stride)
    t = time.perf_counter() - t0
    frac = pairs / total_pairs
    sample = (f"rows i = 0, {stride}, 2*{stride}, ... of the i<j pair loop at N = {n} "
              f"({pairs} unordered pair evaluations = {100 * frac:.3g}% of a full evaluation, {t:.1f} s)")
    return 2.0 * pairs / t, t, sample


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Fortran binary cannot
    be built in this image) on the host cores, same metric/config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, how = load_fast_oracle()
    n = args.n
    pos = make_cloud(n)
    q = np.full(n, -orc.k.q_0)
    m = np.full(n, orc.k.m_0)
    p = orc.params_planar(2000.0, 1000 * NM, (1000 * NM, 1000 * NM, 1000 * NM), 1.0e-16, True, args.nic)
    vals, times, sample = [], [], ""
    for it in range(args.warmup + args.steps):
        v, t, sample = cpu_sample(orc, p, pos, q, m)
        if it >= args.warmup:
            vals.append(v); times.append(t)
    cores = orc.max_threads()
    value = float(np.mean(vals))
    ms_full = n * (n - 1) / value * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_full, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + f"; {how}; ms_per_step is the full-step extrapolation"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"synthetic planar-diode electron cloud N={args.n}, d=1000nm, V=2000V, image_charge, "
                        f"N_ic_max={args.nic}, dt=1e-4ps (BASELINE.json configs[4])",
            "n_particles": args.n, "N_ic_max": args.nic, "image_charge": True,
            "flops_per_pair": flops_per_pair(True, args.nic),
            "parallelism": (f"pair work units x{world}" + ((" + peer-memory (NVLink) exchange of the partial pair sums fused into the finalise kernel"
                                                            if getattr(args, "exchange", "p2p") == "p2p" else " + NCCL all-reduce of the partial pair sums") if world > 1 else ""))
            if (getattr(args, "pair_mode", "auto") == "sym" or (getattr(args, "pair_mode", "auto") == "auto" and args.n >= 3500))
            else (f"i-partition x{world}" + (" + NCCL all-gather of accelerations" if world > 1 else "")),
            "l2": "256 MiB L2-flush write between timed steps (outside the per-step CUDA-event pairs)"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import rumdeed_b200 as rb
    from rumdeed_b200.api import M_0, Q_0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    pos = make_cloud(n)
    q = np.full(n, -Q_0)
    m = np.full(n, M_0)
    from rumdeed_b200.partition import row_partition
    chunk, i0, i1, cap = row_partition(n, world, rank)
    cfg = rb.planar_config(2000.0, 1000 * NM, (1000 * NM, 1000 * NM, 1000 * NM), 1.0e-16, True, args.nic,
                           capacity=cap, device=local)
    hp = rb.HotPath(cfg)
    ngpu = world
    if args.single_process and world == 1 and args.gpus > 1:
        hp.set_devices(list(range(args.gpus)))  # replicas + split pair work inside this one process
        ngpu = args.gpus
    hp.upload(pos, q, m)
    # pair kernel: "sym" = each unordered pair once (default from 3500 particles on), "gather" = ordered pairs
    sym = (args.pair_mode == "sym") or (args.pair_mode == "auto" and n >= 3500)
    hp.set_option("pair_mode", 2 if sym else 1)
    p2p = sym and world > 1 and args.exchange == "p2p"
    if sym and ngpu == world:
        hp.set_pair_rank(rank, world)   # (target superblock, source group) work units dealt round-robin
        if p2p:
            # the one exchange step over NVLink peer memory, fused into the finalise kernel (rb2_p2p.cu): every rank
            # exports its exchange block as a CUDA IPC handle, the handles are gathered once, every rank maps them
            handles = [None] * world
            dist.all_gather_object(handles, hp.p2p_export(n))
            hp.p2p_attach(world, rank, handles)
    else:
        hp.set_partition(i0, i1)        # contiguous i-rows per rank
    ext = torch.cuda.ExternalStream(hp.stream(), device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    class _Alias:  # expose a device buffer of the library to torch without a copy
        def __init__(self, ptr, nelem):
            self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    def dev_tensor(name):
        ptr, nbytes = hp.device_buffer(name)
        return torch.as_tensor(_Alias(ptr, nbytes // 8), device=f"cuda:{local}")

    def exchange():
        """The one exchange step of the path: partial pair sums are all-reduced (pair-symmetric kernel),
        or the acceleration rows are all-gathered in place (gather kernel)."""
        with torch.cuda.stream(ext):
            if sym:
                dist.all_reduce(dev_tensor("raw"), op=dist.ReduceOp.SUM)
            else:
                t = dev_tensor("acc")
                dist.all_gather_into_tensor(t[: 3 * cap], t[3 * i0: 3 * (i0 + chunk)])

    def one_step(step):
        if world == 1 or p2p:
            return hp.Update_Position(step)
        hp.Update_Particle_Position(step)
        if sym:
            hp.accel_partial()
            exchange()
            hp.accel_finalize()
        else:
            hp.Calculate_Acceleration_Particles()
            exchange()
        return hp.Update_Particle_Velocity()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 peak of this GPU (independent DFMA chains), burst and sustained
    peak_burst, _ = hp.fp64_peak(30.0)
    peak_sust, _ = hp.fp64_peak(1500.0)

    def timed_steps(nsteps, first_step):
        """nsteps MD steps, each bracketed by CUDA events on the library's stream, L2 flushed in between."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        acc_ms = []
        barrier()
        t0 = time.perf_counter()
        for k in range(nsteps):
            if not args.no_flush:
                flush.zero_()                  # L2 flush on torch's stream ...
            torch.cuda.current_stream().synchronize()
            ev[k][0].record(ext)               # ... timed region on the library's launching stream
            one_step(first_step + k)
            ev[k][1].record(ext)
            acc_ms.append(hp.last_accel_info()["ms"])
        barrier()
        wall = time.perf_counter() - t0
        return [a.elapsed_time(b) for a, b in ev], acc_ms, wall

    for w in range(args.warmup):
        one_step(w + 1)
    barrier()
    hp.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        clocks.start()
    step_ms_list, accel_ms, t_wall = timed_steps(args.steps, args.warmup + 1)
    clk = clocks.stop() if rank == 0 else None
    launches = hp.launch_count()
    dev_ms = sum(step_ms_list)
    info = hp.last_accel_info()

    # the accelerations of the last step: a checksum for comparing runs (1 GPU vs 8), and -- the replicas of the O(N)
    # state are integrated redundantly -- bit-identical on every rank of this run
    import hashlib
    acc_last = hp.download(("acc",))["acc"]
    acc_sha = hashlib.sha256(acc_last.tobytes()).hexdigest()[:16]
    replicas_identical = None
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, acc_sha)
        replicas_identical = all(s == shas[0] for s in shas)
    checksum = {"sum_abs": float(np.abs(acc_last).sum()), "sum": [float(x) for x in acc_last.sum(axis=0)], "sha256_16": acc_sha,
                "replicas_identical": replicas_identical,
                "what": "particles_cur_accel after the last timed step; sum_abs agrees between 1 and N GPUs to rounding "
                        "(the ranks' partial sums are added in rank order), the hash is identical on all ranks of one run"}

    # end to end through the C ABI with HOST buffers: pinned pos/q/m in, accelerations out
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_pos, h_q, h_m = pin(pos), pin(q), pin(m)
    h_acc = torch.empty((n, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_once():
        if world == 1 or not sym:
            # stateless C-ABI call: host pos/q/m in, this rank's acceleration rows out
            hp.accel_host_ptr(n, h_pos.data_ptr(), h_q.data_ptr(), h_m.data_ptr(), h_acc.data_ptr())
        elif p2p:
            # split pair work, partial sums exchanged over peer memory inside the finalise kernel
            hp.upload(h_pos.numpy(), h_q.numpy(), h_m.numpy())
            hp.Calculate_Acceleration_Particles()
            h_acc.numpy()[:] = hp.download(("acc",))["acc"]
        else:
            # split pair work: upload, partial sums, all-reduce, finalise, rows back to the host
            hp.upload(h_pos.numpy(), h_q.numpy(), h_m.numpy())
            hp.accel_partial()
            exchange()
            hp.accel_finalize()
            h_acc.numpy()[:] = hp.download(("acc",))["acc"]

    e2e_once()  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps

    kern = "sym" if sym else "gather"

    def contract_and_executed(n_, acc_ms_):
        """Fractions of the measured FP64 peak for one acceleration evaluation of n_ particles in acc_ms_ (slowest rank's
        share): the contract figure (99 algorithmic flops per ORDERED pair) and what the kernel executed."""
        pairs_ = float(n_) * float(n_ - 1) / ngpu
        pk = peak_sust if acc_ms_ > 200.0 else peak_burst
        a = flops_per_pair(True, args.nic) * pairs_ / (acc_ms_ * 1e-3) / 1e12
        ex = EXECUTED[kern]
        evals = pairs_ / 2.0 if kern == "sym" else pairs_
        e = ex["flops"] * evals / (acc_ms_ * 1e-3) / 1e12 if args.nic == 1 else None
        u = ex["fp64_instr"] * evals / (acc_ms_ * 1e-3) / (pk * 1e12 / 2.0) if args.nic == 1 else None
        return a, pk, e, u

    # ---- field batches through the C ABI with HOST buffers (M points on z = 0 against the N particles) ----
    field_batch = None
    if ngpu == 1 and not args.no_sweep:
        field_batch = []
        rngp = np.random.default_rng(np.random.PCG64(20261018))
        for M in (1, 32, 256, 4096, 65536):
            pts = np.zeros((M, 3))
            pts[:, 0] = rngp.uniform(-500.0, 500.0, M) * NM
            pts[:, 1] = rngp.uniform(-500.0, 500.0, M) * NM
            row = {"M": M, "N": n}
            for name, fn, fl in (("rb2_field_batch", hp.Calc_Field_at_Batch, flops_per_pair(True, args.nic)),
                                 ("rb2_field_surface_z", hp.field_surface_z, SURFACE_FLOPS_PER_POINT_PARTICLE)):
                fn(pts)  # warm (buffers grow on the first call)
                reps = 3 if M >= 4096 else 10
                ts = []
                for _ in range(reps):
                    flush.zero_(); torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn(pts)
                    ts.append(time.perf_counter() - t0)
                t = float(np.median(ts))
                row[name] = {"ms": t * 1e3, "point_interactions_per_s": M * float(n) / t,
                             "fp64_frac": (fl * M * float(n) / t / 1e12) / peak_burst,
                             "flops_per_point_particle": fl}
            field_batch.append(row)

    # ---- the rest of the contract sweep (N = 1e4, 1e5) on the same context, same step, same timing ----
    sweep = None
    if not args.no_sweep:
        sweep = []
        for n_s, k_s in ((10_000, 30), (100_000, 8)):
            if n_s >= n:
                continue
            pos_s = make_cloud(n_s)
            hp.upload(pos_s, np.full(n_s, -Q_0), np.full(n_s, M_0))
            for w in range(3):
                one_step(w + 1)
            ms_l, acc_l, _ = timed_steps(k_s, 4)
            t3 = torch.tensor([float(np.median(ms_l)), float(np.mean(ms_l)), float(np.median(acc_l))], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            med, mean, acc_med = [float(x) for x in t3.tolist()]
            a, pk, e, u = contract_and_executed(n_s, acc_med)
            sweep.append({"n_particles": n_s, "steps": k_s, "ms_per_step_median": med, "ms_per_step_mean": mean,
                          "md_steps_per_s": 1e3 / med, "pair_interactions_per_s": float(n_s) * (n_s - 1) / (med * 1e-3),
                          "kernel_ms_median": acc_med, "frac": a / pk, "frac_executed": (e / pk) if e else None,
                          "fp64_pipe_util_est": u})

    stats = torch.tensor([dev_ms, t_wall * 1e3, t_e2e * 1e3, float(np.mean(accel_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    dev_ms, wall_ms, e2e_ms, acc_ms = [float(x) for x in stats.tolist()]

    if rank == 0:
        pairs = float(n) * float(n - 1)
        ms_per_step = dev_ms / args.steps
        value = pairs / (ms_per_step * 1e-3)
        fpp = flops_per_pair(True, args.nic)
        # algorithmic flops of the slowest rank's share of the N(N-1) ordered pair interactions
        achieved, peak, executed, pipe_util = contract_and_executed(n, acc_ms)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get(f"k_pair_n{n}", None)
                traffic_src = "static: " + tj.get("source", "profiles/ncu_traffic.json") + " (an ncu capture of this kernel, not measured in this run)"
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, ngpu),
            "processes": world,
            "md_steps_per_s": 1e3 / ms_per_step, "wall_ms_per_step": wall_ms / args.steps,
            "ms_steps_rank0": [round(x, 4) for x in step_ms_list], "accel_ms_steps_rank0": [round(x, 4) for x in accel_ms],
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "achieved_executed": executed, "frac_executed": (executed / peak) if executed else None,
                         "executed_note": "frac is the contract figure (99 ALGORITHMIC flops per ordered pair, N(N-1) of them); "
                                          "frac_executed counts what the kernel ran (%s: %.0f flops per %s)" % (
                                              kern, EXECUTED[kern]["flops"], EXECUTED[kern]["per"]),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": ("k_pair_sym<N_ic_max=%d> + k_sym_reduce per band, k_sym_finalize (each unordered pair "
                                    "evaluated once and applied to both particles, like the reference's CPU pair loop; "
                                    "achieved = ALGORITHMIC flops of the N(N-1) ordered interactions, so it can exceed "
                                    "the executed-flop peak)" % args.nic) if sym
                         else "k_pair<planar, N_ic_max=%d> + k_accel_finalize (ordered pairs)" % args.nic,
                         "kernel_ms": acc_ms, "flops_per_pair": fpp,
                         "fp64_instr_per_pair_evaluation": EXECUTED[kern]["fp64_instr"] if args.nic == 1 else None,
                         "fp64_pipe_util_est": pipe_util,
                         "peak_source": "measured in this run: rb2_fp64_peak independent-DFMA-chain kernel "
                                        f"(burst {peak_burst:.2f}, sustained 1.5 s {peak_sust:.2f} TFLOP/s; nominal 37.2); "
                                        "MEASURED_PEAKS.json holds no FP64 figure",
                         "launch": info},
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 40 * n, "d2h_bytes_per_step": 24 * n,
                    "ms_per_step": e2e_ms,
                    "what": "rb2_accel_host: pinned host pos/q/m -> device, pair kernel, accelerations -> host" if (world == 1 or not sym)
                    else ("rb2_upload_particles (host pos/q/m) -> rb2_accel_only (split pair work, peer-memory exchange) -> rb2_download_particles(acc)" if p2p
                          else "rb2_upload_particles (host pos/q/m) -> rb2_accel_partial -> NCCL all-reduce -> rb2_accel_finalize -> rb2_download_particles(acc)")},
            "pair_kernel": "pair-symmetric" if sym else "gather",
            "gpu_launches": launches, "clocks": clk,
            "acc_checksum": checksum, "sweep": sweep, "field_batch": field_batch,
        }
        if ngpu == 1 and not args.no_cpu:
            try:
                orc, how = load_fast_oracle()
                p = orc.params_planar(2000.0, 1000 * NM, (1000 * NM, 1000 * NM, 1000 * NM), 1.0e-16, True, args.nic)
                v, t, sample = cpu_sample(orc, p, pos, q, m)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": orc.max_threads(), "kind": "port",
                                        "sample": sample + "; " + how}
            except Exception as e:  # the baseline is a reported extra, never the product path
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if p2p:
        barrier()  # nobody unmaps while a peer may still read
        hp.p2p_detach()
    hp.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--particles", "--n", dest="n", type=int, default=1_000_000,
                    help="particles (headline 1e6; 1e4 / 1e5 for the sweep); use --particles under torchrun")
    ap.add_argument("--nic", type=int, default=1, help="N_ic_max")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the N = 1e4 / 1e5 sweep and the field-batch table")
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N and NO torchrun: one process drives the N GPUs (rb2_set_devices), the way the "
                         "single-process Fortran host would")
    ap.add_argument("--no-flush", action="store_true", help="diagnostics only: no L2 flush between steps (not a valid bench line)")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostics only: no nvidia-smi sampling")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange of the partial pair sums: p2p = peer-memory loads fused into the finalise kernel "
                         "(default), nccl = torch.distributed all-reduce between rb2_accel_partial and rb2_accel_finalize")
    ap.add_argument("--pair-mode", default="auto", choices=["auto", "sym", "gather"],
                    help="pair kernel: sym = each unordered pair once (default for N >= 3500), gather = ordered pairs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
