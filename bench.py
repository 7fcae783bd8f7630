#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native FiDiBench hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Metric (BASELINE.json): GCUPS = FP64 cell-updates per second of the 3-D first-order
upwind step (ref: upwind/cxx/upwind.cxx:51-86), whole job over all N GPUs.
One bench "step" = one Upwind::advect(numTimeSteps=T) call (T = 100, BASELINE
config 2) over the resident field.  N = 1 runs 512^3 (configs[1]); N > 1 keeps
512^3 cells per GPU (weak scaling), slabs along axis 0 with an NCCL halo ring:
N=2 -> 1024x512x512, N=4 -> 1024x1024x512, N=8 -> 1024^3 (configs[2]).
`--workload upwind1024` runs 1024^3 on every N instead (strong scaling);
`--workload lap1024` is BASELINE configs[3], the 7-point Laplacian iterate loop on 1024^3.

Prints ONE JSON line on rank 0 (see the keys at the bottom of main()).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 16.0  # one FP64 read + one FP64 write per cell-update
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def workload_dims(workload: str, n: int):
    if workload == "upwind512":      # weak: 512^3 cells per GPU
        dims = {1: (512, 512, 512), 2: (1024, 512, 512), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}.get(n)
        if dims is None:
            dims = (512 * n, 512, 512)
        return dims, "weak"
    if workload == "upwind1024":     # strong: BASELINE configs[2]
        return (1024, 1024, 1024), "strong"
    if workload == "upwind128":      # configs[0], parity-sized (launch-bound)
        return (128, 128, 128), "strong"
    if workload == "upwind2048":     # configs[4], 8 GPUs
        return (2048, 2048, 2048), "strong"
    if workload == "lap1024":        # configs[3]: 7-point Laplacian, 1024^3, 1 and 8 GPUs
        return (1024, 1024, 1024), "strong"
    if workload == "lap512":
        return (512, 512, 512), "strong"
    raise SystemExit(f"unknown workload {workload}")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thread, self.gpu, self.first = [], None, None, gpu_index, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        t0 = time.time()
        while not self.rows and time.time() - t0 < 5.0:  # nvidia-smi takes a moment to print its first row
            time.sleep(0.02)

    def mark(self):
        """Rows from here on were sampled under load."""
        self.first = len(self.rows)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows[self.first:]:
            p = [x.strip() for x in row.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons),
                "window": "warm-up + timed region, nvidia-smi every 20 ms"}


# --------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref = the untouched
    upwind.cxx compiled with OpenMP; else the oracle port), all host threads, on a
    bounded sample of the same workload."""
    if rank != 0:
        return 0
    # the reference's own run-time recipe (scripts/fdi_maui_vs_mahuika.py:9): threads pinned to cores.  Must be in
    # the environment before libgomp initialises, i.e. before the reference library is loaded.
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ.setdefault("OMP_PLACES", "cores")
    import numpy as np
    import oracle
    dims, scaling = workload_dims(args.workload, args.gpus)
    ncpu = os.cpu_count() or 1
    # bounded sample: the same 512^3-cells-per-GPU grid is far too slow on the host at
    # 100 time steps; keep the single-GPU grid (the reference's own published case is
    # 512^3 x 10, pictures/mahuika.py:13) and calibrate the time steps per bench step
    sdims = (512, 512, 512) if dims[0] >= 512 else dims
    cells = float(np.prod(sdims))
    use_ref = oracle.ref_available()
    if use_ref:
        r = oracle.ref()
        r.set_threads(ncpu)
        threads = r.threads()
        run = lambda t: r.upwind_run(sdims, t, want_field=False)["seconds"]
        kind = "reference"
    else:
        threads = oracle.c.num_threads()
        f0 = np.zeros(sdims); f0.reshape(-1)[0] = 1.0
        def run(t):
            t0 = time.perf_counter(); oracle.c.upwind_advect(f0, t); return time.perf_counter() - t0
        kind = "port"
    t1 = run(1)  # calibration (also first-touch)
    budget = args.budget
    tsteps = int(max(1, min(args.tsteps, budget / max(t1, 1e-3) / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        run(tsteps)
    secs = 0.0
    for _ in range(args.steps):
        secs += run(tsteps)
    value = cells * tsteps * args.steps / secs / 1e9
    sample = (f"{sdims[0]}x{sdims[1]}x{sdims[2]} grid, {tsteps} time step(s) per bench step, advect() only, "
              f"{'untouched upwind.cxx -O3 -fopenmp, OMP_PROC_BIND=' + os.environ.get('OMP_PROC_BIND', '') if use_ref else 'oracle/fdb_oracle.c'}")
    line = {
        "impl": "reference", "metric": "GCUPS (FP64 cell-updates/s), upwind 3-D", "value": value, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"upwind3d {dims[0]}x{dims[1]}x{dims[2]} x{args.tsteps} time steps (reference arm: "
                               f"bounded sample {sample})", "parallelism": f"openmp{threads}"},
        "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline(workload, tsteps_hint=10):
    """The reference arm (the untouched upwind.cxx, OpenMP on every host core, threads pinned as the reference's
    scripts do) run once in a process of its own on a bounded sample; reported beside the GPU number, not a target.
    A separate process because OMP_PROC_BIND has to be set before libgomp starts and must not pin this one."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--workload", workload, "--tsteps", str(tsteps_hint), "--budget", "15"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    if p.returncode != 0 or not lines:
        raise RuntimeError("reference arm failed: " + p.stderr[-500:])
    return json.loads(lines[-1])["cpu_baseline"]


def lap_cpu_baseline(niter_hint=2):
    """The reference's Filter path (cxx/Filter.cpp, untouched, single rank against the MPI stub of
    oracle/fakempi) on a bounded sample; else the oracle port."""
    import numpy as np
    import oracle
    sdims = (128, 128, 128)
    off, w = oracle.laplacian_stencil(3)
    x = oracle.c.laplacian_input(sdims)
    if oracle.ref_available():
        kind, cores = "reference", 1
        run = lambda n: oracle.ref().filter_run(sdims, off, w, init=x, niter=n, want_field=False)["seconds"]
    else:
        kind, cores = "port", 1
        def run(n):
            t0 = time.perf_counter(); y = x
            for _ in range(n):
                y = oracle.c.stencil_apply(y, off, w)
            return time.perf_counter() - t0
    t1 = run(1)
    niter = int(max(1, min(niter_hint, 12.0 / max(t1, 1e-3))))
    secs = run(niter)
    return {"value": float(np.prod(sdims)) * niter / secs / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
            "sample": f"128x128x128 x {niter} x (applyFilter; copyOutToIn), "
                      f"{'untouched Filter.cpp, 1 rank (MPI stub)' if kind == 'reference' else 'oracle/fdb_oracle.c'}"}


def laplacian_reference_arm(args, rank):
    if rank != 0:
        return 0
    dims, scaling = workload_dims(args.workload, args.gpus)
    cpu, vals = None, []
    for i in range(args.warmup + args.steps):
        cpu = lap_cpu_baseline(niter_hint=1)   # one bounded sample per bench step
        if i >= args.warmup:
            vals.append(cpu["value"])
    value = len(vals) / sum(1.0 / v for v in vals)   # total cell-applies / total seconds
    cpu["value"] = value
    line = {"impl": "reference", "metric": "GCUPS (FP64 cell-applies/s), laplacian 3-D 7-point", "value": value,
            "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 128.0 ** 3 / value / 1e6, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"laplacian7 {dims[0]}x{dims[1]}x{dims[2]} (reference arm: bounded sample {cpu['sample']})",
                       "parallelism": "1 rank"},
            "cpu_baseline": cpu, "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main_laplacian(args, rank, world, local_rank):
    """BASELINE configs[3]: the 3-D 7-point Laplacian of laplacian/cxx/laplacian.cxx on 1024^3.  One bench
    step = the driver's loop, ITER x (applyFilter; copyOutToIn) (ITER = 10, laplacian.cxx:86-90)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import fidibench_b200 as fb
    import oracle

    if not torch.cuda.is_available() or fb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = fb.Comm.from_torch_distributed(device=local_rank)
    dims, scaling = workload_dims(args.workload, world)
    ITER = 10
    off, w = oracle.laplacian_stencil(3)
    st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
    fl = fb.Filter(dims, [0.0] * 3, [1.0] * 3, st, comm=comm)
    if args.kernel == "generic":
        fl.set_kernel(fb.FDB_KERNEL_GENERIC)
    if args.fuse:
        fl.set_fuse(args.fuse)
    fuse = fl.fuse()
    tiled = fl.kernel() == fb.FDB_KERNEL_TMA
    kernel_name = "stencil_generic_kernel" if not tiled else ("lap7_fused2_kernel" if fuse == 2 else "lap7_tma_kernel")
    nloc = fl.hi - fl.lo
    slab_cells = nloc * dims[1] * dims[2]
    total_cells = float(np.prod(dims))
    # the driver's input function on this rank's planes (ref: laplacian.cxx:22-28, Filter.cpp:103-112),
    # evaluated on the host as the reference does, in pinned memory
    x1 = np.sin(2.0 * np.pi * ((np.arange(dims[0]) + 0.5) * (1.0 / float(dims[0]))))
    host = torch.empty(slab_cells, dtype=torch.float64, pin_memory=True)
    host_np = host.numpy().reshape(nloc, dims[1], dims[2])
    np.multiply(x1[fl.lo:fl.hi, None, None] * x1[None, :dims[1], None], x1[None, None, :dims[2]], out=host_np)
    fl.set_input_slab(host_np)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if sampler:
        sampler.mark()
    for _ in range(args.warmup):
        fl.iterate(ITER)
    barrier()
    launches, ms = 0, 0.0
    for _ in range(args.steps):
        # every step starts from the driver's input again (untimed upload): iterating the Laplacian on and on
        # amplifies roundoff by up to 12x per apply and would overflow after ~280 applies
        fl.set_input_slab(host_np)
        barrier()
        l0 = fb.launch_count()
        fl.iterate(ITER)                       # synchronous; CUDA events on the kernels' stream inside the library
        launches += fb.launch_count() - l0
        ms += fl.last_timing()["gpu_ms"]
    barrier()
    clocks = sampler.stop() if sampler else None
    halo = fl.last_timing()["halo_bytes"]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = total_cells * ITER * args.steps / (ms / 1e3) / 1e9

    # end to end with HOST buffers: upload the input, iterate, read the checksums back
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 3))
        fl.set_input_slab(host_np); fl.iterate(ITER); fl.computeCheckSum("output")
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fl.set_input_slab(host_np)
            fl.iterate(ITER)
            chk = fl.computeCheckSum("output")
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t1], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t1 = float(t.item())
        e2e = {"value": total_cells * ITER * e2e_steps / t1 / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(slab_cells * 8), "d2h_bytes_per_step": int(dims[0] * 8), "steps": e2e_steps,
               "checksum": chk, "ms_per_step": t1 / e2e_steps * 1e3,
               "what": "per step: fdb_stencil_set_input_slab(pinned host) + fdb_stencil_iterate(10) + fdb_stencil_checksum "
                       "(PCIe-bound: the upload of the input field is most of the step)"}

    peak, peak_src = measured_peak()
    per_launch = 2 if fuse == 2 else 1
    n_kernel_launches = (ITER // per_launch + ITER % per_launch) * args.steps
    avg_launch_ms = ms / n_kernel_launches
    algo_bytes_per_launch = slab_cells * ALGO_BYTES_PER_UPDATE * ITER * args.steps / n_kernel_launches
    achieved = algo_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(f"{kernel_name}:{nloc}x{dims[1]}x{dims[2]}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,  # the north star quotes B200's nominal ~8 TB/s as well
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "applies_per_launch": per_launch,
                "dram_frac": (traffic / (avg_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "how": "16 B per cell-apply x cell-applies of one launch / (CUDA-event time / launches); the fused kernel "
                       "does two applies per launch, so the algorithmic figure may exceed the copy roofline -- "
                       "`traffic`/`dram_frac` are the measured DRAM bytes"}
    cpu = lap_cpu_baseline() if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        line = {"metric": "GCUPS (FP64 cell-applies/s), laplacian 3-D 7-point", "value": value, "unit": "GCUPS",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (the driver's prod sin(2 pi x) input, evaluated on the host)",
                "config": {"workload": f"laplacian7 {dims[0]}x{dims[1]}x{dims[2]}, {ITER} x (apply; copyOutToIn) per step",
                           "cells_per_gpu": int(slab_cells), "parallelism": f"slab{world}" if world > 1 else "single",
                           "kernel": kernel_name, "applies_per_sweep": per_launch,
                           "l2": "inputs larger than L2 (2 ping-pong fields of %.2f GiB per GPU)" % (slab_cells * 8 / 2**30)},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "halo_bytes_per_gpu": halo}
        print(json.dumps(line), flush=True)
    fl.close()
    if comm is not None:
        comm.close()
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="upwind512")
    ap.add_argument("--tsteps", type=int, default=100, help="time steps per advect() call (= per bench step)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "tma"])
    ap.add_argument("--fuse", type=int, default=0, help="time steps per sweep (temporal blocking), 0 = library default")
    ap.add_argument("--budget", type=float, default=150.0,
                    help="reference arm: seconds of host work the whole run may take (bounds the time steps per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return laplacian_reference_arm(args, rank) if args.workload.startswith("lap") else reference_arm(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.workload.startswith("lap"):
        if world != args.gpus:
            raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
        return main_laplacian(args, rank, world, local_rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    import fidibench_b200 as fb

    if not torch.cuda.is_available() or fb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = fb.Comm.from_torch_distributed(device=local_rank)

    dims, scaling = workload_dims(args.workload, world)
    base = float(min(dims))
    lengths = [d / base for d in dims]        # same resolution in each direction, as the reference's main()
    T = args.tsteps
    up = fb.Upwind([1.0, 1.0, 1.0], lengths, dims, comm=comm)
    if args.kernel != "auto":
        up.set_kernel(fb.FDB_KERNEL_GENERIC if args.kernel == "generic" else fb.FDB_KERNEL_TMA)
    up.set_fuse(args.fuse)
    fuse = (args.fuse or 3) if up.kernel() == fb.FDB_KERNEL_TMA else 1
    kernel_name = ("upwind_generic_kernel" if up.kernel() != fb.FDB_KERNEL_TMA else
                   "upwind3d_tma_kernel" if fuse == 1 else f"upwind3d_fused_kernel<T={fuse}>")
    dt = up.default_dt()
    slab_cells = up.slab_cells()
    total_cells = float(np.prod(dims))

    # synthetic input: uniform random FP64 field in pinned host memory (one slab per rank)
    gen = torch.Generator().manual_seed(20261017 + rank)
    host = torch.empty(slab_cells, dtype=torch.float64, pin_memory=True)
    host.uniform_(0.0, 1.0, generator=gen)
    host_np = host.numpy()
    up.set_slab(host_np)

    stream = torch.cuda.Stream()
    up.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if sampler:
        sampler.mark()          # load window = warm-up + timed region (same kernels, back to back)
    for _ in range(args.warmup):
        up.advect_async(T, dt)
    up.sync()
    barrier()
    launches0 = fb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        up.advect_async(T, dt)
    e1.record(stream)
    up.sync()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = fb.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    secs = ms / 1e3
    value = total_cells * T * args.steps / secs / 1e9
    halo = up.last_timing()["halo_bytes"]

    # ---- end to end through the public API with HOST buffers ("e2e") ---------------------
    # Every step uploads its input field from pinned host memory (set_slab), advects T time
    # steps and reads the step's result back (checksum).  Two handles are software-pipelined
    # through the asynchronous entry points: the upload of step n+1 (copy engine) overlaps the
    # advect of step n (SMs); each step's H2D and D2H stay inside the timed region.
    e2e = None
    if not args.no_e2e:
        up.set_stream(None)
        up2 = fb.Upwind([1.0, 1.0, 1.0], lengths, dims, comm=comm)
        up2.set_fuse(args.fuse)
        if args.kernel != "auto":
            up2.set_kernel(fb.FDB_KERNEL_GENERIC if args.kernel == "generic" else fb.FDB_KERNEL_TMA)
        pair = [up, up2]
        e2e_steps = max(2, min(args.steps, 6))

        def run_e2e(nsteps):
            chk = None
            pair[0].set_slab_async(host_np); pair[0].advect_async(T, dt)
            for i in range(nsteps):
                if i + 1 < nsteps:
                    nxt = pair[(i + 1) % 2]
                    nxt.set_slab_async(host_np)   # host -> device copy of the next step's field
                    nxt.advect_async(T, dt)
                chk = pair[i % 2].checksum()      # device -> host read of this step's result (syncs it)
            return chk

        run_e2e(2)
        barrier()
        t0 = time.perf_counter()
        chk = run_e2e(e2e_steps)
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t1], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t1 = float(t.item())
        e2e = {"value": total_cells * T * e2e_steps / t1 / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(slab_cells * 8), "d2h_bytes_per_step": int(dims[0] * 8),
               "steps": e2e_steps, "checksum": chk, "ms_per_step": t1 / e2e_steps * 1e3,
               "what": "per step: fdb_upwind_set_slab_async(pinned host) + fdb_upwind_advect_async(T) + "
                       "fdb_upwind_checksum; two handles software-pipelined (upload of step n+1 overlaps advect of step n)"}
        up2.close()

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = measured_peak()
    # one sweep kernel advances `fuse` time steps of the slab; a remainder (T % fuse) runs the
    # single-step kernel.  achieved = algorithmic bytes of all launches / their total duration.
    n_kernel_launches = (T // fuse + (1 if T % fuse else 0)) * args.steps   # remainders: [3,1] runs as [2,2], [2] as one sweep
    avg_launch_ms = ms / n_kernel_launches
    algo_bytes_per_launch = slab_cells * ALGO_BYTES_PER_UPDATE * T * args.steps / n_kernel_launches
    achieved = algo_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        key = f"{kernel_name}:{dims[0] // world}x{dims[1]}x{dims[2]}"
        traffic = tj.get(key)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,  # the north star quotes B200's nominal ~8 TB/s as well
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "time_steps_per_launch": fuse, "dram_frac": (traffic / (avg_launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "how": "16 B per cell-update x cell-updates of one launch / (CUDA-event time of the timed region / "
                       "launches); with temporal blocking one launch advances several time steps, so the algorithmic "
                       "figure may exceed the copy roofline -- `traffic`/`dram_frac` are the measured DRAM bytes"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)

    if rank == 0:
        line = {
            "metric": "GCUPS (FP64 cell-updates/s), upwind 3-D", "value": value, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (uniform random FP64 field)",
            "config": {"workload": f"upwind3d {dims[0]}x{dims[1]}x{dims[2]} x{T} time steps per step",
                       "cells_per_gpu": int(slab_cells), "parallelism": f"slab{world}" if world > 1 else "single",
                       "kernel": kernel_name, "time_steps_per_sweep": fuse,
                       "l2": "inputs larger than L2 (2 ping-pong fields of %.2f GiB per GPU)" % (slab_cells * 8 / 2**30)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "halo_bytes_per_gpu": halo,
        }
        print(json.dumps(line), flush=True)
    up.close()
    if comm is not None:
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
