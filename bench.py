#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native FiDiBench hot path.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Metric (BASELINE.json): GCUPS = FP64 cell-updates per second of the 3-D first-order
upwind step (ref: upwind/cxx/upwind.cxx:51-86), whole job over all N GPUs.
One bench "step" = one Upwind::advect(numTimeSteps=T) call (T = 100, BASELINE
config 2) over the resident field.  N = 1 runs 512^3 (configs[1]); N > 1 keeps
512^3 cells per GPU (weak scaling), slabs along axis 0 with a peer-store halo ring:
N=2 -> 1024x512x512, N=4 -> 1024x1024x512, N=8 -> 1024^3 (configs[2]).
`--workload upwind1024` runs 1024^3 on every N instead (strong scaling);
`--workload lap1024` is BASELINE configs[3], the 7-point Laplacian iterate loop on 1024^3.

Before the timed region every rank runs an UNTIMED parity block through the same
communicator (`"parity"` in the line): a seeded field advected / filtered on
(16 N) x 48 x 256 and compared bit for bit, plane by plane, with SHA-256 digests of what the
untouched reference produced (tests/golden/bench_parity.json, generated from oracle/_ref by
tests/golden/make_golden.py), and the delta corner rule at the bench grid against
tests/golden/upwind_128_s100.npz.  A mismatch aborts with a non-zero exit code.  After the
timed region an `"also"` block measures the other BASELINE configurations that fit N GPUs
(device-timed, roofline per entry, inputs generated on the device).

Prints ONE JSON line on rank 0 (see the keys at the bottom of main()).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
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
SEED = 20261017


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
    if workload == "lap2d8000":      # the reference driver's own default: laplacian -numDims 2 -numCells 8000 (laplacian.cxx:41-42)
        return (8000, 8000), "strong"
    raise SystemExit(f"unknown workload {workload}")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def laplacian_weights():
    """The 3-D stencil laplacian.cxx builds (ref: laplacian/cxx/laplacian.cxx:55-65): -2*ndims on the centre, 1 on
    the six axis neighbours."""
    st = {(0, 0, 0): -6.0}
    for a in range(3):
        for s in (1, -1):
            o = [0, 0, 0]
            o[a] = s
            st[tuple(o)] = 1.0
    return st


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank (and so its pinned host allocations, first touch) to the CPU cores next to its GPU: the
    host->device copies of the end-to-end arm otherwise cross the socket interconnect on a multi-socket box."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


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
# reference arm / CPU baseline: the only part of this file that may touch oracle/
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


def process_walltime():
    """The reference's own published convention (pictures/mahuika.py:10-16): whole-process wall time
    of the driver executable.  drivers/bin/upwindCuda (this repository, the drop-in driver) against
    oracle/_ref/upwindCxx (the untouched reference executable, every host core, threads pinned), same flags.
    Part of the CPU-baseline leg: the only place where a binary under oracle/ is executed."""
    ours = os.path.join(ROOT, "drivers", "bin", "upwindCuda")
    ref = os.path.join(ROOT, "oracle", "_ref", "upwindCxx")
    out = {"convention": "wall time of the whole driver process (CUDA context creation, allocation, init, steps, "
                         "checksum), best of 2; reference: OMP threads = host cores, OMP_PROC_BIND=true"}
    env_ref = dict(os.environ, OMP_PROC_BIND="true", OMP_PLACES="cores", OMP_NUM_THREADS=str(os.cpu_count() or 1))
    for n, s in ((128, 10), (512, 10)):
        flags = ["-numCells", str(n), "-numSteps", str(s)]
        rec = {}
        for name, exe, env in (("ours_s", ours, os.environ), ("reference_s", ref, env_ref)):
            if not os.path.exists(exe):
                rec[name] = None
                continue
            best, text = None, ""
            for _ in range(2):
                t0 = time.perf_counter()
                p = subprocess.run([exe] + flags, capture_output=True, text=True, timeout=600, env=env)
                dt = time.perf_counter() - t0
                if p.returncode == 0 and (best is None or dt < best):
                    best, text = dt, p.stdout
            rec[name] = best
            rec[name.replace("_s", "_checksum_line")] = next((l.strip() for l in text.splitlines() if "check sum" in l), None)
        if rec.get("ours_s") and rec.get("reference_s"):
            rec["speedup"] = rec["reference_s"] / rec["ours_s"]
        out[f"upwind -numCells {n} -numSteps {s}"] = rec
    return out


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


# --------------------------------------------------------------------------------------
# the CUDA arm
# --------------------------------------------------------------------------------------
class Ctx:
    """One rank of the CUDA arm: torch is plumbing (device selection, rendezvous, pinned memory, events)."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        import fidibench_b200 as fb
        self.np, self.torch, self.dist, self.fb, self.args = np, torch, dist, fb, args
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if not torch.cuda.is_available() or fb.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
        if self.world != args.gpus:
            raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={self.world}; launch N>1 with torch.distributed.run")
        self.numa_cpus = bind_to_gpu_numa_node(self.local_rank)
        torch.cuda.set_device(self.local_rank)
        self.comm = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.comm = fb.Comm.from_torch_distributed(device=self.local_rank)
        self.peak, self.peak_src = measured_peak()
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                self.traffic = json.load(fh)
        except Exception:
            self.traffic = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        if self.world == 1:
            return bool(ok)
        t = self.torch.tensor([1.0 if ok else 0.0], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def close(self):
        if self.comm is not None:
            self.comm.close()
            self.dist.destroy_process_group()

    def roofline(self, kernel_name, slab_dims, algo_bytes_per_launch, avg_launch_ms, per_launch_key, per_launch):
        traffic = self.traffic.get(f"{kernel_name}:{slab_dims[0]}x{slab_dims[1]}x{slab_dims[2]}")
        achieved = algo_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                "frac": achieved / self.peak, "traffic": traffic, "peak_source": self.peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,  # the north star quotes B200's nominal ~8 TB/s as well
                "algorithmic_bytes_per_launch": algo_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                per_launch_key: per_launch,
                "dram_frac": (traffic / (avg_launch_ms * 1e-3) / 1e9 / self.peak) if traffic else None,
                "how": "16 B per cell-update x cell-updates of one launch / (CUDA-event time of the timed region / "
                       "launches); with temporal blocking one launch advances several time steps, so the algorithmic "
                       "figure may exceed the copy roofline -- `traffic` (ncu dram bytes per launch, profiles/traffic.json) "
                       "and `dram_frac` are the measured DRAM side"}


def plane_digests(slab):
    return [hashlib.sha256(p.tobytes()).hexdigest() for p in slab]


def corner_rule(ctx, up, dt, full: bool):
    """Delta at cell 0 advected 100 steps on the handle's own grid (any power-of-two extents >= 128 with the same
    spacing in each direction): the 101^3 corner must equal the reference's 128^3 x 100 run bit for bit and every
    other cell must be exactly zero (golden generated from oracle/_ref).  `full` downloads the slab;
    otherwise (slabs too large to copy back) only the per-plane sums and the checksum are compared, to 1e-12."""
    np = ctx.np
    g = np.load(os.path.join(ROOT, "tests", "golden", "upwind_128_s100.npz"))
    corner = g["corner"]
    up.reset()
    up.advect(100, dt)
    res = {}
    sums = up.plane_sums()
    ref_sums = np.zeros_like(sums)
    ref_sums[:101] = corner.reshape(101, -1).sum(axis=1)
    live = ref_sums != 0
    res["plane_sums_rel_err"] = float(np.max(np.abs(sums[live] - ref_sums[live]) / np.abs(ref_sums[live])))
    ok = res["plane_sums_rel_err"] <= 1e-12 and not np.any(sums[~live])
    cs = up.checksum()
    res["checksum"] = cs
    ok = ok and abs(cs - float(g["checksum"])) <= 1e-12 * abs(cs)
    if full:
        slab = up.slab()
        lo, hi = up.lo, up.hi
        if lo < 101:
            n = min(hi, 101) - lo
            same = np.array_equal(slab[:n, :101, :101], corner[lo:lo + n])
            slab[:n, :101, :101] = 0.0
            ok = ok and same
        ok = ok and not slab.any()
        del slab
    res["mode"] = "field bit-exact" if full else "plane sums + checksum (1e-12)"
    res["ok"] = ctx.all_true(ok)
    return res


def parity_block(ctx, up_bench, dt_bench, slab_bytes):
    """Untimed bit-for-bit comparison with the reference through the same ranks, communicator and kernels as
    the timed region."""
    np, fb = ctx.np, ctx.fb
    with open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")) as fh:
        gold = json.load(fh)
    out = {"ranks": ctx.world, "golden": "tests/golden/bench_parity.json + upwind_128_s100.npz (from oracle/_ref = the untouched reference)"}
    key = str(ctx.world)
    if key in gold["upwind"]:
        gu = gold["upwind"][key]
        with fb.Upwind([1.0] * 3, [1.0] * 3, gu["dims"], comm=ctx.comm) as up:
            up.fill_random(gold["seed"])
            up.advect(gold["upwind_steps"], gu["dt"])
            ok = plane_digests(up.slab()) == gu["planes"][up.lo:up.hi]
            cs = up.checksum()
            ok = ok and abs(cs - gu["checksum"]) <= 1e-12 * abs(cs)
            fused = up.kernel() == fb.FDB_KERNEL_TMA
        out["random_upwind_bitexact"] = ctx.all_true(ok)
        gs = gold["stencil"][key]
        with fb.Filter(gs["dims"], [0.0] * 3, [1.0] * 3, laplacian_weights(), comm=ctx.comm) as fl:
            fl.fill_random(gold["seed"])
            fl.iterate(gold["stencil_iters"])
            ok = plane_digests(fl.get_slab(fb.FDB_OUTPUT)) == gs["planes"][fl.lo:fl.hi]
            fused_l = fl.fuse()
        out["random_laplacian_bitexact"] = ctx.all_true(ok)
        out["random_bitexact"] = out["random_upwind_bitexact"] and out["random_laplacian_bitexact"]
        out["random_cases"] = (f"upwind {gu['dims']} x{gold['upwind_steps']} steps ({'fused TMA' if fused else 'generic'} kernels), "
                               f"laplacian7 {gs['dims']} x{gold['stencil_iters']} applies ({fused_l} per sweep); device-side hash field, "
                               "SHA-256 per plane")
    else:
        out["random_bitexact"] = None
        out["random_cases"] = f"no golden for {ctx.world} ranks"
    if up_bench is not None:
        c = corner_rule(ctx, up_bench, dt_bench, full=slab_bytes <= (2 << 30))
        out["corner_bitexact"] = c["ok"]
        out["corner"] = c
    bad = [k for k in ("random_bitexact", "corner_bitexact") if out.get(k) is False]
    if bad:
        if ctx.rank == 0:
            print(json.dumps({"parity": out, "error": "parity mismatch: " + ", ".join(bad)}), flush=True)
        raise SystemExit(3)
    return out


def upwind_kernel_name(fb, up, fuse):
    """The __global__ function the sweeps launch, as ncu lists it (from fdb_upwind_describe), and steps per launch."""
    if up.kernel() != fb.FDB_KERNEL_TMA:
        return "upwind_generic_kernel", 1
    name = up.describe().split(" ")[0]          # e.g. upwind3d_fused_lean_kernel<T=3>
    f = int(name.split("T=")[1].rstrip(">")) if "T=" in name else 1
    return name, f


def time_upwind(ctx, up, T, steps, warmup, dt, sampler=None):
    """Device-resident throughput of `steps` x advect(T): CUDA events on the stream the kernels run on."""
    torch, fb = ctx.torch, ctx.fb
    stream = torch.cuda.Stream()
    up.set_stream(stream.cuda_stream)
    ctx.barrier()
    if sampler:
        sampler.mark()          # load window = warm-up + timed region (same kernels, back to back)
    for _ in range(warmup):
        up.advect_async(T, dt)
    up.sync()
    ctx.barrier()
    l0 = fb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        up.advect_async(T, dt)
    e1.record(stream)
    up.sync()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = fb.launch_count() - l0
    up.set_stream(None)
    return ms, launches


def upwind_entry(ctx, dims, T, steps, warmup, fuse=0, kernel="auto", sampler=None, up=None, dt=None):
    """value / roofline of one upwind configuration on a field already resident in HBM."""
    np, fb = ctx.np, ctx.fb
    own = up is None
    if own:
        base = float(min(dims))
        up = fb.Upwind([1.0] * 3, [d / base for d in dims], dims, comm=ctx.comm)
        if kernel != "auto":
            up.set_kernel(fb.FDB_KERNEL_GENERIC if kernel == "generic" else fb.FDB_KERNEL_TMA)
        up.set_fuse(fuse)
        dt = up.default_dt()
        up.fill_random(SEED)
    name, f = upwind_kernel_name(fb, up, fuse)
    ms, launches = time_upwind(ctx, up, T, steps, warmup, dt, sampler)
    total = float(np.prod(dims))
    slab_cells = up.slab_cells()
    n_launch = (T // f + (1 if T % f else 0)) * steps   # remainders: [3,1] runs as [2,2], [2] as one sweep
    avg = ms / n_launch
    roof = ctx.roofline(name, (dims[0] // ctx.world, dims[1], dims[2]), slab_cells * ALGO_BYTES_PER_UPDATE * T * steps / n_launch,
                        avg, "time_steps_per_launch", f)
    halo = up.last_timing()["halo_bytes"]
    if own:
        up.close()
    return {"value": total * T * steps / (ms / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": ms / steps, "steps": steps,
            "time_steps_per_step": T, "kernel": name, "roofline": roof, "gpu_launches": int(launches),
            "halo_bytes_per_gpu": halo, "cells_per_gpu": int(slab_cells)}


def laplacian_weights_nd(nd):
    st = {tuple([0] * nd): -2.0 * nd}
    for a in range(nd):
        for sgn in (1, -1):
            o = [0] * nd
            o[a] = sgn
            st[tuple(o)] = 1.0
    return st


def laplacian_entry(ctx, dims, steps, warmup, fuse=0, kernel="auto", sampler=None, want_e2e=False):
    """BASELINE configs[3]: one step = the driver's loop, ITER x (applyFilter; copyOutToIn) (ITER = 10,
    laplacian.cxx:86-90), from the driver's input function each time (iterating on and on amplifies roundoff by
    up to 12x per apply).  The input is a product of 1-D factors evaluated with libm on the host and
    multiplied out on the device (fdb_stencil_set_input_separable): same bits as Filter::setInData, no 8 GiB upload.
    dims may be 2-D (the reference driver's default case)."""
    np, fb, torch = ctx.np, ctx.fb, ctx.torch
    ITER = 10
    nd = len(dims)
    fl = fb.Filter(dims, [0.0] * nd, [1.0] * nd, laplacian_weights_nd(nd), comm=ctx.comm)
    if kernel == "generic":
        fl.set_kernel(fb.FDB_KERNEL_GENERIC)
    if fuse:
        fl.set_fuse(fuse)
    fz = fl.fuse()
    tiled = fl.kernel() == fb.FDB_KERNEL_TMA
    name = "stencil_generic_kernel" if not tiled else (fl.describe().split(" ")[1] if fz == 2 else "lap7_tma_kernel")
    nloc = fl.hi - fl.lo
    total = float(np.prod(dims))
    slab_cells = int(total) // dims[0] * nloc
    slab_dims = (nloc,) + tuple(dims[1:]) if nd == 3 else (1, nloc, dims[1])
    factors = fl.laplacian_factors()
    # size-independent parity probe: the input is an eigenfunction of the periodic operator with eigenvalue
    # sum_j (2 cos(2 pi / N_j) - 2), so ONE apply must scale its 2-norm by |lambda| (cancellation noise ~1e-11)
    fl.set_input_separable(factors)
    n_in = fl.sumsq("input")
    fl.applyFilter()
    n_out = fl.sumsq("output")
    lam = sum(2.0 * math.cos(2.0 * math.pi / d) - 2.0 for d in dims)
    ratio_err = abs(math.sqrt(n_out / n_in) / abs(lam) - 1.0)
    tol = 1e-8 if min(dims) <= 2048 else 1e-6   # |lambda| shrinks like 1/N^2: the cancellation noise grows with it
    parity = {"one_apply_norm_ratio_rel_err": ratio_err, "ok": ctx.all_true(ratio_err < tol),
              "what": "||L x|| / ||x|| against the analytic eigenvalue of the driver's input; bitwise parity of the same "
                      "kernels is in `parity` (random field) and tests/"}
    # one applyFilter() alone (lap7_tma_kernel, 16 B of DRAM traffic per cell-apply: the plain HBM roofline case)
    ctx.barrier()
    single_ms = 1e30
    for _ in range(3):
        fl.applyFilter()
        single_ms = min(single_ms, fl.last_timing()["gpu_ms"])
    single_ms = ctx.max_over_ranks(single_ms)
    single = {"value": total / (single_ms / 1e3) / 1e9, "unit": "GCUPS", "ms": single_ms,
              "kernel": "lap7_tma_kernel" if tiled else "stencil_generic_kernel",
              "frac_of_peak": slab_cells * ALGO_BYTES_PER_UPDATE / (single_ms * 1e-3) / 1e9 / ctx.peak}
    ctx.barrier()
    if sampler:
        sampler.mark()
    for _ in range(warmup):
        fl.iterate(ITER)
    launches, ms = 0, 0.0
    for _ in range(steps):
        fl.set_input_separable(factors)       # untimed
        ctx.barrier()
        l0 = fb.launch_count()
        fl.iterate(ITER)                      # synchronous; CUDA events on the kernels' stream inside the library
        launches += fb.launch_count() - l0
        ms += fl.last_timing()["gpu_ms"]
    ctx.barrier()
    halo = fl.last_timing()["halo_bytes"]
    ms = ctx.max_over_ranks(ms)
    per_launch = 2 if fz == 2 else 1
    n_launch = (ITER // per_launch + ITER % per_launch) * steps
    roof = ctx.roofline(name, slab_dims, slab_cells * ALGO_BYTES_PER_UPDATE * ITER * steps / n_launch,
                        ms / n_launch, "applies_per_launch", per_launch)
    out = {"value": total * ITER * steps / (ms / 1e3) / 1e9, "unit": "GCUPS", "ms_per_step": ms / steps, "steps": steps,
           "applies_per_step": ITER, "kernel": name, "roofline": roof, "gpu_launches": int(launches),
           "halo_bytes_per_gpu": halo, "cells_per_gpu": int(slab_cells), "parity": parity, "single_apply": single,
           "kernel_choice": fl.describe()}
    if want_e2e:
        # end to end with HOST buffers: upload the host-evaluated input, iterate, read the checksums back
        host = torch.empty(slab_cells, dtype=torch.float64, pin_memory=True)
        host_np = host.numpy().reshape((nloc,) + tuple(dims[1:]))
        if nd == 3:
            np.multiply(factors[0][fl.lo:fl.hi, None, None] * factors[1][None, :, None], factors[2][None, None, :], out=host_np)
        else:
            np.multiply(factors[0][fl.lo:fl.hi, None], factors[1][None, :], out=host_np)
        e2e_steps = max(2, min(steps, 3))
        fl.set_input_slab(host_np); fl.iterate(ITER); fl.computeCheckSum("output")
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fl.set_input_slab(host_np)
            fl.iterate(ITER)
            chk = fl.computeCheckSum("output")
        torch.cuda.synchronize()
        t1 = ctx.max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": total * ITER * e2e_steps / t1 / 1e9, "unit": "GCUPS",
                      "h2d_bytes_per_step": int(slab_cells * 8), "d2h_bytes_per_step": int(dims[0] * 8), "steps": e2e_steps,
                      "checksum": chk, "ms_per_step": t1 / e2e_steps * 1e3,
                      "what": "per step: fdb_stencil_set_input_slab(pinned host) + fdb_stencil_iterate(10) + fdb_stencil_checksum "
                              "(PCIe-bound: the upload of the input field is most of the step)"}
    fl.close()
    return out


def also_block(ctx, main_workload):
    """The other BASELINE configurations that fit this many GPUs, device-timed with inputs generated on the
    device: lap1024 (configs[3]) at 1 and 8, upwind1024 strong (configs[2]) at 1/2/4, upwind2048 (configs[4]) at
    8, upwind128 (configs[0]) at 1."""
    plan = {1: ["upwind1024", "lap1024", "lap2d8000", "upwind128"], 2: ["upwind1024"], 4: ["upwind1024"], 8: ["lap1024", "upwind2048"]}
    names = [w for w in plan.get(ctx.world, []) if w != main_workload]
    if ctx.args.also:
        names = [w for w in ctx.args.also.split(",") if w]
    out = {}
    for w in names:
        dims, scaling = workload_dims(w, ctx.world)
        t0 = time.perf_counter()
        try:
            if w.startswith("lap"):
                e = laplacian_entry(ctx, dims, steps=3, warmup=1)
            elif w == "upwind128":
                # configs[0]: 128^3 x 10 time steps per advect() -- launch-bound, 4 sweeps of a persistent kernel
                e = upwind_entry(ctx, dims, T=10, steps=50, warmup=5)
            else:
                fb = ctx.fb
                base = float(min(dims))
                up = fb.Upwind([1.0] * 3, [d / base for d in dims], dims, comm=ctx.comm)
                dt = up.default_dt()
                par = corner_rule(ctx, up, dt, full=up.slab_cells() * 8 <= (2 << 30))
                up.fill_random(SEED)
                e = upwind_entry(ctx, dims, T=100, steps=3, warmup=1, up=up, dt=dt)
                up.close()
                e["parity"] = par
            e["workload"] = f"{w} " + "x".join(str(d) for d in dims)
            e["scaling"] = scaling
            e["n_gpus"] = ctx.world
        except Exception as ex:  # an `also` entry never takes the headline down
            e = {"error": f"{type(ex).__name__}: {ex}"}
        e["wall_s"] = time.perf_counter() - t0
        out[w] = e
    return out


def main_laplacian(ctx):
    args = ctx.args
    dims, scaling = workload_dims(args.workload, ctx.world)
    parity = None if args.no_parity else parity_block(ctx, None, None, 0)
    sampler = ClockSampler(ctx.local_rank) if ctx.rank == 0 else None
    if sampler:
        sampler.start()
    e = laplacian_entry(ctx, dims, args.steps, args.warmup, fuse=args.fuse, kernel=args.kernel, sampler=sampler,
                        want_e2e=not args.no_e2e)
    clocks = sampler.stop() if sampler else None
    if parity is not None:
        parity["eigen_probe"] = e["parity"]
    cpu = lap_cpu_baseline() if (ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline) else None
    also = None if args.no_also else also_block(ctx, args.workload)
    if ctx.rank == 0:
        line = {"metric": "GCUPS (FP64 cell-applies/s), laplacian 3-D 7-point", "value": e["value"], "unit": "GCUPS",
                "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": e["ms_per_step"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (the driver's prod sin(2 pi x) input; 1-D factors from the host's libm, multiplied out on the device)",
                "config": {"workload": f"laplacian7 {dims[0]}x{dims[1]}x{dims[2]}, 10 x (apply; copyOutToIn) per step",
                           "cells_per_gpu": e["cells_per_gpu"], "parallelism": f"slab{ctx.world}" if ctx.world > 1 else "single",
                           "kernel": e["kernel"], "applies_per_sweep": e["roofline"]["applies_per_launch"],
                           "l2": "inputs larger than L2 (2 ping-pong fields of %.2f GiB per GPU)" % (e["cells_per_gpu"] * 8 / 2**30)},
                "roofline": e["roofline"], "cpu_baseline": cpu, "e2e": e.get("e2e"), "gpu_launches": e["gpu_launches"],
                "clocks": clocks, "halo_bytes_per_gpu": e["halo_bytes_per_gpu"], "parity": parity, "also": also}
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


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
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed bitwise parity block")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--also", default="", help="comma-separated workloads for the `also` block (default: by N)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return laplacian_reference_arm(args, rank) if args.workload.startswith("lap") else reference_arm(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    ctx = Ctx(args)
    if args.workload.startswith("lap"):
        return main_laplacian(ctx)
    np, torch, fb = ctx.np, ctx.torch, ctx.fb

    dims, scaling = workload_dims(args.workload, world)
    base = float(min(dims))
    lengths = [d / base for d in dims]        # same resolution in each direction, as the reference's main()
    T = args.tsteps
    up = fb.Upwind([1.0, 1.0, 1.0], lengths, dims, comm=ctx.comm)
    if args.kernel != "auto":
        up.set_kernel(fb.FDB_KERNEL_GENERIC if args.kernel == "generic" else fb.FDB_KERNEL_TMA)
    up.set_fuse(args.fuse)
    dt = up.default_dt()
    slab_cells = up.slab_cells()
    total_cells = float(np.prod(dims))

    # ---- untimed parity block, same ranks / communicator / kernels -------------------------
    parity = None if args.no_parity else parity_block(ctx, up, dt, slab_cells * 8)

    # synthetic input: uniform random FP64 field in pinned host memory (one slab per rank)
    gen = torch.Generator().manual_seed(SEED + rank)
    host = torch.empty(slab_cells, dtype=torch.float64, pin_memory=True)
    host.uniform_(0.0, 1.0, generator=gen)
    host_np = host.numpy()
    up.set_slab(host_np)

    # ---- device-resident throughput ("value") ------------------------------------------
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    main_e = upwind_entry(ctx, dims, T, args.steps, args.warmup, fuse=args.fuse, sampler=sampler, up=up, dt=dt)
    clocks = sampler.stop() if sampler else None

    # ---- end to end through the public API with HOST buffers ("e2e") ---------------------
    # Every step uploads its input field from pinned host memory (set_slab), advects T time
    # steps and reads the step's result back (checksum).  Two handles are software-pipelined
    # through the asynchronous entry points: the upload of step n+1 (copy engine) overlaps the
    # advect of step n (SMs); each step's H2D and D2H stay inside the timed region.
    e2e = None
    if not args.no_e2e:
        up2 = fb.Upwind([1.0, 1.0, 1.0], lengths, dims, comm=ctx.comm)
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
        ctx.barrier()
        t0 = time.perf_counter()
        chk = run_e2e(e2e_steps)
        torch.cuda.synchronize()
        t1 = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": total_cells * T * e2e_steps / t1 / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": int(slab_cells * 8), "d2h_bytes_per_step": int(dims[0] * 8),
               "steps": e2e_steps, "checksum": chk, "ms_per_step": t1 / e2e_steps * 1e3,
               "h2d_gbs_per_gpu": slab_cells * 8 * e2e_steps / t1 / 1e9, "cpus_bound_to": ctx.numa_cpus,
               "what": "per step: fdb_upwind_set_slab_async(pinned host) + fdb_upwind_advect_async(T) + "
                       "fdb_upwind_checksum; two handles software-pipelined (upload of step n+1 overlaps advect of step n)"}
        up2.close()
        # variant that also copies the whole field back to the host every step (what a driver that post-processes
        # the field, e.g. -vtk, pays): upload + advect + download, not pipelined
        back = torch.empty(slab_cells, dtype=torch.float64, pin_memory=True)
        back_np = back.numpy()
        fsteps = 2
        def run_field(n):
            for _ in range(n):
                up.set_slab_async(host_np)
                up.advect_async(T, dt)
                up.slab_into(back_np)
        run_field(1)
        ctx.barrier()
        t0 = time.perf_counter()
        run_field(fsteps)
        t1 = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e["with_field_copyback"] = {"value": total_cells * T * fsteps / t1 / 1e9, "unit": "GCUPS", "steps": fsteps,
                                      "h2d_bytes_per_step": int(slab_cells * 8), "d2h_bytes_per_step": int(slab_cells * 8),
                                      "ms_per_step": t1 / fsteps * 1e3, "field_sum": float(back_np[:4096].sum())}
        del back, back_np
    up.close()
    del host, host_np

    cpu = None
    e2e_process = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)
        try:
            e2e_process = process_walltime()
        except Exception as ex:
            e2e_process = {"error": str(ex)}

    also = None if args.no_also else also_block(ctx, args.workload)

    if rank == 0:
        fuse = main_e["roofline"]["time_steps_per_launch"]
        line = {
            "metric": "GCUPS (FP64 cell-updates/s), upwind 3-D", "value": main_e["value"], "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_e["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (uniform random FP64 field)",
            "config": {"workload": f"upwind3d {dims[0]}x{dims[1]}x{dims[2]} x{T} time steps per step",
                       "cells_per_gpu": int(slab_cells), "parallelism": f"slab{world}" if world > 1 else "single",
                       "kernel": main_e["kernel"], "time_steps_per_sweep": fuse,
                       "l2": "inputs larger than L2 (2 ping-pong fields of %.2f GiB per GPU)" % (slab_cells * 8 / 2**30)},
            "roofline": main_e["roofline"], "cpu_baseline": cpu, "e2e": e2e, "e2e_process": e2e_process,
            "gpu_launches": main_e["gpu_launches"], "clocks": clocks, "halo_bytes_per_gpu": main_e["halo_bytes_per_gpu"],
            "parity": parity, "also": also,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
