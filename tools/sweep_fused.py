"""Tuning sweep of the fused (temporal blocking) upwind kernel on the GPU box:
    python tools/sweep_fused.py N > gpurun_out/fused.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(1)
small = rng.random((24, 40, 260))
big = rng.random((N, N, N)) if N <= 512 else None
combos = os.environ.get("SWEEP_FUSED", "2:0,2:1,2:2,2:3,3:0,3:1,3:2,3:3,4:0,4:1,4:2").split(",")
cis = [int(x) for x in os.environ.get("SWEEP_CIS", "0,64,128").split(",")]
impl = "r01l-layout"
for combo in combos:
    T, cfg = (int(x) for x in combo.split(":"))
    os.environ["FDB_FUSED_CFG"] = str(cfg)
    os.environ["FDB_TMA_CI"] = "0"
    steps = 3 * T + 1
    with fb.Upwind([1.0] * 3, [1.0] * 3, small.shape) as up:
        up.set_fuse(T)
        up.set_field(small)
        up.advect(steps, up.default_dt())
        ok = np.array_equal(up.field(), oracle.c.upwind_advect(small, steps))
    with fb.Upwind([1.0] * 3, [1.0] * 3, [N] * 3) as up:
        up.set_fuse(T)
        if big is not None:
            up.set_field(big)
        dt = up.default_dt()
        for ci in cis:
            os.environ["FDB_TMA_CI"] = str(ci)
            nst = 12 * T
            up.advect(nst, dt)
            best = 1e30
            for _ in range(3):
                up.advect(nst, dt)
                best = min(best, up.last_timing()["gpu_ms"] / nst)
            gcups = N ** 3 / best / 1e6
            print(f"impl={impl} T={T} cfg={cfg} ci={ci:3d} N={N} parity={'ok' if ok else 'FAIL'} ms/step={best:.4f} "
                  f"GCUPS={gcups:.1f} x_roofline={gcups * 16 / 6548.5:.3f}", flush=True)
