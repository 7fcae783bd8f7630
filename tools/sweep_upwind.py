"""Tuning sweep of the TMA upwind kernel (run on the GPU box):
    python tools/sweep_upwind.py [N] > gpurun_out/sweep.txt
For every kernel configuration (FDB_TMA_CFG) and i-chunk (FDB_TMA_CI): parity with the
oracle on a small ragged grid, then GCUPS on N^3 from CUDA-event timing of 100 steps."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(x) for x in os.environ.get("SWEEP_CFGS", "0,1,2,3,4,5,6,7").split(",")]
cis = [int(x) for x in os.environ.get("SWEEP_CIS", "0,8,16,32,64,128").split(",")]
rng = np.random.default_rng(1)
small = rng.random((24, 40, 260))
small_ref = oracle.c.upwind_advect(small, 3)
big = rng.random((N, N, N)) if N <= 512 else None

for cfg in cfgs:
    os.environ["FDB_TMA_CFG"] = str(cfg)
    os.environ["FDB_TMA_CI"] = "0"
    with fb.Upwind([1.0] * 3, [1.0] * 3, small.shape) as up:
        up.set_field(small)
        up.advect(3, up.default_dt())
        ok = np.array_equal(up.field(), small_ref)
    with fb.Upwind([1.0] * 3, [1.0] * 3, [N] * 3) as up:
        if big is not None:
            up.set_field(big)
        dt = up.default_dt()
        for ci in cis:
            os.environ["FDB_TMA_CI"] = str(ci)
            up.advect(20, dt)
            best = 1e30
            for _ in range(3):
                up.advect(50, dt)
                best = min(best, up.last_timing()["gpu_ms"] / 50)
            gcups = N ** 3 / best / 1e6
            print(f"cfg={cfg} ci={ci:3d} N={N} parity={'ok' if ok else 'FAIL'} ms/step={best:.4f} GCUPS={gcups:.1f} "
                  f"frac_of_6548.5={gcups * 16 / 6548.5:.3f}", flush=True)
