"""One short run of the default upwind path for ncu / a quick rate check:
    python tools/prof_upwind.py N [vx vy vz]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vel = [float(x) for x in sys.argv[2:5]] if len(sys.argv) >= 5 else [1.0, 1.0, 1.0]
with fb.Upwind(vel, [1.0] * 3, [N] * 3) as up:
    if N <= 512:
        up.set_field(np.random.default_rng(5).random((N, N, N)))
    dt = abs(up.default_dt())
    up.advect(9, dt)
    best = 1e30
    for _ in range(3):
        up.advect(30, dt)
        best = min(best, up.last_timing()["gpu_ms"] / 30)
    print(f"upwind N={N} velocity={vel} kernel={up.kernel()} ms/step={best:.4f} GCUPS={N ** 3 / best / 1e6:.1f}", flush=True)
