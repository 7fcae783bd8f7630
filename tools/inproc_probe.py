"""In-process multi-GPU probe (one process drives `ngpus` devices): python tools/inproc_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402

ndev = fb.device_count()
for halo in ("direct", "nccl"):
    os.environ["FDB_HALO"] = halo
    for ngpus in [g for g in (1, 2, 4, 8) if g <= ndev]:
        dims = (512 * ngpus, 512, 512)
        with fb.Upwind([1.0] * 3, [d / 512 for d in dims], dims, ngpus=ngpus) as up:
            dt = up.default_dt()
            for fuse in (1, 3):
                up.set_fuse(fuse)
                up.advect(12, dt)
                best, wall = 1e30, 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    up.advect(60, dt)
                    wall = min(wall, time.perf_counter() - t0)
                    best = min(best, up.last_timing()["gpu_ms"])
                cells = float(np.prod(dims)) * 60
                print(f"halo={halo} ngpus={ngpus} fuse={fuse} gpu_ms={best:.2f} wall_ms={wall * 1e3:.2f} "
                      f"GCUPS(gpu)={cells / best / 1e6:.1f} GCUPS(wall)={cells / wall / 1e9:.1f}", flush=True)
