"""Laplacian / stencil throughput on the GPU box: python tools/bench_stencil.py N [cfg...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(x) for x in sys.argv[2:]] or [0, 1, 2]
off, w = oracle.laplacian_stencil(3)
st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
rng = np.random.default_rng(3)
x = rng.random((N, N, N)) if N <= 512 else None
for kern in ["generic"] + cfgs:
    if kern != "generic":
        os.environ["FDB_LAP_CFG"] = str(kern)
    with fb.Filter([N] * 3, [0.0] * 3, [1.0] * 3, st) as fl:
        if kern == "generic":
            fl.set_kernel(fb.FDB_KERNEL_GENERIC)
        if x is not None:
            fl.set_input(x)
        for ci in ([0] if kern == "generic" else [int(x) for x in os.environ.get("SWEEP_CIS", "0,16,32,64").split(",")]):
            os.environ["FDB_TMA_CI"] = str(ci)
            fl.iterate(5)
            best = 1e30
            for _ in range(3):
                fl.iterate(10)
                best = min(best, fl.last_timing()["gpu_ms"] / 10)
            print(f"lap7 N={N} kernel={kern} used={fl.kernel()} ci={ci} ms/apply={best:.4f} GCUPS={N**3 / best / 1e6:.1f} "
                  f"frac_of_6548.5={N**3 * 16 / best / 1e6 / 6548.5:.3f}", flush=True)
