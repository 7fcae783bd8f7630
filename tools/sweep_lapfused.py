"""Fused two-apply 7-point kernel on the GPU box: python tools/sweep_lapfused.py N [cfg...]
Prints GCUPS (cell-applies/s) for iterate() with fuse = 1 (lap7_tma_kernel) and fuse = 2 per tile
configuration (FDB_LAPF_CFG), plane-chunk length (FDB_TMA_CI, env SWEEP_CIS) and TMA L2 promotion
(FDB_TMA_L2PROMO, env SWEEP_PROMOS)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(x) for x in sys.argv[2:]] or list(range(12))
cis = [int(x) for x in os.environ.get("SWEEP_CIS", "0,64").split(",")]
promos = os.environ.get("SWEEP_PROMOS", "256").split(",")
ITER = 10
off, w = oracle.laplacian_stencil(3)
st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
if N <= 512:
    field = np.random.default_rng(3).random((N, N, N))
else:  # a random host array of 8 GiB is not worth the box time: one random slab, repeated
    field = np.concatenate([np.random.default_rng(3).random((64, N, N))] * (N // 64))


def run(fl, tag):
    fl.set_input_slab(field)
    fl.iterate(4)
    best = 1e30
    for _ in range(3):
        fl.iterate(ITER)
        best = min(best, fl.last_timing()["gpu_ms"] / ITER)
    print(f"lap7 N={N} {tag} ms/apply={best:.4f} GCUPS={N ** 3 / best / 1e6:.1f} "
          f"x_copy_roofline={N ** 3 * 16 / best / 1e6 / 6548.5:.3f}", flush=True)


for promo in promos:
    os.environ["FDB_TMA_L2PROMO"] = promo   # read when the tensor maps are encoded
    with fb.Filter([N] * 3, [0.0] * 3, [1.0] * 3, st) as fl:
        print(f"-- TMA L2 promotion {promo} B", flush=True)
        os.environ.pop("FDB_LAPF_CFG", None)
        os.environ["FDB_TMA_CI"] = "0"
        fl.set_fuse(1)
        run(fl, "fuse=1")
        fl.set_fuse(2)
        os.environ["FDB_LAPF_GENERAL"] = "1"   # the six unit weights multiplied anyway
        run(fl, "fuse=2 default cfg, general-weights kernel")
        os.environ["FDB_LAPF_GENERAL"] = "0"
        for c in cfgs:
            os.environ["FDB_LAPF_CFG"] = str(c)
            for ci in cis:
                os.environ["FDB_TMA_CI"] = str(ci)
                run(fl, f"fuse=2 cfg={c} ci={ci}")
