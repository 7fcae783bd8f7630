"""One short run of the fused two-apply kernel for ncu: python tools/prof_lapfused.py N"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
off, w = oracle.laplacian_stencil(3)
st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
with fb.Filter([N] * 3, [0.0] * 3, [1.0] * 3, st) as fl:
    slab = np.random.default_rng(3).random((64, N, N))
    fl.set_input_slab(np.concatenate([slab] * (N // 64)))
    fl.iterate(6)
    print("ok", fl.fuse(), fl.last_timing())
