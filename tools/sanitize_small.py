"""Tiny run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

rng = np.random.default_rng(5)
a = rng.random(tuple(int(x) for x in os.environ.get("SAN_SHAPE", "7,40,132").split(",")))
FUSES = [int(x) for x in os.environ.get("SAN_FUSES", "1,2,3,4").split(",") if x]
for fuse in FUSES:
    with fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        up.set_fuse(fuse)
        up.set_field(a)
        up.advect(5, up.default_dt())
        assert np.array_equal(up.field(), oracle.c.upwind_advect(a, 5)), fuse
        up.checksum(); up.std()
if os.environ.get("SAN_ONLY_FUSED"):
    print("sanitize_small ok (fused upwind only)")
    sys.exit(0)
with fb.Upwind([1.0, -1.0, 1.0], [1.0] * 3, (6, 9, 11)) as up:   # generic kernel
    b = rng.random((6, 9, 11))
    up.set_field(b)
    up.advect(3, 0.01)
    assert np.array_equal(up.field(), oracle.c.upwind_advect(b, 3, velocity=[1, -1, 1], dt=0.01))
off, w = oracle.laplacian_stencil(3)
st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
x = rng.random((5, 16, 64))
for kern in (fb.FDB_KERNEL_TMA, fb.FDB_KERNEL_GENERIC):
    with fb.Filter(x.shape, [0.0] * 3, [1.0] * 3, st) as fl:
        fl.set_kernel(kern)
        fl.set_input(x)
        fl.iterate(3)
        r = x
        for _ in range(3):
            r = oracle.c.stencil_apply(r, off, w)
        assert np.array_equal(fl.get(), r)
# fused two-apply 7-point kernel (unit-weight and general kernels), odd count ends on the single-apply kernel
y = rng.random((5, 16, 128)) - 0.5
for weights in (w, rng.standard_normal(7)):
    stw = {tuple(int(v) for v in o): float(c) for o, c in zip(off, weights)}
    with fb.Filter(y.shape, [0.0] * 3, [1.0] * 3, stw) as fl:
        assert fl.fuse() == 2
        fl.set_input(y)
        fl.iterate(3)
        r = y
        for _ in range(3):
            r = oracle.c.stencil_apply(r, off, weights)
        assert np.array_equal(fl.get(), r)
if os.environ.get("SAN_SKIP_MIRROR"):
    print("sanitize_small ok (without the mirrored upwind case)")
    sys.exit(0)
# negative velocities on the mirrored grid (mirror_kernel + tiled kernels)
c = rng.random((6, 10, 36))
with fb.Upwind([-1.0, 1.0, -0.5], [1.0] * 3, c.shape) as up:
    up.set_field(c)
    up.advect(4, 0.004)
    assert np.array_equal(up.field(), oracle.c.upwind_advect(c, 4, velocity=[-1.0, 1.0, -0.5], dt=0.004))
print("sanitize_small ok")
