"""Seeded random sweep of shapes / step counts / fuse depths / kernels against the oracle.
Small grids (the oracle runs each in milliseconds); meant to shake out tile-edge cases:
extents below one tile, ragged tiles, wrap boxes, slabs thinner than the fuse depth."""
import numpy as np
import pytest

import oracle
from conftest import SEED

pytestmark = pytest.mark.gpu
C = oracle.c


def test_random_upwind_cases_bitwise(gpu_fb):
    rng = np.random.default_rng(SEED + 100)
    failures = []
    for case in range(60):
        n0 = int(rng.integers(1, 21))
        n1 = int(rng.integers(2, 71))
        n2 = int(rng.integers(2, 151)) * 2
        steps = int(rng.integers(1, 13))
        fuse = int(rng.integers(0, 5))
        a = rng.random((n0, n1, n2))
        with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
            try:
                up.set_fuse(fuse)
            except gpu_fb.FdbError:
                up.set_fuse(0)  # grid too small for that depth: the library says so, auto still works
            up.set_field(a)
            dt = up.default_dt()
            up.advect(steps, dt)
            out = up.field()
        if not np.array_equal(out, C.upwind_advect(a, steps)):
            failures.append((a.shape, steps, fuse))
    assert not failures, failures


def test_random_velocity_signs_and_dims(gpu_fb):
    rng = np.random.default_rng(SEED + 101)
    for case in range(25):
        nd = int(rng.integers(1, 4))
        shape = tuple(int(rng.integers(2, 40)) for _ in range(nd))
        vel = [float(rng.choice([-2.0, -1.0, 0.5, 1.0, 3.0])) for _ in range(nd)]
        lens = [float(rng.choice([0.5, 1.0, 2.0])) for _ in range(nd)]
        steps = int(rng.integers(1, 9))
        dt = 0.05 * min(l / n for l, n in zip(lens, shape)) / max(abs(v) for v in vel)
        a = rng.random(shape)
        with gpu_fb.Upwind(vel, lens, shape) as up:
            up.set_field(a)
            up.advect(steps, dt)
            out = up.field()
        ref = C.upwind_advect(a, steps, velocity=vel, lengths=lens, dt=dt)
        assert np.array_equal(out, ref), (shape, vel, lens, steps)


def test_random_stencils_bitwise(gpu_fb):
    rng = np.random.default_rng(SEED + 102)
    for case in range(25):
        nd = int(rng.integers(1, 4))
        shape = tuple(int(rng.integers(3, 34)) for _ in range(nd))
        nb = min(int(rng.integers(1, 10)), 5 ** nd)  # only 5^nd distinct offsets exist in [-2, 2]^nd
        offs = set()
        while len(offs) < nb:
            offs.add(tuple(int(x) for x in rng.integers(-2, 3, size=nd)))
        st = {o: float(rng.standard_normal()) for o in offs}
        a = rng.random(shape)
        niter = int(rng.integers(1, 4))
        with gpu_fb.Filter(shape, [0.0] * nd, [1.0] * nd, st) as fl:
            fl.set_input(a)
            fl.iterate(niter)
            out = fl.get()
        ref = a
        for _ in range(niter):
            ref = C.stencil_apply(ref, np.array(list(st.keys()), dtype=np.int32).reshape(len(st), nd),
                                  np.array(list(st.values())))
        assert np.array_equal(out, ref), (shape, st, niter)


def test_random_seven_point_subsets_on_tileable_planes(gpu_fb):
    rng = np.random.default_rng(SEED + 103)
    full = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]
    for case in range(16):
        n0 = int(rng.integers(1, 12))
        n1 = int(rng.choice([8, 16, 32, 48]))
        n2 = int(rng.choice([32, 64, 128, 256]))
        keep = [o for o in full if rng.random() < 0.75] or [(0, 0, 0)]
        st = {o: float(rng.standard_normal()) for o in keep}
        a = rng.random((n0, n1, n2))
        with gpu_fb.Filter(a.shape, [0.0] * 3, [1.0] * 3, st) as fl:
            used = fl.kernel()
            fl.set_input(a)
            fl.iterate(3)
            out = fl.get()
        ref = a
        for _ in range(3):
            ref = C.stencil_apply(ref, np.array(list(st.keys()), dtype=np.int32), np.array(list(st.values())))
        assert np.array_equal(out, ref), (a.shape, st, used)
