"""The shipped persistent kernels where round 1 left holes: (1) every CTA walking MANY work items on non-zero
data (FDB_MAX_CTAS caps the grid, FDB_TMA_CI shortens the chunks so items also chain along the marching axis:
stage/phase/carry hand-over between items), (2) a full-size random field against the oracle, (3) run-to-run
determinism at full size (an mbarrier/proxy ordering bug would show as flaky bits), (4) the device-side inputs."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from conftest import SEED

pytestmark = pytest.mark.gpu
C = oracle.c


@pytest.fixture
def few_ctas():
    old = {k: os.environ.get(k) for k in ("FDB_MAX_CTAS", "FDB_TMA_CI")}
    os.environ["FDB_MAX_CTAS"] = "3"
    os.environ["FDB_TMA_CI"] = "8"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def lap7():
    off, w = oracle.laplacian_stencil(3)
    return off, w, {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}


@pytest.mark.parametrize("fuse", [1, 2, 3, 4])
def test_upwind_many_items_per_cta_bitwise(gpu_fb, few_ctas, fuse):
    rng = np.random.default_rng(SEED + 200 + fuse)
    for shape, steps in (((40, 100, 128), 7), ((37, 45, 300), 5), ((64, 64, 256), 9), ((19, 130, 66), 4)):
        a = rng.random(shape)
        with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, shape) as up:
            up.set_fuse(fuse)
            up.set_field(a)
            up.advect(steps, up.default_dt())
            out = up.field()
        assert np.array_equal(out, C.upwind_advect(a, steps)), (shape, steps, fuse)


@pytest.mark.parametrize("fuse", [1, 2])
def test_laplacian_many_items_per_cta_bitwise(gpu_fb, few_ctas, fuse):
    off, w, st = lap7()
    rng = np.random.default_rng(SEED + 210 + fuse)
    for shape, niter in (((40, 64, 256), 4), ((33, 48, 128), 3), ((24, 96, 384), 5)):
        a = rng.random(shape) - 0.5
        with gpu_fb.Filter(shape, [0.0] * 3, [1.0] * 3, st) as fl:
            fl.set_fuse(fuse)
            assert fl.kernel() == gpu_fb.FDB_KERNEL_TMA
            fl.set_input(a)
            fl.iterate(niter)
            out = fl.get()
        ref = a
        for _ in range(niter):
            ref = C.stencil_apply(ref, off, w)
        assert np.array_equal(out, ref), (shape, niter, fuse)


def test_many_items_per_cta_on_in_process_slabs(gpu_fb, few_ctas):
    if gpu_fb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(SEED + 220)
    a = rng.random((48, 60, 256))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, ngpus=2) as up:
        up.set_field(a)
        up.advect(11, up.default_dt())
        assert np.array_equal(up.field(), C.upwind_advect(a, 11))


def test_device_side_inputs_match_their_host_restatements(gpu_fb):
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, (12, 20, 36)) as up:
        up.fill_random(SEED)
        f = up.field()
        assert np.array_equal(f, oracle.hash_field(SEED, f.shape))
        sums = up.plane_sums()
        assert sums.shape == (12,) and float(np.sum(sums)) == pytest.approx(up.checksum(), rel=1e-15)
        assert np.allclose(sums, f.reshape(12, -1).sum(axis=1), rtol=1e-13)
    with gpu_fb.Upwind([-1.0, 1.0, -1.0], [1.0] * 3, (6, 10, 36)) as up:   # mirrored device grid
        up.fill_random(7)
        assert np.array_equal(up.field(), oracle.hash_field(7, (6, 10, 36)))
        assert np.allclose(up.plane_sums(), up.field().reshape(6, -1).sum(axis=1), rtol=1e-13)
    _, _, st = lap7()
    for dims in ((16, 32, 128), (8, 24, 20)):
        with gpu_fb.Filter(dims, [0.0] * 3, [1.0] * 3, st) as fl:
            fl.set_input_separable(fl.laplacian_factors())
            x = fl.get(gpu_fb.FDB_INPUT)
            assert np.array_equal(x, C.laplacian_input(dims))     # the reference's setInData(func), bit for bit
            assert fl.sumsq("input") == pytest.approx(float(np.sum(x * x)), rel=1e-13)
            fl.fill_random(3)
            assert np.array_equal(fl.get(gpu_fb.FDB_INPUT), oracle.hash_field(3, dims))
    off2, w2 = oracle.laplacian_stencil(2)
    st2 = {tuple(int(v) for v in o): float(c) for o, c in zip(off2, w2)}
    with gpu_fb.Filter((40, 24), [0.0] * 2, [1.0] * 2, st2) as fl:
        fl.set_input_separable(fl.laplacian_factors())
        assert np.array_equal(fl.get(gpu_fb.FDB_INPUT), C.laplacian_input((40, 24)))


def test_fullsize_random_512_cubed_bitwise(gpu_fb):
    """BASELINE configs[1]'s grid with a non-zero field everywhere: 148 CTAs x several items each."""
    dims = (512, 512, 512)
    steps = 7   # one [3, 2, 2] plan
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, dims) as up:
        up.fill_random(SEED)
        up.advect(steps, up.default_dt())
        out = up.field()
    ref = C.upwind_advect(oracle.hash_field(SEED, dims), steps)
    assert np.array_equal(out, ref)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("fuse", [1, 2, 3])
def test_upwind_run_to_run_determinism_512_cubed(gpu_fb, fuse):
    got = []
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, (512, 512, 512)) as up:
        up.set_fuse(fuse)
        for _ in range(2):
            up.fill_random(SEED + fuse)
            up.advect(12, up.default_dt())
            got.append(digest(up.field()))
    assert got[0] == got[1]


def test_laplacian_fused_run_to_run_determinism_512_cubed(gpu_fb):
    _, _, st = lap7()
    got = []
    with gpu_fb.Filter((512, 512, 512), [0.0] * 3, [1.0] * 3, st) as fl:
        assert fl.fuse() == 2
        for _ in range(2):
            fl.fill_random(SEED)
            fl.iterate(4)
            got.append(digest(fl.get()))
    assert got[0] == got[1]


def test_cuda_graph_replay_of_small_plans_is_bitwise_and_counts_launches(gpu_fb):
    """Launch-bound grids replay a captured CUDA graph of the whole sweep plan (runtime.cu: field_run_sweeps)."""
    rng = np.random.default_rng(SEED + 230)
    a = rng.random((32, 40, 64))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        up.set_field(a)
        dt = up.default_dt()
        n0 = gpu_fb.launch_count()
        for _ in range(3):                 # same plan three times: one capture, then replays
            up.advect(10, dt)
        per_call = (gpu_fb.launch_count() - n0) // 3
        assert per_call == 3               # [4, 4, 2]
        up.advect(7, dt)                   # another plan, odd parity start
        up.advect(10, 0.5 * dt)            # same depths, other coefficients: must not replay the old graph
        out = up.field()
    ref = C.upwind_advect(a, 37, dt=dt)
    ref = C.upwind_advect(ref, 10, dt=0.5 * dt)
    assert np.array_equal(out, ref)
    off, w, st = lap7()
    b = rng.random((16, 32, 128)) - 0.5
    with gpu_fb.Filter(b.shape, [0.0] * 3, [1.0] * 3, st) as fl:
        fl.set_input(b)
        fl.iterate(4)
        fl.iterate(4)
        fl.iterate(3)
        r = b
        for _ in range(11):
            r = C.stencil_apply(r, off, w)
        assert np.array_equal(fl.get(), r)


def test_cuda_graph_cache_eviction_with_replays_still_queued(gpu_fb):
    """More than 16 distinct plans on one small field: the graph cache drops its oldest executable graphs while earlier
    replays are still queued on the stream (no sync between the calls).  Bit-exact after 960 time steps
    (profiles/r02gg_graph_cache_eviction.txt)."""
    rng = np.random.default_rng(SEED + 231)
    a = rng.random((16, 24, 64))
    total = 0
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        up.set_field(a)
        dt = up.default_dt()
        for _ in range(2):
            for steps in range(5, 45, 2):   # 20 plans; the field's parity differs between the two rounds for odd totals
                up.advect_async(steps, dt)
                total += steps
        out = up.field()
    assert np.array_equal(out, C.upwind_advect(a, total, dt=dt))
