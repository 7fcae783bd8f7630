"""Parity of the CUDA upwind engine (through the C ABI) with the oracle and the
golden fixtures.  Bit-exact on the field; 1e-12 relative on checksum/std (the
reference sums sequentially, SURVEY.md H4 -- tolerance from north_star)."""
import numpy as np
import pytest

import oracle
from conftest import golden, SEED

pytestmark = pytest.mark.gpu
C = oracle.c
RTOL = 1e-12  # north_star: "global checksum and the full field agree within 1e-12 relative"


def run_gpu(fb, init, steps, velocity=None, lengths=None, dt=None, kernel=None, ngpus=1):
    shape = init.shape
    nd = len(shape)
    velocity = [1.0] * nd if velocity is None else list(velocity)
    lengths = [1.0] * nd if lengths is None else list(lengths)
    with fb.Upwind(velocity, lengths, shape, ngpus=ngpus) as up:
        if kernel is not None:
            up.set_kernel(kernel)
        up.set_field(init)
        if dt is None:
            dt = up.default_dt()
        up.advect(steps, dt)
        return up.field(), up.checksum(), up.std(), up.kernel()


def delta(shape):
    f = np.zeros(shape)
    f.reshape(-1)[0] = 1.0
    return f


def test_config1_128cubed_10_steps_matches_reference_golden(gpu_fb):
    g = golden("upwind_128_s10.npz")
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [128] * 3) as up:  # ctor state = delta at cell 0
        assert up.default_dt() == float(g["dt"])
        assert up.kernel() == gpu_fb.FDB_KERNEL_TMA
        up.advect(10, up.default_dt())
        f = up.field()
        assert np.array_equal(f[:11, :11, :11], g["corner"])
        assert np.count_nonzero(f) == 286
        assert up.checksum() == pytest.approx(float(g["checksum"]), rel=RTOL)
        assert up.std() == pytest.approx(float(g["std"]), rel=RTOL)
        assert f"{up.checksum():g}" == "1"  # the reference's ctest regex "check sum: 1"
    ref = C.upwind_advect(delta((128,) * 3), 10)
    assert np.array_equal(f, ref)


@pytest.mark.parametrize("kernel", ["generic", "tma"])
@pytest.mark.parametrize("shape,steps", [((16, 16, 16), 100), ((64, 64, 64), 5), ((24, 20, 28), 7),
                                         ((96, 64, 128), 3), ((8, 34, 130), 4), ((5, 3, 6), 9),
                                         ((32, 16, 256), 3), ((3, 40, 260), 2)])
def test_random_field_bitwise_vs_oracle(gpu_fb, kernel, shape, steps):
    rng = np.random.default_rng(SEED)
    a = rng.random(shape)
    k = gpu_fb.FDB_KERNEL_GENERIC if kernel == "generic" else gpu_fb.FDB_KERNEL_TMA
    f, cs, sd, used = run_gpu(gpu_fb, a, steps, kernel=k)
    assert used == k
    ref = C.upwind_advect(a, steps)
    assert np.array_equal(f, ref)
    assert cs == pytest.approx(C.checksum(ref), rel=RTOL)
    assert sd == pytest.approx(C.std(ref), rel=RTOL)


@pytest.mark.parametrize("case", ["pos", "mixed", "neg"])
def test_random_field_golden_from_reference(gpu_fb, case):
    g = golden("upwind_random_24x20x28.npz")
    f, cs, sd, _ = run_gpu(gpu_fb, g["init"], int(g[f"{case}_steps"]), velocity=g[f"{case}_vel"],
                           lengths=g[f"{case}_len"], dt=float(g[f"{case}_dt"]))
    assert np.array_equal(f, g[f"{case}_out"])
    assert cs == pytest.approx(float(g[f"{case}_checksum"]), rel=RTOL)
    assert sd == pytest.approx(float(g[f"{case}_std"]), rel=RTOL)


def test_mass_wraps_around_many_times(gpu_fb):
    g = golden("upwind_16_s100.npz")
    for k in (gpu_fb.FDB_KERNEL_GENERIC, gpu_fb.FDB_KERNEL_TMA):
        f, cs, _, _ = run_gpu(gpu_fb, delta((16,) * 3), 100, kernel=k)
        assert np.array_equal(f, g["out"])
        assert cs == pytest.approx(float(g["checksum"]), rel=RTOL)


def test_1d_and_2d_instantiations(gpu_fb):
    g = golden("upwind_1d2d.npz")
    f1, _, _, k1 = run_gpu(gpu_fb, g["init1"], 5)
    f2, _, _, k2 = run_gpu(gpu_fb, g["init2"], 5)
    assert np.array_equal(f1, g["out1"]) and np.array_equal(f2, g["out2"])
    assert k1 == k2 == gpu_fb.FDB_KERNEL_GENERIC


def test_negative_and_zero_velocities(gpu_fb):
    rng = np.random.default_rng(SEED + 3)
    a = rng.random((12, 18, 20))
    for vel in ([-1, 2, 0.5], [1, 1, -1], [0.0, 1, 1]):
        f, _, _, _ = run_gpu(gpu_fb, a, 6, velocity=vel, lengths=[1, 2, 1], dt=0.004)
        assert np.array_equal(f, C.upwind_advect(a, 6, velocity=vel, lengths=[1, 2, 1], dt=0.004))


def test_tma_kernel_rejects_what_it_cannot_run(gpu_fb, monkeypatch):
    monkeypatch.setenv("FDB_NO_FLIP", "1")  # without axis mirroring a negative velocity needs the generic kernel
    with gpu_fb.Upwind([1.0, -1.0, 1.0], [1.0] * 3, [8, 8, 8]) as up:
        assert up.kernel() == gpu_fb.FDB_KERNEL_GENERIC
        with pytest.raises(gpu_fb.FdbError):
            up.set_kernel(gpu_fb.FDB_KERNEL_TMA)
    monkeypatch.delenv("FDB_NO_FLIP")
    with gpu_fb.Upwind([1.0, -1.0, 1.0], [1.0] * 3, [8, 8, 8]) as up:  # mirrored along axis 1 on the device
        assert up.kernel() == gpu_fb.FDB_KERNEL_TMA
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [8, 8, 9]) as up:  # odd rows are not 16-byte pitched
        assert up.kernel() == gpu_fb.FDB_KERNEL_GENERIC


def test_repeated_advect_calls_equal_one_call(gpu_fb):
    rng = np.random.default_rng(SEED + 4)
    a = rng.random((32, 32, 64))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        up.set_field(a)
        dt = up.default_dt()
        for n in (1, 2, 3):
            up.advect(n, dt)
        assert np.array_equal(up.field(), C.upwind_advect(a, 6))
        up.reset()
        up.advect(4, dt)
        assert np.array_equal(up.field(), C.upwind_advect(delta(a.shape), 4))
        t = up.last_timing()
        assert t["gpu_ms"] > 0 and t["cell_updates"] == 4 * a.size


def test_config2_512cubed_100_steps_corner_rule(gpu_fb):
    """BASELINE config 2 at full size.  S=100 < 128, so the non-zero corner is
    bit-identical to the reference's 128^3 x 100 run and every other
    cell is exactly zero; mass is conserved."""
    g = golden("upwind_128_s100.npz")
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [512] * 3) as up:
        assert up.kernel() == gpu_fb.FDB_KERNEL_TMA
        up.advect(100, up.default_dt())
        f = up.field()
        cs, sd = up.checksum(), up.std()
    assert np.array_equal(f[:101, :101, :101], g["corner"])
    assert np.count_nonzero(f) == int(g["nnz"])
    assert cs == pytest.approx(float(g["checksum"]), rel=RTOL)
    assert cs == pytest.approx(1.0, rel=RTOL)
    # std depends on N: sqrt(sum((f-mean)^2)/N^3), recomputed on the host from the field
    mean = cs / f.size
    assert sd == pytest.approx(float(np.sqrt(np.sum((f - mean) ** 2) / f.size)), rel=1e-10)


def test_linearity_at_full_plane_size(gpu_fb):
    """Size-independent property: the step is linear, A(x + 2y) == A(x) + 2 A(y) to rounding."""
    rng = np.random.default_rng(SEED + 5)
    shape = (8, 256, 512)
    x, y = rng.random(shape), rng.random(shape)
    fx = run_gpu(gpu_fb, x, 3)[0]
    fy = run_gpu(gpu_fb, y, 3)[0]
    fxy = run_gpu(gpu_fb, x + 2 * y, 3)[0]
    assert np.allclose(fxy, fx + 2 * fy, rtol=1e-13, atol=0)


def test_caller_stream(gpu_fb):
    import torch
    rng = np.random.default_rng(SEED + 6)
    a = rng.random((16, 32, 64))
    s = torch.cuda.Stream()
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        up.set_stream(s.cuda_stream)
        up.set_field(a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            up.advect_async(5, up.default_dt())
            e1.record()
        up.sync()
        s.synchronize()
        assert e0.elapsed_time(e1) > 0
        assert np.array_equal(up.field(), C.upwind_advect(a, 5))
        up.set_stream(None)


def test_launch_counter_moves(gpu_fb):
    before = gpu_fb.launch_count()
    run_gpu(gpu_fb, delta((16, 16, 16)), 3)
    assert gpu_fb.launch_count() >= before + 3


@pytest.mark.parametrize("ngpus", [2, 4, 8])
@pytest.mark.parametrize("kernel", ["generic", "tma"])
def test_in_process_slabs_match_single_domain_oracle(gpu_fb, ngpus, kernel):
    if gpu_fb.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(SEED + 7)
    a = rng.random((16, 24, 64))
    k = gpu_fb.FDB_KERNEL_GENERIC if kernel == "generic" else gpu_fb.FDB_KERNEL_TMA
    f, cs, _, _ = run_gpu(gpu_fb, a, 20, kernel=k, ngpus=ngpus)
    ref = C.upwind_advect(a, 20)
    assert np.array_equal(f, ref)
    f1, cs1, _, _ = run_gpu(gpu_fb, a, 20, kernel=k, ngpus=1)
    assert cs == cs1  # the reduction is bitwise partition-invariant
    # negative axis-0 velocity: the ghost plane sits above the slab
    dt = C.upwind_dt(a.shape, [1.0] * 3, [1.0] * 3)   # |v|: the reference's signed formula would give dt < 0
    f, _, _, _ = run_gpu(gpu_fb, a, 9, velocity=[-1, 1, 1], dt=dt, ngpus=ngpus)
    assert np.array_equal(f, C.upwind_advect(a, 9, velocity=[-1, 1, 1], dt=dt))


def test_invalid_slab_count_is_a_decomposition_error(gpu_fb):
    with pytest.raises(gpu_fb.FdbError) as e:
        gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [9, 8, 8], ngpus=2) if gpu_fb.device_count() >= 2 else \
            gpu_fb.slab_partition(9, 2, 0)
    assert e.value.code == -5


@pytest.mark.parametrize("fuse", [1, 2, 3, 4])
@pytest.mark.parametrize("shape,steps", [((16, 16, 16), 50), ((24, 20, 28), 7), ((64, 64, 64), 12),
                                         ((9, 34, 130), 8), ((32, 16, 256), 6), ((5, 40, 260), 9),
                                         ((40, 100, 128), 5)])
def test_fused_sweeps_are_bitwise_equal_to_single_steps(gpu_fb, fuse, shape, steps):
    """Temporal blocking (T steps per sweep) must not change a single bit: halo rows/columns,
    periodic wrap in all three axes, ragged tiles, remainders (steps % T != 0)."""
    rng = np.random.default_rng(SEED + 20)
    a = rng.random(shape)
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, shape) as up:
        up.set_fuse(fuse)
        up.set_field(a)
        up.advect(steps, up.default_dt())
        f = up.field()
    assert np.array_equal(f, C.upwind_advect(a, steps))


def test_fused_config1_golden(gpu_fb):
    g = golden("upwind_128_s100.npz")
    for fuse in (2, 3, 4):
        with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [128] * 3) as up:
            up.set_fuse(fuse)
            up.advect(100, up.default_dt())
            f = up.field()
            assert np.array_equal(f[:101, :101, :101], g["corner"])
            assert np.count_nonzero(f) == int(g["nnz"])
            assert up.checksum() == pytest.approx(float(g["checksum"]), rel=RTOL)


def test_fuse_argument_validation(gpu_fb, monkeypatch):
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [16, 16, 16]) as up:
        with pytest.raises(gpu_fb.FdbError):
            up.set_fuse(-1)
        with pytest.raises(gpu_fb.FdbError):
            up.set_fuse(9)
        up.set_fuse(0)  # auto
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [16, 16, 15]) as up:
        with pytest.raises(gpu_fb.FdbError):
            up.set_fuse(2)  # odd last extent: no TMA path, no fused path
    with gpu_fb.Upwind([1.0, -1.0, 1.0], [1.0] * 3, [16, 16, 16]) as up:
        up.set_fuse(2)      # negative velocity: mirrored on the device, the fused kernel runs
    monkeypatch.setenv("FDB_NO_FLIP", "1")
    with gpu_fb.Upwind([1.0, -1.0, 1.0], [1.0] * 3, [16, 16, 16]) as up:
        with pytest.raises(gpu_fb.FdbError):
            up.set_fuse(2)  # ... unless mirroring is switched off: generic kernel only


@pytest.mark.parametrize("ngpus", [2, 4])
@pytest.mark.parametrize("fuse", [2, 3, 4])
def test_fused_in_process_slabs(gpu_fb, ngpus, fuse):
    if gpu_fb.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(SEED + 21)
    a = rng.random((8 * ngpus, 24, 64))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, ngpus=ngpus) as up:
        up.set_fuse(fuse)
        up.set_field(a)
        up.advect(8 * ngpus + 5, up.default_dt())
        f = up.field()
    assert np.array_equal(f, C.upwind_advect(a, 8 * ngpus + 5))


def test_config3_1024cubed_100_steps_corner_rule(gpu_fb):
    """BASELINE config 3 on one GPU (8 GiB per field): bitwise corner + exact zeros."""
    g = golden("upwind_128_s100.npz")
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, [1024] * 3) as up:
        up.advect(100, up.default_dt())
        cs = up.checksum()
        f = up.field()
    assert np.array_equal(f[:101, :101, :101], g["corner"])
    assert np.count_nonzero(f[101:]) == 0 and np.count_nonzero(f[:101, 101:]) == 0 \
        and np.count_nonzero(f[:101, :101, 101:]) == 0
    assert cs == pytest.approx(float(g["checksum"]), rel=RTOL)


@pytest.mark.parametrize("ngpus", [2, 4])
def test_repeated_advect_calls_on_slabs_refresh_deeper_ghosts(gpu_fb, ngpus):
    """advect(4) ends on a one-step remainder sweep (one fresh ghost plane); the next call starts
    with a three-step sweep that reads three -- the runtime must refresh them first."""
    if gpu_fb.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(SEED + 30)
    a = rng.random((8 * ngpus, 24, 64))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, ngpus=ngpus) as up:
        up.set_field(a)
        dt = up.default_dt()
        for n in (4, 3, 7, 1, 6):
            up.advect(n, dt)
        f = up.field()
    assert np.array_equal(f, C.upwind_advect(a, 21))


@pytest.mark.parametrize("vel", [[-1.0, 1.0, 1.0], [1.0, -0.5, 2.0], [1.0, 1.0, -1.0], [-1.0, -2.0, -0.25]])
@pytest.mark.parametrize("fuse", [1, 3])
def test_negative_velocities_run_the_tiled_kernels_on_a_mirrored_grid(gpu_fb, vel, fuse):
    """A negative velocity along an axis is a positive one on the mirrored grid with the same coefficient
    bits; uploads, downloads and the ctor's delta are mirrored, so the caller never sees it."""
    rng = np.random.default_rng(SEED + 31)
    shape, lengths, dt = (16, 24, 64), [1.0, 1.5, 2.0], 0.002
    a = rng.random(shape)
    with gpu_fb.Upwind(vel, lengths, shape) as up:
        assert up.kernel() == gpu_fb.FDB_KERNEL_TMA
        up.set_fuse(fuse)
        up.set_field(a)
        assert np.array_equal(up.field(), a)                      # upload + download round trip
        up.advect(7, dt)
        ref = C.upwind_advect(a, 7, velocity=vel, lengths=lengths, dt=dt)
        assert np.array_equal(up.field(), ref)
        assert np.array_equal(up.slab(), ref)
        assert abs(up.checksum() - C.checksum(ref)) <= 1e-12 * abs(C.checksum(ref))
        assert abs(up.std() - C.std(ref)) <= 1e-12 * C.std(ref)
        up.advect(2, dt)                                          # carries on from the mirrored state
        assert np.array_equal(up.field(), C.upwind_advect(a, 9, velocity=vel, lengths=lengths, dt=dt))
        up.reset()                                                # the delta sits at reference cell 0
        assert np.array_equal(up.field(), delta(shape))
        up.advect(5, dt)
        assert np.array_equal(up.field(), C.upwind_advect(delta(shape), 5, velocity=vel, lengths=lengths, dt=dt))


def test_default_dt_is_positive_for_negative_velocities(gpu_fb):
    """the ABI helper uses |v| like drivers/upwindCuda.cxx (same bits as the reference for v > 0)."""
    rng = np.random.default_rng(SEED + 50)
    a = rng.random((6, 10, 36))
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
        assert up.default_dt() == C.upwind_dt(a.shape, [1.0] * 3, [1.0] * 3)
    vel = [-1.0, 2.0, -0.5]
    with gpu_fb.Upwind(vel, [1.0] * 3, a.shape) as up:
        dt = up.default_dt()
        assert dt > 0 and dt == C.upwind_dt(a.shape, [abs(v) for v in vel], [1.0] * 3)
        up.set_field(a)
        up.advect(6, dt)
        out = up.field()
    assert np.array_equal(out, C.upwind_advect(a, 6, velocity=vel, dt=dt))
    assert abs(out).max() <= abs(a).max()   # a stable (monotone) step, not an anti-diffusive one


@pytest.mark.parametrize("halo", ["single", "copy"])
@pytest.mark.parametrize("ngpus", [2, 4])
def test_slab_ring_halo_transports_agree(gpu_fb, ngpus, halo, monkeypatch):
    """The default on a one-sided ring is a boundary launch that pushes its planes with peer stores + an interior
    launch; FDB_HALO=single runs ONE launch per sweep (top chunk first, device-side ACK / ghost waits and ghost flag),
    FDB_HALO=copy the copy engines.  All three must give the single-domain oracle's bits, over plans that mix sweep
    depths and repeated calls."""
    if gpu_fb.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(SEED + 60)
    a = rng.random((40 * ngpus, 45, 260))

    def run():
        with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, ngpus=ngpus) as up:
            up.set_field(a)
            dt = up.default_dt()
            for n in (10, 3, 7, 1, 12):
                up.advect(n, dt)
            return up.field(), dt

    ref_default, dt = run()
    assert np.array_equal(ref_default, C.upwind_advect(a, 33, dt=dt))
    monkeypatch.setenv("FDB_HALO", halo)
    other, _ = run()
    assert np.array_equal(other, ref_default)


@pytest.mark.parametrize("ngpus", [2, 4])
@pytest.mark.parametrize("vel", [(-1.0, 1.0, 1.0), (-1.0, -0.5, 2.0), (1.0, -1.0, -1.0)])
def test_negative_velocities_on_slab_rings_run_the_tiled_kernels(gpu_fb, ngpus, vel):
    """VERDICT r1 (missing 3): on several slabs a negative velocity used to mean the generic kernel.  Now every slab is
    mirrored in place and, for the slab axis, the ring runs backwards (ref: upDirection, upwind.cxx:34-35)."""
    if gpu_fb.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(SEED + 70)
    a = rng.random((16 * ngpus, 40, 132))
    dt = C.upwind_dt(a.shape, [abs(v) for v in vel], [1.0] * 3)
    with gpu_fb.Upwind(list(vel), [1.0] * 3, a.shape, ngpus=ngpus) as up:
        assert up.kernel() == gpu_fb.FDB_KERNEL_TMA and "upwind3d_fused" in up.describe() and "mirrored" in up.describe()
        assert up.default_dt() == dt
        up.set_field(a)
        for n in (7, 3, 5):
            up.advect(n, dt)
        out = up.field()
        ref = C.upwind_advect(a, 15, velocity=list(vel), dt=dt)
        assert np.array_equal(out, ref)
        cs, sums = up.checksum(), up.plane_sums()
        assert np.allclose(sums, ref.reshape(ref.shape[0], -1).sum(axis=1), rtol=1e-13)
        with gpu_fb.Upwind(list(vel), [1.0] * 3, a.shape) as single:
            single.set_field(a)
            for n in (7, 3, 5):
                single.advect(n, dt)
            assert np.array_equal(single.field(), ref)
            assert single.checksum() == cs          # the reduction stays partition-invariant on mirrored fields
        up.reset()                                   # the ctor's delta at logical cell 0
        up.advect(4, dt)
        d = np.zeros(a.shape); d.reshape(-1)[0] = 1.0
        assert np.array_equal(up.field(), C.upwind_advect(d, 4, velocity=list(vel), dt=dt))
        up.fill_random(11)
        assert np.array_equal(up.field(), oracle.hash_field(11, a.shape))


def test_describe_names_the_kernel_and_the_reason_for_the_generic_one(gpu_fb):
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, (16, 16, 64)) as up:
        assert up.describe().startswith("upwind3d_fused_lean_kernel<T=4>")
        up.set_fuse(1)
        assert up.describe().startswith("upwind3d_tma_kernel")
    with gpu_fb.Upwind([1.0] * 3, [1.0] * 3, (8, 8, 9)) as up:
        assert "upwind_generic_kernel" in up.describe() and "odd" in up.describe()
    with gpu_fb.Upwind([1.0] * 2, [1.0] * 2, (8, 8)) as up:
        assert "upwind_generic_kernel" in up.describe() and "3-D" in up.describe()
    off, w = oracle.laplacian_stencil(3)
    st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
    with gpu_fb.Filter((8, 16, 128), [0.0] * 3, [1.0] * 3, st) as fl:
        assert "lap7_fused2_lean_kernel" in fl.describe()
    with gpu_fb.Filter((8, 10, 12), [0.0] * 3, [1.0] * 3, st) as fl:
        assert fl.describe().startswith("lap7_tma_kernel")      # ragged tiles
    with gpu_fb.Filter((8, 10, 13), [0.0] * 3, [1.0] * 3, st) as fl:
        assert "stencil_generic_kernel" in fl.describe() and "odd" in fl.describe()
