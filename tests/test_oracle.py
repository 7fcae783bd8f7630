"""The CPU oracle against the reference: golden fixtures generated from the
untouched reference build (tests/golden/make_golden.py), the reference build
itself when oracle/_ref is present, analytic solutions, and the published
checksum contract.  No GPU needed."""
import math

import numpy as np
import pytest

import oracle
from conftest import golden, SEED

C = oracle.c


def delta(shape):
    f = np.zeros(shape)
    f.reshape(-1)[0] = 1.0
    return f


def test_config1_golden_corner_checksum_std():
    g = golden("upwind_128_s10.npz")
    out = C.upwind_advect(delta((128,) * 3), 10)
    assert np.array_equal(out[:11, :11, :11], g["corner"])
    assert np.count_nonzero(out) == int(g["nnz"]) == 286
    assert C.checksum(out) == float(g["checksum"]) == 1.0000000000000011
    assert C.std(out) == float(g["std"]) == 0.00011898287050402086
    assert C.upwind_dt([128] * 3, [1.0] * 3, [1.0] * 3) == float(g["dt"])
    # the survey's spot values
    assert out[0, 0, 0] == 0.028247524900000005
    assert out[0, 0, 1] == 0.040353607 and out[1, 0, 0] == 0.040353607
    assert out[0, 1, 0] == 0.040353606999999993
    assert out[1, 1, 1] == 0.059295095999999999 == out.max()


def test_reference_ctest_contract_checksum_prints_as_one():
    # upwind/cxx/CMakeLists.txt:54-60 matches "check sum: 1" at 6 significant digits
    out = C.upwind_advect(delta((32,) * 3), 10)
    assert f"{C.checksum(out):g}" == "1"


def test_corner_block_is_independent_of_n():
    # for power-of-two N the coefficient is exactly the same double
    ref = golden("upwind_128_s10.npz")["corner"]
    for n in (16, 32, 64):
        out = C.upwind_advect(delta((n,) * 3), 10)
        assert np.array_equal(out[:11, :11, :11], ref)


def test_100_steps_golden():
    g = golden("upwind_128_s100.npz")
    out = C.upwind_advect(delta((128,) * 3), 100)
    assert np.array_equal(out[:101, :101, :101], g["corner"])
    assert np.count_nonzero(out) == int(g["nnz"])
    assert C.checksum(out) == float(g["checksum"])
    assert C.std(out) == float(g["std"])


def test_analytic_multinomial():
    # f[i,j,k](S) = S!/(i!j!k!(S-i-j-k)!) 0.1^(i+j+k) 0.7^(S-i-j-k)
    S = 10
    out = C.upwind_advect(delta((32,) * 3), S)
    for (i, j, k) in [(0, 0, 0), (1, 0, 0), (1, 1, 1), (2, 3, 1), (4, 4, 2), (10, 0, 0)]:
        m = S - i - j - k
        exact = math.factorial(S) / (math.factorial(i) * math.factorial(j) * math.factorial(k) *
                                     math.factorial(m)) * 0.1 ** (i + j + k) * 0.7 ** m
        assert out[i, j, k] == pytest.approx(exact, rel=1e-13)


@pytest.mark.parametrize("case", ["pos", "mixed", "neg"])
def test_random_field_golden(case):
    g = golden("upwind_random_24x20x28.npz")
    out = C.upwind_advect(g["init"], int(g[f"{case}_steps"]), velocity=g[f"{case}_vel"],
                          lengths=g[f"{case}_len"], dt=float(g[f"{case}_dt"]))
    assert np.array_equal(out, g[f"{case}_out"])
    assert C.checksum(out) == float(g[f"{case}_checksum"])
    assert C.std(out) == float(g[f"{case}_std"])
    npout = oracle.np_upwind_advect(g["init"], int(g[f"{case}_steps"]), velocity=g[f"{case}_vel"],
                                    lengths=g[f"{case}_len"], dt=float(g[f"{case}_dt"]))
    assert np.array_equal(npout, out)


def test_wraps_many_times_golden():
    g = golden("upwind_16_s100.npz")
    out = C.upwind_advect(delta((16,) * 3), 100)
    assert np.array_equal(out, g["out"])
    assert C.checksum(out) == float(g["checksum"])


def test_1d_and_2d_golden():
    g = golden("upwind_1d2d.npz")
    assert np.array_equal(C.upwind_advect(g["init1"], 5), g["out1"])
    assert np.array_equal(C.upwind_advect(g["init2"], 5), g["out2"])
    assert np.array_equal(oracle.np_upwind_advect(g["init2"], 5), g["out2"])


def test_laplacian_golden():
    g = golden("laplacian_16.npz")
    off, w = oracle.laplacian_stencil(3)
    so, sw = C.sort_branches(off, w)
    # the reference's std::map order
    assert so.tolist() == [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0]]
    assert sw.tolist() == [1, 1, 1, -6, 1, 1, 1]
    x = C.laplacian_input([16] * 3)
    assert np.array_equal(x, g["input"])
    y = C.stencil_apply(x, off, w)
    assert np.array_equal(y, g["out1"])
    assert np.array_equal(oracle.np_stencil_apply(x, off, w), g["out1"])
    for _ in range(9):
        y = C.stencil_apply(y, off, w)
    assert np.array_equal(y, g["out10"])  # roundoff-amplifying: only bit-exact order survives (H1)
    y32 = C.stencil_apply(C.laplacian_input([32] * 3), off, w)
    assert np.abs(y32).max() == float(g["max32"]) == 0.11363088994787285
    assert y32[8, 8, 8] == float(g["probe32"])


def test_laplacian_2d_golden():
    g = golden("laplacian2d_32.npz")
    off, w = oracle.laplacian_stencil(2)
    x = C.laplacian_input([32, 32])
    assert np.array_equal(x, g["input"])
    assert np.array_equal(C.stencil_apply(x, off, w), g["out1"])
    # non power-of-two extent: the reference's wrap of index -1 is not periodic;
    # the oracle reproduces it on request, the product implements the true periodic wrap
    x24 = C.laplacian_input([24, 24])
    assert np.array_equal(x24, g["input24"])
    assert np.array_equal(C.stencil_apply(x24, off, w, ref_wrap_quirk=True), g["out24_quirk"])
    assert not np.array_equal(C.stencil_apply(x24, off, w), g["out24_quirk"])


def test_upwindmpi_stencil_golden():
    g = golden("upwindmpi_16.npz")
    off, w = oracle.upwind_filter_stencil(16)
    assert np.array_equal(off, g["offsets"]) and np.array_equal(w, g["weights"])
    x = g["init"]
    for _ in range(3):
        x = C.stencil_apply(x, off, w)
    assert np.array_equal(x, g["out3"])


def test_stencil2d_golden():
    g = golden("stencil2d_8.npz")
    assert np.array_equal(C.stencil_apply(g["init"], g["offsets"], g["weights"]), g["out"])


def test_two_formulations_differ_within_1e12():
    # upwind.cxx vs upwindMpi.cxx+Filter agree to ~1e-14, not bitwise
    rng = np.random.default_rng(SEED)
    a = rng.random((16, 16, 16))
    off, w = oracle.upwind_filter_stencil(16)
    x = a
    for _ in range(20):
        x = C.stencil_apply(x, off, w)
    y = C.upwind_advect(a, 20)
    assert np.max(np.abs(x - y) / np.abs(y)) < 1e-12


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built on this machine")
class TestAgainstCompiledReference:
    def test_upwind_bitwise(self):
        r = oracle.ref()
        rng = np.random.default_rng(SEED + 1)
        for shape, vel, lens in [((12, 10, 14), [1, 1, 1], [1, 1, 1]), ((6, 30, 8), [0.3, -1, 2], [2, 1, 0.5]),
                                 ((40,), [1], [1]), ((7, 9), [-1, 1], [1, 1])]:
            a = rng.random(shape)
            dt = 0.05 * min(l / n for l, n in zip(lens, shape))
            g = r.upwind_run(shape, 6, init=a, velocity=vel, lengths=lens, dt=dt)
            out = C.upwind_advect(a, 6, velocity=vel, lengths=lens, dt=dt)
            assert np.array_equal(out, g["field"])
            assert C.checksum(out) == g["checksum"] and C.std(out) == g["std"]

    def test_filter_bitwise_including_wrap_quirk(self):
        r = oracle.ref()
        rng = np.random.default_rng(SEED + 2)
        off, w = oracle.laplacian_stencil(3)
        a = rng.random((16, 16, 16))
        assert np.array_equal(C.stencil_apply(a, off, w), r.filter_run([16] * 3, off, w, init=a)["field"])
        # non power-of-two extents: the reference's (int %= size_t) wrap is not periodic (H2)
        b = rng.random((12, 12, 12))
        gq = r.filter_run([12] * 3, off, w, init=b)["field"]
        assert np.array_equal(C.stencil_apply(b, off, w, ref_wrap_quirk=True), gq)
        assert not np.array_equal(C.stencil_apply(b, off, w), gq)

    def test_reference_cli_output(self, tmp_path):
        import subprocess
        r = oracle.ref()
        p = subprocess.run([r.upwind_exe(), "-numCells", "32", "-numSteps", "10", "-std"],
                           capture_output=True, text=True, env={"OMP_NUM_THREADS": "2"})
        assert "number of cells:  32 32 32" in p.stdout
        assert "number of time steps: 10" in p.stdout
        assert "check sum: 1\n" in p.stdout
        assert "std      : 0.000951381" in p.stdout
