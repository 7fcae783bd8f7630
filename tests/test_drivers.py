"""The C++ drivers: the reference's command lines and stdout contract (ref:
upwind/cxx/upwind.cxx:139-217, laplacian/cxx/laplacian.cxx:30-129, upwind/cxx/upwindMpi.cxx:30-166)."""
import os
import re
import subprocess

import pytest

import numpy as np

import oracle
from conftest import ROOT, SEED


@pytest.fixture(scope="module")
def drivers(lib_built):
    from drivers import build as dbuild
    return {os.path.basename(p): p for p in dbuild.build()}


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)


def test_bad_option_prints_help_and_exits_zero_like_the_reference(drivers):
    p = run(drivers["upwindCuda"], "-bogus", "3")
    assert p.returncode == 0  # ref: upwind.cxx:207-216 returns 0 after printing help
    assert "-bogus is not a valid option." in p.stdout
    assert "ERROR when parsing command line arguments" in p.stderr
    assert "Usage:" in p.stdout and "-numCells <int#> Number of cells along each axis (128)" in p.stdout
    assert "-numSteps <int#> Number of time steps (10)" in p.stdout
    assert "-std Print out spread of solution (0)" in p.stdout


def test_help_flag(drivers):
    p = run(drivers["laplacianCuda"], "-h")
    assert p.returncode == 0
    assert "Purpose: benchmark finite difference operations." in p.stdout
    assert "-numCells <int#> Number of cells along each axis (8000)" in p.stdout
    assert "-numDims <int#> Number of dimensions (2)" in p.stdout


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built on this machine")
def test_help_text_option_lines_match_the_reference_binary(drivers):
    ref = run(oracle.ref().upwind_exe(), "-h").stdout
    ours = run(drivers["upwindCuda"], "-h").stdout
    ref_opts = [l for l in ref.splitlines() if l.startswith("\t-")]
    for line in ref_opts:  # every reference option appears verbatim (ours adds more)
        assert line in ours.splitlines()


@pytest.mark.gpu
def test_upwind_driver_config1_output(drivers, gpu_fb):
    p = run(drivers["upwindCuda"], "-numCells", "128", "-numSteps", "10", "-std", "-timing")
    assert p.returncode == 0, p.stderr
    out = p.stdout
    assert "number of cells:  128 128 128\n" in out
    assert "number of time steps: 10\n" in out
    assert "check sum: 1\n" in out
    assert "std      : 0.000118983\n" in out
    assert re.search(r"[Cc]heck sum:[ ]*[1|0\.999]", out)  # the reference's ctest regex
    m = re.search(r"check sum \(17 digits\): (\S+)", out)
    assert abs(float(m.group(1)) - 1.0000000000000011) < 1e-12
    # positional tokens are ignored, exactly as the reference does (it then runs 128^3)
    p2 = run(drivers["upwindCuda"], "32", "10")
    assert "number of cells:  128 128 128" in p2.stdout


@pytest.mark.gpu
def test_upwind_driver_vtk_files(drivers, gpu_fb, tmp_path):
    p = subprocess.run([drivers["upwindCuda"], "-numCells", "8", "-numSteps", "3", "-vtk"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    up0, up1 = (tmp_path / "up0.vtk").read_text(), (tmp_path / "up1.vtk").read_text()
    assert up0.startswith("# vtk DataFile Version 2.0\nupwind.cxx\nASCII\nDATASET RECTILINEAR_GRID\nDIMENSIONS 9 9 9\n")
    assert "CELL_DATA 512\nSCALARS f double 1\nLOOKUP_TABLE default\n1 0 0 0 0 0 0 0 0 0 \n" in up0
    assert "0.343 0.147 0.021 0.001 0 0 0 0 " in up1  # 0.7^3, 3*0.7^2*0.1, 3*0.7*0.01, 0.001
    if oracle.ref_available():
        q = subprocess.run([oracle.ref().upwind_exe(), "-numCells", "8", "-numSteps", "3", "-vtk"],
                           cwd=tmp_path / "..", capture_output=True, text=True, env={"OMP_NUM_THREADS": "1"})
        ref1 = (tmp_path / ".." / "up1.vtk").read_text()
        assert ref1 == up1  # byte-identical VTK dump


@pytest.mark.gpu
def test_laplacian_and_upwindmpi_drivers(drivers, gpu_fb):
    p = run(drivers["laplacianCuda"], "-numDims", "3", "-numCells", "32")
    assert p.returncode == 0, p.stderr
    assert "Number of procs: 1\nglobal dimensions: 32 32 32 \nDomain decomp 1 1 1 \n" in p.stdout
    assert "Laplace times min/max/avg:" in p.stdout
    m = re.search(r"Check sums: input = (\S+) output = (\S+)", p.stdout)
    assert abs(float(m.group(1))) < 1e-9 and abs(float(m.group(2))) < 1e-9  # ~0 by symmetry
    p = run(drivers["laplacianCuda"], "-numCells", "64")  # default -numDims 2
    assert "global dimensions: 64 64 \n" in p.stdout and "Check sums:" in p.stdout
    p = run(drivers["upwindMpiCuda"], "-numCells", "32", "-numSteps", "4")
    assert p.returncode == 0, p.stderr
    for i in range(4):
        assert f"iter {i} check sum  in/out = 1 / 1\n" in p.stdout
    assert "Check sum: 1\n" in p.stdout


@pytest.mark.gpu
def test_invalid_decomposition_message(drivers, gpu_fb):
    if gpu_fb.device_count() < 2:
        pytest.skip("needs 2 GPUs to ask for a non-dividing slab count")
    p = run(drivers["laplacianCuda"], "-numDims", "3", "-numCells", "33", "-ngpus", "2")
    assert "No valid domain decomposition could be found" in p.stderr
    assert "Decomposition is invalid" in p.stderr


@pytest.mark.gpu
def test_stencil2d_driver_prints_the_reference_values(drivers, gpu_fb):
    """ref: laplacian/cxx/testStencil2d.cxx -- out(i,j) = in(i+1,j) - in(i,j-1), in = 1 where i*j == 0."""
    p = run(drivers["testStencil2dCuda"])
    assert p.returncode == 0, p.stderr
    vals = {}
    for m in re.finditer(r"inds = (\d+) (\d+)\s+outData = (\S+)", p.stderr):
        vals[(int(m.group(1)), int(m.group(2)))] = float(m.group(3))
    assert len(vals) == 64
    f = lambda i, j: 1.0 if (i % 8) * (j % 8) == 0 else 0.0
    for (i, j), v in vals.items():
        assert v == f(i + 1, j) - f(i, j - 1)


@pytest.mark.gpu
@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built on this machine")
@pytest.mark.parametrize("flags", [("-numDims", "3", "-numCells", "8"), ("-numDims", "2", "-numCells", "16"),
                                   ("-numDims", "2", "-numCells", "12", "-refwrap")])
def test_laplacian_driver_vtk_is_byte_identical_to_the_reference(drivers, gpu_fb, tmp_path, flags):
    """Filter::saveVTK (Filter.cpp:487-538 + cxx/writeVTK.cpp:12-95) through the untouched
    laplacian.cxx main() against drivers/Filter.hpp::saveVTK through laplacianCuda, same flags.  12 is not a power
    of two: there the reference's index wrap is not periodic (H2) and -refwrap reproduces it."""
    ours_dir, ref_dir = tmp_path / "ours", tmp_path / "ref"
    ours_dir.mkdir(); ref_dir.mkdir()
    ref_flags = [f for f in flags if f != "-refwrap"]
    q = oracle.ref().run_main("laplacian", list(ref_flags) + ["-vtk"], ref_dir)
    assert q.returncode == 0, q.stderr
    p = subprocess.run([drivers["laplacianCuda"], *flags, "-vtk", "-raw", "out.bin"], cwd=ours_dir, capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert (ours_dir / "laplacian.vtk").read_bytes() == (ref_dir / "laplacian.vtk").read_bytes()
    # the lines both drivers print
    for key in ("global dimensions:", "Check sums:", "Data will be written to file laplacian.vtk"):
        ours = [l for l in p.stdout.splitlines() if key in l]
        ref = [l for l in q.stdout.splitlines() if key in l]
        assert ours and len(ours) == len(ref)
    # the raw dump carries the same field at full precision
    nd, n = int(flags[1]), int(flags[3])
    raw = np.fromfile(ours_dir / "out.bin", dtype=np.float64).reshape((n,) * nd)
    off, w = oracle.laplacian_stencil(nd)
    r = oracle.c.laplacian_input((n,) * nd)
    for _ in range(10):
        r = oracle.c.stencil_apply(r, off, w, ref_wrap_quirk="-refwrap" in flags)
    assert np.array_equal(raw, r)


@pytest.mark.gpu
@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built on this machine")
def test_upwindmpi_driver_vtk_and_stdout_match_the_reference(drivers, gpu_fb, tmp_path):
    ours_dir, ref_dir = tmp_path / "ours", tmp_path / "ref"
    ours_dir.mkdir(); ref_dir.mkdir()
    flags = ["-numCells", "8", "-numSteps", "3", "-vtk"]
    q = oracle.ref().run_main("upwindmpi", flags, ref_dir)
    assert q.returncode == 0, q.stderr
    p = subprocess.run([drivers["upwindMpiCuda"], *flags, "-raw", "out.bin"], cwd=ours_dir, capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr
    assert (ours_dir / "upMpi.vtk").read_bytes() == (ref_dir / "upMpi.vtk").read_bytes()
    keep = lambda out: [l for l in out.splitlines() if l.startswith("iter ") or l.startswith("Check sum:")]
    assert keep(p.stdout) == keep(q.stdout) and len(keep(p.stdout)) == 4
    assert np.fromfile(ours_dir / "out.bin", dtype=np.float64).size == 512


@pytest.mark.gpu
def test_upwind_driver_raw_dump_and_anisotropic_flags(drivers, gpu_fb, tmp_path):
    """-raw round trip at full precision; -nx/-ny/-nz/-lx/-ly/-lz expose what the class supports and the reference's
    main() does not (upwind.cxx:174,182-183)."""
    p = subprocess.run([drivers["upwindCuda"], "-numCells", "16", "-numSteps", "7", "-raw", "f.bin"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    f0 = np.zeros((16, 16, 16)); f0.reshape(-1)[0] = 1.0
    assert np.array_equal(np.fromfile(tmp_path / "f.bin", dtype=np.float64).reshape(16, 16, 16), oracle.c.upwind_advect(f0, 7))
    p = subprocess.run([drivers["upwindCuda"], "-nx", "12", "-ny", "20", "-nz", "36", "-lx", "1.5", "-lz", "0.75", "-vy", "-2",
                        "-numSteps", "5", "-raw", "g.bin"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "number of cells:  12 20 36\n" in p.stdout
    shape, vel, lens = (12, 20, 36), [1.0, -2.0, 1.0], [1.5, 1.0, 0.75]
    g0 = np.zeros(shape); g0.reshape(-1)[0] = 1.0
    dt = oracle.c.upwind_dt(shape, [abs(v) for v in vel], lens)
    ref = oracle.c.upwind_advect(g0, 5, velocity=vel, lengths=lens, dt=dt)
    assert np.array_equal(np.fromfile(tmp_path / "g.bin", dtype=np.float64).reshape(shape), ref)
