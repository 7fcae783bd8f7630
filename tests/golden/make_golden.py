"""Regenerates tests/golden/*.npz from the UNTOUCHED reference compiled into
oracle/_ref (see oracle/Makefile).  Run in the authoring container, where
/root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures are what pins oracle/fdb_oracle.c (and through it the CUDA path)
to the reference on machines where the reference cannot be rebuilt.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

SEED = 20261017


def bench_parity(r):
    # bench.py's untimed parity block (every N in 1, 2, 4, 8): the device-generated hash field (oracle.hash_field
    # restates it) on (16 N) x 48 x 256 advected 11 steps by the untouched upwind.cxx, and on (16 N) x 32 x 256
    # through 5 x (applyFilter; copyOutToIn) of the untouched Filter.cpp with the 7-point Laplacian; stored as one
    # SHA-256 per plane of the reference's output (the fields themselves would be 25 MB)
    import hashlib
    import json
    par = {"seed": SEED, "upwind_steps": 11, "stencil_iters": 5, "upwind": {}, "stencil": {}}
    off3, w3 = oracle.laplacian_stencil(3)
    for n in (1, 2, 4, 8):
        shape = (16 * n, 48, 256)
        a = oracle.hash_field(SEED, shape)
        g = r.upwind_run(shape, par["upwind_steps"], init=a)
        par["upwind"][str(n)] = {"dims": list(shape), "dt": g["dt"], "checksum": g["checksum"],
                                 "planes": [hashlib.sha256(p.tobytes()).hexdigest() for p in g["field"]]}
        shape = (16 * n, 32, 256)
        a = oracle.hash_field(SEED, shape)
        g = r.filter_run(shape, off3, w3, init=a, niter=par["stencil_iters"])
        par["stencil"][str(n)] = {"dims": list(shape),
                                  "planes": [hashlib.sha256(p.tobytes()).hexdigest() for p in g["field"]]}
    with open(os.path.join(HERE, "bench_parity.json"), "w") as fh:
        json.dump(par, fh)



def main():
    r = oracle.ref()
    rng = np.random.default_rng(SEED)

    # config 1: upwindCxx -numCells 128 -numSteps 10 (delta at cell 0); the field
    # is zero outside the (S+1)^3 corner, which is stored whole
    g = r.upwind_run([128] * 3, 10)
    f = g["field"]
    assert np.count_nonzero(f[11:, :, :]) == 0 and np.count_nonzero(f[:, 11:, :]) == 0 \
        and np.count_nonzero(f[:, :, 11:]) == 0
    np.savez_compressed(os.path.join(HERE, "upwind_128_s10.npz"), corner=f[:11, :11, :11].copy(),
                        checksum=g["checksum"], std=g["std"], nnz=np.count_nonzero(f), dt=g["dt"])

    # same run at 32^3 and 256^3: identical corner, different std
    for n in (32, 256):
        g2 = r.upwind_run([n] * 3, 10)
        assert np.array_equal(g2["field"][:11, :11, :11], f[:11, :11, :11])
    g100 = r.upwind_run([128] * 3, 100)
    np.savez_compressed(os.path.join(HERE, "upwind_128_s100.npz"),
                        corner=g100["field"][:101, :101, :101].astype(np.float64),
                        checksum=g100["checksum"], std=g100["std"], nnz=np.count_nonzero(g100["field"]))

    # halo/wrap stress: random field, anisotropic box, mixed-sign velocities
    cases = {}
    a = rng.random((24, 20, 28))
    for name, vel, lens, steps in (("pos", [1, 1, 1], [1, 1, 1], 7),
                                   ("mixed", [1, -2, 0.5], [1, 2, 3], 5),
                                   ("neg", [-1, -1, -1], [1, 1, 1], 4)):
        dt = oracle.c.upwind_dt(a.shape, [abs(v) for v in vel], lens)
        g = r.upwind_run(a.shape, steps, init=a, velocity=vel, lengths=lens, dt=dt)
        cases[f"{name}_out"] = g["field"]
        cases[f"{name}_vel"] = np.array(vel, dtype=np.float64)
        cases[f"{name}_len"] = np.array(lens, dtype=np.float64)
        cases[f"{name}_steps"] = steps
        cases[f"{name}_dt"] = dt
        cases[f"{name}_checksum"] = g["checksum"]
        cases[f"{name}_std"] = g["std"]
    np.savez_compressed(os.path.join(HERE, "upwind_random_24x20x28.npz"), init=a, **cases)

    # wrap several times: 16^3 x 100 steps from the delta
    g = r.upwind_run([16] * 3, 100)
    np.savez_compressed(os.path.join(HERE, "upwind_16_s100.npz"), out=g["field"], checksum=g["checksum"],
                        std=g["std"])

    # 1-D and 2-D instantiations of the class template
    a1, a2 = rng.random((37,)), rng.random((9, 20))
    np.savez_compressed(os.path.join(HERE, "upwind_1d2d.npz"), init1=a1, out1=r.upwind_run(a1.shape, 5, init=a1)["field"],
                        init2=a2, out2=r.upwind_run(a2.shape, 5, init=a2)["field"])

    # Laplacian driver: input function, 1 apply and 10 x (apply; copyOutToIn)
    off, w = oracle.laplacian_stencil(3)
    g1 = r.filter_run([16] * 3, off, w, init=None, niter=1, want_input=True)
    g10 = r.filter_run([16] * 3, off, w, init=None, niter=10)
    g32 = r.filter_run([32] * 3, off, w, init=None, niter=1)
    np.savez_compressed(os.path.join(HERE, "laplacian_16.npz"), input=g1["input"], out1=g1["field"],
                        out10=g10["field"], sums1=np.array([g1["in_sum"], g1["out_sum"]]),
                        sums10=np.array([g10["in_sum"], g10["out_sum"]]),
                        max32=np.abs(g32["field"]).max(), probe32=g32["field"][8, 8, 8])
    off2, w2 = oracle.laplacian_stencil(2)
    h1 = r.filter_run([32, 32], off2, w2, init=None, niter=1, want_input=True)
    # 24 is not a power of two: there the reference's (int %= size_t) wrap is not periodic
    hq = r.filter_run([24, 24], off2, w2, init=None, niter=1, want_input=True)
    np.savez_compressed(os.path.join(HERE, "laplacian2d_32.npz"), input=h1["input"], out1=h1["field"],
                        input24=hq["input"], out24_quirk=hq["field"])

    # upwindMpi.cxx's stencil through Filter, random field, 3 steps
    offu, wu = oracle.upwind_filter_stencil(16)
    a = rng.random((16, 16, 16))
    gu = r.filter_run([16] * 3, offu, wu, init=a, niter=3)
    np.savez_compressed(os.path.join(HERE, "upwindmpi_16.npz"), init=a, out3=gu["field"], offsets=offu, weights=wu)

    # testStencil2d.cxx's stencil (laplacian/cxx/testStencil2d.cxx:63-75)
    offt = np.array([[0, 0], [1, 0], [0, -1]], dtype=np.int32)
    wt = np.array([0.0, 1.0, -1.0])
    a = rng.random((8, 8))
    gt = r.filter_run([8, 8], offt, wt, init=a, niter=1)
    np.savez_compressed(os.path.join(HERE, "stencil2d_8.npz"), init=a, out=gt["field"], offsets=offt, weights=wt)

    bench_parity(r)

    # CubeDecomp's choices, for the record (we replace it with slabs)
    dec = {f"p{p}": np.array(r.cubedecomp(p, [128] * 3) or (0, 0, 0)) for p in (1, 2, 3, 4, 8, 16)}
    np.savez_compressed(os.path.join(HERE, "cubedecomp_128.npz"), **dec)

    # the whole CubeDecomp surface on a spread of grids: chosen process grid, every rank's block and its
    # neighbours in a fixed set of directions (fdb_cube_* must reproduce all of it, quirks included)
    cases = []
    for dims in ([128] * 3, [96, 64, 128], [12, 18], [8000, 8000], [30], [1, 8, 8], [64, 1, 64], [7, 5, 3], [1]):
        for p in (1, 2, 3, 4, 6, 8, 12, 16, 32, 64):
            cases.append((p, dims))
    rec = []
    for p, dims in cases:
        nd = len(dims)
        dirs = [[(1 if a == j else 0) * s for a in range(nd)] for j in range(nd) for s in (1, -1)] + [[1] * nd, [-1] * nd]
        dec = r.cubedecomp(p, dims)
        row = dict(nprocs=p, dims=list(dims), decomp=list(dec) if dec else None, ranks=[])
        if dec:
            for rk in sorted({0, 1 % p, p // 2, p - 1}):
                lo, hi, nb = r.cubedecomp_rank(p, dims, rk, dirs)
                row["ranks"].append(dict(rank=rk, lo=list(lo), hi=list(hi), dirs=dirs, nbr=list(nb)))
        rec.append(row)
    import json
    with open(os.path.join(HERE, "cubedecomp_cases.json"), "w") as fh:
        json.dump(rec, fh)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    if sys.argv[1:] == ["bench_parity"]:   # only tests/golden/bench_parity.json
        bench_parity(oracle.ref())
    else:
        main()
