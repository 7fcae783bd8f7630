"""CPU oracle for the FiDiBench finite-difference hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker.  The product (``fidibench_b200``) never does.

Three layers, all bit-identical to one another on the cases in
``tests/test_oracle.py``:

* ``ref``    -- the UNTOUCHED reference sources compiled into ``oracle/_ref/``
                (``oracle/Makefile``; only available where it was built or
                shipped, i.e. not re-buildable on the GPU box),
* ``c``      -- ``oracle/fdb_oracle.c``, our plain-C restatement,
* ``np_*``   -- numpy restatements (``np.roll``), for readability and as a
                third opinion.

Parity status: PINNED (``c`` == ``ref`` bit-for-bit, and both == the committed
``tests/golden`` fixtures that were generated from ``ref``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = C.POINTER(C.c_double)
_c_i64_p = C.POINTER(C.c_int64)
_c_int_p = C.POINTER(C.c_int)


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _i64(seq):
    return (C.c_int64 * len(seq))(*[int(x) for x in seq])


def _f64(seq):
    return (C.c_double * len(seq))(*[float(x) for x in seq])


def build(with_ref: bool = True) -> None:
    """Compile the C restatement and (where /root/reference exists) oracle/_ref."""
    targets = ["oracle"] + (["ref"] if with_ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


# --------------------------------------------------------------------------- #
# our C restatement
# --------------------------------------------------------------------------- #
class _COracle:
    def __init__(self):
        path = os.path.join(_HERE, "libfdb_oracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        L = C.CDLL(path)
        L.fdb_oracle_upwind_dt.restype = C.c_double
        L.fdb_oracle_upwind_dt.argtypes = [C.c_int, _c_i64_p, _c_double_p, _c_double_p]
        L.fdb_oracle_upwind_advect.restype = None
        L.fdb_oracle_upwind_advect.argtypes = [C.c_int, _c_i64_p, _c_double_p, _c_double_p,
                                               _c_double_p, _c_double_p, C.c_int64, C.c_double]
        L.fdb_oracle_checksum.restype = C.c_double
        L.fdb_oracle_checksum.argtypes = [_c_double_p, C.c_int64]
        L.fdb_oracle_std.restype = C.c_double
        L.fdb_oracle_std.argtypes = [_c_double_p, C.c_int64]
        L.fdb_oracle_stencil_apply.restype = None
        L.fdb_oracle_stencil_apply.argtypes = [C.c_int, _c_i64_p, C.c_int, _c_int_p, _c_double_p,
                                               _c_double_p, _c_double_p, C.c_int]
        L.fdb_oracle_sort_branches.restype = None
        L.fdb_oracle_sort_branches.argtypes = [C.c_int, C.c_int, _c_int_p, _c_double_p]
        L.fdb_oracle_laplacian_input.restype = None
        L.fdb_oracle_laplacian_input.argtypes = [C.c_int, _c_i64_p, _c_double_p, _c_double_p, _c_double_p]
        L.fdb_oracle_num_threads.restype = C.c_int
        self.L = L

    def num_threads(self) -> int:
        return int(self.L.fdb_oracle_num_threads())

    def upwind_dt(self, num_cells, velocity, lengths) -> float:
        nd = len(num_cells)
        return float(self.L.fdb_oracle_upwind_dt(nd, _i64(num_cells), _f64(velocity), _f64(lengths)))

    def upwind_advect(self, field, num_steps, velocity=None, lengths=None, dt=None):
        """field: ndarray of shape numCells (row-major). Returns the advected copy."""
        f = np.ascontiguousarray(field, dtype=np.float64).copy()
        nd = f.ndim
        velocity = [1.0] * nd if velocity is None else list(velocity)
        lengths = [1.0] * nd if lengths is None else list(lengths)
        if dt is None:
            dt = self.upwind_dt(f.shape, velocity, lengths)
        scratch = np.empty_like(f)
        self.L.fdb_oracle_upwind_advect(nd, _i64(f.shape), _f64(velocity), _f64(lengths),
                                        _dp(f), _dp(scratch), int(num_steps), float(dt))
        return f

    def checksum(self, field) -> float:
        f = np.ascontiguousarray(field, dtype=np.float64)
        return float(self.L.fdb_oracle_checksum(_dp(f), f.size))

    def std(self, field) -> float:
        f = np.ascontiguousarray(field, dtype=np.float64)
        return float(self.L.fdb_oracle_std(_dp(f), f.size))

    def sort_branches(self, offsets, weights):
        off = np.ascontiguousarray(offsets, dtype=np.int32).copy()
        w = np.ascontiguousarray(weights, dtype=np.float64).copy()
        nb, nd = off.shape
        self.L.fdb_oracle_sort_branches(nd, nb, off.ctypes.data_as(_c_int_p), _dp(w))
        return off, w

    def stencil_apply(self, field, offsets, weights, ref_wrap_quirk=False, presorted=False):
        """One Filter::applyFilter on a single periodic domain (row-major)."""
        f = np.ascontiguousarray(field, dtype=np.float64)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if not presorted:
            off, w = self.sort_branches(off, w)
        out = np.empty_like(f)
        self.L.fdb_oracle_stencil_apply(f.ndim, _i64(f.shape), off.shape[0],
                                        off.ctypes.data_as(_c_int_p), _dp(w), _dp(f), _dp(out),
                                        1 if ref_wrap_quirk else 0)
        return out

    def laplacian_input(self, dims, xmins=None, xmaxs=None):
        nd = len(dims)
        xmins = [0.0] * nd if xmins is None else xmins
        xmaxs = [1.0] * nd if xmaxs is None else xmaxs
        out = np.empty(tuple(int(d) for d in dims), dtype=np.float64)
        self.L.fdb_oracle_laplacian_input(nd, _i64(dims), _f64(xmins), _f64(xmaxs), _dp(out))
        return out


# --------------------------------------------------------------------------- #
# the untouched reference, compiled into oracle/_ref
# --------------------------------------------------------------------------- #
def ref_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", n))
               for n in ("libref_upwind.so", "libref_filter.so"))


class _Ref:
    def __init__(self):
        d = os.path.join(_HERE, "_ref")
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        U = C.CDLL(os.path.join(d, "libref_upwind.so"))
        U.fdb_ref_upwind_run.restype = C.c_double
        U.fdb_ref_upwind_run.argtypes = [C.c_int, C.POINTER(C.c_longlong), _c_double_p, _c_double_p,
                                         _c_double_p, C.c_int, C.c_double, _c_double_p, _c_double_p, _c_double_p]
        U.fdb_ref_upwind_threads.restype = C.c_int
        U.fdb_ref_upwind_set_threads.argtypes = [C.c_int]
        F = C.CDLL(os.path.join(d, "libref_filter.so"))
        F.fdb_ref_filter_run.restype = C.c_double
        F.fdb_ref_filter_run.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.c_int, _c_int_p, _c_double_p,
                                         _c_double_p, C.c_int, _c_double_p, _c_double_p, _c_double_p]
        F.fdb_ref_filter_branch_order.argtypes = [C.c_int, C.c_int, _c_int_p, _c_double_p]
        F.fdb_ref_cubedecomp.restype = C.c_int
        F.fdb_ref_cubedecomp.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        self.U, self.F, self.dir = U, F, d

    @staticmethod
    def _ll(seq):
        return (C.c_longlong * len(seq))(*[int(x) for x in seq])

    def threads(self) -> int:
        return int(self.U.fdb_ref_upwind_threads())

    def set_threads(self, n: int) -> None:
        self.U.fdb_ref_upwind_set_threads(int(n))

    def upwind_run(self, num_cells, num_steps, init=None, velocity=None, lengths=None, dt=None,
                   want_field=True):
        """Upwind<ndims>(...).advect(num_steps, dt) from upwind/cxx/upwind.cxx.
        Returns dict(field, checksum, std, seconds)."""
        nd = len(num_cells)
        velocity = [1.0] * nd if velocity is None else list(velocity)
        lengths = [1.0] * nd if lengths is None else list(lengths)
        if dt is None:
            dt = c.upwind_dt(num_cells, velocity, lengths)
        init_a = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
        out = np.empty(tuple(int(n) for n in num_cells), dtype=np.float64) if want_field else None
        cs, sd = C.c_double(), C.c_double()
        secs = self.U.fdb_ref_upwind_run(nd, self._ll(num_cells), _f64(velocity), _f64(lengths),
                                         _dp(init_a), int(num_steps), float(dt), _dp(out),
                                         C.byref(cs), C.byref(sd))
        if secs < 0:
            raise ValueError("bad ndims")
        return dict(field=out, checksum=cs.value, std=sd.value, seconds=float(secs), dt=dt)

    def filter_run(self, dims, offsets, weights, init=None, niter=1, want_field=True, want_input=False):
        """niter x {Filter::applyFilter; copyOutToIn} from cxx/Filter.cpp (single rank)."""
        nd = len(dims)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        init_a = None if init is None else np.ascontiguousarray(init, dtype=np.float64)
        shape = tuple(int(n) for n in dims)
        out = np.empty(shape, dtype=np.float64) if want_field else None
        inp = np.empty(shape, dtype=np.float64) if want_input else None
        sums = np.zeros(2, dtype=np.float64)
        secs = self.F.fdb_ref_filter_run(nd, self._ll(dims), off.shape[0], off.ctypes.data_as(_c_int_p),
                                         _dp(w), _dp(init_a), int(niter), _dp(out), _dp(inp), _dp(sums))
        if secs < 0:
            raise ValueError("reference found no valid decomposition")
        return dict(field=out, input=inp, in_sum=float(sums[0]), out_sum=float(sums[1]), seconds=float(secs))

    def branch_order(self, offsets, weights):
        off = np.ascontiguousarray(offsets, dtype=np.int32).copy()
        w = np.ascontiguousarray(weights, dtype=np.float64).copy()
        self.F.fdb_ref_filter_branch_order(off.shape[1], off.shape[0], off.ctypes.data_as(_c_int_p), _dp(w))
        return off, w

    def cubedecomp(self, nprocs, dims):
        dec = (C.c_longlong * len(dims))()
        ok = self.F.fdb_ref_cubedecomp(int(nprocs), len(dims), self._ll(dims), dec)
        return tuple(int(x) for x in dec) if ok else None

    def cubedecomp_rank(self, nprocs, dims, rank, dirs):
        """(lo, hi, neighbour ranks) of `rank` from CubeDecomp::getBegIndices/getEndIndices/getNeighborRank."""
        nd = len(dims)
        lo, hi = (C.c_longlong * nd)(), (C.c_longlong * nd)()
        d = np.ascontiguousarray(dirs, dtype=np.int32).reshape(-1, nd)
        nbr = np.zeros(d.shape[0], dtype=np.int32)
        fn = self.F.fdb_ref_cubedecomp_rank
        fn.restype = C.c_int
        ok = fn(int(nprocs), nd, self._ll(dims), int(rank), lo, hi, d.shape[0], d.ctypes.data_as(_c_int_p),
                nbr.ctypes.data_as(_c_int_p))
        if not ok:
            return None
        return tuple(int(x) for x in lo), tuple(int(x) for x in hi), tuple(int(x) for x in nbr)

    def upwind_exe(self) -> str:
        return os.path.join(self.dir, "upwindCxx")

    def run_main(self, which, argv, cwd):
        """The reference's own laplacian.cxx / upwindMpi.cxx main() (untouched, single rank against the MPI stub),
        run in a process of its own inside `cwd` (they write laplacian.vtk / upMpi.vtk there).
        which: "laplacian" | "upwindmpi".  Returns the CompletedProcess (stdout = the driver's output)."""
        import sys
        code = (
            "import ctypes, sys\n"
            f"L = ctypes.CDLL({os.path.join(self.dir, 'libref_filter.so')!r})\n"
            "argv = [b'ref'] + [a.encode() for a in sys.argv[1:]]\n"
            "arr = (ctypes.c_char_p * (len(argv) + 1))(*argv, None)\n"
            f"rc = L.fdb_ref_{which}_main(len(argv), arr)\n"
            "sys.stdout.flush(); sys.exit(rc)\n")
        return subprocess.run([sys.executable, "-c", code] + [str(a) for a in argv], cwd=cwd, capture_output=True,
                              text=True, timeout=600)


# --------------------------------------------------------------------------- #
# numpy restatements
# --------------------------------------------------------------------------- #
def np_upwind_advect(field, num_steps, velocity=None, lengths=None, dt=None):
    """upwind/cxx/upwind.cxx:51-86 with np.roll."""
    f = np.array(field, dtype=np.float64, copy=True)
    nd = f.ndim
    velocity = [1.0] * nd if velocity is None else list(velocity)
    lengths = [1.0] * nd if lengths is None else list(lengths)
    up = [(+1 if v < 0.0 else -1) for v in velocity]
    deltas = [np.float64(lengths[j]) / np.float64(f.shape[j]) for j in range(nd)]
    if dt is None:
        dt = min(np.float64(0.1) * deltas[j] / np.float64(velocity[j]) for j in range(nd))
    dt = np.float64(dt)
    for _ in range(int(num_steps)):
        old = f.copy()
        for j in range(nd):
            coeff = dt * np.float64(velocity[j]) * np.float64(up[j]) / deltas[j]
            f -= coeff * (np.roll(old, -up[j], axis=j) - old)
    return f


def np_stencil_apply(field, offsets, weights):
    """cxx/Filter.cpp:191-263 on one periodic domain, true periodic wrap.
    Branches are applied in std::map (lexicographic) order starting from 0."""
    f = np.asarray(field, dtype=np.float64)
    order = sorted(range(len(weights)), key=lambda b: tuple(int(x) for x in offsets[b]))
    out = np.zeros_like(f)
    for b in order:
        src = f
        for j, o in enumerate(offsets[b]):
            if o:
                src = np.roll(src, -int(o), axis=j)
        out += np.float64(weights[b]) * src
    return out


def hash_field(seed, shape):
    """numpy restatement of the device-side synthetic fill (fidibench_b200/csrc/kernels_generic.cu: hash_u01):
    cell g of the row-major field = (splitmix64(seed + (g+1)*0x9E3779B97F4A7C15) >> 11) * 2^-53."""
    n = int(np.prod(shape, dtype=np.int64))
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53).reshape(tuple(int(x) for x in shape))


def laplacian_stencil(ndims):
    """laplacian/cxx/laplacian.cxx:55-65 (insertion order; the map sorts it)."""
    offs, w = [[0] * ndims], [-2.0 * ndims]
    for i in range(ndims):
        for s in (1, -1):
            o = [0] * ndims
            o[i] = s
            offs.append(o)
            w.append(1.0)
    return np.array(offs, dtype=np.int32), np.array(w, dtype=np.float64)


def upwind_filter_stencil(num_cells, ndims=3):
    """upwind/cxx/upwindMpi.cxx:62-92: the upwind step as a Filter stencil."""
    v = [1.0] * ndims
    courant = 0.1
    deltas, signs, dt = [], [], np.finfo(np.float64).max
    for j in range(ndims):
        dx = np.float64(1.0) / np.float64(num_cells)
        deltas.append(dx)
        val = np.float64(courant) * dx / np.float64(v[j])
        dt = val if val < dt else dt
        signs.append(1 if v[j] > 0 else -1)
    diag = np.float64(1.0)
    for i in range(ndims):
        diag -= signs[i] * dt * v[i] / deltas[i]
    offs, w = [[0] * ndims], [float(diag)]
    for i in range(ndims):
        o = [0] * ndims
        o[i] = -signs[i]
        offs.append(o)
        w.append(float(signs[i] * dt * v[i] / deltas[i]))
    return np.array(offs, dtype=np.int32), np.array(w, dtype=np.float64)


c = _COracle()
_ref_singleton = None


def ref() -> _Ref:
    global _ref_singleton
    if _ref_singleton is None:
        _ref_singleton = _Ref()
    return _ref_singleton
