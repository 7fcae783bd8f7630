"""Host-side mirror of the reference's `template<size_t NDIMS> class Upwind`
(ref: upwind/cxx/upwind.cxx:19-135) on top of the C ABI.

Same constructor arguments, same method names and meaning:
    Upwind(velocity, lengths, numCells); advect(numTimeSteps, deltaTime);
    checksum(); std(); saveVTK(filename); print()
plus what a test harness needs (set_field / field, timing).  The arithmetic runs
in libfidib200.so on the GPU; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, arr_i64, arr_dbl


class Upwind:
    def __init__(self, velocity, lengths, numCells, ngpus: int = 1, comm=None):
        self.ndims = len(numCells)
        if not (len(velocity) == len(lengths) == self.ndims):
            raise ValueError("velocity, lengths and numCells must have the same length")
        self.numCells = tuple(int(n) for n in numCells)
        self.v = tuple(float(x) for x in velocity)
        self.lengths = tuple(float(x) for x in lengths)
        self.deltas = tuple(l / n for l, n in zip(self.lengths, self.numCells))
        self.ntot = int(np.prod(self.numCells, dtype=np.int64))
        self.comm = comm
        self._h = C.c_void_p()
        if comm is None:
            check(lib.fdb_upwind_create(self.ndims, arr_i64(self.numCells), arr_dbl(self.v),
                                        arr_dbl(self.lengths), int(ngpus), C.byref(self._h)))
        else:
            check(lib.fdb_upwind_create_dist(self.ndims, arr_i64(self.numCells), arr_dbl(self.v),
                                             arr_dbl(self.lengths), comm._h, C.byref(self._h)))
        lo, hi = C.c_int64(), C.c_int64()
        check(lib.fdb_upwind_local_range(self._h, C.byref(lo), C.byref(hi)))
        self.lo, self.hi = int(lo.value), int(hi.value)

    # -- the reference's public surface -------------------------------------------------
    def advect(self, numTimeSteps: int, deltaTime: float) -> None:
        check(lib.fdb_upwind_advect(self._h, int(numTimeSteps), float(deltaTime)))

    def checksum(self) -> float:
        out = C.c_double()
        check(lib.fdb_upwind_checksum(self._h, C.byref(out)))
        return float(out.value)

    def std(self) -> float:
        out = C.c_double()
        check(lib.fdb_upwind_std(self._h, C.byref(out)))
        return float(out.value)

    def plane_sums(self) -> np.ndarray:
        """Per-plane sums along axis 0 (the partial sums behind checksum())."""
        n = C.c_int64()
        check(lib.fdb_upwind_plane_sums(self._h, None, 0, C.byref(n)))
        out = np.zeros(int(n.value), dtype=np.float64)
        check(lib.fdb_upwind_plane_sums(self._h, out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        return out

    def saveVTK(self, filename: str) -> None:
        """ASCII rectilinear-grid dump with the layout of upwind/cxx/saveVTK.h."""
        f = self.field().reshape(-1)
        nc, d = self.numCells, self.deltas
        n2 = nc[2] + 1 if self.ndims > 2 else 1
        n1 = nc[1] + 1 if self.ndims > 1 else 1
        with open(filename, "w") as fh:
            fh.write("# vtk DataFile Version 2.0\nupwind.cxx\nASCII\nDATASET RECTILINEAR_GRID\n")
            fh.write(f"DIMENSIONS {n2} {n1} {nc[0] + 1}\n")
            for name, axis in (("X", 2), ("Y", 1), ("Z", 0)):
                if self.ndims > axis or axis == 0:
                    a = axis if axis < self.ndims else 0
                    fh.write(f"{name}_COORDINATES {nc[a] + 1} double\n")
                    fh.write("".join(f" {0.0 + d[a] * i:g}" for i in range(nc[a] + 1)) + "\n")
                else:
                    fh.write(f"{name}_COORDINATES 1 double\n0.0\n")
            fh.write(f"CELL_DATA {self.ntot}\nSCALARS f double 1\nLOOKUP_TABLE default\n")
            for i in range(0, self.ntot, 10):
                fh.write(" ".join(f"{x:g}" for x in f[i:i + 10]) + " \n")
            fh.write("\n")

    def print(self) -> None:
        for i, x in enumerate(self.field().reshape(-1)):
            print(i, f"{x:g}")

    # -- harness ---------------------------------------------------------------------------
    def default_dt(self) -> float:
        """main()'s dt = min_j 0.1*dx_j/v_j (ref: upwind.cxx:186-192)."""
        out = C.c_double()
        check(lib.fdb_upwind_default_dt(self._h, C.byref(out)))
        return float(out.value)

    def set_field(self, field: np.ndarray) -> None:
        a = np.ascontiguousarray(field, dtype=np.float64)
        if a.size != self.ntot:
            raise ValueError(f"expected {self.ntot} cells, got {a.size}")
        check(lib.fdb_upwind_set_field(self._h, a.ctypes.data_as(C.c_void_p)))

    def set_field_ptr(self, host_ptr: int) -> None:
        """Whole-domain upload from a raw host address (e.g. pinned torch memory)."""
        check(lib.fdb_upwind_set_field(self._h, C.c_void_p(host_ptr)))

    def set_slab(self, slab: np.ndarray) -> None:
        a = np.ascontiguousarray(slab, dtype=np.float64)
        if a.size != self.slab_cells():
            raise ValueError(f"expected a slab of {self.slab_cells()} cells, got {a.size}")
        check(lib.fdb_upwind_set_slab(self._h, a.ctypes.data_as(C.c_void_p)))

    def set_slab_async(self, slab: np.ndarray) -> None:
        """Enqueue the upload only (keep `slab` alive and unchanged until the next sync)."""
        if slab.dtype != np.float64 or not slab.flags["C_CONTIGUOUS"] or slab.size != self.slab_cells():
            raise ValueError("set_slab_async needs a C-contiguous float64 array of slab_cells() elements")
        check(lib.fdb_upwind_set_slab_async(self._h, slab.ctypes.data_as(C.c_void_p)))

    def slab_cells(self) -> int:
        if self.ndims == 1:
            return self.ntot
        return (self.hi - self.lo) * (self.ntot // self.numCells[0])

    def reset(self) -> None:
        check(lib.fdb_upwind_reset(self._h))

    def fill_random(self, seed: int) -> None:
        """Seeded synthetic field generated on the device (see fdb_upwind_fill_random)."""
        check(lib.fdb_upwind_fill_random(self._h, C.c_uint64(seed)))

    def field(self) -> np.ndarray:
        """Whole-domain field (row-major); in dist mode only planes [lo,hi) are filled."""
        out = np.zeros(self.numCells, dtype=np.float64)
        check(lib.fdb_upwind_get_field(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def slab(self) -> np.ndarray:
        shape = self.numCells if self.ndims == 1 else (self.hi - self.lo,) + self.numCells[1:]
        out = np.zeros(shape, dtype=np.float64)
        check(lib.fdb_upwind_get_slab(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def slab_into(self, out: np.ndarray) -> None:
        """Copy this handle's planes into a caller-owned C-contiguous float64 array (e.g. pinned memory)."""
        if out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"] or out.size != self.slab_cells():
            raise ValueError("slab_into needs a C-contiguous float64 array of slab_cells() elements")
        check(lib.fdb_upwind_get_slab(self._h, out.ctypes.data_as(C.c_void_p)))

    def advect_async(self, numTimeSteps: int, deltaTime: float) -> None:
        check(lib.fdb_upwind_advect_async(self._h, int(numTimeSteps), float(deltaTime)))

    def sync(self) -> None:
        check(lib.fdb_upwind_sync(self._h))

    def describe(self) -> str:
        """Which kernel runs and, for the generic one, why the tiled kernels do not apply."""
        buf = C.create_string_buffer(512)
        check(lib.fdb_upwind_describe(self._h, buf, 512))
        return buf.value.decode()

    def set_kernel(self, kernel: int) -> None:
        check(lib.fdb_upwind_set_kernel(self._h, int(kernel)))

    def kernel(self) -> int:
        k = C.c_int()
        check(lib.fdb_upwind_get_kernel(self._h, C.byref(k)))
        return int(k.value)

    def set_fuse(self, steps_per_sweep: int) -> None:
        """Time steps advanced per sweep by the fused (temporal blocking) TMA kernel, 1..4."""
        check(lib.fdb_upwind_set_fuse(self._h, int(steps_per_sweep)))

    def set_stream(self, cuda_stream: int | None) -> None:
        check(lib.fdb_upwind_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def last_timing(self) -> dict:
        ms, upd, halo = C.c_double(), C.c_double(), C.c_double()
        check(lib.fdb_upwind_last_timing(self._h, C.byref(ms), C.byref(upd), C.byref(halo)))
        return dict(gpu_ms=ms.value, cell_updates=upd.value, halo_bytes=halo.value)

    def close(self) -> None:
        if self._h:
            lib.fdb_upwind_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
