"""Host-side mirror of the reference's `class Filter` (ref: cxx/Filter.h:36-170)
on top of the C ABI: an offset->weight stencil applied to a periodic,
slab-decomposed field.  Same method names and meaning as the reference:

    Filter(globalDims, xmins, xmaxs, stencil)      stencil: {offset tuple: weight}
    setInData(f)            f(position list) -> value      (Filter.cpp:131-159)
    setInDataByIndices(f)   f(global index list) -> value  (Filter.cpp:161-188)
    applyFilter(); copyOutToIn(); computeCheckSum("input"|"output")
    getRank(); getNumProcs(); isDecompValid(); saveVTK(filename)

Callbacks are evaluated on the host (as the reference does) and uploaded; the
stencil itself runs in libfidib200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np

from . import _lib
from ._lib import lib, check, arr_i64, FdbError


class Filter:
    def __init__(self, globalDims, xmins, xmaxs, stencil: dict, ngpus: int = 1, comm=None):
        self.ndims = len(globalDims)
        self.globalDims = tuple(int(n) for n in globalDims)
        self.xmins = [float(x) for x in xmins]
        self.xmaxs = [float(x) for x in xmaxs]
        self.stencil = {tuple(int(o) for o in k): float(v) for k, v in stencil.items()}
        self.comm = comm
        self.ngpus = int(ngpus)
        self._h = C.c_void_p()
        self.validDecomp = False
        offs = np.array(list(self.stencil.keys()), dtype=np.int32).reshape(len(self.stencil), self.ndims)
        w = np.array(list(self.stencil.values()), dtype=np.float64)
        try:
            if comm is None:
                rc = lib.fdb_stencil_create(self.ndims, arr_i64(self.globalDims), len(w),
                                            offs.ctypes.data_as(_lib.p_i32), w.ctypes.data_as(_lib.p_dbl),
                                            int(ngpus), C.byref(self._h))
            else:
                rc = lib.fdb_stencil_create_dist(self.ndims, arr_i64(self.globalDims), len(w),
                                                 offs.ctypes.data_as(_lib.p_i32), w.ctypes.data_as(_lib.p_dbl),
                                                 comm._h, C.byref(self._h))
            check(rc)
        except FdbError as e:
            if e.code != _lib.FDB_E_DECOMP:
                raise
            # ref: Filter.cpp:27-34 -- report and leave an invalid object behind
            if self.getRank() == 0:
                sys.stderr.write("ERROR: No valid domain decomposition could be found. Adjust the number\n"
                                 "of processes and/or the domain dimensions.\n")
            return
        self.validDecomp = True
        self._nparts = comm.nranks if comm is not None else int(ngpus)
        lo, hi = C.c_int64(), C.c_int64()
        check(lib.fdb_stencil_local_range(self._h, C.byref(lo), C.byref(hi)))
        self.lo, self.hi = int(lo.value), int(hi.value)

    # -- the reference's public surface -------------------------------------------------
    def getRank(self) -> int:
        return self.comm.rank if self.comm is not None else 0

    def getNumProcs(self) -> int:
        """Slabs of the ring: the comm's ranks, or the devices of an in-process handle (as drivers/Filter.hpp)."""
        return self.comm.nranks if self.comm is not None else self.ngpus

    def isDecompValid(self) -> bool:
        return self.validDecomp

    def getPosition(self, globalInds):
        """ref: Filter.cpp:103-112 (cell centred)."""
        pos = []
        for i in range(self.ndims):
            delta = (self.xmaxs[i] - self.xmins[i]) / float(self.globalDims[i])
            pos.append(self.xmins[i] + (globalInds[i] + 0.5) * delta)
        return pos

    def setInData(self, f) -> None:
        self.setInDataByIndices(lambda inds: f(self.getPosition(inds)))

    def setInDataByIndices(self, f) -> None:
        a = np.empty(self.globalDims, dtype=np.float64)
        for idx in np.ndindex(*self.globalDims):
            a[idx] = f(list(idx))
        self.set_input(a)

    def applyFilter(self) -> None:
        check(lib.fdb_stencil_apply(self._h))

    def copyOutToIn(self) -> None:
        check(lib.fdb_stencil_swap(self._h))

    def computeCheckSum(self, inOrOut: str) -> float:
        which = _lib.FDB_INPUT if inOrOut == "input" else _lib.FDB_OUTPUT  # ref: Filter.cpp:470-482
        out = C.c_double()
        check(lib.fdb_stencil_checksum(self._h, which, C.byref(out)))
        return float(out.value)

    def sumsq(self, inOrOut: str = "output") -> float:
        which = _lib.FDB_INPUT if inOrOut == "input" else _lib.FDB_OUTPUT
        out = C.c_double()
        check(lib.fdb_stencil_sumsq(self._h, which, C.byref(out)))
        return float(out.value)

    # -- harness ---------------------------------------------------------------------------
    def set_input(self, field: np.ndarray, layout: int = _lib.FDB_ROW_MAJOR) -> None:
        a = np.ascontiguousarray(field, dtype=np.float64)
        if a.size != int(np.prod(self.globalDims, dtype=np.int64)):
            raise ValueError("field has the wrong size")
        check(lib.fdb_stencil_set_input(self._h, a.ctypes.data_as(C.c_void_p), layout))

    def set_input_slab(self, slab: np.ndarray) -> None:
        """Only this handle's planes [lo,hi) (one-process-per-GPU runs on grids too big for one host array)."""
        a = np.ascontiguousarray(slab, dtype=np.float64)
        if a.size != self.slab_cells():
            raise ValueError(f"expected a slab of {self.slab_cells()} cells, got {a.size}")
        check(lib.fdb_stencil_set_input_slab(self._h, a.ctypes.data_as(C.c_void_p)))

    def slab_cells(self) -> int:
        ntot = int(np.prod(self.globalDims, dtype=np.int64))
        return ntot if self.ndims == 1 else (self.hi - self.lo) * (ntot // self.globalDims[0])

    def set_input_separable(self, factors) -> None:
        """Input = product over the axes (reference order) of 1-D factors, evaluated on the device: what
        setInData(func) yields for laplacian.cxx's func when factors[j][i] = sin(2 pi x_j(i))."""
        arrs = [np.ascontiguousarray(f, dtype=np.float64) for f in factors]
        if len(arrs) != self.ndims or any(a.size != n for a, n in zip(arrs, self.globalDims)):
            raise ValueError("one factor array of globalDims[j] values per axis")
        ptrs = (C.c_void_p * self.ndims)(*[a.ctypes.data for a in arrs])
        check(lib.fdb_stencil_set_input_separable(self._h, ptrs))

    def laplacian_factors(self):
        """The 1-D factors of the laplacian driver's input function, evaluated on the host with libm's sin
        (ref: laplacian.cxx:22-28 on Filter::getPosition, Filter.cpp:103-112)."""
        import math
        out = []
        for j in range(self.ndims):
            delta = (self.xmaxs[j] - self.xmins[j]) / float(self.globalDims[j])
            out.append(np.array([math.sin(2.0 * math.pi * (self.xmins[j] + (i + 0.5) * delta))
                                 for i in range(self.globalDims[j])]))
        return out

    def fill_random(self, seed: int) -> None:
        check(lib.fdb_stencil_fill_random(self._h, C.c_uint64(seed)))

    def get_slab(self, which: int = _lib.FDB_OUTPUT) -> np.ndarray:
        shape = self.globalDims if self.ndims == 1 else (self.hi - self.lo,) + self.globalDims[1:]
        out = np.zeros(shape, dtype=np.float64)
        check(lib.fdb_stencil_get_slab(self._h, which, out.ctypes.data_as(C.c_void_p)))
        return out

    def saveVTK(self, filename: str) -> None:
        """ASCII structured-grid dump of the output data, layout of cxx/writeVTK.cpp:12-95."""
        nd = self.ndims
        if nd > 3:
            sys.stderr.write("WARNING: writeVTK does not support more than 3 dimensions\n")
            return
        field = self.get(_lib.FDB_OUTPUT).reshape(-1, order="C").reshape(self.globalDims).reshape(-1, order="F")
        cells = list(self.globalDims) + [1] * (3 - nd)
        lo = self.xmins + [0.0] * (3 - nd)
        hi = self.xmaxs + [1.0] * (3 - nd)
        nodes = [cells[a] + 1 if a < nd else 2 for a in range(3)]
        with open(filename, "w") as fh:
            fh.write("# vtk DataFile Version 2.0\nproduced by laplacian\nASCII\nDATASET STRUCTURED_GRID\n")
            fh.write(f"DIMENSIONS {nodes[0]} {nodes[1]} {nodes[2]}\nPOINTS {nodes[0] * nodes[1] * nodes[2]} float\n")
            for k in range(nodes[2]):
                for j in range(nodes[1]):
                    for i in range(nodes[0]):
                        x = lo[0] + (hi[0] - lo[0]) * i / float(cells[0])
                        y = lo[1] + (hi[1] - lo[1]) * j / float(cells[1])
                        z = lo[2] + (hi[2] - lo[2]) * k / float(cells[2])
                        fh.write(f"{x:g} {y:g} {z:g}\n")
            fh.write(f"CELL_DATA {cells[0] * cells[1] * cells[2]}\nSCALARS outData float\nLOOKUP_TABLE default\n")
            fh.write("".join(f"{v:g}\n" for v in field))

    def iterate(self, niter: int) -> None:
        """niter x { applyFilter(); copyOutToIn() } (ref: laplacian.cxx:86-90)."""
        check(lib.fdb_stencil_iterate(self._h, int(niter)))

    def get(self, which: int = _lib.FDB_OUTPUT, layout: int = _lib.FDB_ROW_MAJOR) -> np.ndarray:
        out = np.zeros(self.globalDims, dtype=np.float64)
        check(lib.fdb_stencil_get(self._h, which, out.ctypes.data_as(C.c_void_p), layout))
        return out

    def set_ref_wrap(self, on: bool) -> None:
        """Reproduce Filter.cpp:240's (int %= size_t) wrap (non-periodic unless the extent is a power of two)."""
        check(lib.fdb_stencil_set_ref_wrap(self._h, 1 if on else 0))

    def describe(self) -> str:
        """Which kernel runs and, for the generic one, why the tiled kernels do not apply."""
        buf = C.create_string_buffer(512)
        check(lib.fdb_stencil_describe(self._h, buf, 512))
        return buf.value.decode()

    def set_kernel(self, kernel: int) -> None:
        check(lib.fdb_stencil_set_kernel(self._h, int(kernel)))

    def kernel(self) -> int:
        k = C.c_int()
        check(lib.fdb_stencil_get_kernel(self._h, C.byref(k)))
        return int(k.value)

    def set_fuse(self, applies_per_sweep: int) -> None:
        """Applies fused per sweep by iterate(): 1, 2 (3-D 7-point stencil, temporal blocking) or 0 = auto."""
        check(lib.fdb_stencil_set_fuse(self._h, int(applies_per_sweep)))

    def fuse(self) -> int:
        n = C.c_int()
        check(lib.fdb_stencil_get_fuse(self._h, C.byref(n)))
        return int(n.value)

    def set_stream(self, cuda_stream: int | None) -> None:
        check(lib.fdb_stencil_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def last_timing(self) -> dict:
        ms, upd, halo = C.c_double(), C.c_double(), C.c_double()
        check(lib.fdb_stencil_last_timing(self._h, C.byref(ms), C.byref(upd), C.byref(halo)))
        return dict(gpu_ms=ms.value, cell_updates=upd.value, halo_bytes=halo.value)

    def close(self) -> None:
        if self._h:
            lib.fdb_stencil_destroy(self._h)
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
