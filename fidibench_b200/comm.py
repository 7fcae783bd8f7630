"""Communicator for one-process-per-GPU runs (replaces MPI_COMM_WORLD + CubeDecomp).

The data path is NCCL send/recv over NVLink inside libfidib200.so; this module
only ships the 128-byte NCCL id between ranks, through torch.distributed when it
is initialised (torch is plumbing here) or any `broadcast` callable.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from ._lib import lib, check


class Comm:
    def __init__(self, rank: int, nranks: int, id_bytes: bytes | None, device: int):
        self._h = C.c_void_p()
        buf = C.create_string_buffer(id_bytes, _lib.FDB_COMM_ID_BYTES) if id_bytes else None
        check(lib.fdb_comm_create(rank, nranks, buf, device, C.byref(self._h)))
        self.rank, self.nranks, self.device = rank, nranks, device

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.FDB_COMM_ID_BYTES)
        check(lib.fdb_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int | None = None) -> "Comm":
        """Build from an initialised torch.distributed process group (any backend)."""
        import torch
        import torch.distributed as dist
        rank, nranks = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", rank))
        ids = [cls.unique_id() if (rank == 0 and nranks > 1) else None]
        if nranks > 1:
            dist.broadcast_object_list(ids, src=0)
        return cls(rank, nranks, ids[0], device)

    def barrier(self) -> None:
        check(lib.fdb_comm_barrier(self._h))

    def max(self, value: float) -> float:
        v = C.c_double(value)
        check(lib.fdb_comm_max(self._h, C.byref(v)))
        return float(v.value)

    def close(self) -> None:
        """Destroy the communicator; engine handles created on it must have been closed first (FDB_E_STATE)."""
        if self._h:
            check(lib.fdb_comm_destroy(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def slab_partition(n0: int, nparts: int, part: int) -> tuple[int, int]:
    """Planes [lo, hi) of axis 0 owned by `part` (ref: CubeDecomp::getBegIndices/getEndIndices)."""
    lo, hi = C.c_int64(), C.c_int64()
    check(lib.fdb_slab_partition(n0, nparts, part, C.byref(lo), C.byref(hi)))
    return int(lo.value), int(hi.value)


class CubeDecomp:
    """Host-side mirror of the reference's `class CubeDecomp` (ref: cxx/CubeDecomp.h, CubeDecomp.cpp:11-131):
    the process grid the reference would choose for `nprocs` ranks on `dims`, its blocks and neighbours.
    The CUDA engines partition in slabs (slab_partition); this is for parity with MPI runs of the reference."""

    def __init__(self):
        self.nprocs, self.dims, self.decomp = 0, (), ()

    def build(self, nprocs: int, dims) -> bool:
        self.nprocs, self.dims = int(nprocs), tuple(int(d) for d in dims)
        grid = (C.c_int64 * len(self.dims))()
        rc = lib.fdb_cube_decomp(self.nprocs, len(self.dims), _lib.arr_i64(self.dims), grid)
        if rc == _lib.FDB_E_DECOMP:
            self.decomp = ()
            return False
        check(rc)
        self.decomp = tuple(int(g) for g in grid)
        return True

    def getDecomp(self):
        return self.decomp

    def _block(self, rk: int):
        nd = len(self.dims)
        lo, hi = (C.c_int64 * nd)(), (C.c_int64 * nd)()
        check(lib.fdb_cube_block(self.nprocs, nd, _lib.arr_i64(self.dims), int(rk), lo, hi))
        return tuple(int(x) for x in lo), tuple(int(x) for x in hi)

    def getBegIndices(self, rk: int):
        return self._block(rk)[0]

    def getEndIndices(self, rk: int):
        return self._block(rk)[1]

    def getNeighborRank(self, rk: int, direction) -> int:
        nd = len(self.dims)
        d = (C.c_int * nd)(*[int(x) for x in direction])
        out = C.c_int()
        check(lib.fdb_cube_neighbor(self.nprocs, nd, _lib.arr_i64(self.dims), int(rk), d, C.byref(out)))
        return int(out.value)
