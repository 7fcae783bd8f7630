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
        if self._h:
            lib.fdb_comm_destroy(self._h)
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
