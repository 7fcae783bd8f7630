"""ctypes binding of libfidib200.so -- exactly the entry points of include/fidib200.h.

There is no fallback of any kind: if the CUDA library is not built, importing
this module raises, and if it is built but no GPU is usable every create call
returns FDB_E_CUDA, surfaced as FdbError.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libfidib200.so")

FDB_OK = 0
FDB_E_INVALID, FDB_E_CUDA, FDB_E_NCCL, FDB_E_OOM, FDB_E_DECOMP, FDB_E_STATE = -1, -2, -3, -4, -5, -6
FDB_ROW_MAJOR, FDB_COL_MAJOR = 0, 1
FDB_INPUT, FDB_OUTPUT = 0, 1
FDB_KERNEL_AUTO, FDB_KERNEL_GENERIC, FDB_KERNEL_TMA = 0, 1, 2
FDB_COMM_ID_BYTES = 128

_ERR_NAMES = {-1: "FDB_E_INVALID", -2: "FDB_E_CUDA", -3: "FDB_E_NCCL", -4: "FDB_E_OOM",
              -5: "FDB_E_DECOMP", -6: "FDB_E_STATE"}


class FdbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m fidibench_b200.build` "
            "(nvcc, sm_100a). fidibench_b200 has no CPU or PyTorch fallback.")
    # torch (when the process uses it) must own libnccl.so.2: import it first so the
    # dynamic loader resolves our NCCL dependency to the copy torch already mapped.
    try:
        import torch  # noqa: F401
    except Exception:  # pragma: no cover - torch is plumbing, not a requirement
        pass
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
p_i64, p_i32, p_dbl, p_vp = C.POINTER(i64), C.POINTER(i32), C.POINTER(dbl), C.POINTER(vp)

_SIGS = {
    "fdb_last_error": (C.c_char_p, []),
    "fdb_version": (i32, [p_i32, p_i32]),
    "fdb_device_count": (i32, [p_i32]),
    "fdb_launch_count": (i32, [p_i64]),
    "fdb_slab_partition": (i32, [i64, i32, i32, p_i64, p_i64]),
    "fdb_cube_decomp": (i32, [i32, i32, p_i64, p_i64]),
    "fdb_cube_block": (i32, [i32, i32, p_i64, i32, p_i64, p_i64]),
    "fdb_cube_neighbor": (i32, [i32, i32, p_i64, i32, p_i32, p_i32]),
    "fdb_comm_unique_id": (i32, [vp]),
    "fdb_comm_create": (i32, [i32, i32, vp, i32, p_vp]),
    "fdb_comm_rank": (i32, [vp, p_i32, p_i32]),
    "fdb_comm_barrier": (i32, [vp]),
    "fdb_comm_max": (i32, [vp, p_dbl]),
    "fdb_comm_destroy": (i32, [vp]),
    "fdb_upwind_create": (i32, [i32, p_i64, p_dbl, p_dbl, i32, p_vp]),
    "fdb_upwind_create_dist": (i32, [i32, p_i64, p_dbl, p_dbl, vp, p_vp]),
    "fdb_upwind_local_range": (i32, [vp, p_i64, p_i64]),
    "fdb_upwind_set_field": (i32, [vp, vp]),
    "fdb_upwind_set_slab": (i32, [vp, vp]),
    "fdb_upwind_set_slab_async": (i32, [vp, vp]),
    "fdb_upwind_reset": (i32, [vp]),
    "fdb_upwind_fill_random": (i32, [vp, C.c_uint64]),
    "fdb_upwind_advect": (i32, [vp, i64, dbl]),
    "fdb_upwind_advect_async": (i32, [vp, i64, dbl]),
    "fdb_upwind_sync": (i32, [vp]),
    "fdb_upwind_default_dt": (i32, [vp, p_dbl]),
    "fdb_upwind_checksum": (i32, [vp, p_dbl]),
    "fdb_upwind_std": (i32, [vp, p_dbl]),
    "fdb_upwind_plane_sums": (i32, [vp, vp, i64, p_i64]),
    "fdb_upwind_get_field": (i32, [vp, vp]),
    "fdb_upwind_get_slab": (i32, [vp, vp]),
    "fdb_upwind_set_kernel": (i32, [vp, i32]),
    "fdb_upwind_get_kernel": (i32, [vp, p_i32]),
    "fdb_upwind_describe": (i32, [vp, C.c_char_p, C.c_size_t]),
    "fdb_upwind_set_fuse": (i32, [vp, i32]),
    "fdb_upwind_set_stream": (i32, [vp, vp]),
    "fdb_upwind_last_timing": (i32, [vp, p_dbl, p_dbl, p_dbl]),
    "fdb_upwind_destroy": (i32, [vp]),
    "fdb_stencil_create": (i32, [i32, p_i64, i32, p_i32, p_dbl, i32, p_vp]),
    "fdb_stencil_create_dist": (i32, [i32, p_i64, i32, p_i32, p_dbl, vp, p_vp]),
    "fdb_stencil_local_range": (i32, [vp, p_i64, p_i64]),
    "fdb_stencil_set_input": (i32, [vp, vp, i32]),
    "fdb_stencil_set_input_slab": (i32, [vp, vp]),
    "fdb_stencil_set_input_separable": (i32, [vp, p_vp]),
    "fdb_stencil_fill_random": (i32, [vp, C.c_uint64]),
    "fdb_stencil_apply": (i32, [vp]),
    "fdb_stencil_swap": (i32, [vp]),
    "fdb_stencil_iterate": (i32, [vp, i64]),
    "fdb_stencil_checksum": (i32, [vp, i32, p_dbl]),
    "fdb_stencil_sumsq": (i32, [vp, i32, p_dbl]),
    "fdb_stencil_get": (i32, [vp, i32, vp, i32]),
    "fdb_stencil_get_slab": (i32, [vp, i32, vp]),
    "fdb_stencil_set_ref_wrap": (i32, [vp, i32]),
    "fdb_stencil_set_kernel": (i32, [vp, i32]),
    "fdb_stencil_get_kernel": (i32, [vp, p_i32]),
    "fdb_stencil_set_fuse": (i32, [vp, i32]),
    "fdb_stencil_get_fuse": (i32, [vp, p_i32]),
    "fdb_stencil_describe": (i32, [vp, C.c_char_p, C.c_size_t]),
    "fdb_stencil_set_stream": (i32, [vp, vp]),
    "fdb_stencil_last_timing": (i32, [vp, p_dbl, p_dbl, p_dbl]),
    "fdb_stencil_destroy": (i32, [vp]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int) -> None:
    if rc != FDB_OK:
        raise FdbError(rc, lib.fdb_last_error().decode("utf-8", "replace"))


def arr_i64(seq):
    return (i64 * len(seq))(*[int(x) for x in seq])


def arr_dbl(seq):
    return (dbl * len(seq))(*[float(x) for x in seq])


def device_count() -> int:
    n = i32(0)
    rc = lib.fdb_device_count(C.byref(n))
    return int(n.value) if rc == FDB_OK else 0


def launch_count() -> int:
    n = i64(0)
    check(lib.fdb_launch_count(C.byref(n)))
    return int(n.value)
