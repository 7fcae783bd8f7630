"""fidibench_b200 -- B200-native finite-difference engines behind the reference's
own class interfaces (pletzer/fidibench: Upwind, Filter).

The product is `lib/libfidib200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/fidib200.h`); this package is the thin host mirror used by tests and
bench.py.  There is no CPU fallback: importing fails if the library is not built.
"""
from ._lib import (FdbError, FDB_COL_MAJOR, FDB_INPUT, FDB_KERNEL_AUTO, FDB_KERNEL_GENERIC,
                   FDB_KERNEL_TMA, FDB_OUTPUT, FDB_ROW_MAJOR, device_count, launch_count, LIB_PATH)
from .comm import Comm, slab_partition
from .filter import Filter
from .upwind import Upwind

__all__ = ["Upwind", "Filter", "Comm", "slab_partition", "FdbError", "device_count", "launch_count",
           "FDB_ROW_MAJOR", "FDB_COL_MAJOR", "FDB_INPUT", "FDB_OUTPUT", "FDB_KERNEL_AUTO",
           "FDB_KERNEL_GENERIC", "FDB_KERNEL_TMA", "LIB_PATH"]
