"""fidibench_b200 -- B200-native finite-difference engines behind the reference's
own class interfaces (pletzer/fidibench: Upwind, Filter).

The product is `lib/libfidib200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/fidib200.h`); this package is the thin host mirror used by tests and
bench.py.  There is no CPU fallback: touching any engine name fails if the library
is not built.  Names resolve lazily so that `fidibench_b200.build` can (re)build the
library without loading it first.
"""
import importlib

_EXPORTS = {
    "Upwind": "upwind", "Filter": "filter", "Comm": "comm", "slab_partition": "comm", "CubeDecomp": "comm",
    "FdbError": "_lib", "device_count": "_lib", "launch_count": "_lib", "LIB_PATH": "_lib",
    "FDB_ROW_MAJOR": "_lib", "FDB_COL_MAJOR": "_lib", "FDB_INPUT": "_lib", "FDB_OUTPUT": "_lib",
    "FDB_KERNEL_AUTO": "_lib", "FDB_KERNEL_GENERIC": "_lib", "FDB_KERNEL_TMA": "_lib",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        mod = importlib.import_module(f".{_EXPORTS[name]}", __name__)
        value = getattr(mod, name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
