"""The drop-in boundary without a GPU: the library builds, loads, exports every
symbol include/fidib200.h declares, and fails loudly (never falls back) when no
device is usable.  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fidib200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("fdb_upwind_create", "fdb_upwind_advect", "fdb_upwind_checksum", "fdb_upwind_std",
                 "fdb_upwind_get_field", "fdb_stencil_create", "fdb_stencil_apply", "fdb_stencil_swap",
                 "fdb_stencil_checksum", "fdb_comm_create", "fdb_slab_partition", "fdb_last_error"):
        assert must in syms
    assert len(syms) >= 40


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in fidib200.h but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol(fb):
    from fidibench_b200 import _lib
    assert sorted(_lib._SIGS) == declared_symbols()


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "fidib200.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def test_sass_has_tma_and_no_fma(lib_built):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass, "the TMA kernel must issue cp.async.bulk.tensor"
    assert "DFMA" not in sass, "an FMA in a stencil body breaks bit parity with the reference"


def test_slab_partition_matches_block_decomposition(fb):
    # ref: CubeDecomp::getBegIndices/getEndIndices with a (P,1,1) process grid
    assert fb.slab_partition(1024, 8, 0) == (0, 128)
    assert fb.slab_partition(1024, 8, 7) == (896, 1024)
    assert fb.slab_partition(128, 1, 0) == (0, 128)
    with pytest.raises(fb.FdbError) as e:
        fb.slab_partition(128, 3, 0)  # the reference finds no decomposition for P=3 either
    assert e.value.code == -5 and "No valid domain decomposition" in str(e.value)
    with pytest.raises(fb.FdbError):
        fb.slab_partition(128, 4, 4)


def test_fails_loudly_without_a_device(fb):
    if fb.device_count() > 0:
        pytest.skip("a GPU is visible; covered by the -m gpu tests")
    with pytest.raises(fb.FdbError) as e:
        fb.Upwind([1.0] * 3, [1.0] * 3, [8, 8, 8])
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(fb.FdbError):
        fb.Filter([8, 8], [0, 0], [1, 1], {(0, 0): -4.0, (1, 0): 1.0})


def test_argument_validation_happens_before_any_device_work(fb):
    with pytest.raises(fb.FdbError) as e:
        fb.Upwind([1.0] * 4, [1.0] * 4, [4, 4, 4, 4])
    assert e.value.code == -1
    with pytest.raises(fb.FdbError) as e:
        fb.Upwind([1.0] * 3, [1.0, 0.0, 1.0], [4, 4, 4])
    assert e.value.code == -1
    with pytest.raises(fb.FdbError) as e:
        fb.Filter([8, 8], [0, 0], [1, 1], {})
    assert e.value.code == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fidibench_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h", ".cpp", ".cxx", ".hpp")):
                src = open(os.path.join(base, fn)).read()
                if re.search(r"^\s*(import|from)\s+oracle|fdb_oracle|oracle/", src, flags=re.M):
                    offenders.append(os.path.join(base, fn))
    assert not offenders, f"product files reference the oracle: {offenders}"


def test_bench_touches_the_oracle_only_in_its_cpu_baseline_leg():
    """bench.py may import oracle only inside the reference arm / cpu_baseline functions; drivers never."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"reference_arm", "lap_cpu_baseline"}
    offenders = []

    def visit(node, fn):
        for child in ast.iter_child_nodes(node):
            name = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(child, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in child.names):
                if fn not in allowed:
                    offenders.append((fn, child.lineno))
            if isinstance(child, ast.ImportFrom) and (child.module or "").split(".")[0] == "oracle":
                if fn not in allowed:
                    offenders.append((fn, child.lineno))
            visit(child, name)

    visit(tree, "<module>")
    assert not offenders, offenders
    for fn in os.listdir(os.path.join(ROOT, "drivers")):
        if fn.endswith((".cxx", ".hpp", ".py")):
            assert "oracle" not in open(os.path.join(ROOT, "drivers", fn)).read(), fn
