"""Build recipe of libfidib200.so (the C-ABI library of sm_100a kernels).

Explicit nvcc, in-tree output (fidibench_b200/lib/libfidib200.so) so that the
built library travels with the repository snapshot to the GPU box.  nvcc
cross-compiles for sm_100a without a GPU.

    python -m fidibench_b200.build            # build if stale
    python -m fidibench_b200.build --force
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libfidib200.so")
SOURCES = ["runtime.cu", "kernels_generic.cu", "kernels_tma.cu", "kernels_fused.cu", "kernels_lapfused.cu", "decomp.cu", "capi.cu"]
HEADERS = [os.path.join(CSRC, "fdb_internal.h"), os.path.join(CSRC, "tma_ptx.cuh"), os.path.join(ROOT, "include", "fidib200.h")]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# --fmad=false: the parity contract forbids contracting a*b+c;
# the kernels use __d*_rn intrinsics as well, this is the belt to those braces.
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "--fmad=false", "-Xcompiler", "-fPIC,-O2",
                     "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# experiments only (e.g. FDB_NVCC_EXTRA=-DFDB_DBG_NO_LANDED_WAIT for the synccheck probe of profiles/r02s_*)
NVCC_FLAGS += [f for f in os.environ.get("FDB_NVCC_EXTRA", "").split() if f]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def host_compiler_args() -> list[str]:
    # the image exports CC/CXX=/opt/gcc/bin/*, wrappers that nvcc should not pick up
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


STAMP = LIB + ".srchash"


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive snapshots)."""
    import hashlib
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).replace(ROOT, "").encode())
    return h.hexdigest()


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as fh:
        return fh.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    cc = nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [cc] + host_compiler_args() + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(objdir, "nvcc.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see fidibench_b200/build/nvcc.log")
    link = [cc] + host_compiler_args() + ARCH + ["-shared", "-o", LIB] + objs + ["-lnccl"]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
