"""Builds the C++ drivers (host code only; they link libfidib200.so through its C ABI).

    python -m drivers.build      ->  drivers/bin/{upwindCuda,laplacianCuda,upwindMpiCuda}
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "bin")
LIBDIR = os.path.join(ROOT, "fidibench_b200", "lib")
TARGETS = ["upwindCuda", "laplacianCuda", "upwindMpiCuda", "testStencil2dCuda"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _digest(paths) -> str:
    import hashlib
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode() + b"\0" + fh.read())
    return h.hexdigest()


def build(force: bool = False) -> list[str]:
    from fidibench_b200 import build as fbuild
    fbuild.build()
    os.makedirs(BIN, exist_ok=True)
    outs = []
    headers = [os.path.join(HERE, h) for h in ("cmdline.hpp", "Upwind.hpp", "Filter.hpp")] + \
              [os.path.join(ROOT, "include", "fidib200.h")]
    for t in TARGETS:
        src, out = os.path.join(HERE, t + ".cxx"), os.path.join(BIN, t)
        outs.append(out)
        stamp = out + ".srchash"
        digest = _digest([src] + headers)
        if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
            continue
        cmd = [CXX, "-std=c++11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HERE, src,
               "-o", out, "-L", LIBDIR, "-lfidib200", "-Wl,-rpath,$ORIGIN/../../fidibench_b200/lib",
               "-Wl,--allow-shlib-undefined"]
        subprocess.run(cmd, check=True)
        with open(stamp, "w") as fh:
            fh.write(digest + "\n")
    return outs


if __name__ == "__main__":
    print("\n".join(build(force="--force" in sys.argv)))
