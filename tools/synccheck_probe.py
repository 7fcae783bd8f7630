"""Does compute-sanitizer synccheck's "Missing init" depend on how many mbarriers a CTA initialises?
Single-apply 7-point kernel (consumers wait directly on the TMA barrier; 2 x STAGES mbarriers) at 3 / 5 / 8 stages:
    FDB_LAP_CFG=<n> compute-sanitizer --tool synccheck python tools/synccheck_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402

st = {(0, 0, 0): -6.0}
for a in range(3):
    for s in (1, -1):
        o = [0, 0, 0]; o[a] = s; st[tuple(o)] = 1.0
x = np.random.default_rng(1).random((6, 32, 128))
with fb.Filter(x.shape, [0.0] * 3, [1.0] * 3, st) as fl:
    fl.set_fuse(1)
    fl.set_input(x)
    fl.iterate(3)
    print("probe ok", fl.describe(), float(fl.get().sum()))
