#!/bin/bash
# round 2, call 29 (1 GPU): ragged tiles for the single-apply 7-point kernel -- parity tests, then the rate on 1000^3 / 500^3 / 8000^2-like planes
out=gpurun_out; mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_stencil_gpu.py tests/test_random_gpu.py tests/test_upwind_gpu.py -m gpu -q --timeout 120 > $out/r02cc_tests.log 2>&1; echo "rc=$?"; tail -6 $out/r02cc_tests.log
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import fidibench_b200 as fb
st = {(0, 0, 0): -6.0}
for a in range(3):
    for s in (1, -1):
        o = [0, 0, 0]; o[a] = s; st[tuple(o)] = 1.0
for dims in ((1000, 1000, 1000), (500, 500, 500), (250, 1000, 1000)):
    for env in ("0", "1"):
        import os
        os.environ["FDB_LAP_NO_RAGGED"] = env
        with fb.Filter(dims, [0.0] * 3, [1.0] * 3, st) as fl:
            fl.fill_random(1)
            fl.iterate(2)
            best = 1e9
            for _ in range(3):
                fl.iterate(4); best = min(best, fl.last_timing()["gpu_ms"] / 4)
            print(dims, "FDB_LAP_NO_RAGGED=" + env, fl.describe()[:40], "GCUPS=%.1f" % (np.prod(dims) / best / 1e6), flush=True)
PY
