#!/bin/bash
# round 2, call 15 (2 GPUs): mirrored axes on slab rings, describe(), the whole multi-GPU suite
out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --timeout 90 > $out/r02p_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -12 $out/r02p_tests_n2.log
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import fidibench_b200 as fb
for vel in ([1.0,1.0,1.0], [-1.0,1.0,1.0]):
    with fb.Upwind(vel, [2.0,1.0,1.0], (1024,512,512), ngpus=2) as up:
        up.fill_random(3)
        dt = up.default_dt()
        up.advect(30, dt)
        best = 1e9
        for _ in range(3):
            up.advect(99, dt); best = min(best, up.last_timing()["gpu_ms"]/99)
        print(vel, up.describe()[:60], "GCUPS=%.1f" % (1024*512*512/best/1e6))
PY
