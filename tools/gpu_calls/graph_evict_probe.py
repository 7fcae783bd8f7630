"""More than 16 distinct sweep plans on one small field: the graph cache (runtime.cu: field_run_sweeps) evicts its oldest
executable graphs while earlier replays may still be queued.  Field must equal the oracle bit for bit.  No torch import
(ctypes + numpy only) so the call costs seconds."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import fidibench_b200 as fb
import oracle

t0 = time.time()
rng = np.random.default_rng(7)
a = rng.random((16, 24, 64))
ref = a
with fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as up:
    up.set_field(a)
    dt = up.default_dt()
    n0 = fb.launch_count()
    total = 0
    for rnd in range(2):
        for steps in range(5, 45, 2):        # 20 plans, each at both parities over the two rounds
            up.advect_async(steps, dt) if hasattr(up, "advect_async") else up.advect(steps, dt)
            total += steps
    out = up.field()
    print("launches", fb.launch_count() - n0, "time steps", total)
ref = oracle.c.upwind_advect(a, total, dt=dt)
ok = np.array_equal(out, ref)
print("graph eviction probe:", "BITEXACT" if ok else "MISMATCH", "%.1f s" % (time.time() - t0))
sys.exit(0 if ok else 1)
