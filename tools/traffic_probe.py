"""One short run of a sweep kernel on a single-slab field of a given shape, for
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:<kernel> -s <n> -c 1 python tools/traffic_probe.py upwind 128 1024 1024
(the per-launch DRAM traffic of a slab shape bench.py runs on several GPUs is the same on one GPU: same kernel, same grid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fidibench_b200 as fb  # noqa: E402

kind = sys.argv[1]
dims = [int(x) for x in sys.argv[2:5]]
if kind == "upwind":
    with fb.Upwind([1.0] * 3, [1.0] * 3, dims) as up:
        up.fill_random(1)
        up.advect(12, up.default_dt())
        print("ok", up.describe())
else:
    st = {(0, 0, 0): -6.0}
    for a in range(3):
        for s in (1, -1):
            o = [0, 0, 0]; o[a] = s; st[tuple(o)] = 1.0
    with fb.Filter(dims, [0.0] * 3, [1.0] * 3, st) as fl:
        fl.fill_random(1)
        fl.iterate(8)
        print("ok", fl.describe())
