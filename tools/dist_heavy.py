"""Full-size multi-GPU runs under torchrun (one rank per GPU), parity by size-independent rules:

  torchrun ... tools/dist_heavy.py upwind N S     upwind N^3 x S steps from the delta: the (S+1)^3 corner must equal
                                                  the reference's 128^3 x 100 golden, zeros elsewhere
  torchrun ... tools/dist_heavy.py lap N ITER     laplacian N^3, ITER x (apply; swap): throughput + partition-invariant checksum
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

mode, N, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
comm = fb.Comm.from_torch_distributed(device=local)

if mode == "upwind":
    g = np.load(os.path.join(ROOT, "tests", "golden", "upwind_128_s100.npz"))
    assert S == 100 and N >= 128
    up = fb.Upwind([1.0] * 3, [1.0] * 3, [N] * 3, comm=comm)
    dt = up.default_dt()
    up.advect(S, dt)           # warm-up (also the run that is checked: reset + repeat below)
    up.reset()
    t0 = time.perf_counter()
    up.advect(S, dt)
    wall = time.perf_counter() - t0
    ms = up.last_timing()["gpu_ms"]
    cs = up.checksum()
    slab = up.slab()
    lo, hi = up.lo, up.hi
    corner = g["corner"]
    nz = int(np.count_nonzero(slab))
    ok = True
    if lo < 101:
        n_in = min(hi, 101) - lo
        ok &= bool(np.array_equal(slab[:n_in, :101, :101], corner[lo:lo + n_in]))
        expect_nz = int(np.count_nonzero(corner[lo:lo + n_in]))
    else:
        expect_nz = 0
    ok &= (nz == expect_nz)
    oks = [None] * world
    dist.all_gather_object(oks, (ok, nz))
    if rank == 0:
        total_nz = sum(o[1] for o in oks)
        print(f"upwind {N}^3 x {S} on {world} GPUs: bitwise corner/zeros {'OK' if all(o[0] for o in oks) else 'FAIL'}; "
              f"nnz={total_nz} (golden {int(g['nnz'])}); checksum={cs!r} (golden {float(g['checksum'])!r}); "
              f"gpu_ms={ms:.2f} GCUPS={N ** 3 * S / ms / 1e6:.1f} wall_ms={wall * 1e3:.1f}", flush=True)
    assert all(o[0] for o in oks)
    up.close()
else:
    off, w = oracle.laplacian_stencil(3)
    st = {tuple(int(v) for v in o): float(c) for o, c in zip(off, w)}
    fl = fb.Filter([N] * 3, [0.0] * 3, [1.0] * 3, st, comm=comm)
    lo, hi = fl.lo, fl.hi
    # the driver's input function on this rank's planes only (ref: laplacian.cxx:22-28, Filter.cpp:103-112)
    x1 = np.sin(2.0 * np.pi * ((np.arange(N) + 0.5) * (1.0 / float(N))))
    slab = (x1[lo:hi, None, None] * x1[None, :, None]) * x1[None, None, :]
    from fidibench_b200._lib import lib, check
    import ctypes as C
    check(lib.fdb_stencil_set_input_slab(fl._h, np.ascontiguousarray(slab).ctypes.data_as(C.c_void_p)))
    fl.iterate(2)
    best = 1e30
    for _ in range(3):
        fl.iterate(S)
        best = min(best, fl.last_timing()["gpu_ms"] / S)
    cs = fl.computeCheckSum("output")
    if rank == 0:
        print(f"laplacian {N}^3 on {world} GPUs kernel={fl.kernel()} ms/apply={best:.4f} GCUPS={N ** 3 / best / 1e6:.1f} "
              f"checksum(out)={cs!r}", flush=True)
    fl.close()
dist.barrier()
comm.close()
dist.destroy_process_group()
