"""Host->device copy ceiling of one box: every rank copies a pinned 1 GiB buffer to its GPU,
first one rank at a time, then all ranks at once.  Run under torch.distributed.run with one rank per GPU.
The end-to-end arm of bench.py uploads 1 GiB per GPU per step, so `all at once` is its ceiling."""
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << 27  # doubles = 1 GiB
host = torch.empty(n, dtype=torch.float64, pin_memory=True)
host.fill_(1.0)
dev = torch.empty(n, dtype=torch.float64, device="cuda")
back = torch.empty(n, dtype=torch.float64, pin_memory=True)


def copy_gbs(reps=6, d2h=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if d2h:
            back.copy_(dev, non_blocking=True)
        else:
            dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n * 8 / (time.perf_counter() - t0) / 1e9


copy_gbs(2)
alone = torch.zeros(world, dtype=torch.float64, device="cuda")
for r in range(world):
    dist.barrier()
    if r == rank:
        alone[r] = copy_gbs()
dist.all_reduce(alone)
dist.barrier()
both = torch.zeros(world, dtype=torch.float64, device="cuda")
both[rank] = copy_gbs()
dist.all_reduce(both)
dist.barrier()
d2h = torch.zeros(world, dtype=torch.float64, device="cuda")
d2h[rank] = copy_gbs(d2h=True)
dist.all_reduce(d2h)
if rank == 0:
    f = lambda t: " ".join(f"{x:.1f}" for x in t.tolist())
    print(f"ranks={world} cpus={os.cpu_count()}")
    print(f"H2D GB/s, one rank at a time : {f(alone)}")
    print(f"H2D GB/s, all ranks at once  : {f(both)}   aggregate {both.sum().item():.1f}")
    print(f"D2H GB/s, all ranks at once  : {f(d2h)}   aggregate {d2h.sum().item():.1f}")
dist.destroy_process_group()
