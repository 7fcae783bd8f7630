"""Worker run under `python -m torch.distributed.run` by tests/test_dist_*.py.

    dist_worker.py cpu   -- gloo, no GPU: id plumbing + a host model of the slab/halo protocol
    dist_worker.py gpu   -- nccl, one rank per GPU: the real fdb_*_create_dist engines vs the oracle
Each rank prints one line "RANK r OK ..." on success; any assertion kills the job.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fidibench_b200 as fb  # noqa: E402
import oracle  # noqa: E402

SEED = 20261017


def host_model_of_the_slab_ring(rank, world):
    """The protocol libfidib200 runs on the device (runtime.cu: field_sweep/field_exchange),
    restated with numpy slabs and gloo send/recv: each rank owns planes [lo,hi) plus one ghost
    plane below; per step it updates its top plane first, ships it to rank+1's ghost for the
    NEXT step, then updates the rest.  Must equal the single-domain oracle bit for bit."""
    n0, n1, n2, steps = 8 * world, 6, 10, 2 * 8 * world + 3  # long enough to wrap the ring
    rng = np.random.default_rng(SEED)
    full = rng.random((n0, n1, n2))
    lo, hi = fb.slab_partition(n0, world, rank)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    cur = full[lo:hi].copy()
    ghost = full[(lo - 1) % n0].copy()
    c = -0.1  # ((dt*v)*up)/dx for the default run, any power-of-two resolution

    def update(plane, below):
        t = plane - c * (below - plane)
        t = t - c * (np.roll(plane, 1, axis=0) - plane)
        t = t - c * (np.roll(plane, 1, axis=1) - plane)
        return t

    for _ in range(steps):
        new = np.empty_like(cur)
        new[-1] = update(cur[-1], cur[-2])                     # boundary plane first
        send = dist.isend(torch.from_numpy(new[-1].copy()), nxt)  # its halo starts travelling
        buf = torch.empty((n1, n2), dtype=torch.float64)
        recv = dist.irecv(buf, prv)
        new[0] = update(cur[0], ghost)                         # interior overlaps the exchange
        for i in range(1, cur.shape[0] - 1):
            new[i] = update(cur[i], cur[i - 1])
        send.wait(); recv.wait()
        cur, ghost = new, buf.numpy().copy()
    # single-domain result, same operation order as upwind.cxx:64-84 (axes 0, 1, 2)
    ref = full.copy()
    for _ in range(steps):
        old = ref.copy()
        for j in range(3):
            ref = ref - c * (np.roll(old, 1, axis=j) - old)
    assert np.array_equal(cur, ref[lo:hi]), "host slab-ring model differs from the single-domain result"


def host_model_of_the_reversed_ring(rank, world):
    """Negative velocity along the slab axis (capi.cu: flip[0] on several slabs, runtime.cu: Field::ring_reversed): every
    rank holds ITS OWN global planes mirrored in place, runs the positive-velocity update on them, and the ring runs
    backwards -- the top device plane (the lowest global plane) goes to rank-1's ghost below.  With a ghost depth of 2
    and sweeps of two fused steps, restated with numpy slabs and gloo send/recv; must equal the single-domain result
    for v0 < 0 bit for bit."""
    n0, n1, n2 = 6 * world, 5, 8
    T, steps = 2, 4 * world + 2
    rng = np.random.default_rng(SEED + 9)
    full = rng.random((n0, n1, n2))
    lo, hi = fb.slab_partition(n0, world, rank)
    nxt, prv = (rank - 1) % world, (rank + 1) % world       # reversed: my top planes go to rank - 1
    c = -0.1                                                 # ((dt * -v) * -1) / dx == ((dt * v) * +1) / dx: same bits
    dev = full[lo:hi][::-1].copy()                           # mirrored in place: device plane i = global plane hi-1-i
    ghost = np.stack([full[(hi + T - 1 - g) % n0] for g in range(T)])   # device planes -T..-1 = global hi+T-1 .. hi

    def update(plane, below):
        t = plane - c * (below - plane)
        t = t - c * (np.roll(plane, 1, axis=0) - plane)
        t = t - c * (np.roll(plane, 1, axis=1) - plane)
        return t

    for _ in range(steps // T):
        ext = np.concatenate([ghost, dev])
        for _ in range(T):                                   # each fused step eats one plane at the bottom
            ext = np.stack([update(ext[i], ext[i - 1]) for i in range(1, ext.shape[0])])
        new = ext
        assert new.shape[0] == dev.shape[0]
        send = dist.isend(torch.from_numpy(new[-T:].copy()), nxt)
        buf = torch.empty((T, n1, n2), dtype=torch.float64)
        recv = dist.irecv(buf, prv)
        send.wait(); recv.wait()
        dev, ghost = new, buf.numpy().copy()
    # single-domain result for v0 < 0: the upwind neighbour along axis 0 is i + 1 (ref: upwind.cxx:34-35,75-76)
    ref = full.copy()
    for _ in range(steps):
        old = ref.copy()
        ref = ref - c * (np.roll(old, -1, axis=0) - old)        # coeff = ((dt * -1) * +1) / dx: the same -0.1
        ref = ref - c * (np.roll(old, 1, axis=1) - old)
        ref = ref - c * (np.roll(old, 1, axis=2) - old)
    assert np.array_equal(dev[::-1], ref[lo:hi]), "reversed slab ring differs from the single-domain result"


def host_model_of_the_two_sided_ring_with_fused_pairs(rank, world):
    """The stencil engine's multi-slab protocol (runtime.cu: field_run_sweeps / sweep_device_direct, capi.cu:
    fdb_stencil_iterate) restated with numpy slabs and gloo send/recv: G = 2 ghost planes on both sides, a plan of
    sweeps that do two applies each (reading two ghost planes, refreshing two) and an odd apply at the end (one
    plane), the depth of the last exchange tracked, and a full-depth refresh when the next plan starts deeper than
    the ghosts it finds.  Must equal the single-domain 7-point applies bit for bit."""
    G = 2
    n0, n1, n2 = 6 * world, 5, 8
    rng = np.random.default_rng(SEED + 7)
    full = rng.random((n0, n1, n2))
    w = [1.0, 1.0, 1.0, -6.0, 1.0, 1.0, 1.0]        # std::map order of the offsets (Filter.cpp:202)
    lo, hi = fb.slab_partition(n0, world, rank)
    nloc = hi - lo
    nxt, prv = (rank + 1) % world, (rank - 1) % world

    def apply(ext):
        """one apply on every plane of `ext` that has both neighbours: returns ext[1:-1] applied"""
        c, below, above = ext[1:-1], ext[:-2], ext[2:]
        acc = 0.0 + w[0] * below
        acc = acc + w[1] * np.roll(c, 1, axis=1)
        acc = acc + w[2] * np.roll(c, 1, axis=2)
        acc = acc + w[3] * c
        acc = acc + w[4] * np.roll(c, -1, axis=2)
        acc = acc + w[5] * np.roll(c, -1, axis=1)
        acc = acc + w[6] * above
        return acc

    def exchange(body, depth):
        """push `depth` boundary planes to both neighbours; returns (ghost_lo, ghost_hi), nearest planes valid"""
        reqs = [dist.isend(torch.from_numpy(body[-depth:].copy()), nxt),   # my top planes -> next's ghosts below
                dist.isend(torch.from_numpy(body[:depth].copy()), prv)]    # my bottom planes -> prev's ghosts above
        glo_t = torch.empty((depth, n1, n2), dtype=torch.float64)
        ghi_t = torch.empty((depth, n1, n2), dtype=torch.float64)
        reqs += [dist.irecv(glo_t, prv), dist.irecv(ghi_t, nxt)]
        for r in reqs:
            r.wait()
        glo = np.full((G, n1, n2), np.nan); ghi = np.full((G, n1, n2), np.nan)   # planes beyond `depth` are stale
        glo[G - depth:] = glo_t.numpy(); ghi[:depth] = ghi_t.numpy()
        return glo, ghi

    body = full[lo:hi].copy()
    glo, ghi = exchange(body, G)        # publish after the upload: full depth
    ghost_depth = G
    applied = 0
    for plan in ([2, 2, 1], [2, 2], [1], [2, 1]):
        assert all(a >= b for a, b in zip(plan, plan[1:])), "a plan never deepens"
        if plan[0] > ghost_depth:       # the previous plan ended on a one-plane exchange
            glo, ghi = exchange(body, G)
            ghost_depth = G
        for depth in plan:
            ext = np.concatenate([glo[G - depth:], body, ghi[:depth]])
            assert not np.isnan(ext).any(), "sweep read a stale ghost plane"
            for _ in range(depth):
                ext = apply(ext)        # each apply eats one plane on both ends
            assert ext.shape[0] == nloc
            body = ext
            glo, ghi = exchange(body, depth)
            ghost_depth = depth
            applied += depth
    ref = full.copy()
    for _ in range(applied):
        ref = apply(np.concatenate([ref[-1:], ref, ref[:1]]))
    assert np.array_equal(body, ref[lo:hi]), "two-sided slab ring differs from the single-domain applies"


def cpu_main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. the NCCL id travels over whatever torch.distributed backend is up
    ids = [fb.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    gathered = [None] * world
    dist.all_gather_object(gathered, ids[0])
    assert all(g == gathered[0] and len(g) == 128 for g in gathered)
    # 2. without a GPU the communicator must refuse loudly, never fall back
    if fb.device_count() == 0:
        try:
            fb.Comm(rank, world, ids[0], 0)
            raise AssertionError("Comm() succeeded without a CUDA device")
        except fb.FdbError as e:
            assert e.code == -2
    # 3. slabs tile the axis exactly
    ranges = [None] * world
    dist.all_gather_object(ranges, fb.slab_partition(16 * world, world, rank))
    assert ranges[0][0] == 0 and ranges[-1][1] == 16 * world
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    # 4. the halo protocol itself
    host_model_of_the_slab_ring(rank, world)
    host_model_of_the_reversed_ring(rank, world)
    host_model_of_the_two_sided_ring_with_fused_pairs(rank, world)
    dist.barrier()
    print(f"RANK {rank} OK cpu", flush=True)
    dist.destroy_process_group()


def gpu_main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = fb.Comm.from_torch_distributed(device=local)
    rng = np.random.default_rng(SEED)

    # upwind, both kernels, halo ring exercised by a random field and many steps
    a = rng.random((8 * world, 24, 64))
    for kernel in (fb.FDB_KERNEL_GENERIC, fb.FDB_KERNEL_TMA):
        up = fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, comm=comm)
        up.set_kernel(kernel)
        up.set_field(a)
        lo, hi = up.lo, up.hi
        up.advect(8 * world + 5, up.default_dt())
        ref = oracle.c.upwind_advect(a, 8 * world + 5)
        assert np.array_equal(up.slab(), ref[lo:hi]), f"rank {rank}: slab differs (kernel {kernel})"
        cs = up.checksum()
        with fb.Upwind([1.0] * 3, [1.0] * 3, a.shape) as single:
            single.set_field(a)
            single.advect(8 * world + 5, single.default_dt())
            assert cs == single.checksum(), "checksum is not partition-invariant"
        sd = up.std()
        assert abs(sd - oracle.c.std(ref)) <= 1e-12 * sd
        up.close()

    # negative velocity along the slab axis: ghost plane above
    up = fb.Upwind([-1.0, 1.0, 1.0], [1.0] * 3, a.shape, comm=comm)
    assert up.kernel() == fb.FDB_KERNEL_TMA, up.describe()   # mirrored in place, the ring runs backwards
    up.set_field(a)
    up.advect(7, 0.1 / a.shape[1])
    assert np.array_equal(up.slab(), oracle.c.upwind_advect(a, 7, velocity=[-1, 1, 1], dt=0.1 / a.shape[1])[up.lo:up.hi])
    with fb.Upwind([-1.0, 1.0, 1.0], [1.0] * 3, a.shape) as single:
        single.set_field(a)
        single.advect(7, 0.1 / a.shape[1])
        assert single.checksum() == up.checksum(), "checksum of a mirrored field is not partition-invariant"
    up.close()

    # delta on the last plane of slab 0 (SURVEY.md T1 ii)
    d = np.zeros((4 * world, 16, 32)); d[3, 15, 31] = 1.0
    up = fb.Upwind([1.0] * 3, [1.0] * 3, d.shape, comm=comm)
    up.set_field(d)
    up.advect(9, up.default_dt())
    assert np.array_equal(up.slab(), oracle.c.upwind_advect(d, 9)[up.lo:up.hi])
    up.close()

    # Laplacian (two-sided halo)
    off, w = oracle.laplacian_stencil(3)
    b = rng.random((4 * world, 12, 32))
    fl = fb.Filter(b.shape, [0.0] * 3, [1.0] * 3, {tuple(int(x) for x in o): float(v) for o, v in zip(off, w)}, comm=comm)
    fl.set_input(b)
    fl.iterate(5)
    ref = b
    for _ in range(5):
        ref = oracle.c.stencil_apply(ref, off, w)
    out = fl.get()
    assert np.array_equal(out[fl.lo:fl.hi], ref[fl.lo:fl.hi]), f"rank {rank}: laplacian slab differs"
    assert abs(fl.computeCheckSum("output") - oracle.c.checksum(ref)) < 1e-9
    fl.close()

    # two applies per sweep across ranks (ghost depth 2 on both sides), odd count, then more calls
    b2 = rng.random((4 * world, 16, 128))
    fl = fb.Filter(b2.shape, [0.0] * 3, [1.0] * 3, {tuple(int(x) for x in o): float(v) for o, v in zip(off, w)}, comm=comm)
    assert fl.fuse() == 2
    fl.set_input(b2)
    fl.iterate(5)
    fl.iterate(2)
    ref = b2
    for _ in range(7):
        ref = oracle.c.stencil_apply(ref, off, w)
    out = fl.get()
    assert np.array_equal(out[fl.lo:fl.hi], ref[fl.lo:fl.hi]), f"rank {rank}: fused laplacian slab differs"
    fl.close()

    # consecutive advect calls: a call ending on a remainder sweep leaves shallow ghosts behind
    up = fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, comm=comm)
    up.set_field(a)
    for n in (4, 3, 7):
        up.advect(n, up.default_dt())
    assert np.array_equal(up.slab(), oracle.c.upwind_advect(a, 14)[up.lo:up.hi]), f"rank {rank}: repeated advect differs"
    up.close()

    # the communicator refuses to go away under a live handle (its teardown is a collective over it)
    up = fb.Upwind([1.0] * 3, [1.0] * 3, a.shape, comm=comm)
    try:
        comm.close()
        raise AssertionError("Comm.close() succeeded with a live engine handle")
    except fb.FdbError as e:
        assert e.code == -6
    up.close()

    dist.barrier()
    print(f"RANK {rank} OK gpu", flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    cpu_main() if sys.argv[1] == "cpu" else gpu_main()
