"""world_size-2 (and one world_size-3) gloo runs on CPU: the host side of the N>1 path (id plumbing, slab
partition, the halo-ring protocol modelled on the host) and bench.py's reference arm."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def torchrun(nproc, script_args, port, timeout=600, env_extra=None):
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": "2", "MASTER_ADDR": "127.0.0.1"})
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_two_ranks_gloo_host_plumbing_and_halo_protocol(lib_built):
    p = torchrun(2, [os.path.join(ROOT, "tests", "dist_worker.py"), "cpu"], 29611)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "RANK 0 OK cpu" in p.stdout and "RANK 1 OK cpu" in p.stdout


def test_three_ranks_gloo_ring_direction(lib_built):
    """With two ranks the previous and the next slab are the same rank; three make the ring's direction matter (the
    one-sided ring, the reversed ring of mirrored slabs, the two-sided ring with fused pairs)."""
    p = torchrun(3, [os.path.join(ROOT, "tests", "dist_worker.py"), "cpu"], 29615)
    assert p.returncode == 0, p.stdout + p.stderr
    assert all(f"RANK {r} OK cpu" in p.stdout for r in range(3))


def test_bench_reference_arm_under_torchrun_only_rank0_reports():
    p = torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                     "--warmup", "0", "--workload", "upwind128", "--tsteps", "2"], 29613)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "GCUPS" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
