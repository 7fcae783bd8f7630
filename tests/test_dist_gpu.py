"""One process per GPU over NCCL (the torchrun layout bench.py uses): the dist engines
against the single-domain oracle.  Needs >= 2 GPUs on the box."""
import os
import sys

import pytest

from conftest import ROOT
from test_dist_cpu import torchrun

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_nccl_slab_ring_matches_oracle(gpu_fb, nproc):
    if gpu_fb.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    p = torchrun(nproc, [os.path.join(ROOT, "tests", "dist_worker.py"), "gpu"], 29620 + nproc, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for r in range(nproc):
        assert f"RANK {r} OK gpu" in p.stdout
