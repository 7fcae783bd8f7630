"""CPU check of the fused upwind tile pipeline: the host model that mirrors upwind3d_fused_kernel
(tests/host_model_fused.py) must equal T oracle time steps bit for bit -- ragged tiles on both
in-plane axes, ragged plane chunks, single slab and slab of a ring."""
import numpy as np
import pytest

import oracle
from host_model_fused import Cfg, fused_steps

SEED = 20261017


@pytest.mark.parametrize("T,CJ,R,BK,shape,ci", [
    (3, 21, 3, 128, (5, 40, 260), 3),    # ragged j (19-row tiles), ragged k (260 = 2 x 128 + 4), ragged chunks
    (3, 18, 6, 128, (4, 16, 128), 8),    # one tile per axis: halo rows and wrap columns of the tile itself
    (2, 16, 2, 128, (3, 24, 136), 2),
    (2, 18, 6, 128, (4, 34, 256), 4),
    (4, 21, 3, 128, (5, 20, 132), 5),
    (4, 18, 6, 128, (4, 8, 16), 4),      # smallest supported plane
    (3, 18, 3, 64, (4, 20, 200), 4),     # 64-cell tiles
    (3, 24, 8, 128, (3, 48, 128), 3),
])
@pytest.mark.parametrize("split", [False, True])
def test_model_single_slab(T, CJ, R, BK, shape, ci, split):
    """split = the exchange-tile layout of the lean formulation (the default kernel for T >= 3), with the chunks walked
    top chunk first as the single-launch ring sweeps do."""
    rng = np.random.default_rng(SEED)
    x = rng.random(shape)
    out = np.full(shape, np.nan)
    # the driver's dt and the engine's coefficients ((dt*v)*up)/dx, ref: upwind.cxx:72,186-192
    dt = oracle.c.upwind_dt(shape, [1.0] * 3, [1.0] * 3)
    c = [((dt * 1.0) * -1) / (1.0 / shape[j]) for j in range(3)]
    fused_steps(x, c, 0, shape[0], T, Cfg(T, CJ, R, BK), ci, 0, shape[0], out, split=split, reverse=split)
    assert np.array_equal(out, oracle.c.upwind_advect(x, T))


def test_model_slab_of_a_ring():
    """Slab [4,8) of 12 planes, ghost depth 4 (the upwind engine's), launched as the runtime does: top
    planes first, then the rest; anisotropic coefficients."""
    rng = np.random.default_rng(SEED + 1)
    x = rng.random((12, 24, 64))
    v, lengths = [1.0, 0.5, 2.0], [1.0, 2.0, 0.5]
    dt = 0.01
    c = [((dt * v[j]) * -1) / (lengths[j] / x.shape[j]) for j in range(3)]
    ref = oracle.c.upwind_advect(x, 3, velocity=v, lengths=lengths, dt=dt)
    lo, hi = 4, 8
    out = np.full((hi - lo, 24, 64), np.nan)
    for ibeg, iend in ((1, 4), (0, 1)):
        fused_steps(x, c, lo, hi, 4, Cfg(3, 21, 3), 64, ibeg, iend, out)
    assert np.array_equal(out, ref[lo:hi])
