"""CPU check of the single-apply 7-point tile pipeline, ragged tiles included: the host model that mirrors
lap7_tma_kernel<C, RAGGED> (tests/host_model_lap7.py) must equal one oracle apply bit for bit."""
import numpy as np
import pytest

import oracle
from host_model_lap7 import Cfg, ORDER, single_apply

SEED = 20261017


@pytest.mark.parametrize("shape,BJ,BK,R,ci", [
    ((4, 32, 256), 16, 128, 4, 3),     # exact tiling, ragged second chunk
    ((3, 10, 14), 8, 32, 4, 2),        # one ragged tile per axis: both wraps inside the tile
    ((3, 20, 70), 16, 64, 4, 8),       # 2 x 2 tiles, ragged last row of tiles and last column of tiles
    ((2, 17, 130), 16, 128, 4, 2),     # one row and one cell pair past a full tile
    ((3, 2, 4), 8, 32, 4, 3),          # smallest supported plane
    ((2, 48, 100), 16, 128, 4, 2),     # rows exact, columns ragged
])
@pytest.mark.parametrize("subset", [None, (True, True, True, True, True, True, False), (False, True, True, True, True, True, True)])
def test_model_matches_the_oracle(shape, BJ, BK, R, ci, subset):
    rng = np.random.default_rng(SEED + 90)
    x = rng.random(shape) - 0.5
    w = rng.standard_normal(7)
    has = [True] * 7 if subset is None else list(subset)
    out = np.full(shape, np.nan)
    single_apply(x, w, Cfg(BJ, BK, R), ci, out, has)
    off = np.array([o for o, h in zip(ORDER, has) if h], dtype=np.int32)
    ref = oracle.c.stencil_apply(x, off, np.array([v for v, h in zip(w, has) if h]), presorted=True)
    assert np.array_equal(out, ref)
