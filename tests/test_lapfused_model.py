"""CPU check of the fused two-apply 7-point tile pipeline: the host model that mirrors
lap7_fused2_kernel's shared-memory layout and thread mapping (tests/host_model_lapfused.py) must equal
two oracle applies bit for bit -- single slab, slab of a ring (ghost tensors), ragged plane chunks,
one and several tiles per axis, non-unit weights."""
import numpy as np
import pytest

import oracle
from host_model_lapfused import Cfg, fused_two_applies

SEED = 20261017
ORDER = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]  # std::map order


def two_applies(x, w):
    off = np.array(ORDER, dtype=np.int32)
    y = oracle.c.stencil_apply(x, off, np.array(w), presorted=True)
    return oracle.c.stencil_apply(y, off, np.array(w), presorted=True)


@pytest.mark.parametrize("shape,BJ,R,ci", [
    ((6, 32, 256), 16, 3, 4),     # 2 x 2 tiles, ragged second chunk
    ((5, 16, 128), 16, 6, 8),     # one tile per axis: every halo wraps onto the tile itself
    ((4, 16, 256), 8, 5, 2),      # 8-row tiles
    ((2, 32, 128), 16, 2, 2),     # thinnest slab
])
@pytest.mark.parametrize("unit", [False, True])
@pytest.mark.parametrize("lean", [False, True])
def test_model_single_slab(shape, BJ, R, ci, unit, lean):
    """lean = lap7_fused2_lean_kernel (the default): split exchange tile, register sets never reset between items."""
    rng = np.random.default_rng(SEED)
    x = rng.random(shape) - 0.5
    w = [1.0, 1.0, 1.0, -6.0, 1.0, 1.0, 1.0]
    out = np.full(shape, np.nan)
    fused_two_applies(x, w, 0, shape[0], 2, Cfg(BJ, R), ci, 0, shape[0], out, unit=unit, lean=lean)
    assert np.array_equal(out, two_applies(x, w))


@pytest.mark.parametrize("shape,BJ,R", [((4, 32, 256), 16, 6), ((3, 16, 128), 16, 3)])
def test_model_shuffled_k_neighbours(shape, BJ, R):
    """The experimental SHFL variant (k-1 / k+1 from the adjacent lanes, loads only at warp and row edges)."""
    rng = np.random.default_rng(SEED + 3)
    x = rng.random(shape)
    w = [1.0, 1.0, 1.0, -6.0, 1.0, 1.0, 1.0]
    out = np.full(shape, np.nan)
    fused_two_applies(x, w, 0, shape[0], 2, Cfg(BJ, R), 4, 0, shape[0], out, unit=True, shfl=True)
    assert np.array_equal(out, two_applies(x, w))


def test_model_64_cell_tiles():
    rng = np.random.default_rng(SEED + 2)
    x = rng.random((4, 16, 192))
    w = [1.0, 1.0, 1.0, -6.0, 1.0, 1.0, 1.0]
    out = np.full(x.shape, np.nan)
    fused_two_applies(x, w, 0, 4, 2, Cfg(16, 3, BK=64), 4, 0, 4, out, unit=True)
    assert np.array_equal(out, two_applies(x, w))


def test_model_slab_of_a_ring_with_boundary_split():
    """Slab [4,8) of 12 planes, ghost depth 2, launched as the runtime does: bottom planes, top planes,
    interior (runtime.cu: split_slab) -- and with weights that are not 1 or -6."""
    rng = np.random.default_rng(SEED + 1)
    x = rng.random((12, 16, 128))
    w = [0.25, -0.5, 1.5, -6.0, 0.75, 2.0, -1.25]
    ref = two_applies(x, w)
    lo, hi = 4, 8
    out = np.full((hi - lo, 16, 128), np.nan)
    for ibeg, iend in ((0, 2), (2, 4)):
        fused_two_applies(x, w, lo, hi, 2, Cfg(16, 3), 64, ibeg, iend, out)
    assert np.array_equal(out, ref[lo:hi])
    out = np.full((hi - lo, 16, 128), np.nan)
    for ibeg, iend in ((0, 2), (2, 4)):
        fused_two_applies(x, w, lo, hi, 2, Cfg(16, 3), 64, ibeg, iend, out, lean=True)
    assert np.array_equal(out, ref[lo:hi])
    # ghost tensors deeper than the sweep (G = 3): the planes nearest the body are the ones read
    out = np.full((hi - lo, 16, 128), np.nan)
    fused_two_applies(x, w, lo, hi, 3, Cfg(16, 3), 3, 0, 4, out)
    assert np.array_equal(out, ref[lo:hi])
