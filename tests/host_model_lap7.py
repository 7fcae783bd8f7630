"""Host model of lap7_tma_kernel<C, RAGGED> (fidibench_b200/csrc/kernels_tma.cu): the single-apply 7-point tile
pipeline with the kernel's constants, shared-memory offsets, thread-to-cell mapping and -- for planes no tile divides
-- its ragged-tile rules: TMA zero-fills outside the tensor, the row below the plane's last row comes from the wrap
row (`bot`) for whichever thread row holds global row n1-1, the cell right of the plane's last pair from the wrap
columns for whichever thread holds cells n2-2, n2-1, stores are masked.  Memory nobody wrote is NaN.  Must equal one
oracle apply bit for bit (tests/test_lap7_model.py)."""
from __future__ import annotations

import numpy as np

from host_model_lapfused import tma_box


class Cfg:
    """Lap7Cfg<BJ, BK, R, STAGES>"""

    def __init__(self, BJ: int, BK: int, R: int):
        self.BJ, self.BK, self.R = BJ, BK, R
        self.BKH = BK + 4
        self.TX, self.TY = BK // 2, BJ // R
        self.CONSUMERS = self.TX * self.TY
        self.ROW_BYTES = self.BKH * 8
        self.ROW_SLOT = (self.ROW_BYTES + 127) // 128 * 128
        self.BODY_BYTES = BJ * self.ROW_BYTES
        self.WRAP_BYTES = BJ * 16
        self.TOP_OFF = 0
        self.BODY_OFF = self.ROW_SLOT
        self.BOT_OFF = self.BODY_OFF + self.BODY_BYTES
        self.WRAPL_OFF = self.BOT_OFF + self.ROW_SLOT
        self.WRAPR_OFF = self.WRAPL_OFF + (self.WRAP_BYTES + 127) // 128 * 128
        self.STAGE_BYTES = self.WRAPR_OFF + (self.WRAP_BYTES + 127) // 128 * 128
        assert BJ % R == 0 and self.CONSUMERS % 32 == 0 and self.BODY_BYTES % 128 == 0


ORDER = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]  # std::map order


def single_apply(x: np.ndarray, w, cfg: Cfg, ci: int, out: np.ndarray, has=None) -> None:
    """One launch on a single periodic slab: out <- stencil(x); w[7] in application order, has[7] which branches exist."""
    C = cfg
    n0, n1, n2 = x.shape
    assert n1 >= 2 and n2 >= 4 and n2 % 2 == 0
    has = [True] * 7 if has is None else list(has)
    skip_lo, skip_hi = not has[0], not has[6]
    ragged = n1 % C.BJ != 0 or n2 % C.BK != 0
    njt, nkt = -(-n1 // C.BJ), -(-n2 // C.BK)
    w = [np.float64(v) for v in w]
    tid = np.arange(C.CONSUMERS)
    tx, ty = tid % C.TX, tid // C.TX
    r0 = ty * C.R
    col = (2 + 2 * tx) * 8

    def lds_v2(mem, addr):
        assert np.all(addr % 16 == 0)
        return np.stack([mem[addr // 8], mem[addr // 8 + 1]], axis=-1)

    def lds_f64(mem, addr):
        return mem[addr // 8]

    def acc(a, wt, v):
        return a + wt * v

    nchunk = (n0 + ci - 1) // ci
    smem = np.full(C.STAGE_BYTES // 8, np.nan)
    for wi in range(njt * nkt * nchunk):
        kt, jt, ic = wi % nkt, (wi // nkt) % njt, wi // (nkt * njt)
        i0, i1 = ic * ci, min(ic * ci + ci, n0)
        k0, j0 = kt * C.BK - 2, jt * C.BJ
        jm = n1 - 1 if j0 == 0 else j0 - 1
        jp = 0 if j0 + C.BJ >= n1 else j0 + C.BJ
        first_k, last_k = kt == 0, kt == nkt - 1
        k = kt * C.BK + 2 * tx
        j = jt * C.BJ + r0
        wrapl = first_k & (tx == 0)
        wrapr = last_k & (k + 2 == n2)
        last_row = min(C.BJ, n1 - jt * C.BJ) - 1 if ragged else C.BJ - 1
        k_ok = (k < n2) if ragged else np.ones_like(k, dtype=bool)

        def body_only(i):
            smem[:] = np.nan
            tma_box(smem, C.BODY_OFF, x, k0, j0, i % n0, C.BKH, C.BJ)   # the ghost tensors alias the far planes
            return np.stack([lds_v2(smem, C.BODY_OFF + (r0 + r) * C.ROW_BYTES + col) for r in range(C.R)])

        def finish(ip, partial, above):
            for r in range(C.R):
                vx, vy = partial[r, :, 0], partial[r, :, 1]
                if has[6]:
                    vx, vy = acc(vx, w[6], above[r, :, 0]), acc(vy, w[6], above[r, :, 1])
                ok = k_ok & (r0 + r <= last_row) if ragged else np.ones_like(k, dtype=bool)
                rows, cols = (j + r)[ok], k[ok]
                assert np.all((rows < n1) & (cols + 1 < n2))
                assert np.all(np.isnan(out[ip, rows, cols])), "cell stored twice"
                out[ip, rows, cols], out[ip, rows, cols + 1] = vx[ok], vy[ok]

        below = np.zeros((C.R, C.CONSUMERS, 2)) if skip_lo else body_only(i0 - 1)
        partial = None
        for i in range(i0, i1):
            smem[:] = np.nan
            tma_box(smem, C.BODY_OFF, x, k0, j0, i, C.BKH, C.BJ)
            tma_box(smem, C.TOP_OFF, x, k0, jm, i, C.BKH, 1)
            tma_box(smem, C.BOT_OFF, x, k0, jp, i, C.BKH, 1)
            if first_k:
                tma_box(smem, C.WRAPL_OFF, x, n2 - 2, j0, i, 2, C.BJ)
            if last_k:
                tma_box(smem, C.WRAPR_OFF, x, 0, j0, i, 2, C.BJ)
            up_row = np.where(r0 == 0, C.TOP_OFF, C.BODY_OFF + (r0 - 1) * C.ROW_BYTES)
            dn_row = np.where(r0 + C.R == C.BJ, C.BOT_OFF, C.BODY_OFF + (r0 + C.R) * C.ROW_BYTES)
            up, dn = lds_v2(smem, up_row + col), lds_v2(smem, dn_row + col)
            bot = lds_v2(smem, C.BOT_OFF + col) if ragged else dn
            ctr = np.empty((C.R, C.CONSUMERS, 2))
            km1 = np.empty((C.R, C.CONSUMERS))
            kp1 = np.empty((C.R, C.CONSUMERS))
            for r in range(C.R):
                row = C.BODY_OFF + (r0 + r) * C.ROW_BYTES
                ctr[r] = lds_v2(smem, row + col)
                km1[r] = lds_f64(smem, np.where(wrapl, C.WRAPL_OFF + (r0 + r) * 16 + 8, row + col - 8))
                kp1[r] = lds_f64(smem, np.where(wrapr, C.WRAPR_OFF + (r0 + r) * 16, row + col + 16))
            if not skip_hi and i > i0:
                finish(i - 1, partial, ctr)
            partial = np.empty((C.R, C.CONSUMERS, 2))
            for r in range(C.R):
                jmv = up if r == 0 else ctr[r - 1]
                jpv = dn if r == C.R - 1 else ctr[r + 1]
                if ragged:
                    jpv = np.where((r0 + r == last_row)[:, None], bot, jpv)
                xx, yy = np.zeros(C.CONSUMERS), np.zeros(C.CONSUMERS)
                if has[0]: xx, yy = acc(xx, w[0], below[r, :, 0]), acc(yy, w[0], below[r, :, 1])
                if has[1]: xx, yy = acc(xx, w[1], jmv[:, 0]), acc(yy, w[1], jmv[:, 1])
                if has[2]: xx, yy = acc(xx, w[2], km1[r]), acc(yy, w[2], ctr[r, :, 0])
                if has[3]: xx, yy = acc(xx, w[3], ctr[r, :, 0]), acc(yy, w[3], ctr[r, :, 1])
                if has[4]: xx, yy = acc(xx, w[4], ctr[r, :, 1]), acc(yy, w[4], kp1[r])
                if has[5]: xx, yy = acc(xx, w[5], jpv[:, 0]), acc(yy, w[5], jpv[:, 1])
                partial[r, :, 0], partial[r, :, 1] = xx, yy
            below = ctr
            if skip_hi:
                finish(i, partial, ctr)
        if not skip_hi:
            finish(i1 - 1, partial, body_only(i1))
