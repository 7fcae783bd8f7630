"""Host model of lap7_fused2_kernel (fidibench_b200/csrc/kernels_lapfused.cu).

Restates, with numpy and the SAME constants, shared-memory offsets, thread-to-cell mapping,
clamping rules and wrap-area selects as the CUDA kernel, what every consumer thread does for one
work item: TMA boxes with zero out-of-bounds fill land in a flat "shared memory" array, the
threads (vectorised over the thread index) read their pairs / neighbours at the kernel's byte
offsets, run the two-plane-deep pipeline and store the interior of level 2.  The result must
equal two applies of the oracle bit for bit; this pins the tiling/halo/wrap design on machines
without a GPU (the CUDA kernel itself is checked against the oracle by the -m gpu tests).
"""
from __future__ import annotations

import numpy as np


def align128(x: int) -> int:
    return (x + 127) // 128 * 128


class Cfg:
    """LapFusedCfg<BJ, R, STAGES>"""

    def __init__(self, BJ: int, R: int):
        self.BJ, self.BK, self.R = BJ, 128, R
        self.CJ = BJ + 2
        self.CK = self.BK + 4
        self.IN_ROWS = BJ + 4
        self.TX = self.CK // 2
        self.TY = self.CJ // R
        assert self.CJ % R == 0 and BJ % 8 == 0
        self.WORKERS = self.TX * self.TY
        self.CONSUMERS = (self.WORKERS + 31) // 32 * 32
        self.ROW_BYTES = self.CK * 8
        self.BODY_OFF = align128(2 * self.ROW_BYTES)
        self.ROW_SKEW = self.BODY_OFF - 2 * self.ROW_BYTES
        self.BOT_OFF = self.BODY_OFF + BJ * self.ROW_BYTES
        self.MAIN_BYTES = align128(self.BOT_OFF + 2 * self.ROW_BYTES)
        self.WBODY_OFF = 128
        self.WSKEW = self.WBODY_OFF - 32
        self.WBOT_OFF = self.WBODY_OFF + BJ * 16
        self.WRAP_BYTES = align128(self.WBOT_OFF + 32)
        self.WL_OFF = self.MAIN_BYTES
        self.WR_OFF = self.WL_OFF + self.WRAP_BYTES
        self.STAGE_BYTES = self.WR_OFF + self.WRAP_BYTES
        self.XP = self.CK * 8
        self.X_BYTES = align128(self.CJ * self.XP)
        assert self.BOT_OFF % 128 == 0 and self.WBOT_OFF % 128 == 0


def tma_box(smem: np.ndarray, dst: int, tensor: np.ndarray, c0: int, c1: int, c2: int, box0: int, box1: int) -> None:
    """cp.async.bulk.tensor.3d: box {box0, box1, 1} of `tensor` (planes, rows, cols) at (c0 col, c1 row,
    c2 plane) written densely at byte offset dst; out-of-bounds elements are zero."""
    assert dst % 128 == 0
    n0, n1, n2 = tensor.shape
    assert 0 <= c2 < n0
    blk = np.zeros((box1, box0))
    for r in range(box1):
        for c in range(box0):
            jj, kk = c1 + r, c0 + c
            if 0 <= jj < n1 and 0 <= kk < n2:
                blk[r, c] = tensor[c2, jj, kk]
    smem[dst // 8: dst // 8 + box0 * box1] = blk.ravel()


def fused_two_applies(x: np.ndarray, w, lo: int, hi: int, G: int, cfg: Cfg, ci: int, ibeg: int, iend: int,
                      out: np.ndarray) -> None:
    """One launch of the kernel on the slab [lo,hi) of the periodic field x: writes local output
    planes [ibeg,iend) of `out` (shape of the slab).  Ghost tensors hold G planes each."""
    C = cfg
    n0, n1, n2 = x.shape
    nloc = hi - lo
    assert n1 % C.BJ == 0 and n2 % C.BK == 0 and G >= 2 and nloc >= 2
    body = x[lo:hi]
    glo = np.stack([x[(lo - G + g) % n0] for g in range(G)])
    ghi = np.stack([x[(hi + g) % n0] for g in range(G)])
    njt, nkt = n1 // C.BJ, n2 // C.BK
    w0, w1, w2, w3, w4, w5, w6 = [np.float64(v) for v in w]
    # thread constants (workers only; the padding threads of the last warp compute and discard)
    tid = np.arange(C.WORKERS)
    tx, ty = tid % C.TX, tid // C.TX
    q0 = ty * C.R
    cb = tx * 16
    kmb = np.where(tx == 0, cb, cb - 8)
    kpb = np.where(tx == C.TX - 1, cb + 8, cb + 16)
    x_own = q0 * C.XP + cb
    x_up = np.where(q0 == 0, 0, q0 - 1) * C.XP + cb
    x_dn = np.where(q0 + C.R >= C.CJ, C.CJ - 1, q0 + C.R) * C.XP + cb
    x_km = q0 * C.XP + kmb
    x_kp = q0 * C.XP + kpb

    def main_row(s):
        return s * C.ROW_BYTES + np.where(s >= 2, C.ROW_SKEW, 0)

    def wrap_row(s):
        return s * 16 + np.where(s >= 2, C.WSKEW, 0)

    def acc(a, wt, v):
        return a + wt * v  # numpy float64: separately rounded multiply and add

    smem = np.full(C.STAGE_BYTES // 8, np.nan)       # one stage (the ring only reorders time)
    xbuf = [np.full(C.X_BYTES // 8, np.nan), np.full(C.X_BYTES // 8, np.nan)]
    xsel = 0
    planes = iend - ibeg
    nchunk = (planes + ci - 1) // ci
    for wi in range(njt * nkt * nchunk):
        kt = wi % nkt
        jt = (wi // nkt) % njt
        ic = wi // (nkt * njt)
        i0 = ibeg + ic * ci
        i1 = min(i0 + ci, iend)
        kb = kt * C.BK - 2
        j0 = jt * C.BJ
        jtop = n1 - 2 if j0 == 0 else j0 - 2
        jbot = 0 if j0 + C.BJ >= n1 else j0 + C.BJ
        first_k, last_k = kt == 0, kt == nkt - 1
        k = kt * C.BK - 2 + 2 * tx
        j = jt * C.BJ - 1 + q0
        own_wl = first_k & (tx == 0)
        own_wr = last_k & (tx == C.TX - 1)
        km_wl = first_k & (tx == 1)
        kp_wr = last_k & (tx == C.TX - 2)
        store_cols = (tx >= 1) & (tx <= C.TX - 2)
        z = np.zeros((C.R, C.WORKERS, 2))
        below0, part1, below1, part2 = z.copy(), z.copy(), z.copy(), z.copy()
        for p in range(i0 - 2, i1 + 2):
            # ---- producer
            smem[:] = np.nan
            if p < 0:
                t, pl = glo, G + p
            elif p >= nloc:
                t, pl = ghi, p - nloc
            else:
                t, pl = body, p
            tma_box(smem, 0, t, kb, jtop, pl, C.CK, 2)
            tma_box(smem, C.BODY_OFF, t, kb, j0, pl, C.CK, C.BJ)
            tma_box(smem, C.BOT_OFF, t, kb, jbot, pl, C.CK, 2)
            if first_k:
                tma_box(smem, C.WL_OFF, t, n2 - 2, jtop, pl, 2, 2)
                tma_box(smem, C.WL_OFF + C.WBODY_OFF, t, n2 - 2, j0, pl, 2, C.BJ)
                tma_box(smem, C.WL_OFF + C.WBOT_OFF, t, n2 - 2, jbot, pl, 2, 2)
            if last_k:
                tma_box(smem, C.WR_OFF, t, 0, jtop, pl, 2, 2)
                tma_box(smem, C.WR_OFF + C.WBODY_OFF, t, 0, j0, pl, 2, C.BJ)
                tma_box(smem, C.WR_OFF + C.WBOT_OFF, t, 0, jbot, pl, 2, 2)

            # ---- consumers
            def lds_v2(mem, addr):
                assert np.all(addr % 16 == 0)
                return np.stack([mem[addr // 8], mem[addr // 8 + 1]], axis=-1)

            def lds_f64(mem, addr):
                return mem[addr // 8]

            def pair_at(s):
                ad = np.where(own_wl, C.WL_OFF + wrap_row(s), np.where(own_wr, C.WR_OFF + wrap_row(s), main_row(s) + cb))
                return lds_v2(smem, ad)

            up = pair_at(q0)
            dn = pair_at(q0 + C.R + 1)
            c = np.empty((C.R, C.WORKERS, 2))
            km = np.empty((C.R, C.WORKERS))
            kp = np.empty((C.R, C.WORKERS))
            for r in range(C.R):
                s = q0 + 1 + r
                c[r] = pair_at(s)
                km[r] = lds_f64(smem, np.where(km_wl, C.WL_OFF + wrap_row(s) + 8, main_row(s) + kmb))
                kp[r] = lds_f64(smem, np.where(kp_wr, C.WR_OFF + wrap_row(s), main_row(s) + kpb))
            l1 = np.empty_like(c)
            for r in range(C.R):
                l1[r, :, 0] = acc(part1[r, :, 0], w6, c[r, :, 0])
                l1[r, :, 1] = acc(part1[r, :, 1], w6, c[r, :, 1])
            if p >= i0 + 2:
                for r in range(C.R):
                    q = q0 + r
                    ok = store_cols & (q >= 1) & (q <= C.CJ - 2)
                    vx = acc(part2[r, :, 0], w6, l1[r, :, 0])
                    vy = acc(part2[r, :, 1], w6, l1[r, :, 1])
                    rows, cols = (j + r)[ok], k[ok]
                    assert np.all((rows >= 0) & (rows < n1) & (cols >= 0) & (cols + 1 < n2 + 0 + 1))
                    assert np.all(np.isnan(out[p - 2, rows, cols])), "cell stored twice"
                    out[p - 2, rows, cols] = vx[ok]
                    out[p - 2, rows, cols + 1] = vy[ok]
            xb = xbuf[xsel]
            xsel ^= 1
            xb[:] = np.nan
            for r in range(C.R):
                ad = (x_own + r * C.XP) // 8
                xb[ad] = l1[r, :, 0]
                xb[ad + 1] = l1[r, :, 1]
            for r in range(C.R):
                jm = up if r == 0 else c[r - 1]
                jp = dn if r == C.R - 1 else c[r + 1]
                xx = acc(0.0, w0, below0[r, :, 0]); yy = acc(0.0, w0, below0[r, :, 1])
                xx = acc(xx, w1, jm[:, 0]);          yy = acc(yy, w1, jm[:, 1])
                xx = acc(xx, w2, km[r]);             yy = acc(yy, w2, c[r, :, 0])
                xx = acc(xx, w3, c[r, :, 0]);        yy = acc(yy, w3, c[r, :, 1])
                xx = acc(xx, w4, c[r, :, 1]);        yy = acc(yy, w4, kp[r])
                xx = acc(xx, w5, jp[:, 0]);          yy = acc(yy, w5, jp[:, 1])
                part1[r, :, 0], part1[r, :, 1] = xx, yy
            below0 = c.copy()
            # named barrier
            up1 = lds_v2(xb, x_up)
            dn1 = lds_v2(xb, x_dn)
            for r in range(C.R):
                km[r] = lds_f64(xb, x_km + r * C.XP)
                kp[r] = lds_f64(xb, x_kp + r * C.XP)
            for r in range(C.R):
                jm = up1 if r == 0 else l1[r - 1]
                jp = dn1 if r == C.R - 1 else l1[r + 1]
                xx = acc(0.0, w0, below1[r, :, 0]); yy = acc(0.0, w0, below1[r, :, 1])
                xx = acc(xx, w1, jm[:, 0]);          yy = acc(yy, w1, jm[:, 1])
                xx = acc(xx, w2, km[r]);             yy = acc(yy, w2, l1[r, :, 0])
                xx = acc(xx, w3, l1[r, :, 0]);       yy = acc(yy, w3, l1[r, :, 1])
                xx = acc(xx, w4, l1[r, :, 1]);       yy = acc(yy, w4, kp[r])
                xx = acc(xx, w5, jp[:, 0]);          yy = acc(yy, w5, jp[:, 1])
                part2[r, :, 0], part2[r, :, 1] = xx, yy
            below1 = l1.copy()
