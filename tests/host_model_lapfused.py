"""Host model of lap7_fused2_kernel (fidibench_b200/csrc/kernels_lapfused.cu).

Restates, with numpy and the SAME constants, shared-memory offsets and thread-to-cell mapping
as the CUDA kernel, what the loader warp and every consumer thread do for one work item: TMA
boxes with zero out-of-bounds fill land in a flat "shared memory" array, the loader copies the
periodic wrap columns into the tile rows, the threads (vectorised over the thread index) read
their pairs / neighbours at the kernel's byte offsets, run the two-plane-deep pipeline and store
the interior of level 2.  Memory nobody wrote is NaN here, so a read that matters shows up.  The result must
equal two applies of the oracle bit for bit; this pins the tiling/halo/wrap design on machines
without a GPU (the CUDA kernel itself is checked against the oracle by the -m gpu tests).
"""
from __future__ import annotations

import numpy as np


def align128(x: int) -> int:
    return (x + 127) // 128 * 128


class Cfg:
    """LapFusedCfg<BJ, R, STAGES, BK>"""

    def __init__(self, BJ: int, R: int, BK: int = 128):
        self.BJ, self.BK, self.R = BJ, BK, R
        self.CJ = BJ + 2
        self.CK = BK + 4
        self.IN_ROWS = BJ + 4
        self.TX = self.CK // 2
        self.TY = self.CJ // R
        assert self.CJ % R == 0 and BJ % 2 == 0
        self.WORKERS = self.TX * self.TY
        self.BKP = BK + 8
        self.PITCH = self.BKP * 8
        self.BODY_OFF = 2 * self.PITCH
        self.BOT_OFF = (2 + BJ) * self.PITCH
        self.MAIN_BYTES = self.IN_ROWS * self.PITCH
        self.WPITCH = 64
        self.WL_OFF = align128(self.MAIN_BYTES)
        self.WR_OFF = self.WL_OFF + self.IN_ROWS * self.WPITCH
        self.STAGE_BYTES = self.WR_OFF + self.IN_ROWS * self.WPITCH
        self.X_BYTES = align128((self.CJ + 2) * self.PITCH)
        assert (2 * self.PITCH) % 128 == 0 and (self.IN_ROWS * self.WPITCH) % 128 == 0 and self.IN_ROWS <= 32


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
                      out: np.ndarray, unit: bool = False, shfl: bool = False, lean: bool = False) -> None:
    """One launch of the kernel on the slab [lo,hi) of the periodic field x: writes local output
    planes [ibeg,iend) of `out` (shape of the slab).  Ghost tensors hold G planes each.
    lean: lap7_fused2_lean_kernel -- the exchange tile keeps even cells (x) and odd cells (y) of a row in two halves
    (8 bytes per thread), and the register sets are NOT reset between work items: here they are poisoned with NaN at
    every item start, so a stale value reaching a stored cell fails the comparison."""
    C = cfg
    n0, n1, n2 = x.shape
    nloc = hi - lo
    assert n1 % C.BJ == 0 and n2 % C.BK == 0 and G >= 2 and nloc >= 2
    body = x[lo:hi]
    glo = np.stack([x[(lo - G + g) % n0] for g in range(G)])
    ghi = np.stack([x[(hi + g) % n0] for g in range(G)])
    njt, nkt = n1 // C.BJ, n2 // C.BK
    w0, w1, w2, w3, w4, w5, w6 = [np.float64(v) for v in w]
    if unit:
        assert all(v == 1.0 for v in (w0, w1, w2, w4, w5, w6))
    # thread constants (workers only; the padding threads of the last warp compute and discard)
    tid = np.arange(C.WORKERS)
    tx, ty = tid % C.TX, tid // C.TX
    q0 = ty * C.R
    tb = q0 * C.PITCH + 16 + tx * 16
    P = C.PITCH
    xs = q0 * P + 8 + tx * 8   # lean: row above the thread's first row, x half, this thread's slot
    YOFF = P // 2
    assert not lean or (C.TX + 2) * 8 <= YOFF
    rowmask = [(q0 + r >= 1) & (q0 + r <= C.CJ - 2) for r in range(C.R)]
    store_cols = (tx >= 1) & (tx <= C.TX - 2)
    # SHFL variant: the k-1 / k+1 cells come from the adjacent LANE's registers except at a warp's edge
    # lanes, a row's end threads and the last worker (whose next lane is a padding thread)
    lane = tid & 31
    edge_lo = (lane == 0) | (tx == 0)
    edge_hi = (lane == 31) | (tx == C.TX - 1) | (tid >= C.WORKERS - 1)

    def nbr_cells(pairs, lo_loaded, hi_loaded):
        """pairs: (WORKERS, 2) registers of one row; *_loaded: what the shared-memory load returns"""
        if not shfl:
            return lo_loaded, hi_loaded
        up_lane = np.concatenate([[np.nan], pairs[:-1, 1]])    # __shfl_up(y, 1)
        dn_lane = np.concatenate([pairs[1:, 0], [np.nan]])     # __shfl_down(x, 1)
        return np.where(edge_lo, lo_loaded, up_lane), np.where(edge_hi, hi_loaded, dn_lane)

    def acc(a, wt, v, u=False):
        # numpy float64: separately rounded multiply and add; the unit kernel skips the multiply
        return a + v if (unit and u) else a + wt * v

    smem = np.full(C.STAGE_BYTES // 8, np.nan)       # one stage (the ring only reorders time)
    xbuf = [np.full(C.X_BYTES // 8, np.nan), np.full(C.X_BYTES // 8, np.nan)]
    xsel = 0
    planes = iend - ibeg
    nchunk = (planes + ci - 1) // ci

    def lds_v2(mem, addr):
        assert np.all(addr % 16 == 0) and np.all(addr >= 0)
        return np.stack([mem[addr // 8], mem[addr // 8 + 1]], axis=-1)

    def lds_f64(mem, addr):
        assert np.all(addr >= 0)
        return mem[addr // 8]

    for wi in range(njt * nkt * nchunk):
        kt = wi % nkt
        jt = (wi // nkt) % njt
        ic = wi // (nkt * njt)
        i0 = ibeg + ic * ci
        i1 = min(i0 + ci, iend)
        kb = kt * C.BK - 4
        j0 = jt * C.BJ
        jtop = n1 - 2 if j0 == 0 else j0 - 2
        jbot = 0 if j0 + C.BJ >= n1 else j0 + C.BJ
        first_k, last_k = kt == 0, kt == nkt - 1
        k = kt * C.BK - 2 + 2 * tx
        j = jt * C.BJ - 1 + q0
        z = np.full((C.R, C.WORKERS, 2), np.nan) if lean else np.zeros((C.R, C.WORKERS, 2))
        below0, part1, below1, part2 = z.copy(), z.copy(), z.copy(), z.copy()
        for p in range(i0 - 2, i1 + 2):
            # ---- loader warp: TMA boxes, then the wrap columns copied into the tile rows
            smem[:] = np.nan
            if p < 0:
                t, pl = glo, G + p
            elif p >= nloc:
                t, pl = ghi, p - nloc
            else:
                t, pl = body, p
            tma_box(smem, 0, t, kb, jtop, pl, C.BKP, 2)
            tma_box(smem, C.BODY_OFF, t, kb, j0, pl, C.BKP, C.BJ)
            tma_box(smem, C.BOT_OFF, t, kb, jbot, pl, C.BKP, 2)
            if first_k:
                tma_box(smem, C.WL_OFF, t, n2 - 8, jtop, pl, 8, 2)
                tma_box(smem, C.WL_OFF + 2 * C.WPITCH, t, n2 - 8, j0, pl, 8, C.BJ)
                tma_box(smem, C.WL_OFF + (2 + C.BJ) * C.WPITCH, t, n2 - 8, jbot, pl, 8, 2)
            if last_k:
                tma_box(smem, C.WR_OFF, t, 0, jtop, pl, 8, 2)
                tma_box(smem, C.WR_OFF + 2 * C.WPITCH, t, 0, j0, pl, 8, C.BJ)
                tma_box(smem, C.WR_OFF + (2 + C.BJ) * C.WPITCH, t, 0, jbot, pl, 8, 2)
            lanes = np.arange(C.IN_ROWS)
            if first_k:
                v = lds_v2(smem, C.WL_OFF + lanes * C.WPITCH + 48)
                ad = (lanes * C.PITCH + 16) // 8
                smem[ad], smem[ad + 1] = v[:, 0], v[:, 1]
            if last_k:
                v = lds_v2(smem, C.WR_OFF + lanes * C.WPITCH)
                ad = (lanes * C.PITCH + C.CK * 8) // 8
                smem[ad], smem[ad + 1] = v[:, 0], v[:, 1]

            # ---- consumers
            sb = tb
            up = lds_v2(smem, sb)
            dn = lds_v2(smem, sb + (C.R + 1) * P)
            c = np.empty((C.R, C.WORKERS, 2))
            km = np.empty((C.R, C.WORKERS))
            kp = np.empty((C.R, C.WORKERS))
            for r in range(C.R):
                c[r] = lds_v2(smem, sb + (1 + r) * P)
                km[r], kp[r] = nbr_cells(c[r], lds_f64(smem, sb + (1 + r) * P - 8), lds_f64(smem, sb + (1 + r) * P + 16))
            l1 = np.empty_like(c)
            for r in range(C.R):
                l1[r, :, 0] = acc(part1[r, :, 0], w6, c[r, :, 0], True)
                l1[r, :, 1] = acc(part1[r, :, 1], w6, c[r, :, 1], True)
            if p >= i0 + 2:
                for r in range(C.R):
                    ok = store_cols & rowmask[r]
                    vx = acc(part2[r, :, 0], w6, l1[r, :, 0], True)
                    vy = acc(part2[r, :, 1], w6, l1[r, :, 1], True)
                    rows, cols = (j + r)[ok], k[ok]
                    assert np.all((rows >= 0) & (rows < n1) & (cols >= 0) & (cols + 1 < n2 + 1))
                    assert np.all(np.isnan(out[p - 2, rows, cols])), "cell stored twice"
                    out[p - 2, rows, cols] = vx[ok]
                    out[p - 2, rows, cols + 1] = vy[ok]
            xb = xbuf[xsel]
            xsel ^= 1
            xb[:] = np.nan
            for r in range(C.R):
                if lean:
                    xb[(xs + (1 + r) * P) // 8] = l1[r, :, 0]
                    xb[(xs + (1 + r) * P + YOFF) // 8] = l1[r, :, 1]
                else:
                    ad = (tb + (1 + r) * P) // 8
                    xb[ad] = l1[r, :, 0]
                    xb[ad + 1] = l1[r, :, 1]
            for r in range(C.R):
                jm = up if r == 0 else c[r - 1]
                jp = dn if r == C.R - 1 else c[r + 1]
                xx = acc(0.0, w0, below0[r, :, 0], True); yy = acc(0.0, w0, below0[r, :, 1], True)
                xx = acc(xx, w1, jm[:, 0], True);          yy = acc(yy, w1, jm[:, 1], True)
                xx = acc(xx, w2, km[r], True);             yy = acc(yy, w2, c[r, :, 0], True)
                xx = acc(xx, w3, c[r, :, 0]);              yy = acc(yy, w3, c[r, :, 1])
                xx = acc(xx, w4, c[r, :, 1], True);        yy = acc(yy, w4, kp[r], True)
                xx = acc(xx, w5, jp[:, 0], True);          yy = acc(yy, w5, jp[:, 1], True)
                part1[r, :, 0], part1[r, :, 1] = xx, yy
            below0 = c.copy()
            # named barrier
            if lean:
                up1 = np.stack([lds_f64(xb, xs), lds_f64(xb, xs + YOFF)], axis=-1)
                dn1 = np.stack([lds_f64(xb, xs + (C.R + 1) * P), lds_f64(xb, xs + (C.R + 1) * P + YOFF)], axis=-1)
                for r in range(C.R):
                    km[r] = lds_f64(xb, xs + (1 + r) * P + YOFF - 8)   # y of the thread to the left
                    kp[r] = lds_f64(xb, xs + (1 + r) * P + 8)          # x of the thread to the right
            else:
                up1 = lds_v2(xb, tb)
                dn1 = lds_v2(xb, tb + (C.R + 1) * P)
                for r in range(C.R):
                    km[r], kp[r] = nbr_cells(l1[r], lds_f64(xb, tb + (1 + r) * P - 8), lds_f64(xb, tb + (1 + r) * P + 16))
            for r in range(C.R):
                jm = up1 if r == 0 else l1[r - 1]
                jp = dn1 if r == C.R - 1 else l1[r + 1]
                xx = acc(0.0, w0, below1[r, :, 0], True); yy = acc(0.0, w0, below1[r, :, 1], True)
                xx = acc(xx, w1, jm[:, 0], True);          yy = acc(yy, w1, jm[:, 1], True)
                xx = acc(xx, w2, km[r], True);             yy = acc(yy, w2, l1[r, :, 0], True)
                xx = acc(xx, w3, l1[r, :, 0]);             yy = acc(yy, w3, l1[r, :, 1])
                xx = acc(xx, w4, l1[r, :, 1], True);       yy = acc(yy, w4, kp[r], True)
                xx = acc(xx, w5, jp[:, 0], True);          yy = acc(yy, w5, jp[:, 1], True)
                part2[r, :, 0], part2[r, :, 1] = xx, yy
            below1 = l1.copy()
