"""Host model of upwind3d_fused_kernel (fidibench_b200/csrc/kernels_fused.cu): same constants,
shared-memory offsets and thread-to-cell mapping as the CUDA kernel, numpy instead of threads (see
tests/host_model_lapfused.py for the idea).  Memory nobody wrote is NaN."""
from __future__ import annotations

import numpy as np

from host_model_lapfused import align128, tma_box


class Cfg:
    """FusedCfg<T, CJ, R, STAGES, BK>"""

    def __init__(self, T: int, CJ: int, R: int, BK: int = 128):
        self.T, self.CJ, self.R, self.BK = T, CJ, R, BK
        self.HKC = 2 * (T // 2)
        self.CK = BK + self.HKC
        self.TX = self.CK // 2
        self.TY = CJ // R
        assert CJ % R == 0 and 2 <= T <= 4 and CJ > T
        self.WORKERS = self.TX * self.TY
        self.BJ = CJ - (T - 1)
        self.HR = (T + 1) // 2 * 2
        self.IN_ROWS = self.HR + self.BJ
        self.BKP = BK + 8
        self.LEFT = 8 - self.HKC
        self.PITCH = self.BKP * 8
        self.BODY_OFF = self.HR * self.PITCH
        self.MAIN_BYTES = self.IN_ROWS * self.PITCH
        self.WPITCH = 64
        self.W_OFF = align128(self.MAIN_BYTES)
        self.STAGE_BYTES = align128(self.W_OFF + self.IN_ROWS * self.WPITCH)
        self.X_BYTES = align128((CJ + 1) * self.PITCH)
        assert (self.HR * self.PITCH) % 128 == 0 and self.IN_ROWS <= 32


def fused_steps(x: np.ndarray, c, lo: int, hi: int, G: int, cfg: Cfg, ci: int, ibeg: int, iend: int,
                out: np.ndarray, split: bool = False, reverse: bool = False) -> None:
    """One launch on the slab [lo,hi) of the periodic field x: local output planes [ibeg,iend) of `out`
    receive the field T time steps later.  c = (c0, c1, c2).
    split: the lean formulation's exchange-tile layout (Lean<C, PUSH, SPLIT = true>): even cells (x) and odd cells
    (y) of a tile row in two halves of the row, 8 bytes per thread, and only what a neighbour reads is stored
    (every row's y, the last row's x).  reverse: chunks walked top chunk first (single-launch ring sweeps)."""
    C = cfg
    T = C.T
    n0, n1, n2 = x.shape
    nloc = hi - lo
    assert G >= T and n2 % 2 == 0 and n2 >= 16 and n1 >= 8
    body = x[lo:hi]
    glo = np.stack([x[(lo - G + g) % n0] for g in range(G)])
    njt, nkt = -(-n1 // C.BJ), -(-n2 // C.BK)
    c0, c1, c2 = [np.float64(v) for v in c]
    tid = np.arange(C.WORKERS)
    tx, ty = tid % C.TX, tid // C.TX
    q0 = ty * C.R
    P = C.PITCH
    xt = q0 * P + (C.LEFT + 2 * tx) * 8
    tb = xt + (C.HR - T) * P
    xs = q0 * P + 8 + tx * 8          # split layout: row above the thread's first row, x half, this thread's slot
    YOFF = P // 2
    assert not split or (C.TX + 1) * 8 <= YOFF
    rowmask = [(q0 + r >= T - 1) for r in range(C.R)]

    def cell(ctr, im1, jm1, km1):
        t = ctr
        t = t - c0 * (im1 - ctr)
        t = t - c1 * (jm1 - ctr)
        t = t - c2 * (km1 - ctr)
        return t

    def lds_v2(mem, addr):
        assert np.all(addr % 16 == 0) and np.all(addr >= 0)
        return np.stack([mem[addr // 8], mem[addr // 8 + 1]], axis=-1)

    def lds_f64(mem, addr):
        assert np.all(addr >= 0)
        return mem[addr // 8]

    smem = np.full(C.STAGE_BYTES // 8, np.nan)
    xbuf = [np.full(C.X_BYTES // 8, np.nan), np.full(C.X_BYTES // 8, np.nan)]
    xsel = 0
    planes = iend - ibeg
    nchunk = (planes + ci - 1) // ci
    for wi in range(njt * nkt * nchunk):
        kt = wi % nkt
        jt = (wi // nkt) % njt
        ic = wi // (nkt * njt)
        if reverse:
            ic = nchunk - 1 - ic
        i0 = ibeg + ic * ci
        i1 = min(i0 + ci, iend)
        kb = kt * C.BK - 8
        j0 = jt * C.BJ
        jh = j0 - C.HR + n1 if j0 - C.HR < 0 else j0 - C.HR
        first_k = kt == 0
        k = kt * C.BK - C.HKC + 2 * tx
        j = jt * C.BJ - (T - 1) + q0
        store_cols = (2 * tx >= C.HKC) & (k < n2)
        rmask = [rowmask[r] & (j + r < n1) for r in range(C.R)]
        carry = np.zeros((T, C.R, C.WORKERS, 2))
        for p in range(i0 - T, i1):
            # ---- loader
            smem[:] = np.nan
            t, pl = (glo, G + p) if p < 0 else (body, p)
            tma_box(smem, 0, t, kb, jh, pl, C.BKP, C.HR)
            tma_box(smem, C.BODY_OFF, t, kb, j0, pl, C.BKP, C.BJ)
            if first_k:
                tma_box(smem, C.W_OFF, t, n2 - 8, jh, pl, 8, C.HR)
                tma_box(smem, C.W_OFF + C.HR * C.WPITCH, t, n2 - 8, j0, pl, 8, C.BJ)
                lanes = np.arange(C.IN_ROWS)
                for off in (16, 32, 48):
                    v = lds_v2(smem, C.W_OFF + lanes * C.WPITCH + off)
                    ad = (lanes * C.PITCH + off) // 8
                    smem[ad], smem[ad + 1] = v[:, 0], v[:, 1]
            # ---- consumers
            up = lds_v2(smem, tb)
            v = np.empty((C.R, C.WORKERS, 2))
            km = np.empty((C.R, C.WORKERS))
            for r in range(C.R):
                v[r] = lds_v2(smem, tb + (1 + r) * P)
                km[r] = lds_f64(smem, tb + (1 + r) * P - 8)
            for s in range(T):
                nv = np.empty_like(v)
                for r in range(C.R):
                    jm = up if r == 0 else v[r - 1]
                    nv[r, :, 0] = cell(v[r, :, 0], carry[s, r, :, 0], jm[:, 0], km[r])
                    nv[r, :, 1] = cell(v[r, :, 1], carry[s, r, :, 1], jm[:, 1], v[r, :, 0])
                carry[s] = v
                if s == T - 1:
                    if p >= i0:
                        for r in range(C.R):
                            ok = store_cols & rmask[r]
                            rows, cols = (j + r)[ok], k[ok]
                            assert np.all((rows >= 0) & (cols >= 0) & (cols + 1 < n2 + 1))
                            assert np.all(np.isnan(out[p, rows, cols])), "cell stored twice"
                            out[p, rows, cols] = nv[r, ok, 0]
                            out[p, rows, cols + 1] = nv[r, ok, 1]
                else:
                    xb = xbuf[xsel]
                    xsel ^= 1
                    xb[:] = np.nan
                    if split:
                        for r in range(C.R):
                            xb[(xs + (1 + r) * P + YOFF) // 8] = nv[r, :, 1]
                        xb[(xs + C.R * P) // 8] = nv[C.R - 1, :, 0]
                        up = np.stack([lds_f64(xb, xs), lds_f64(xb, xs + YOFF)], axis=-1)
                        for r in range(C.R):
                            km[r] = lds_f64(xb, xs + (1 + r) * P + YOFF - 8)
                    else:
                        for r in range(C.R):
                            ad = (xt + (1 + r) * P) // 8
                            xb[ad], xb[ad + 1] = nv[r, :, 0], nv[r, :, 1]
                        up = lds_v2(xb, xt)
                        for r in range(C.R):
                            km[r] = lds_f64(xb, xt + (1 + r) * P - 8)
                    v = nv
