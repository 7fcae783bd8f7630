// kernels_lapfused.cu -- temporal blocking for the 3-D 7-point stencil: TWO applies per sweep.
//
// The reference's Laplacian driver iterates  applyFilter(); copyOutToIn();  (ref:
// laplacian/cxx/laplacian.cxx:86-90).  One apply is pinned to the HBM roofline at 16 B per
// cell-apply (lap7_tma_kernel, kernels_tma.cu); fusing two applies halves the DRAM traffic.
// Unlike the upwind stencil this one looks both ways on every axis, so the pipeline along the
// marching axis runs two planes behind the input:
//
//   level-0 plane p lands in shared memory (TMA)            p = i0-2 .. i1+1
//     -> level 1 of plane p-1 is finished  (its (+1,0,0) branch is this plane's centre)
//     -> level 2 of plane p-2 is finished  (its (+1,0,0) branch is level 1 of plane p-1): STORE
//     -> first six branches of level 1 of plane p      (needs level 0 of planes p-1, p)
//     -> first six branches of level 2 of plane p-1    (needs level 1 of planes p-2, p-1)
//
//   * same TMA / mbarrier producer-consumer ring as the other tiled kernels;
//   * each consumer thread owns R rows x 2 cells of the CJ x CK level-1 tile (the BJ x BK output
//     tile plus one ring: rows j0-1..j0+BJ, columns k0-2..k0+BK+1 so that pairs stay 16-byte
//     aligned) and carries four planes in registers: level 0 of p-1, the level-1 partial of p,
//     level 1 of p-2 and the level-2 partial of p-1;
//   * in-plane neighbours of level 1 travel through two ping-pong exchange tiles in shared memory
//     with ONE named barrier among the consumer warps per plane;
//   * periodic wrap: two halo rows above and below the tile come from their own TMA boxes at
//     (j0-2) mod N1 and (j0+BJ) mod N1, the wrap columns of the first / last k-tile from 2-cell
//     boxes at N2-2 and 0; planes outside the slab from the ghost tensors (depth 2), which alias
//     the far planes on a single device.  The ring cells of level 1 are recomputed by the
//     neighbouring tiles from the same inputs in the same order, so they are the same bits.
//
// Arithmetic per level is exactly lap7_tma_kernel's: acc = 0; acc = acc + w*in[...] for the seven
// offsets in std::map order, multiply and add rounded separately (ref: Filter.cpp:202,247-251).
// Two fused applies are therefore bit-identical to two single applies.
#include "fdb_internal.h"
#include "tma_ptx.cuh"

namespace fdb {

namespace {

using namespace ptx;

constexpr int lf_align128(int x) { return (x + 127) / 128 * 128; }

// UNIT: the six off-centre weights are exactly 1.0 (the Laplacian, laplacian.cxx:55-65): their
// multiplies are skipped, which is exact (1.0 * v == v for every v), so the bits do not change.
template <int BJ_, int R_, int STAGES_, int BK_ = 128, int MINB_ = 1, bool SHFL_ = false>
struct LapFusedCfg {
  static constexpr int BJ = BJ_, BK = BK_, R = R_, STAGES = STAGES_, MINB = MINB_;
  // SHFL (experimental, not the default): the k-1 / k+1 neighbours come from the adjacent lanes' registers by
  // warp shuffle instead of 64-bit shared-memory loads at a 16-byte stride (4 wavefronts each, 60 % of the
  // kernel's shared-memory wavefronts, profiles/r01k_*); only a warp's edge lanes and a row's end threads load
  static constexpr bool SHFL = SHFL_;
  static constexpr int CJ = BJ + 2;       // level-1 rows:    global j0-1 .. j0+BJ
  static constexpr int CK = BK + 4;       // level-1 columns: global k0-2 .. k0+BK+1
  static constexpr int IN_ROWS = BJ + 4;  // level-0 rows:    global j0-2 .. j0+BJ+1 (stage row s)
  static constexpr int TX = CK / 2;       // threads per row (2 cells each)
  static constexpr int TY = CJ / R;
  static constexpr int WORKERS = TX * TY;
  static constexpr int CONSUMERS = (WORKERS + 31) / 32 * 32;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;  // + the loader warp
  // Stage layout: IN_ROWS rows of BKP = BK + 8 cells, global columns k0-4 .. k0+BK+3 (two
  // unused cells on each side keep every in-row neighbour inside the row and make two rows a
  // multiple of 128 bytes, so the three TMA boxes -- 2 halo rows, BJ tile rows, 2 halo rows --
  // land back to back with ONE pitch).  Thread tx owns byte offset 16 + 16*tx of a row.
  static constexpr int BKP = BK + 8;
  static constexpr int PITCH = BKP * 8;
  static constexpr int BODY_OFF = 2 * PITCH;
  static constexpr int BOT_OFF = (2 + BJ) * PITCH;
  static constexpr int MAIN_BYTES = IN_ROWS * PITCH;
  // wrap columns of the first / last k-tile: 8-cell boxes (pitch 64 B) at columns N2-8 and 0; the
  // loader warp copies the pair each row needs into the row's zero-filled halo cells
  static constexpr int WPITCH = 64;
  static constexpr int WL_OFF = lf_align128(MAIN_BYTES);
  static constexpr int WR_OFF = WL_OFF + IN_ROWS * WPITCH;
  static constexpr int STAGE_BYTES = WR_OFF + IN_ROWS * WPITCH;
  static constexpr int TX_MAIN = IN_ROWS * PITCH;
  static constexpr int TX_WRAP = IN_ROWS * WPITCH;
  // exchange tile (level 1): same pitch and column offsets as a stage, one spare row above and
  // below, so a thread addresses both with the same base: row q of level 1 sits at row q+1
  static constexpr int X_BYTES = lf_align128((CJ + 2) * PITCH);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * X_BYTES + 3 * STAGES * 8 + 128;
  static constexpr int LAG = STAGES - 2;  // planes the loader keeps in flight behind the hand-over
  static_assert(CJ % R == 0, "rows per thread must divide the level-1 tile");
  static_assert(BJ % 2 == 0 && (2 * PITCH) % 128 == 0 && (IN_ROWS * WPITCH) % 128 == 0, "misaligned TMA destination");
  static_assert(BKP <= 256 && BJ <= 256, "TMA box limit");
  static_assert(IN_ROWS <= 32, "one loader lane per stage row");
  static_assert(STAGES >= 3, "the loader needs two stages of slack");
  static_assert(THREADS <= 1024, "too many threads");
};

struct LapFusedArgs {
  double* out;          // local plane 0 of the output field (level 2)
  int64_t n1, n2, nloc;
  int64_t ibeg, iend;   // local output planes
  int ci, njt, nkt;
  int64_t nwork;
  int G;                // planes in each ghost tensor; local plane p < 0 is plane G + p of the low one
  double w[7];          // weights in application order (kernels_tma.cu: lap7_slot)
};

// Tensor maps: m[4*t + b], t = 0 local planes, 1 ghost planes below, 2 ghost planes above;
// box b = 0 {BKP, 2} halo rows, 1 {BKP, BJ} tile rows, 2 {8, 2} wrap corner, 3 {8, BJ} wrap columns
struct LapFusedMaps {
  CUtensorMap m[12];
};

// acc + w*v, rounded separately (ref: Filter.cpp:247-251); UNIT: w is exactly 1.0
template <bool UNIT>
__device__ __forceinline__ double lf_acc(double acc, double w, double v) {
  return UNIT ? __dadd_rn(acc, v) : __dadd_rn(acc, __dmul_rn(w, v));
}

// position of the loader in the sequence of (work item, plane) pairs of this CTA
template <class C>
struct LapFusedCursor {
  int64_t w;       // work item
  int64_t p, i1;   // plane, end of the chunk
  int kt, jt;
  __device__ __forceinline__ void open(const LapFusedArgs& a) {
    kt = (int)(w % a.nkt);
    jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    p = i0 - 2;
  }
  __device__ __forceinline__ bool valid(const LapFusedArgs& a) const { return w < a.nwork; }
  __device__ __forceinline__ void next(const LapFusedArgs& a) {
    if (++p > i1 + 1) {
      w += gridDim.x;
      if (w < a.nwork) open(a);
    }
  }
};

// ===================== loader warp (shared by both consumer formulations) =====================
// Lane 0 issues the TMA boxes of plane n; LAG planes behind, the warp waits for a plane to
// land, copies the periodic wrap columns of the first / last k-tile into the tile rows (one
// lane per row) and hands the stage to the consumers.  Issue waits for the consumers to free
// the stage of plane n - STAGES while planes up to n - STAGES + 1 are already handed over, so
// the two never wait for each other.
template <class C>
__device__ __forceinline__ void lapf_loader_warp(const LapFusedMaps& maps, const LapFusedArgs& a, uint32_t smem,
                                                 uint32_t landed, uint32_t full, uint32_t empty, int lane) {
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < 12; ++m) prefetch_tmap(&maps.m[m]);
    }
    LapFusedCursor<C> ci, cf;  // issue / hand-over
    ci.w = blockIdx.x;
    if (ci.valid(a)) ci.open(a);
    cf = ci;
    int si = 0, sf = 0;        // stage of the next issue / hand-over
    uint32_t phi = 0, phf = 0;
    int ahead = 0;             // planes issued and not yet handed over
    while (cf.valid(a)) {
      if (ci.valid(a)) {
        mbar_wait(empty + 8 * si, phi ^ 1);
        if (lane == 0) {
          const uint32_t st = smem + si * C::STAGE_BYTES;
          const uint32_t lb = landed + 8 * si;
          const int kb = ci.kt * C::BK - 4;  // first stage column (negative for the first k-tile: zero fill)
          const int j0 = ci.jt * C::BJ;
          const int jtop = (j0 == 0) ? (int)a.n1 - 2 : j0 - 2;          // periodic rows j0-2, j0-1
          const int jbot = (j0 + C::BJ >= (int)a.n1) ? 0 : j0 + C::BJ;  // periodic rows j0+BJ, j0+BJ+1
          const bool first_k = (ci.kt == 0), last_k = (ci.kt == a.nkt - 1);
          const int64_t p = ci.p;
          const int g = (p < 0) ? 4 : (p >= a.nloc ? 8 : 0);
          const int pl = (p < 0) ? a.G + (int)p : (p >= a.nloc ? (int)(p - a.nloc) : (int)p);
          mbar_expect_tx(lb, C::TX_MAIN + (first_k ? C::TX_WRAP : 0) + (last_k ? C::TX_WRAP : 0));
          tma_load_3d(st, &maps.m[g + 0], lb, kb, jtop, pl);
          tma_load_3d(st + C::BODY_OFF, &maps.m[g + 1], lb, kb, j0, pl);
          tma_load_3d(st + C::BOT_OFF, &maps.m[g + 0], lb, kb, jbot, pl);
          if (first_k) {
            const uint32_t wl = st + C::WL_OFF;
            tma_load_3d(wl, &maps.m[g + 2], lb, (int)a.n2 - 8, jtop, pl);
            tma_load_3d(wl + 2 * C::WPITCH, &maps.m[g + 3], lb, (int)a.n2 - 8, j0, pl);
            tma_load_3d(wl + (2 + C::BJ) * C::WPITCH, &maps.m[g + 2], lb, (int)a.n2 - 8, jbot, pl);
          }
          if (last_k) {
            const uint32_t wr = st + C::WR_OFF;
            tma_load_3d(wr, &maps.m[g + 2], lb, 0, jtop, pl);
            tma_load_3d(wr + 2 * C::WPITCH, &maps.m[g + 3], lb, 0, j0, pl);
            tma_load_3d(wr + (2 + C::BJ) * C::WPITCH, &maps.m[g + 2], lb, 0, jbot, pl);
          }
        }
        ci.next(a);
        if (++si == C::STAGES) { si = 0; phi ^= 1; }
        ++ahead;
      }
      if (ahead > C::LAG || !ci.valid(a)) {
        mbar_wait(landed + 8 * sf, phf);
        const uint32_t st = smem + sf * C::STAGE_BYTES;
        if (lane < C::IN_ROWS) {
          if (cf.kt == 0) {  // columns -2, -1 of the tile rows <- columns N2-2, N2-1
            const double2 v = lds_v2(st + C::WL_OFF + lane * C::WPITCH + 48);
            sts_v2(st + lane * C::PITCH + 16, v.x, v.y);
          }
          if (cf.kt == a.nkt - 1) {  // columns N2, N2+1 <- columns 0, 1
            const double2 v = lds_v2(st + C::WR_OFF + lane * C::WPITCH);
            sts_v2(st + lane * C::PITCH + C::CK * 8, v.x, v.y);
          }
        }
        fence_proxy_async_shared();  // the patched cells are rewritten by a later TMA box
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * sf);
        cf.next(a);
        if (++sf == C::STAGES) { sf = 0; phf ^= 1; }
        --ahead;
      }
    }
}

template <class C, bool UNIT>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    lap7_fused2_kernel(const __grid_constant__ LapFusedMaps maps, const LapFusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t landed = xbuf + 2 * C::X_BYTES;      // TMA bytes of the stage have arrived
  const uint32_t full = landed + C::STAGES * 8;       // ... and its wrap columns are in place
  const uint32_t empty = full + C::STAGES * 8;        // every consumer warp has read the stage

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(landed + 8 * s, 1);
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, C::CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == C::CONSUMER_WARPS) {
    lapf_loader_warp<C>(maps, a, smem, landed, full, empty, lane);
    return;
  }

  // ===================== consumer warps =====================
  const bool worker = tid < C::WORKERS;  // threads past the tile only keep the barriers company
  const int wid = worker ? tid : 0;
  const int tx = wid % C::TX;
  const int ty = wid / C::TX;
  const int q0 = ty * C::R;  // first level-1 row of this thread (level-1 row q = global row j0-1+q)
  int stage = 0;
  uint32_t phase = 0;
  uint32_t xsel = 0;
  const double w0 = a.w[0], w1 = a.w[1], w2 = a.w[2], w3 = a.w[3], w4 = a.w[4], w5 = a.w[5], w6 = a.w[6];
  // One base serves the stage (level 0: stage row s = level-1 row q + 1) and the exchange tile
  // (level 1: row q at tile row q + 1): the row above the thread's first row, its own pair.
  const uint32_t tb = q0 * C::PITCH + 16 + tx * 16;
  constexpr uint32_t P = C::PITCH;
  // rows of this thread that belong to the output tile (level-1 rows 1 .. CJ-2)
  uint32_t rowmask = 0;
#pragma unroll
  for (int r = 0; r < C::R; ++r)
    if (q0 + r >= 1 && q0 + r <= C::CJ - 2) rowmask |= 1u << r;
  const bool store_cols = worker && tx >= 1 && tx <= C::TX - 2;
  // SHFL: lanes whose left / right neighbour cell is not in the adjacent lane (warp edge, row end, and the
  // last worker, whose next lane is a padding thread)
  const bool edge_lo = lane == 0 || tx == 0;
  const bool edge_hi = lane == 31 || tx == C::TX - 1 || tid >= C::WORKERS - 1;
  const int64_t plane = a.n1 * a.n2;

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - 2 + 2 * tx;  // global column of this thread's first cell
    const int64_t j = (int64_t)jt * C::BJ - 1 + q0;      // global row of this thread's first row
    // offset of (plane p-2, row j, column k) in the output, advanced by one plane per iteration
    int64_t ooff = ((i0 - 5) * a.n1 + j) * a.n2 + k;

    double2 below0[C::R], part1[C::R], below1[C::R], part2[C::R];
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      below0[r] = make_double2(0.0, 0.0);
      part1[r] = make_double2(0.0, 0.0);
      below1[r] = make_double2(0.0, 0.0);
      part2[r] = make_double2(0.0, 0.0);
    }

    for (int64_t p = i0 - 2; p <= i1 + 1; ++p) {
      ooff += plane;
      mbar_wait(full + 8 * stage, phase);  // the loader's hand-over: TMA bytes landed, wrap columns patched
      const uint32_t sb = smem + stage * C::STAGE_BYTES + tb;
      double2 c[C::R];
      double km[C::R], kp[C::R];
      const double2 up = lds_v2(sb);
      const double2 dn = lds_v2(sb + (C::R + 1) * P);
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        c[r] = lds_v2(sb + (1 + r) * P);
        if (C::SHFL) {
          km[r] = __shfl_up_sync(0xffffffffu, c[r].y, 1);
          kp[r] = __shfl_down_sync(0xffffffffu, c[r].x, 1);
          if (edge_lo) km[r] = lds_f64(sb + (1 + r) * P - 8);
          if (edge_hi) kp[r] = lds_f64(sb + (1 + r) * P + 16);
        } else {
          km[r] = lds_f64(sb + (1 + r) * P - 8);
          kp[r] = lds_f64(sb + (1 + r) * P + 16);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }

      // level 1 of plane p-1: the (+1,0,0) branch is this plane's centre
      double2 l1[C::R];
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        l1[r].x = lf_acc<UNIT>(part1[r].x, w6, c[r].x);
        l1[r].y = lf_acc<UNIT>(part1[r].y, w6, c[r].y);
      }
      // level 2 of plane p-2: the (+1,0,0) branch is level 1 of plane p-1 -- done, store it
      if (p >= i0 + 2 && store_cols) {
        double* orow = a.out + ooff;
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          if ((rowmask >> r) & 1u)
            st_global_v2(orow + (int64_t)r * a.n2, lf_acc<UNIT>(part2[r].x, w6, l1[r].x),
                         lf_acc<UNIT>(part2[r].y, w6, l1[r].y));
        }
      }
      // hand level 1 of plane p-1 to the neighbours
      const uint32_t xb = xbuf + xsel * C::X_BYTES + tb;
      xsel ^= 1;
      if (worker) {
#pragma unroll
        for (int r = 0; r < C::R; ++r) sts_v2(xb + (1 + r) * P, l1[r].x, l1[r].y);
      }
      // first six branches of level 1 of plane p (keeps the FP64 pipe busy while warps gather)
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const double2 jm = (r == 0) ? up : c[r - 1];
        const double2 jp = (r == C::R - 1) ? dn : c[r + 1];
        double x = 0.0, y = 0.0;
        x = lf_acc<UNIT>(x, w0, below0[r].x);  y = lf_acc<UNIT>(y, w0, below0[r].y);
        x = lf_acc<UNIT>(x, w1, jm.x);         y = lf_acc<UNIT>(y, w1, jm.y);
        x = lf_acc<UNIT>(x, w2, km[r]);        y = lf_acc<UNIT>(y, w2, c[r].x);
        x = lf_acc<false>(x, w3, c[r].x);      y = lf_acc<false>(y, w3, c[r].y);
        x = lf_acc<UNIT>(x, w4, c[r].y);       y = lf_acc<UNIT>(y, w4, kp[r]);
        x = lf_acc<UNIT>(x, w5, jp.x);         y = lf_acc<UNIT>(y, w5, jp.y);
        part1[r] = make_double2(x, y);
        below0[r] = c[r];
      }
      named_bar_sync(1, C::CONSUMERS);
      // first six branches of level 2 of plane p-1 (rows outside the tile read the spare rows of
      // the exchange tile: whatever is there only reaches level-2 rows that are never stored)
      {
        const double2 up1 = lds_v2(xb);
        const double2 dn1 = lds_v2(xb + (C::R + 1) * P);
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          if (C::SHFL) {
            km[r] = __shfl_up_sync(0xffffffffu, l1[r].y, 1);
            kp[r] = __shfl_down_sync(0xffffffffu, l1[r].x, 1);
            if (edge_lo) km[r] = lds_f64(xb + (1 + r) * P - 8);
            if (edge_hi) kp[r] = lds_f64(xb + (1 + r) * P + 16);
          } else {
            km[r] = lds_f64(xb + (1 + r) * P - 8);
            kp[r] = lds_f64(xb + (1 + r) * P + 16);
          }
        }
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const double2 jm = (r == 0) ? up1 : l1[r - 1];
          const double2 jp = (r == C::R - 1) ? dn1 : l1[r + 1];
          double x = 0.0, y = 0.0;
          x = lf_acc<UNIT>(x, w0, below1[r].x);  y = lf_acc<UNIT>(y, w0, below1[r].y);
          x = lf_acc<UNIT>(x, w1, jm.x);         y = lf_acc<UNIT>(y, w1, jm.y);
          x = lf_acc<UNIT>(x, w2, km[r]);        y = lf_acc<UNIT>(y, w2, l1[r].x);
          x = lf_acc<false>(x, w3, l1[r].x);     y = lf_acc<false>(y, w3, l1[r].y);
          x = lf_acc<UNIT>(x, w4, l1[r].y);      y = lf_acc<UNIT>(y, w4, kp[r]);
          x = lf_acc<UNIT>(x, w5, jp.x);         y = lf_acc<UNIT>(y, w5, jp.y);
          part2[r] = make_double2(x, y);
          below1[r] = l1[r];
        }
      }
    }
  }
}

// ---- lean consumer formulation (round 2) ------------------------------------------------------------------------
// Same tile pipeline and the same arithmetic as lap7_fused2_kernel above, restated after the fused upwind kernel's
// findings on B200 (profiles/r02e_*, r02j_*): FP64 instructions hold the issue port for two cycles and nothing
// co-issues, and the shared-memory pipe is the second limiter.  So
//   * the plane loop is unrolled by two and the register sets swap roles (level 0 of plane p / p-1, level 1 of
//     plane p-1 / p-2): none of the 48 register moves per plane;
//   * the exchange tile keeps the even cells (x) and the odd cells (y) of a row in two contiguous halves, so the
//     k-1 / k+1 operands are conflict-free 8-byte loads (2 wavefronts instead of the 4 of a 16-byte lane stride);
//   * 32-bit plane counters, output rows as pointers that advance by a plane, idle threads duplicate thread 0.
template <class C, bool UNIT>
struct LapLean {
  static constexpr int R = C::R;
  static constexpr uint32_t P = C::PITCH;
  static constexpr uint32_t YOFF = C::PITCH / 2;  // y half of an exchange-tile row
  static_assert((C::TX + 2) * 8 <= C::PITCH / 2, "exchange row halves too narrow");

  struct State {
    uint32_t st, bar, par, st0, bar0, bar_end;
    uint32_t xt0, xt1;     // exchange tiles: row above the thread's first row, x half, this thread's slot
    double* orow[C::R];    // output rows of plane p-2
    int64_t plane_elems;
    uint32_t smask;
    int lane;
  };

  // one plane: level-0 plane p is in the stage; Cp = level 0 of plane p-1, Cc <- level 0 of plane p;
  // Lp = level 1 of plane p-2, Lc <- level 1 of plane p-1; part1 / part2 = the six-branch partial sums
  template <int PARITY>
  static __device__ __forceinline__ void step(State& z, const LapFusedArgs& a, double2 (&Cp)[C::R], double2 (&Cc)[C::R],
                                              double2 (&Lp)[C::R], double2 (&Lc)[C::R], double2 (&part1)[C::R],
                                              double2 (&part2)[C::R], bool store) {
    const double w0 = a.w[0], w1 = a.w[1], w2 = a.w[2], w3 = a.w[3], w4 = a.w[4], w5 = a.w[5], w6 = a.w[6];
    // the loader warp's hand-over: it waited for the TMA bytes (acquire on `landed`), patched the wrap columns and
    // arrived on `full` (release); this acquire orders all of it before the loads below (see kernels_fused.cu)
    mbar_wait(z.bar, z.par);
    double km[R], kp[R];
    const double2 up = lds_v2(z.st);
    const double2 dn = lds_v2(z.st + (R + 1) * P);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      Cc[r] = lds_v2(z.st + (1 + r) * P);
      km[r] = lds_f64(z.st + (1 + r) * P - 8);
      kp[r] = lds_f64(z.st + (1 + r) * P + 16);
    }
    __syncwarp();
    if (z.lane == 0) mbar_arrive(z.bar + 8 * C::STAGES);
    z.st += C::STAGE_BYTES;
    z.bar += 8;
    if (z.bar == z.bar_end) { z.st = z.st0; z.bar = z.bar0; z.par ^= 1; }

    // level 1 of plane p-1: the (+1,0,0) branch is this plane's centre
#pragma unroll
    for (int r = 0; r < R; ++r) {
      Lc[r].x = lf_acc<UNIT>(part1[r].x, w6, Cc[r].x);
      Lc[r].y = lf_acc<UNIT>(part1[r].y, w6, Cc[r].y);
    }
    // level 2 of plane p-2: the (+1,0,0) branch is level 1 of plane p-1 -- done, store it
    const uint32_t m = store ? z.smask : 0u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((m >> r) & 1u)
        st_global_v2(z.orow[r], lf_acc<UNIT>(part2[r].x, w6, Lc[r].x), lf_acc<UNIT>(part2[r].y, w6, Lc[r].y));
      z.orow[r] += z.plane_elems;
    }
    // hand level 1 of plane p-1 to the neighbours (both halves: the left neighbour reads x, the right one y)
    const uint32_t xb = (PARITY == 0) ? z.xt0 : z.xt1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      sts_f64(xb + (1 + r) * P, Lc[r].x);
      sts_f64(xb + (1 + r) * P + YOFF, Lc[r].y);
    }
    // first six branches of level 1 of plane p (keeps the FP64 pipe busy while warps gather)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double2 jm = (r == 0) ? up : Cc[r - 1];
      const double2 jp = (r == R - 1) ? dn : Cc[r + 1];
      double x = 0.0, y = 0.0;
      x = lf_acc<UNIT>(x, w0, Cp[r].x);    y = lf_acc<UNIT>(y, w0, Cp[r].y);
      x = lf_acc<UNIT>(x, w1, jm.x);       y = lf_acc<UNIT>(y, w1, jm.y);
      x = lf_acc<UNIT>(x, w2, km[r]);      y = lf_acc<UNIT>(y, w2, Cc[r].x);
      x = lf_acc<false>(x, w3, Cc[r].x);   y = lf_acc<false>(y, w3, Cc[r].y);
      x = lf_acc<UNIT>(x, w4, Cc[r].y);    y = lf_acc<UNIT>(y, w4, kp[r]);
      x = lf_acc<UNIT>(x, w5, jp.x);       y = lf_acc<UNIT>(y, w5, jp.y);
      part1[r] = make_double2(x, y);
    }
    named_bar_sync(1, C::CONSUMERS);
    // first six branches of level 2 of plane p-1 (rows outside the tile read the spare rows of the exchange tile:
    // whatever is there only reaches level-2 rows that are never stored)
    double2 up1, dn1;
    up1.x = lds_f64(xb);
    up1.y = lds_f64(xb + YOFF);
    dn1.x = lds_f64(xb + (R + 1) * P);
    dn1.y = lds_f64(xb + (R + 1) * P + YOFF);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      km[r] = lds_f64(xb + (1 + r) * P + YOFF - 8);  // y of the thread to the left
      kp[r] = lds_f64(xb + (1 + r) * P + 8);         // x of the thread to the right
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double2 jm = (r == 0) ? up1 : Lc[r - 1];
      const double2 jp = (r == R - 1) ? dn1 : Lc[r + 1];
      double x = 0.0, y = 0.0;
      x = lf_acc<UNIT>(x, w0, Lp[r].x);    y = lf_acc<UNIT>(y, w0, Lp[r].y);
      x = lf_acc<UNIT>(x, w1, jm.x);       y = lf_acc<UNIT>(y, w1, jm.y);
      x = lf_acc<UNIT>(x, w2, km[r]);      y = lf_acc<UNIT>(y, w2, Lc[r].x);
      x = lf_acc<false>(x, w3, Lc[r].x);   y = lf_acc<false>(y, w3, Lc[r].y);
      x = lf_acc<UNIT>(x, w4, Lc[r].y);    y = lf_acc<UNIT>(y, w4, kp[r]);
      x = lf_acc<UNIT>(x, w5, jp.x);       y = lf_acc<UNIT>(y, w5, jp.y);
      part2[r] = make_double2(x, y);
    }
  }
};

template <class C, bool UNIT>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    lap7_fused2_lean_kernel(const __grid_constant__ LapFusedMaps maps, const LapFusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t landed = xbuf + 2 * C::X_BYTES;
  const uint32_t full = landed + C::STAGES * 8;
  const uint32_t empty = full + C::STAGES * 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(landed + 8 * s, 1);
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, C::CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == C::CONSUMER_WARPS) {
    lapf_loader_warp<C>(maps, a, smem, landed, full, empty, lane);
    return;
  }

  using L = LapLean<C, UNIT>;
  const bool worker = tid < C::WORKERS;
  const int wid = worker ? tid : 0;  // threads past the tile repeat thread 0's work and store nothing
  const int tx = wid % C::TX;
  const int ty = wid / C::TX;
  const int q0 = ty * C::R;
  typename L::State z;
  z.st0 = smem + q0 * C::PITCH + 16 + tx * 16;
  z.bar0 = full;
  z.bar_end = full + 8 * C::STAGES;
  z.st = z.st0;
  z.bar = z.bar0;
  z.par = 0;
  z.xt0 = xbuf + q0 * C::PITCH + 8 + tx * 8;
  z.xt1 = z.xt0 + C::X_BYTES;
  z.plane_elems = a.n1 * a.n2;
  z.lane = lane;
  uint32_t rowmask = 0;
#pragma unroll
  for (int r = 0; r < C::R; ++r)
    if (worker && tx >= 1 && tx <= C::TX - 2 && q0 + r >= 1 && q0 + r <= C::CJ - 2) rowmask |= 1u << r;
  z.smask = rowmask;

  double2 C0[C::R], C1[C::R], L0[C::R], L1[C::R], part1[C::R], part2[C::R];
#pragma unroll
  for (int r = 0; r < C::R; ++r)
    C0[r] = C1[r] = L0[r] = L1[r] = part1[r] = part2[r] = make_double2(0.0, 0.0);

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - 2 + 2 * tx;
    const int64_t j = (int64_t)jt * C::BJ - 1 + q0;
    // rows of plane i0 - 4: the step that loads plane p stores plane p - 2, and the first step loads plane i0 - 2
#pragma unroll
    for (int r = 0; r < C::R; ++r) z.orow[r] = a.out + ((i0 - 4) * a.n1 + j + r) * a.n2 + k;
    // planes i0-2 .. i1+1; the step q (from 0) stores plane i0 - 4 + q, valid from q = 4 on.  The register sets are
    // not reset between items: a level is only read once every plane it depends on has been loaded by this item.
    const int np = (int)(i1 - i0) + 4;
    int q = 0;
    for (; q + 2 <= np; q += 2) {
      L::template step<0>(z, a, C0, C1, L0, L1, part1, part2, q >= 4);
      L::template step<1>(z, a, C1, C0, L1, L0, part1, part2, q + 1 >= 4);
    }
    if (q < np) {
      L::template step<0>(z, a, C0, C1, L0, L1, part1, part2, q >= 4);
      // odd plane count: the sets swap by value (once per item), and the next item's first plane must not write
      // the exchange tile this plane's neighbours may still be reading
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        double2 t = C0[r]; C0[r] = C1[r]; C1[r] = t;
        t = L0[r]; L0[r] = L1[r]; L1[r] = t;
      }
      named_bar_sync(1, C::CONSUMERS);
    }
  }
}

// ---- configurations ---------------------------------------------------------------------
typedef void (*LapFusedKernel)(const LapFusedMaps, const LapFusedArgs);
struct LapFusedConfig {
  int BJ, BK, BKP, threads, smem;
  LapFusedKernel kernel_unit, kernel_general;  // off-centre weights all exactly 1.0 / any weights
  LapFusedKernel lean_unit, lean_general;      // lean consumer formulation (FDB_LAPF_IMPL=2)
  const char* name;
};
template <class C>
constexpr LapFusedConfig make_lapf(const char* name) {
  return LapFusedConfig{C::BJ, C::BK, C::BKP, C::THREADS, C::SMEM_BYTES, lap7_fused2_kernel<C, true>,
                        lap7_fused2_kernel<C, false>, lap7_fused2_lean_kernel<C, true>, lap7_fused2_lean_kernel<C, false>,
                        name};
}
// kDefaultLapFused is the default (round-1 sweeps on a B200, profiles/r01j_*: six rows per thread,
// 7 consumer warps, 242 registers -- 639 GCUPS at 1024^3, 611 at 512^3); the rest are tuning
// alternatives (env FDB_LAPF_CFG)
const LapFusedConfig kLapFused[] = {
    make_lapf<LapFusedCfg<16, 3, 4>>("bj16_r3_s4"),
    make_lapf<LapFusedCfg<16, 3, 3>>("bj16_r3_s3"),
    make_lapf<LapFusedCfg<16, 6, 4>>("bj16_r6_s4"),
    make_lapf<LapFusedCfg<16, 2, 4>>("bj16_r2_s4"),
    make_lapf<LapFusedCfg<8, 5, 4, 128, 2>>("bj8_r5_s4_2cta"),
    make_lapf<LapFusedCfg<8, 2, 6>>("bj8_r2_s6"),
    make_lapf<LapFusedCfg<16, 3, 6>>("bj16_r3_s6"),
    make_lapf<LapFusedCfg<16, 6, 6>>("bj16_r6_s6"),
    make_lapf<LapFusedCfg<16, 3, 4, 64, 2>>("bj16_r3_s4_bk64_2cta"),
    make_lapf<LapFusedCfg<16, 3, 6, 64, 2>>("bj16_r3_s6_bk64_2cta"),
    make_lapf<LapFusedCfg<16, 2, 4, 64, 2>>("bj16_r2_s4_bk64_2cta"),
    make_lapf<LapFusedCfg<8, 5, 6>>("bj8_r5_s6"),
    // experimental (compiled, pinned by the host model, not yet timed on a B200): shuffled k-neighbours
    make_lapf<LapFusedCfg<16, 6, 6, 128, 1, true>>("bj16_r6_s6_shfl"),
    make_lapf<LapFusedCfg<16, 3, 4, 128, 1, true>>("bj16_r3_s4_shfl"),
};
constexpr int kNumLapFused = sizeof(kLapFused) / sizeof(kLapFused[0]);
constexpr int kDefaultLapFused = 7;  // bj16_r6_s6
// measured on B200 (profiles/r02k_*): both formulations run 652 GCUPS at 1024^3 in every tile configuration that used to
// differ (580..653 before) -- the kernel sits at 82 % of the measured copy bandwidth in DRAM traffic; the lean one
// issues 20 % fewer instructions for it
constexpr int kDefaultLapFusedImpl = 2;

int lf_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// the first configuration (from the forced or default one on) whose tile divides the plane
const LapFusedConfig* lapf_pick(const Field& f) {
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  int first = lf_env_int("FDB_LAPF_CFG", kDefaultLapFused);
  if (first < 0 || first >= kNumLapFused) first = kDefaultLapFused;
  for (int t = 0; t < kNumLapFused; ++t) {
    const LapFusedConfig& C = kLapFused[(first + t) % kNumLapFused];
    if (n1 % C.BJ == 0 && n2 % C.BK == 0) return &C;
  }
  return nullptr;
}

struct LapFusedAttr {
  LapFusedKernel fn = nullptr;
  int ctas_per_sm = 1;
  int sms = 148;
};
LapFusedAttr g_lapf_attr[16];

int lapf_slot(const int* o) {
  static const int order[7][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
  for (int b = 0; b < 7; ++b)
    if (o[0] == order[b][0] && o[1] == order[b][1] && o[2] == order[b][2]) return b;
  return -1;
}

}  // namespace

// Two applies per sweep: the full 3-D 7-point offset set (any weights), ghost depth >= 2, slabs of
// at least two planes and a plane the tile divides.
bool stencil_lap7_fused_supported(const Field& f, const StencilBranches& b) {
  if (f.geo.ndims != 3 || f.G < 2) return false;
  if (b.nbranch != 7) return false;
  int seen = 0;
  for (int i = 0; i < 7; ++i) {
    const int slot = lapf_slot(b.off[i]);
    if (slot < 0 || (seen >> slot) & 1) return false;
    seen |= 1 << slot;
  }
  for (const Slab& s : f.slabs)
    if (s.nloc() < 2) return false;
  return lapf_pick(f) != nullptr;
}

static int lapf_maps(const Field& f, int d, const LapFusedConfig& C, int p, LapFusedMaps* out) {
  const Slab& s = f.slabs[d];
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  const double* base[3] = {f.body(d, p), f.ghost_lo(d, p), f.ghost_hi(d, p)};
  const int64_t planes[3] = {s.nloc(), f.G, f.G};
  const int boxes[4][2] = {{C.BKP, 2}, {C.BKP, C.BJ}, {8, 2}, {8, C.BJ}};
  for (int t = 0; t < 3; ++t)
    for (int b = 0; b < 4; ++b)
      FDB_TRY(encode_tensor_map_3d(&out->m[4 * t + b], base[t], n2, n1, planes[t], boxes[b][0], boxes[b][1]));
  return FDB_OK;
}

// one-time, per-device set-up (function attributes, occupancy, tensor maps); see SweepLauncher::prepare
static int lapf_prepare(Field& f, int d, const StencilBranches& b, const LapFusedConfig** Cout, LapFusedKernel* fnout,
                        LapFusedAttr** atout) {
  Slab& sl = f.slabs[d];
  const LapFusedConfig* C = lapf_pick(f);
  if (!C) return set_error(FDB_E_INVALID, "no fused 7-point tile divides a %lld x %lld plane",
                           (long long)f.geo.n[1], (long long)f.geo.n[2]);
  // the Laplacian's six unit weights need no multiply (exact); FDB_LAPF_GENERAL=1 keeps them anyway
  bool unit = lf_env_int("FDB_LAPF_GENERAL", 0) == 0;
  for (int i = 0; i < 7; ++i)
    if (lapf_slot(b.off[i]) != 3 && b.w[i] != 1.0) unit = false;
  const int impl = lf_env_int("FDB_LAPF_IMPL", kDefaultLapFusedImpl);  // 1 = first consumer formulation, 2 = lean
  const LapFusedKernel fn = (impl == 2) ? (unit ? C->lean_unit : C->lean_general) : (unit ? C->kernel_unit : C->kernel_general);
  LapFusedAttr& at = g_lapf_attr[sl.device & 15];
  if (at.fn != fn) {
    FDB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, C->smem));
    int nb = 0;
    FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, C->threads, C->smem));
    if (nb < 1) return set_error(FDB_E_CUDA, "fused 7-point kernel %s does not fit on an SM", C->name);
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, sl.device));
    at.ctas_per_sm = nb;
    at.sms = prop.multiProcessorCount;
    at.fn = fn;
  }
  if (sl.lapf_cfg != (const void*)C) {
    for (int p = 0; p < 2; ++p) FDB_TRY(lapf_maps(f, d, *C, p, reinterpret_cast<LapFusedMaps*>(sl.lapf_maps[p])));
    sl.lapf_cfg = (const void*)C;
  }
  if (Cout) *Cout = C;
  if (fnout) *fnout = fn;
  if (atout) *atout = &at;
  return FDB_OK;
}

int stencil_lap7_fused_prepare(Field& f, int d, const StencilBranches& b) { return lapf_prepare(f, d, b, nullptr, nullptr, nullptr); }

int launch_stencil_lap7_fused(Field& f, int d, int X, int64_t ibeg, int64_t iend, const StencilBranches& b,
                              cudaStream_t s) {
  if (iend <= ibeg) return FDB_OK;
  const LapFusedConfig* C = nullptr;
  LapFusedKernel fn = nullptr;
  LapFusedAttr* atp = nullptr;
  FDB_TRY(lapf_prepare(f, d, b, &C, &fn, &atp));
  LapFusedAttr& at = *atp;
  Slab& sl = f.slabs[d];
  LapFusedArgs a;
  a.out = f.body(d, 1 - X);
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.nloc = sl.nloc();
  a.ibeg = ibeg;
  a.iend = iend;
  a.njt = (int)(a.n1 / C->BJ);
  a.nkt = (int)(a.n2 / C->BK);
  a.G = f.G;
  for (int i = 0; i < 7; ++i) a.w[lapf_slot(b.off[i])] = b.w[i];
  const int64_t grid_max = (int64_t)at.ctas_per_sm * at.sms;
  const int64_t tiles = (int64_t)a.njt * a.nkt;
  const int64_t planes = iend - ibeg;
  int64_t ci = lf_env_int("FDB_TMA_CI", 0);
  if (ci <= 0) {
    // Static round-robin: the sweep takes ceil(items / CTAs) rounds of (chunk + 4 warm-up planes)
    // plane-steps.  Long chunks amortise the warm-up, short ones fill the last round: take the
    // cheapest of a few lengths (measured: 64 beats 128 at 512^3, 128 beats 64 at 1024^3).
    static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 24, 16, 8};
    int64_t best_cost = -1;
    for (int c : cand) {
      const int64_t cc = c < planes ? c : planes;
      const int64_t items = tiles * ((planes + cc - 1) / cc);
      const int64_t cost = ((items + grid_max - 1) / grid_max) * (cc + 4);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; ci = cc; }
    }
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nwork = tiles * ((planes + ci - 1) / ci);
  int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  // FDB_MAX_CTAS (tests): fewer CTAs than the device holds, so every CTA walks many work items
  if (const int cap = lf_env_int("FDB_MAX_CTAS", 0); cap > 0 && grid > cap) grid = cap;
  fn<<<(unsigned)grid, C->threads, C->smem, s>>>(*reinterpret_cast<const LapFusedMaps*>(sl.lapf_maps[X]), a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

const char* stencil_lap7_fused_kernel_name() {
  return lf_env_int("FDB_LAPF_IMPL", kDefaultLapFusedImpl) == 2 ? "lap7_fused2_lean_kernel" : "lap7_fused2_kernel";
}

const char* stencil_lap7_fused_name(const Field& f) {
  const LapFusedConfig* C = lapf_pick(f);
  return C ? C->name : "";
}

}  // namespace fdb
