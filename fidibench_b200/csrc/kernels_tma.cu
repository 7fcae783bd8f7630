// kernels_tma.cu -- the sm_100a hot kernels: TMA-staged shared-memory tile pipelines
// for the 3-D upwind step and the 3-D 7-point Laplacian apply.
//
// Shape of both kernels (HBM-bound FP64 stencils, no tensor cores):
//   * persistent CTAs, grid = resident CTAs per SM x SM count, static round-robin
//     over work items (i-chunk, j-tile, k-tile);
//   * one producer warp: an elected lane streams the tile of each i-plane into a
//     ring of shared-memory stages with cp.async.bulk.tensor (TMA) and arms the
//     stage's "full" mbarrier with the expected byte count;
//   * consumer warps: wait on "full", pull the plane's tile into registers with
//     128-bit LDS, release the stage ("empty" mbarrier), then compute and store
//     coalesced 128-bit rows.  Planes i-1 (and i+1 for the Laplacian) are carried
//     in registers while the CTA marches along axis 0, so every cell of the field
//     crosses L2->SM once (plus tile halos);
//   * periodic wrap: rows/columns that fall off the tile grid are fetched by
//     separate small TMA boxes from the far side of the domain (TMA's own
//     out-of-bounds fill is zero, not periodic); planes below/above the slab come
//     from the ghost tensors (which alias the far planes on a single device).
//
// Arithmetic contract: as kernels_generic.cu (separately rounded mul/add, reference
// order), so this path is bit-identical to the generic one and to the oracle.
#include "fdb_internal.h"
#include "tma_ptx.cuh"

namespace fdb {

namespace {

using namespace ptx;

// ---- upwind tile configuration ------------------------------------------------
// Tile = BJ rows x BK cells of one i-plane; each consumer thread owns R
// consecutive rows x 2 consecutive cells (one double2 per row).
template <int BJ_, int BK_, int R_, int STAGES_>
struct UpwindCfg {
  static constexpr int BJ = BJ_, BK = BK_, R = R_, STAGES = STAGES_;
  static constexpr int BKH = BK + 2;              // row pitch in doubles: 2 halo cells + BK
  static constexpr int TX = BK / 2;               // threads along k
  static constexpr int TY = BJ / R;               // thread rows
  static constexpr int CONSUMERS = TX * TY;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;  // + producer warp
  static constexpr int ROW_BYTES = BKH * 8;
  static constexpr int HALO_BYTES = (ROW_BYTES + 127) / 128 * 128;  // slot of row j0-1
  static constexpr int BODY_BYTES = BJ * ROW_BYTES;                 // rows j0..j0+BJ-1
  static constexpr int WRAP_BYTES = BJ * 16;                        // cells N2-2,N2-1 of each row
  static constexpr int BODY_OFF = HALO_BYTES;
  static constexpr int WRAP_OFF = (BODY_OFF + BODY_BYTES + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = (WRAP_OFF + WRAP_BYTES + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
  static_assert(BJ % R == 0 && BK % 2 == 0 && CONSUMERS % 32 == 0, "bad tile");
  static_assert(BODY_BYTES % 128 == 0, "TMA destination must stay 128-byte aligned");
  static_assert(BKH <= 256 && BJ <= 256, "TMA box limit");
};

struct UpwindTmaArgs {
  double* out;          // local plane 0 of the output field
  int64_t n1, n2;       // plane extents
  int64_t ibeg, iend;   // local planes to compute
  int ci;               // planes per work item
  int njt, nkt;         // tiles per plane
  int64_t nwork;        // work items
  int G;                // ghost depth of the ghost tensor (its last plane is local plane -1)
  double c0, c1, c2;
  double* peer_out;     // next slab's ghost planes (peer-mapped) for planes >= peer_from; null = off
  int64_t peer_from;
};

// Tensor maps of the input field, by box shape:
//   tm_body : local planes,  box {BKH, BJ, 1}   (tile rows)
//   tm_row  : local planes,  box {BKH, 1, 1}    (halo row j0-1, wrapped)
//   tm_col  : local planes,  box {2, BJ, 1}     (cells N2-2..N2-1: wrap of k = -1)
//   tm_glo  : ghost planes below local plane 0, box {BKH, BJ, 1}
template <class C>
__global__ void __launch_bounds__(C::THREADS)
    upwind3d_tma_kernel(const __grid_constant__ CUtensorMap tm_body,
                        const __grid_constant__ CUtensorMap tm_row,
                        const __grid_constant__ CUtensorMap tm_col,
                        const __grid_constant__ CUtensorMap tm_glo, const UpwindTmaArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to 128 by hand
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full = smem + C::STAGES * C::STAGE_BYTES;  // STAGES x 8-byte mbarriers
  const uint32_t empty = full + C::STAGES * 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, C::CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == C::CONSUMER_WARPS) {
    // ===================== producer warp =====================
    if ((tid & 31) == 0) {
      prefetch_tmap(&tm_body);
      prefetch_tmap(&tm_row);
      prefetch_tmap(&tm_col);
      prefetch_tmap(&tm_glo);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        const int kt = (int)(w % a.nkt);
        const int jt = (int)((w / a.nkt) % a.njt);
        const int64_t ic = w / ((int64_t)a.nkt * a.njt);
        const int64_t i0 = a.ibeg + ic * a.ci;
        const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
        const int k0 = kt * C::BK - 2;  // box starts 2 cells left of the tile (16 B aligned)
        const int j0 = jt * C::BJ;
        const int jm = (j0 == 0) ? (int)a.n1 - 1 : j0 - 1;  // periodic row j0-1
        for (int64_t i = i0 - 1; i < i1; ++i) {
          mbar_wait(empty + 8 * stage, phase ^ 1);
          const uint32_t st = smem + stage * C::STAGE_BYTES;
          const uint32_t fb = full + 8 * stage;
          if (i == i0 - 1) {
            // plane below the chunk: only its tile rows are needed (register carry)
            mbar_expect_tx(fb, C::BODY_BYTES);
            if (i < 0)
              tma_load_3d(st + C::BODY_OFF, &tm_glo, fb, k0, j0, a.G - 1);
            else
              tma_load_3d(st + C::BODY_OFF, &tm_body, fb, k0, j0, (int)i);
          } else {
            const uint32_t bytes = C::BODY_BYTES + C::ROW_BYTES + (kt == 0 ? C::WRAP_BYTES : 0);
            mbar_expect_tx(fb, bytes);
            tma_load_3d(st + C::BODY_OFF, &tm_body, fb, k0, j0, (int)i);
            tma_load_3d(st, &tm_row, fb, k0, jm, (int)i);
            if (kt == 0) tma_load_3d(st + C::WRAP_OFF, &tm_col, fb, (int)a.n2 - 2, j0, (int)i);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int tx = tid % C::TX;
  const int ty = tid / C::TX;
  const int r0 = ty * C::R;
  const int lane = tid & 31;
  int stage = 0;
  uint32_t phase = 0;
  const double c0 = a.c0, c1 = a.c1, c2 = a.c2;

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK + 2 * tx;
    const int64_t j = (int64_t)jt * C::BJ + r0;
    const bool k_ok = k < a.n2;
    const bool wrap_lane = (kt == 0) && (tx == 0);

    double2 below[C::R];  // plane i-1, this thread's cells
    {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
#pragma unroll
      for (int r = 0; r < C::R; ++r)
        below[r] = lds_v2(st + C::BODY_OFF + (r0 + r) * C::ROW_BYTES + (2 + 2 * tx) * 8);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }

    for (int64_t i = i0; i < i1; ++i) {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
      double2 ctr[C::R];
      double km1[C::R];
      // row j-1 of the thread's first row: the halo slot for the tile's first row
      const uint32_t up_row = (r0 == 0) ? st : st + C::BODY_OFF + (r0 - 1) * C::ROW_BYTES;
      const double2 up = lds_v2(up_row + (2 + 2 * tx) * 8);
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const uint32_t row = st + C::BODY_OFF + (r0 + r) * C::ROW_BYTES;
        ctr[r] = lds_v2(row + (2 + 2 * tx) * 8);
        km1[r] = lds_f64(wrap_lane ? st + C::WRAP_OFF + (r0 + r) * 16 + 8 : row + (1 + 2 * tx) * 8);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }

      double* orow = a.out + (i * a.n1 + j) * a.n2 + k;
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const double2 jm = (r == 0) ? up : ctr[r - 1];
        // ref: upwind.cxx:72-80, axes 0,1,2 in order, no FMA
        double x = ctr[r].x;
        x = __dsub_rn(x, __dmul_rn(c0, __dsub_rn(below[r].x, ctr[r].x)));
        x = __dsub_rn(x, __dmul_rn(c1, __dsub_rn(jm.x, ctr[r].x)));
        x = __dsub_rn(x, __dmul_rn(c2, __dsub_rn(km1[r], ctr[r].x)));
        double y = ctr[r].y;
        y = __dsub_rn(y, __dmul_rn(c0, __dsub_rn(below[r].y, ctr[r].y)));
        y = __dsub_rn(y, __dmul_rn(c1, __dsub_rn(jm.y, ctr[r].y)));
        y = __dsub_rn(y, __dmul_rn(c2, __dsub_rn(ctr[r].x, ctr[r].y)));
        if (k_ok && (j + r) < a.n1) {
          st_global_v2(orow + (int64_t)r * a.n2, x, y);
          if (a.peer_out != nullptr && i >= a.peer_from)
            st_global_v2(a.peer_out + (((i - a.peer_from) * a.n1 + j + r) * a.n2 + k), x, y);
        }
        below[r] = ctr[r];
      }
    }
  }
}


// ---- 7-point (3-D, radius 1, axis-aligned) stencil tile configuration -----------------
// Tile = BJ rows x BK cells; smem rows carry 2 halo cells on each side (16-byte aligned
// boxes): pitch BK + 4.  Two extra row slots hold rows j0-1 and j0+BJ (periodic).
template <int BJ_, int BK_, int R_, int STAGES_>
struct Lap7Cfg {
  static constexpr int BJ = BJ_, BK = BK_, R = R_, STAGES = STAGES_;
  static constexpr int BKH = BK + 4;
  static constexpr int TX = BK / 2;
  static constexpr int TY = BJ / R;
  static constexpr int CONSUMERS = TX * TY;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;
  static constexpr int ROW_BYTES = BKH * 8;
  static constexpr int ROW_SLOT = (ROW_BYTES + 127) / 128 * 128;
  static constexpr int BODY_BYTES = BJ * ROW_BYTES;
  static constexpr int WRAP_BYTES = BJ * 16;
  static constexpr int TOP_OFF = 0;                       // row j0-1
  static constexpr int BODY_OFF = ROW_SLOT;               // rows j0..j0+BJ-1
  static constexpr int BOT_OFF = BODY_OFF + BODY_BYTES;   // row j0+BJ
  static constexpr int WRAPL_OFF = BOT_OFF + ROW_SLOT;    // cells N2-2,N2-1 (left of k = 0)
  static constexpr int WRAPR_OFF = WRAPL_OFF + (WRAP_BYTES + 127) / 128 * 128;  // cells 0,1
  static constexpr int STAGE_BYTES = WRAPR_OFF + (WRAP_BYTES + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
  static_assert(BJ % R == 0 && BK % 2 == 0 && CONSUMERS % 32 == 0, "bad tile");
  static_assert(BODY_BYTES % 128 == 0, "TMA destination must stay 128-byte aligned");
  static_assert(BKH <= 256 && BJ <= 256, "TMA box limit");
};

// Branch slots in the reference's application order (std::map order of the offsets,
// ref: Filter.cpp:202): (-1,0,0) (0,-1,0) (0,0,-1) (0,0,0) (0,0,1) (0,1,0) (1,0,0)
struct Lap7Args {
  double* out;
  int64_t n1, n2, nloc;
  int64_t ibeg, iend;
  int ci, njt, nkt;
  int64_t nwork;
  int G;
  double w[7];
  int has[7];  // branch present in the stencil map (absent branches are skipped, not added as 0)
  int one[7];  // weight is exactly 1.0
  int skip_lo, skip_hi;  // no (-1,0,0) / (1,0,0) branch: the planes outside a chunk are not even loaded
};

// acc + w*v, both operations rounded separately.  (Skipping the multiply for w == 1.0 is exact but
// the per-branch select cost 35 % on a B200 -- measured, profiles/r01g_lap7_ab.txt -- so it multiplies.)
__device__ __forceinline__ double acc_branch(double acc, double w, int /*one*/, double v) {
  return __dadd_rn(acc, __dmul_rn(w, v));
}

// acc = 0; for each present branch in order: acc = acc + w * in[...]   (ref: Filter.cpp:247-251)
// The last branch (i+1) arrives one plane later, so the first six are accumulated when
// plane i is in shared memory and the sum is finished when plane i+1 lands.
// RAGGED: the plane is not a whole number of tiles (any n1 >= 2, any even n2 >= 4).  TMA zero-fills what lies outside
// the tensor; the periodic neighbours of the plane's last row / last cell pair then sit INSIDE the last tiles, so the
// row below comes from the wrap row (`bot`) for whichever thread row holds global row n1-1, the cell to the right from
// the wrap columns for whichever thread holds cells n2-2, n2-1, and stores are masked.  Planes that tile exactly keep
// the unmasked instantiation.
template <class C, bool RAGGED = false>
__global__ void __launch_bounds__(C::THREADS)
    lap7_tma_kernel(const __grid_constant__ CUtensorMap tm_body, const __grid_constant__ CUtensorMap tm_row,
                    const __grid_constant__ CUtensorMap tm_col, const __grid_constant__ CUtensorMap tm_glo,
                    const __grid_constant__ CUtensorMap tm_ghi, const Lap7Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t full = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t empty = full + C::STAGES * 8;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, C::CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == C::CONSUMER_WARPS) {
    // ===================== producer warp =====================
    if ((tid & 31) == 0) {
      prefetch_tmap(&tm_body);
      prefetch_tmap(&tm_row);
      prefetch_tmap(&tm_col);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        const int kt = (int)(w % a.nkt);
        const int jt = (int)((w / a.nkt) % a.njt);
        const int64_t ic = w / ((int64_t)a.nkt * a.njt);
        const int64_t i0 = a.ibeg + ic * a.ci;
        const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
        const int k0 = kt * C::BK - 2;
        const int j0 = jt * C::BJ;
        const int jm = (j0 == 0) ? (int)a.n1 - 1 : j0 - 1;
        const int jp = (j0 + C::BJ >= (int)a.n1) ? 0 : j0 + C::BJ;
        const bool first_k = (kt == 0), last_k = (kt == a.nkt - 1);
        for (int64_t i = (a.skip_lo ? i0 : i0 - 1); i <= (a.skip_hi ? i1 - 1 : i1); ++i) {
          mbar_wait(empty + 8 * stage, phase ^ 1);
          const uint32_t st = smem + stage * C::STAGE_BYTES;
          const uint32_t fb = full + 8 * stage;
          if (i == i0 - 1 || i == i1) {
            // planes just outside the chunk: only their tile cells are needed
            mbar_expect_tx(fb, C::BODY_BYTES);
            if (i < 0)
              tma_load_3d(st + C::BODY_OFF, &tm_glo, fb, k0, j0, a.G - 1);
            else if (i >= a.nloc)
              tma_load_3d(st + C::BODY_OFF, &tm_ghi, fb, k0, j0, 0);
            else
              tma_load_3d(st + C::BODY_OFF, &tm_body, fb, k0, j0, (int)i);
          } else {
            const uint32_t bytes = C::BODY_BYTES + 2 * C::ROW_BYTES + (first_k ? C::WRAP_BYTES : 0) +
                                   (last_k ? C::WRAP_BYTES : 0);
            mbar_expect_tx(fb, bytes);
            tma_load_3d(st + C::BODY_OFF, &tm_body, fb, k0, j0, (int)i);
            tma_load_3d(st + C::TOP_OFF, &tm_row, fb, k0, jm, (int)i);
            tma_load_3d(st + C::BOT_OFF, &tm_row, fb, k0, jp, (int)i);
            if (first_k) tma_load_3d(st + C::WRAPL_OFF, &tm_col, fb, (int)a.n2 - 2, j0, (int)i);
            if (last_k) tma_load_3d(st + C::WRAPR_OFF, &tm_col, fb, 0, j0, (int)i);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const int tx = tid % C::TX;
  const int ty = tid / C::TX;
  const int r0 = ty * C::R;
  const int lane = tid & 31;
  int stage = 0;
  uint32_t phase = 0;
  const uint32_t col = (2 + 2 * tx) * 8;  // byte offset of this thread's two cells in a row

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK + 2 * tx;
    const int64_t j = (int64_t)jt * C::BJ + r0;
    const bool wrapl_lane = (kt == 0) && (tx == 0);
    const bool wrapr_lane = (kt == a.nkt - 1) && (k + 2 == a.n2);  // the thread that holds the plane's last two cells
    // tile row that holds the plane's last row (the tile's last row unless the tile is ragged)
    const int last_row = RAGGED ? (int)((a.n1 - (int64_t)jt * C::BJ < C::BJ ? a.n1 - (int64_t)jt * C::BJ : C::BJ)) - 1 : C::BJ - 1;
    const bool k_ok = !RAGGED || k < a.n2;

    double2 below[C::R];    // plane i-1
    double2 partial[C::R];  // first six branches of plane i-1's output (waiting for plane i)
    if (!a.skip_lo) {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
#pragma unroll
      for (int r = 0; r < C::R; ++r) below[r] = lds_v2(st + C::BODY_OFF + (r0 + r) * C::ROW_BYTES + col);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    } else {
#pragma unroll
      for (int r = 0; r < C::R; ++r) below[r] = make_double2(0.0, 0.0);
    }

    // finish plane `ip` with the (i+1) branch taken from `above`, and store it
    auto finish = [&](int64_t ip, const double2* above) {
      double* orow = a.out + (ip * a.n1 + j) * a.n2 + k;
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        double x = partial[r].x, y = partial[r].y;
        if (a.has[6]) {
          x = acc_branch(x, a.w[6], a.one[6], above[r].x);
          y = acc_branch(y, a.w[6], a.one[6], above[r].y);
        }
        if (!RAGGED || (k_ok && r0 + r <= last_row)) st_global_v2(orow + (int64_t)r * a.n2, x, y);
      }
    };

    for (int64_t i = i0; i < i1; ++i) {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
      double2 ctr[C::R];
      double km1[C::R], kp1[C::R];
      const uint32_t up_row = (r0 == 0) ? st + C::TOP_OFF : st + C::BODY_OFF + (r0 - 1) * C::ROW_BYTES;
      const uint32_t dn_row =
          (r0 + C::R == C::BJ) ? st + C::BOT_OFF : st + C::BODY_OFF + (r0 + C::R) * C::ROW_BYTES;
      const double2 up = lds_v2(up_row + col);
      const double2 dn = lds_v2(dn_row + col);
      double2 bot = dn;
      if (RAGGED) bot = lds_v2(st + C::BOT_OFF + col);  // global row 0 (periodic), wherever the plane's last row sits
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const uint32_t row = st + C::BODY_OFF + (r0 + r) * C::ROW_BYTES;
        ctr[r] = lds_v2(row + col);
        km1[r] = lds_f64(wrapl_lane ? st + C::WRAPL_OFF + (r0 + r) * 16 + 8 : row + col - 8);
        kp1[r] = lds_f64(wrapr_lane ? st + C::WRAPR_OFF + (r0 + r) * 16 : row + col + 16);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }

      // plane i-1 was waiting for this plane's centre
      if (!a.skip_hi && i > i0) finish(i - 1, ctr);
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const double2 jm = (r == 0) ? up : ctr[r - 1];
        double2 jp = (r == C::R - 1) ? dn : ctr[r + 1];
        if (RAGGED && r0 + r == last_row) jp = bot;
        double x = 0.0, y = 0.0;
        if (a.has[0]) { x = acc_branch(x, a.w[0], a.one[0], below[r].x); y = acc_branch(y, a.w[0], a.one[0], below[r].y); }
        if (a.has[1]) { x = acc_branch(x, a.w[1], a.one[1], jm.x); y = acc_branch(y, a.w[1], a.one[1], jm.y); }
        if (a.has[2]) { x = acc_branch(x, a.w[2], a.one[2], km1[r]); y = acc_branch(y, a.w[2], a.one[2], ctr[r].x); }
        if (a.has[3]) { x = acc_branch(x, a.w[3], a.one[3], ctr[r].x); y = acc_branch(y, a.w[3], a.one[3], ctr[r].y); }
        if (a.has[4]) { x = acc_branch(x, a.w[4], a.one[4], ctr[r].y); y = acc_branch(y, a.w[4], a.one[4], kp1[r]); }
        if (a.has[5]) { x = acc_branch(x, a.w[5], a.one[5], jp.x); y = acc_branch(y, a.w[5], a.one[5], jp.y); }
        partial[r].x = x;
        partial[r].y = y;
        below[r] = ctr[r];
      }
      if (a.skip_hi) finish(i, ctr);  // no (i+1) branch: the sum is complete
    }
    if (!a.skip_hi) {
      // plane above the chunk: finishes the last output plane
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
      double2 ctr[C::R];
#pragma unroll
      for (int r = 0; r < C::R; ++r) ctr[r] = lds_v2(st + C::BODY_OFF + (r0 + r) * C::ROW_BYTES + col);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      finish(i1 - 1, ctr);
    }
  }
}

// ---- configurations -------------------------------------------------------------
typedef void (*UpwindTmaKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap,
                                const CUtensorMap, const UpwindTmaArgs);
struct UpwindTmaConfig {
  int BJ, BK, BKH, threads, smem;
  UpwindTmaKernel kernel;
  const char* name;
};
template <class C>
constexpr UpwindTmaConfig make_cfg(const char* name) {
  return UpwindTmaConfig{C::BJ, C::BK, C::BKH, C::THREADS, C::SMEM_BYTES, upwind3d_tma_kernel<C>, name};
}
// kDefaultUpCfg is the default (best of the round-1 sweeps at 512^3 and 1024^3, profiles/);
// the others exist for tuning sweeps (env FDB_TMA_CFG)
const UpwindTmaConfig kUpCfgs[] = {
    make_cfg<UpwindCfg<16, 128, 4, 6>>("bj16_bk128_r4_s6"),
    make_cfg<UpwindCfg<16, 128, 4, 4>>("bj16_bk128_r4_s4"),
    make_cfg<UpwindCfg<8, 128, 4, 8>>("bj8_bk128_r4_s8"),
    make_cfg<UpwindCfg<32, 128, 4, 3>>("bj32_bk128_r4_s3"),
    make_cfg<UpwindCfg<16, 128, 2, 6>>("bj16_bk128_r2_s6"),
    make_cfg<UpwindCfg<16, 64, 4, 8>>("bj16_bk64_r4_s8"),
    make_cfg<UpwindCfg<8, 128, 2, 8>>("bj8_bk128_r2_s8"),
    make_cfg<UpwindCfg<16, 128, 8, 6>>("bj16_bk128_r8_s6"),
    make_cfg<UpwindCfg<32, 128, 8, 3>>("bj32_bk128_r8_s3"),
    make_cfg<UpwindCfg<32, 128, 4, 4>>("bj32_bk128_r4_s4"),
    make_cfg<UpwindCfg<32, 128, 4, 6>>("bj32_bk128_r4_s6"),
    make_cfg<UpwindCfg<64, 128, 8, 3>>("bj64_bk128_r8_s3"),
    make_cfg<UpwindCfg<32, 128, 8, 6>>("bj32_bk128_r8_s6"),
    make_cfg<UpwindCfg<64, 128, 16, 3>>("bj64_bk128_r16_s3"),
    make_cfg<UpwindCfg<48, 128, 8, 4>>("bj48_bk128_r8_s4"),
};
constexpr int kNumUpCfgs = sizeof(kUpCfgs) / sizeof(kUpCfgs[0]);
constexpr int kDefaultUpCfg = 3;  // bj32_bk128_r4_s3: 1 CTA/SM, 512 consumer threads, 3 x 34 KB stages

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

struct KernelAttr {
  bool done = false;
  int ctas_per_sm = 1;
  int sms = 148;
};
KernelAttr g_up_attr[16][kNumUpCfgs];  // per device, per config

}  // namespace

// Encode every tensor map slab d needs (both ping-pong buffers).  Leaves
// have_tma = false when the row pitch is not a multiple of 16 bytes (odd n2).
int tma_encode_slab(Field* f, int d) {
  Slab& s = f->slabs[d];
  s.have_tma = false;
  int ci = env_int("FDB_TMA_CFG", kDefaultUpCfg);
  if (ci < 0 || ci >= kNumUpCfgs) ci = kDefaultUpCfg;
  s.tma_cfg = ci;
  const UpwindTmaConfig& C = kUpCfgs[ci];
  const int64_t n1 = f->geo.n[1], n2 = f->geo.n[2];
  if (f->geo.ndims != 3 || n2 % 2 != 0 || n2 < 4) return FDB_OK;
  FDB_CUDA(cudaSetDevice(s.device));
  for (int p = 0; p < 2; ++p) {
    FDB_TRY(encode_tensor_map_3d(&s.tm_body[p], f->body(d, p), n2, n1, s.nloc(), C.BKH, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_row[p], f->body(d, p), n2, n1, s.nloc(), C.BKH, 1));
    FDB_TRY(encode_tensor_map_3d(&s.tm_col[p], f->body(d, p), n2, n1, s.nloc(), 2, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_glo[p], f->ghost_lo(d, p), n2, n1, f->G, C.BKH, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_ghi[p], f->ghost_hi(d, p), n2, n1, f->G, C.BKH, C.BJ));
  }
  s.have_tma = true;
  return FDB_OK;
}

bool upwind_tma_supported(const Field& f, const UpwindCoeffs& k) {
  if (f.geo.ndims != 3) return false;
  if (!f.slabs.empty() && !f.slabs[0].have_tma) return false;
  if (k.up[0] != -1 || k.up[1] != -1 || k.up[2] != -1) return false;
  if (f.geo.n[2] % 2 != 0 || f.geo.n[2] < 4) return false;  // 16-byte row pitch, wrap box of 2 cells
  if (f.geo.n[1] < 2) return false;
  return true;
}

// One-time, per-device set-up of the kernel (function attributes, occupancy: calls that may load the module and
// synchronise the device).  Also reachable through SweepLauncher::prepare, which multi-device fields call BEFORE
// they enqueue anything that waits on another device's counters.
int upwind_tma_prepare(const Field& f, int d) {
  const Slab& sl = f.slabs[d];
  if (!sl.have_tma) return FDB_OK;
  const UpwindTmaConfig& C = kUpCfgs[sl.tma_cfg];
  KernelAttr& at = g_up_attr[sl.device & 15][sl.tma_cfg];
  if (!at.done) {
    FDB_CUDA(cudaFuncSetAttribute(C.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem));
    int nb = 0;
    FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, C.kernel, C.threads, C.smem));
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, sl.device));
    at.ctas_per_sm = nb < 1 ? 1 : nb;
    at.sms = prop.multiProcessorCount;
    at.done = true;
  }
  return FDB_OK;
}

int launch_upwind_tma(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                      cudaStream_t s, double* peer_out, int64_t peer_from) {
  if (iend <= ibeg) return FDB_OK;
  const Slab& sl = f.slabs[d];
  const UpwindTmaConfig& C = kUpCfgs[sl.tma_cfg];
  KernelAttr& at = g_up_attr[sl.device & 15][sl.tma_cfg];
  FDB_TRY(upwind_tma_prepare(f, d));
  UpwindTmaArgs a;
  a.out = f.body(d, 1 - X);
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.ibeg = ibeg;
  a.iend = iend;
  a.njt = (int)((a.n1 + C.BJ - 1) / C.BJ);
  a.nkt = (int)((a.n2 + C.BK - 1) / C.BK);
  a.G = f.G;
  a.c0 = k.c[0];
  a.c1 = k.c[1];
  a.c2 = k.c[2];
  a.peer_out = peer_out;
  a.peer_from = peer_from;
  // i-chunk: long enough to amortise the extra plane each work item reads,
  // short enough that every CTA gets several items (static round-robin).
  const int reserve = f.single() ? 0 : env_int("FDB_COMM_SMS", 0);  // SMs left free for NCCL halo kernels (see kernels_fused.cu)
  int64_t grid_max = (int64_t)at.ctas_per_sm * (at.sms - reserve);
  if (grid_max < 1) grid_max = 1;
  const int64_t tiles = (int64_t)a.njt * a.nkt;
  const int64_t planes = iend - ibeg;
  int64_t ci = env_int("FDB_TMA_CI", 0);
  if (ci <= 0) {
    // HBM-bound: a ragged last round costs little (the remaining CTAs still saturate DRAM),
    // re-reading a plane per work item costs 1/ci of the reads -- favour long chunks
    ci = 64;
    while (ci > 4 && tiles * ((planes + ci - 1) / ci) < 3 * grid_max) ci /= 2;
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nwork = tiles * ((planes + ci - 1) / ci);
  int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  // FDB_MAX_CTAS (tests): fewer CTAs than the device holds, so every CTA walks many work items
  if (const int cap = env_int("FDB_MAX_CTAS", 0); cap > 0 && grid > cap) grid = cap;
  const int p = X;
  C.kernel<<<(unsigned)grid, C.threads, C.smem, s>>>(sl.tm_body[p], sl.tm_row[p], sl.tm_col[p],
                                                     sl.tm_glo[p], a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

// ---- 7-point stencil: configs, tensor maps, launcher ---------------------------------------
namespace {
typedef void (*Lap7Kernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const CUtensorMap, const Lap7Args);
struct Lap7Config {
  int BJ, BK, BKH, threads, smem;
  Lap7Kernel kernel;
  Lap7Kernel kernel_ragged;  // for planes the tile does not divide (null: this configuration only runs exact tilings)
  const char* name;
};
template <class C>
constexpr Lap7Config make_lap_cfg(const char* name) {
  return Lap7Config{C::BJ, C::BK, C::BKH, C::THREADS, C::SMEM_BYTES, lap7_tma_kernel<C, false>, nullptr, name};
}
template <class C>
constexpr Lap7Config make_lap_cfg_ragged(const char* name) {
  return Lap7Config{C::BJ, C::BK, C::BKH, C::THREADS, C::SMEM_BYTES, lap7_tma_kernel<C, false>, lap7_tma_kernel<C, true>, name};
}
const Lap7Config kLapCfgs[] = {
    make_lap_cfg_ragged<Lap7Cfg<16, 128, 4, 5>>("bj16_bk128_r4_s5"),
    make_lap_cfg<Lap7Cfg<32, 128, 4, 3>>("bj32_bk128_r4_s3"),
    make_lap_cfg<Lap7Cfg<32, 128, 8, 3>>("bj32_bk128_r8_s3"),
    make_lap_cfg_ragged<Lap7Cfg<16, 64, 4, 6>>("bj16_bk64_r4_s6"),
    make_lap_cfg_ragged<Lap7Cfg<8, 32, 4, 6>>("bj8_bk32_r4_s6"),
    make_lap_cfg<Lap7Cfg<16, 128, 4, 6>>("bj16_bk128_r4_s6"),
    make_lap_cfg<Lap7Cfg<16, 128, 4, 8>>("bj16_bk128_r4_s8"),
    make_lap_cfg<Lap7Cfg<16, 128, 2, 6>>("bj16_bk128_r2_s6"),
    make_lap_cfg<Lap7Cfg<32, 128, 4, 4>>("bj32_bk128_r4_s4"),
    make_lap_cfg<Lap7Cfg<8, 128, 4, 8>>("bj8_bk128_r4_s8"),
    make_lap_cfg<Lap7Cfg<16, 128, 8, 5>>("bj16_bk128_r8_s5"),
    make_lap_cfg<Lap7Cfg<32, 128, 2, 3>>("bj32_bk128_r2_s3"),
};
constexpr int kNumLapCfgs = sizeof(kLapCfgs) / sizeof(kLapCfgs[0]);
KernelAttr g_lap_attr[16][kNumLapCfgs];

// slot of an internal-axis offset in the reference's application order, or -1
int lap7_slot(const int* o) {
  static const int order[7][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
  for (int b = 0; b < 7; ++b)
    if (o[0] == order[b][0] && o[1] == order[b][1] && o[2] == order[b][2]) return b;
  return -1;
}

// the largest configured tile that divides the plane; failing that (round 2) the ragged-tile configuration that wastes
// the fewest tile cells -- any n1 >= 2 and any even n2 >= 4.  *ragged tells which kernel instantiation runs.
int lap7_pick_cfg(const Field& f, bool* ragged) {
  const int forced = env_int("FDB_LAP_CFG", -1);
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  *ragged = false;
  // round-1 sweeps (profiles/r01g_lap7_cfg_sweep.txt): on 1024^2 planes the 8-rows-per-thread
  // tile (index 10) is ~6 % ahead, on 512^2 planes the default 4-rows-per-thread one
  if (forced < 0 && n1 * n2 >= 768 * 768 && n1 % kLapCfgs[10].BJ == 0 && n2 % kLapCfgs[10].BK == 0) return 10;
  for (int c = 0; c < kNumLapCfgs; ++c) {
    if (forced >= 0 && c != forced) continue;
    if (n1 % kLapCfgs[c].BJ == 0 && n2 % kLapCfgs[c].BK == 0) return c;
  }
  if (env_int("FDB_LAP_NO_RAGGED", 0) != 0 || n1 < 2 || n2 < 4 || n2 % 2 != 0) return -1;
  int best = -1;
  int64_t best_cells = 0;
  for (int c = 0; c < kNumLapCfgs; ++c) {
    if (!kLapCfgs[c].kernel_ragged || (forced >= 0 && c != forced)) continue;
    const int64_t bj = kLapCfgs[c].BJ, bk = kLapCfgs[c].BK;
    const int64_t cells = ((n1 + bj - 1) / bj * bj) * ((n2 + bk - 1) / bk * bk);
    if (best < 0 || cells < best_cells) { best = c; best_cells = cells; }
  }
  *ragged = best >= 0;
  return best;
}
}  // namespace

int tma_encode_slab_lap7(Field* f, int d) {
  Slab& s = f->slabs[d];
  s.have_tma = false;
  const bool plane2d = f->geo.ndims == 2 && f->geo.n[0] == 1;
  if (f->geo.ndims != 3 && !plane2d) return FDB_OK;
  bool ragged = false;
  const int c = lap7_pick_cfg(*f, &ragged);
  if (c < 0) return FDB_OK;
  s.tma_cfg = c;
  s.lap_ragged = ragged;
  const Lap7Config& C = kLapCfgs[c];
  const int64_t n1 = f->geo.n[1], n2 = f->geo.n[2];
  FDB_CUDA(cudaSetDevice(s.device));
  for (int p = 0; p < 2; ++p) {
    FDB_TRY(encode_tensor_map_3d(&s.tm_body[p], f->body(d, p), n2, n1, s.nloc(), C.BKH, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_row[p], f->body(d, p), n2, n1, s.nloc(), C.BKH, 1));
    FDB_TRY(encode_tensor_map_3d(&s.tm_col[p], f->body(d, p), n2, n1, s.nloc(), 2, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_glo[p], f->ghost_lo(d, p), n2, n1, f->G, C.BKH, C.BJ));
    FDB_TRY(encode_tensor_map_3d(&s.tm_ghi[p], f->ghost_hi(d, p), n2, n1, f->G, C.BKH, C.BJ));
  }
  s.have_tma = true;
  return FDB_OK;
}

bool stencil_lap7_supported(const Field& f, const StencilBranches& b) {
  const bool plane2d = f.geo.ndims == 2 && f.geo.n[0] == 1;
  if (f.geo.ndims != 3 && !plane2d) return false;
  if (f.slabs.empty() || !f.slabs[0].have_tma) return false;
  if (b.nbranch < 1 || b.nbranch > 7) return false;
  int seen = 0;
  for (int i = 0; i < b.nbranch; ++i) {
    const int slot = lap7_slot(b.off[i]);
    if (slot < 0 || (seen >> slot) & 1) return false;
    seen |= 1 << slot;
  }
  return true;
}

int stencil_lap7_prepare(const Field& f, int d) {
  const Slab& sl = f.slabs[d];
  if (!sl.have_tma) return FDB_OK;
  const Lap7Config& C = kLapCfgs[sl.tma_cfg];
  KernelAttr& at = g_lap_attr[sl.device & 15][sl.tma_cfg];
  if (!at.done) {
    FDB_CUDA(cudaFuncSetAttribute(C.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem));
    if (C.kernel_ragged) FDB_CUDA(cudaFuncSetAttribute(C.kernel_ragged, cudaFuncAttributeMaxDynamicSharedMemorySize, C.smem));
    int nb = 0;
    FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, C.kernel, C.threads, C.smem));
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, sl.device));
    at.ctas_per_sm = nb < 1 ? 1 : nb;
    at.sms = prop.multiProcessorCount;
    at.done = true;
  }
  return FDB_OK;
}

int launch_stencil_lap7(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const StencilBranches& b,
                        cudaStream_t s) {
  if (iend <= ibeg) return FDB_OK;
  const Slab& sl = f.slabs[d];
  const Lap7Config& C = kLapCfgs[sl.tma_cfg];
  KernelAttr& at = g_lap_attr[sl.device & 15][sl.tma_cfg];
  FDB_TRY(stencil_lap7_prepare(f, d));
  Lap7Args a;
  a.out = f.body(d, 1 - X);
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.nloc = sl.nloc();
  a.ibeg = ibeg;
  a.iend = iend;
  a.njt = (int)((a.n1 + C.BJ - 1) / C.BJ);
  a.nkt = (int)((a.n2 + C.BK - 1) / C.BK);
  a.G = f.G;
  for (int i = 0; i < 7; ++i) { a.w[i] = 0.0; a.has[i] = 0; a.one[i] = 0; }
  for (int i = 0; i < b.nbranch; ++i) {
    const int slot = lap7_slot(b.off[i]);
    a.w[slot] = b.w[i];
    a.has[slot] = 1;
    a.one[slot] = (b.w[i] == 1.0) ? 1 : 0;
  }
  a.skip_lo = !a.has[0];
  a.skip_hi = !a.has[6];
  const int reserve = f.single() ? 0 : env_int("FDB_COMM_SMS", 0);  // SMs left free for NCCL halo kernels (see kernels_fused.cu)
  int64_t grid_max = (int64_t)at.ctas_per_sm * (at.sms - reserve);
  if (grid_max < 1) grid_max = 1;
  const int64_t tiles = (int64_t)a.njt * a.nkt;
  const int64_t planes = iend - ibeg;
  int64_t ci = env_int("FDB_TMA_CI", 0);
  if (ci <= 0) {
    ci = 64;
    while (ci > 4 && tiles * ((planes + ci - 1) / ci) < 4 * grid_max) ci /= 2;
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nwork = tiles * ((planes + ci - 1) / ci);
  int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  // FDB_MAX_CTAS (tests): fewer CTAs than the device holds, so every CTA walks many work items
  if (const int cap = env_int("FDB_MAX_CTAS", 0); cap > 0 && grid > cap) grid = cap;
  const int p = X;
  const Lap7Kernel kern = sl.lap_ragged ? C.kernel_ragged : C.kernel;
  kern<<<(unsigned)grid, C.threads, C.smem, s>>>(sl.tm_body[p], sl.tm_row[p], sl.tm_col[p], sl.tm_glo[p], sl.tm_ghi[p], a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

}  // namespace fdb
