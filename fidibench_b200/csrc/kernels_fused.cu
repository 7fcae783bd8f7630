// kernels_fused.cu -- temporal blocking for the 3-D upwind step: T time steps per sweep.
//
// The single-step kernel (kernels_tma.cu) is pinned to the HBM roofline: 16 B of DRAM
// traffic per cell-update.  The upwind stencil only looks at LOWER indices (i-1, j-1, k-1),
// so T steps can be fused with one-sided halos and no look-ahead along the marching axis:
// when plane i of the input (level 0) lands in shared memory, level 1 of plane i follows from
// level 0 of planes i and i-1, level 2 from level 1 of planes i and i-1, ... up to level T,
// which is the only one written back.  DRAM traffic per cell-update drops to ~16/T bytes.
//
//   * same TMA / mbarrier producer-consumer ring as the single-step kernel;
//   * each consumer thread owns R rows x 2 cells of a CJ x CK compute tile and keeps plane
//     i-1 of every level 0..T-1 in registers (T*R double2);
//   * in-plane neighbours (j-1, k-1) of levels >= 1 travel through two ping-pong exchange
//     tiles in shared memory, one named barrier among the consumer warps per level;
//   * the tile computes T-1 halo rows/columns redundantly on its low sides (level s is valid
//     from compute row/column s-1 on); only the BJ x BK interior of level T is stored, so
//     output tiles stay 128-byte aligned along k;
//   * periodic wrap: halo rows come from a separate TMA box at (j0-T) mod N1, the wrap
//     columns of the first k-tile from boxes at N2-HKI; planes below the slab from the ghost
//     tensor (depth T), which aliases the far planes on a single device;
//   * a work item warms up on T extra planes below its chunk (levels become valid one plane
//     after the other), nothing is stored for them.
//
// Arithmetic per level is exactly the single step's (separately rounded, reference order,
// ref: upwind/cxx/upwind.cxx:72-80), so T fused steps are bit-identical to T single steps.
#include "fdb_internal.h"
#include "tma_ptx.cuh"

namespace fdb {

namespace {

using namespace ptx;

constexpr int align128(int x) { return (x + 127) / 128 * 128; }

template <int T_, int CJ_, int R_, int STAGES_, bool SHFL_ = false, int BK_ = 128, int MINB_ = 1>
struct FusedCfg {
  static constexpr int T = T_, CJ = CJ_, R = R_, STAGES = STAGES_;
  static constexpr bool SHFL = SHFL_;                  // k-1 neighbour by warp shuffle instead of LDS.64
  static constexpr int BK = BK_;                       // output cells per tile row
  static constexpr int MINB = MINB_;                   // CTAs per SM the register budget is sized for
  static constexpr int HKC = 2 * (T / 2);              // redundant compute columns (even, >= T-1)
  static constexpr int HKI = HKC + 2;                  // input halo columns (even, >= T)
  static constexpr int CK = BK + HKC;                  // compute columns
  static constexpr int TX = CK / 2;                    // threads per row
  static constexpr int TY = CJ / R;
  static constexpr int WORKERS = TX * TY;
  static constexpr int CONSUMERS = (WORKERS + 31) / 32 * 32;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;
  static constexpr int BJ = CJ - (T - 1);              // output rows per tile
  static constexpr int IN_ROWS = CJ + 1;               // input rows j0-T .. j0+BJ-1
  static constexpr int BODY_ROWS = IN_ROWS - T;
  static constexpr int BKH = BK + HKI;                 // input row pitch in doubles
  static constexpr int ROW_BYTES = BKH * 8;
  static constexpr int WROW_BYTES = HKI * 8;           // row pitch of the wrap-column area
  static constexpr int HALO_OFF = 0;
  static constexpr int BODY_OFF = align128(T * ROW_BYTES);
  static constexpr int WH_OFF = align128(BODY_OFF + BODY_ROWS * ROW_BYTES);
  static constexpr int WB_OFF = align128(WH_OFF + T * WROW_BYTES);
  static constexpr int STAGE_BYTES = align128(WB_OFF + BODY_ROWS * WROW_BYTES);
  static constexpr int TX_BYTES_MAIN = IN_ROWS * ROW_BYTES;
  static constexpr int TX_BYTES_WRAP = IN_ROWS * WROW_BYTES;
  static constexpr int XP = CK * 8;                    // exchange tile row pitch
  static constexpr int X_BYTES = align128(CJ * XP);
  static constexpr int NX = (T > 1) ? 2 : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NX * X_BYTES + 2 * STAGES * 8 + 128;
  static_assert(CJ % R == 0 && T >= 1 && HKC >= T - 1 && CJ > T, "bad fused tile");
  static_assert(BKH <= 256 && BODY_ROWS <= 256, "TMA box limit");
};

struct FusedArgs {
  double* out;
  int64_t n1, n2;
  int64_t ibeg, iend;
  int ci, njt, nkt;
  int64_t nwork;
  int G;  // planes in the ghost tensor; local plane p < 0 is its plane G + p
  double c0, c1, c2;
  // fused halo push: output planes p >= peer_from are ALSO stored through `peer_out`, the next
  // slab's ghost planes mapped into this GPU's address space (NVLink peer stores); null = off
  double* peer_out;
  int64_t peer_from;
};

// one upwind update of a cell (ref: upwind.cxx:72-80)
__device__ __forceinline__ double upwind_cell(double ctr, double im1, double jm1, double km1, double c0,
                                              double c1, double c2) {
  double t = ctr;
  t = __dsub_rn(t, __dmul_rn(c0, __dsub_rn(im1, ctr)));
  t = __dsub_rn(t, __dmul_rn(c1, __dsub_rn(jm1, ctr)));
  t = __dsub_rn(t, __dmul_rn(c2, __dsub_rn(km1, ctr)));
  return t;
}

// Tensor maps: m[0..3] over the local planes, m[4..7] over the ghost planes, box shapes
//   0/4: {BKH, T}  halo rows      1/5: {BKH, BODY_ROWS}  tile rows
//   2/6: {HKI, T}  wrap corner    3/7: {HKI, BODY_ROWS}  wrap columns
struct FusedMaps {
  CUtensorMap m[8];
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    upwind3d_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t full = xbuf + C::NX * C::X_BYTES;
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
#pragma unroll
      for (int m = 0; m < 8; ++m) prefetch_tmap(&maps.m[m]);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        const int kt = (int)(w % a.nkt);
        const int jt = (int)((w / a.nkt) % a.njt);
        const int64_t ic = w / ((int64_t)a.nkt * a.njt);
        const int64_t i0 = a.ibeg + ic * a.ci;
        const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
        const int kb = kt * C::BK - C::HKI;  // first input column (negative for the first k-tile)
        const int j0 = jt * C::BJ;
        const int jh = (j0 - C::T < 0) ? j0 - C::T + (int)a.n1 : j0 - C::T;  // periodic halo rows
        const uint32_t bytes = C::TX_BYTES_MAIN + (kt == 0 ? C::TX_BYTES_WRAP : 0);
        for (int64_t p = i0 - C::T; p < i1; ++p) {
          mbar_wait(empty + 8 * stage, phase ^ 1);
          const uint32_t st = smem + stage * C::STAGE_BYTES;
          const uint32_t fb = full + 8 * stage;
          const int g = (p < 0) ? 4 : 0;                 // ghost tensor below the slab
          const int pl = (p < 0) ? a.G + (int)p : (int)p;
          mbar_expect_tx(fb, bytes);
          tma_load_3d(st + C::HALO_OFF, &maps.m[g + 0], fb, kb, jh, pl);
          tma_load_3d(st + C::BODY_OFF, &maps.m[g + 1], fb, kb, j0, pl);
          if (kt == 0) {
            tma_load_3d(st + C::WH_OFF, &maps.m[g + 2], fb, (int)a.n2 - C::HKI, jh, pl);
            tma_load_3d(st + C::WB_OFF, &maps.m[g + 3], fb, (int)a.n2 - C::HKI, j0, pl);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  const bool worker = tid < C::WORKERS;  // threads past the tile only keep the barriers company
  const int wid = worker ? tid : 0;
  const int tx = wid % C::TX;
  const int ty = wid / C::TX;
  const int q0 = ty * C::R;  // first compute row of this thread
  const int lane = tid & 31;
  int stage = 0;
  uint32_t phase = 0;
  uint32_t xsel = 0;
  const double c0 = a.c0, c1 = a.c1, c2 = a.c2;
  // exchange-tile addresses (levels >= 1): own cells, the row above, the cell to the left
  const uint32_t x_own = q0 * C::XP + tx * 16;
  const uint32_t x_up = (q0 == 0 ? 0 : (q0 - 1) * C::XP) + tx * 16;
  const uint32_t x_km = q0 * C::XP + (tx == 0 ? 0 : tx * 16 - 8);

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - C::HKC + 2 * tx;       // global column of this thread's first cell
    const int64_t j = (int64_t)jt * C::BJ - (C::T - 1) + q0;       // global row of this thread's first row
    const bool store_cols = worker && (2 * tx >= C::HKC) && (k < a.n2);
    // level-0 source of this thread's cells / left neighbour: main tile or wrap-column area
    const bool own_wrap = (kt == 0) && (2 * tx + 2 < C::HKI);
    const bool km_wrap = (kt == 0) && (2 * tx + 1 < C::HKI);

    double2 carry[C::T][C::R];  // plane i-1 of levels 0..T-1
#pragma unroll
    for (int s = 0; s < C::T; ++s)
#pragma unroll
      for (int r = 0; r < C::R; ++r) carry[s][r] = make_double2(0.0, 0.0);

    for (int64_t p = i0 - C::T; p < i1; ++p) {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
      double2 v[C::R];
      double km[C::R];
      double2 up;
      {
        // stage row s = compute row + 1; rows 0..T-1 sit in the halo area
        auto main_row = [&](int s) -> uint32_t {
          return st + (s < C::T ? C::HALO_OFF + s * C::ROW_BYTES : C::BODY_OFF + (s - C::T) * C::ROW_BYTES);
        };
        auto wrap_row = [&](int s) -> uint32_t {
          return st + (s < C::T ? C::WH_OFF + s * C::WROW_BYTES : C::WB_OFF + (s - C::T) * C::WROW_BYTES);
        };
        up = lds_v2((own_wrap ? wrap_row(q0) : main_row(q0)) + (2 * tx + 2) * 8);
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const int s = q0 + r + 1;
          v[r] = lds_v2((own_wrap ? wrap_row(s) : main_row(s)) + (2 * tx + 2) * 8);
          if (C::SHFL) {
            // the cell to the left is the previous lane's second cell; only a warp's first
            // lane (and a row's first thread) reads it from shared memory
            km[r] = __shfl_up_sync(0xffffffffu, v[r].y, 1);
            if (lane == 0 || tx == 0) km[r] = lds_f64((km_wrap ? wrap_row(s) : main_row(s)) + (2 * tx + 1) * 8);
          } else {
            km[r] = lds_f64((km_wrap ? wrap_row(s) : main_row(s)) + (2 * tx + 1) * 8);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }

#pragma unroll
      for (int s = 0; s < C::T; ++s) {
        double2 nv[C::R];
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const double2 jm = (r == 0) ? up : v[r - 1];
          nv[r].x = upwind_cell(v[r].x, carry[s][r].x, jm.x, km[r], c0, c1, c2);
          nv[r].y = upwind_cell(v[r].y, carry[s][r].y, jm.y, v[r].x, c0, c1, c2);
          carry[s][r] = v[r];
        }
        if (s == C::T - 1) {
          if (p >= i0 && store_cols) {
            double* orow = a.out + (p * a.n1 + j) * a.n2 + k;
            double* prow = (a.peer_out != nullptr && p >= a.peer_from)
                               ? a.peer_out + ((p - a.peer_from) * a.n1 + j) * a.n2 + k
                               : nullptr;
#pragma unroll
            for (int r = 0; r < C::R; ++r) {
              const int64_t jr = j + r;
              if (q0 + r >= C::T - 1 && jr < a.n1) {
                st_global_v2(orow + (int64_t)r * a.n2, nv[r].x, nv[r].y);
                if (prow) st_global_v2(prow + (int64_t)r * a.n2, nv[r].x, nv[r].y);
              }
            }
          }
        } else {
          // hand level s+1 of this plane to the neighbours through the exchange tile
          const uint32_t xb = xbuf + xsel * C::X_BYTES;
          xsel ^= 1;
          if (worker) {
#pragma unroll
            for (int r = 0; r < C::R; ++r) sts_v2(xb + x_own + r * C::XP, nv[r].x, nv[r].y);
          }
          if (C::SHFL) {
#pragma unroll
            for (int r = 0; r < C::R; ++r) km[r] = __shfl_up_sync(0xffffffffu, nv[r].y, 1);
          }
          named_bar_sync(1, C::CONSUMERS);
          up = lds_v2(xb + x_up);
#pragma unroll
          for (int r = 0; r < C::R; ++r) {
            if (!C::SHFL || lane == 0) km[r] = lds_f64(xb + x_km + r * C::XP);
            v[r] = nv[r];
          }
        }
      }
    }
  }
}

// ---- configurations ---------------------------------------------------------------------
typedef void (*FusedKernel)(const FusedMaps, const FusedArgs);
struct FusedConfig {
  int T, CJ, BJ, BK, BKH, HKI, body_rows, threads, smem;
  FusedKernel kernel;
  const char* name;
};
template <class C>
constexpr FusedConfig make_fused(const char* name) {
  return FusedConfig{C::T, C::CJ, C::BJ, C::BK, C::BKH, C::HKI, C::BODY_ROWS, C::THREADS, C::SMEM_BYTES,
                     upwind3d_fused_kernel<C>, name};
}
// per T: index 0 is the default, the rest are tuning alternatives (env FDB_FUSED_CFG)
const FusedConfig kFused2[] = {
    make_fused<FusedCfg<2, 16, 2, 4>>("t2_cj16_r2_s4"),  // 750 GCUPS at 512^3 (DRAM-bound again)
    make_fused<FusedCfg<2, 32, 4, 4>>("t2_cj32_r4_s4"),
    make_fused<FusedCfg<2, 14, 2, 5>>("t2_cj14_r2_s5"),
    make_fused<FusedCfg<2, 16, 4, 5>>("t2_cj16_r4_s5"),
    make_fused<FusedCfg<2, 28, 4, 3>>("t2_cj28_r4_s3"),
    make_fused<FusedCfg<2, 16, 2, 6>>("t2_cj16_r2_s6"),
    make_fused<FusedCfg<2, 24, 3, 4>>("t2_cj24_r3_s4"),
    make_fused<FusedCfg<2, 21, 3, 5>>("t2_cj21_r3_s5"),
    make_fused<FusedCfg<2, 16, 2, 4, true>>("t2_cj16_r2_s4_shfl"),
    make_fused<FusedCfg<2, 21, 3, 5, true>>("t2_cj21_r3_s5_shfl"),
};
const FusedConfig kFused3[] = {
    make_fused<FusedCfg<3, 21, 3, 4>>("t3_cj21_r3_s4"),  // round-1 best: 900 GCUPS at 512^3 / 1024^3
    make_fused<FusedCfg<3, 16, 2, 4>>("t3_cj16_r2_s4"),
    make_fused<FusedCfg<3, 14, 2, 5>>("t3_cj14_r2_s5"),
    make_fused<FusedCfg<3, 28, 4, 3>>("t3_cj28_r4_s3"),
    make_fused<FusedCfg<3, 16, 4, 4>>("t3_cj16_r4_s4"),
    make_fused<FusedCfg<3, 21, 3, 3>>("t3_cj21_r3_s3"),
    make_fused<FusedCfg<3, 21, 3, 5>>("t3_cj21_r3_s5"),
    make_fused<FusedCfg<3, 24, 3, 4>>("t3_cj24_r3_s4"),
    make_fused<FusedCfg<3, 18, 3, 5>>("t3_cj18_r3_s5"),
    make_fused<FusedCfg<3, 21, 3, 4, true>>("t3_cj21_r3_s4_shfl"),
    make_fused<FusedCfg<3, 16, 2, 4, true>>("t3_cj16_r2_s4_shfl"),
    make_fused<FusedCfg<3, 18, 3, 4, false, 64, 2>>("t3_cj18_r3_s4_bk64_2cta"),
    make_fused<FusedCfg<3, 18, 3, 4, true, 64, 2>>("t3_cj18_r3_s4_bk64_2cta_shfl"),
    make_fused<FusedCfg<3, 21, 3, 3, false, 64, 2>>("t3_cj21_r3_s3_bk64_2cta"),
    make_fused<FusedCfg<3, 21, 3, 5, true>>("t3_cj21_r3_s5_shfl"),
};
const FusedConfig kFused4[] = {
    make_fused<FusedCfg<4, 21, 3, 4>>("t4_cj21_r3_s4"),
    make_fused<FusedCfg<4, 16, 2, 4>>("t4_cj16_r2_s4"),
    make_fused<FusedCfg<4, 14, 2, 4>>("t4_cj14_r2_s4"),
    make_fused<FusedCfg<4, 28, 4, 3>>("t4_cj28_r4_s3"),
    make_fused<FusedCfg<4, 16, 4, 4>>("t4_cj16_r4_s4"),
    make_fused<FusedCfg<4, 21, 3, 3>>("t4_cj21_r3_s3"),
    make_fused<FusedCfg<4, 21, 3, 4, true>>("t4_cj21_r3_s4_shfl"),
    make_fused<FusedCfg<4, 16, 2, 4, true>>("t4_cj16_r2_s4_shfl"),
};

const FusedConfig* fused_table(int T, int* count) {
  switch (T) {
    case 2: *count = sizeof(kFused2) / sizeof(kFused2[0]); return kFused2;
    case 3: *count = sizeof(kFused3) / sizeof(kFused3[0]); return kFused3;
    case 4: *count = sizeof(kFused4) / sizeof(kFused4[0]); return kFused4;
  }
  *count = 0;
  return nullptr;
}

int env_int2(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

const FusedConfig* fused_pick(int T) {
  int n = 0;
  const FusedConfig* tab = fused_table(T, &n);
  if (!tab) return nullptr;
  int c = env_int2("FDB_FUSED_CFG", 0);
  if (c < 0 || c >= n) c = 0;
  return &tab[c];
}

struct FusedAttr {
  const FusedConfig* cfg = nullptr;
  int ctas_per_sm = 1;
  int sms = 148;
};
FusedAttr g_fused_attr[16][kMaxFuse + 1];

}  // namespace

int upwind_fused_max() { return kMaxFuse; }

bool upwind_fused_supported(const Field& f, const UpwindCoeffs& k, int T) {
  if (T < 2 || T > kMaxFuse) return false;
  if (!upwind_tma_supported(f, k)) return false;
  if (f.G < T) return false;
  if (f.geo.n[1] < 8 || f.geo.n[2] < 16) return false;
  for (const Slab& s : f.slabs)
    if (s.nloc() < T) return false;
  return true;
}

// tensor maps of slab d for fuse depth T, buffer parity p (encoded on first use)
static int fused_maps(const Field& f, int d, int T, const FusedConfig& C, int p, FusedMaps* out) {
  const Slab& s = f.slabs[d];
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  const double* body = f.body(d, p);
  const double* glo = f.ghost_lo(d, p);
  const int boxes[4][2] = {{C.BKH, T}, {C.BKH, C.body_rows}, {C.HKI, T}, {C.HKI, C.body_rows}};
  for (int b = 0; b < 4; ++b) {
    FDB_TRY(encode_tensor_map_3d(&out->m[b], body, n2, n1, s.nloc(), boxes[b][0], boxes[b][1]));
    FDB_TRY(encode_tensor_map_3d(&out->m[4 + b], glo, n2, n1, f.G, boxes[b][0], boxes[b][1]));
  }
  return FDB_OK;
}

int launch_upwind_fused(Field& f, int d, int X, int T, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                        cudaStream_t s, double* peer_out, int64_t peer_from) {
  if (iend <= ibeg) return FDB_OK;
  Slab& sl = f.slabs[d];
  FusedAttr& at = g_fused_attr[sl.device & 15][T];
  const FusedConfig* C = fused_pick(T);
  if (!C) return set_error(FDB_E_INVALID, "no fused kernel for %d steps per sweep", T);
  if (at.cfg != C) {
    FDB_CUDA(cudaFuncSetAttribute(C->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C->smem));
    int nb = 0;
    FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, C->kernel, C->threads, C->smem));
    if (nb < 1) return set_error(FDB_E_CUDA, "fused kernel %s does not fit on an SM", C->name);
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, sl.device));
    at.ctas_per_sm = nb;
    at.sms = prop.multiProcessorCount;
    at.cfg = C;
  }
  // tensor maps are cached per (slab, T, config, parity)
  if (sl.fused_T != T || sl.fused_cfg != (const void*)C) {
    for (int p = 0; p < 2; ++p)
      FDB_TRY(fused_maps(f, d, T, *C, p, reinterpret_cast<FusedMaps*>(sl.fused_maps[p])));
    sl.fused_T = T;
    sl.fused_cfg = (const void*)C;
  }
  FusedArgs a;
  a.out = f.body(d, 1 - X);
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.ibeg = ibeg;
  a.iend = iend;
  a.njt = (int)((a.n1 + C->BJ - 1) / C->BJ);
  a.nkt = (int)((a.n2 + C->BK - 1) / C->BK);
  a.G = f.G;
  a.c0 = k.c[0];
  a.c1 = k.c[1];
  a.c2 = k.c[2];
  a.peer_out = peer_out;
  a.peer_from = peer_from;
  // The direct halo transport runs on copy engines and needs no SM.  With the NCCL transport
  // the send/recv kernels must find free SMs while the persistent interior kernel runs:
  // FDB_COMM_SMS leaves some unoccupied (mind the extra round a smaller grid can cost).
  const int reserve = f.single() ? 0 : env_int2("FDB_COMM_SMS", 0);
  int64_t grid_max = (int64_t)at.ctas_per_sm * (at.sms - reserve);
  if (grid_max < 1) grid_max = 1;
  const int64_t tiles = (int64_t)a.njt * a.nkt;
  const int64_t planes = iend - ibeg;
  int64_t ci = env_int2("FDB_TMA_CI", 0);
  if (ci <= 0) {
    // every work item warms up on T extra planes: favour long chunks
    ci = 128;
    while (ci > 8 && tiles * ((planes + ci - 1) / ci) < 2 * grid_max) ci /= 2;
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nwork = tiles * ((planes + ci - 1) / ci);
  const int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  C->kernel<<<(unsigned)grid, C->threads, C->smem, s>>>(*reinterpret_cast<const FusedMaps*>(sl.fused_maps[X]), a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

}  // namespace fdb
