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

template <int BJ_, int R_, int STAGES_, int MINB_ = 1>
struct LapFusedCfg {
  static constexpr int BJ = BJ_, BK = 128, R = R_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int CJ = BJ + 2;       // level-1 rows:    global j0-1 .. j0+BJ
  static constexpr int CK = BK + 4;       // level-1 columns: global k0-2 .. k0+BK+1
  static constexpr int IN_ROWS = BJ + 4;  // level-0 rows:    global j0-2 .. j0+BJ+1 (stage row s)
  static constexpr int TX = CK / 2;       // threads per row (2 cells each)
  static constexpr int TY = CJ / R;
  static constexpr int WORKERS = TX * TY;
  static constexpr int CONSUMERS = (WORKERS + 31) / 32 * 32;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;
  static constexpr int ROW_BYTES = CK * 8;
  // Stage layout.  Three TMA boxes (2 halo rows, BJ tile rows, 2 halo rows) land back to back;
  // every TMA destination must be 128-byte aligned, so the tile rows start at BODY_OFF and stage
  // row s >= 2 sits ROW_SKEW bytes past s * ROW_BYTES.
  static constexpr int BODY_OFF = lf_align128(2 * ROW_BYTES);
  static constexpr int ROW_SKEW = BODY_OFF - 2 * ROW_BYTES;
  static constexpr int BOT_OFF = BODY_OFF + BJ * ROW_BYTES;
  static constexpr int MAIN_BYTES = lf_align128(BOT_OFF + 2 * ROW_BYTES);
  // wrap-column areas: 2 cells per row (pitch 16 B), same three boxes, same skew rule
  static constexpr int WBODY_OFF = 128;
  static constexpr int WSKEW = WBODY_OFF - 2 * 16;
  static constexpr int WBOT_OFF = WBODY_OFF + BJ * 16;
  static constexpr int WRAP_BYTES = lf_align128(WBOT_OFF + 2 * 16);
  static constexpr int WL_OFF = MAIN_BYTES;            // columns N2-2, N2-1 (left of k = 0)
  static constexpr int WR_OFF = WL_OFF + WRAP_BYTES;   // columns 0, 1 (right of k = N2-1)
  static constexpr int STAGE_BYTES = WR_OFF + WRAP_BYTES;
  static constexpr int TX_MAIN = IN_ROWS * ROW_BYTES;
  static constexpr int TX_WRAP = IN_ROWS * 16;
  static constexpr int XP = CK * 8;                    // exchange tile row pitch
  static constexpr int X_BYTES = lf_align128(CJ * XP);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * X_BYTES + 2 * STAGES * 8 + 128;
  static_assert(CJ % R == 0, "rows per thread must divide the level-1 tile");
  static_assert(BJ % 8 == 0, "TMA destinations of the tile and wrap boxes must stay 128-byte aligned");
  static_assert(BOT_OFF % 128 == 0 && WBOT_OFF % 128 == 0, "misaligned TMA destination");
  static_assert(CK <= 256 && BJ <= 256, "TMA box limit");
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
// box b = 0 {CK, 2} halo rows, 1 {CK, BJ} tile rows, 2 {2, 2} wrap corner, 3 {2, BJ} wrap columns
struct LapFusedMaps {
  CUtensorMap m[12];
};

// acc + w*v, rounded separately (ref: Filter.cpp:247-251)
__device__ __forceinline__ double lf_acc(double acc, double w, double v) {
  return __dadd_rn(acc, __dmul_rn(w, v));
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    lap7_fused2_kernel(const __grid_constant__ LapFusedMaps maps, const LapFusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t full = xbuf + 2 * C::X_BYTES;
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
      for (int m = 0; m < 12; ++m) prefetch_tmap(&maps.m[m]);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        const int kt = (int)(w % a.nkt);
        const int jt = (int)((w / a.nkt) % a.njt);
        const int64_t ic = w / ((int64_t)a.nkt * a.njt);
        const int64_t i0 = a.ibeg + ic * a.ci;
        const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
        const int kb = kt * C::BK - 2;  // first level-0 column (-2 for the first k-tile: zero fill)
        const int j0 = jt * C::BJ;
        const int jtop = (j0 == 0) ? (int)a.n1 - 2 : j0 - 2;            // periodic rows j0-2, j0-1
        const int jbot = (j0 + C::BJ >= (int)a.n1) ? 0 : j0 + C::BJ;    // periodic rows j0+BJ, j0+BJ+1
        const bool first_k = (kt == 0), last_k = (kt == a.nkt - 1);
        const uint32_t bytes = C::TX_MAIN + (first_k ? C::TX_WRAP : 0) + (last_k ? C::TX_WRAP : 0);
        for (int64_t p = i0 - 2; p <= i1 + 1; ++p) {
          mbar_wait(empty + 8 * stage, phase ^ 1);
          const uint32_t st = smem + stage * C::STAGE_BYTES;
          const uint32_t fb = full + 8 * stage;
          const int g = (p < 0) ? 4 : (p >= a.nloc ? 8 : 0);
          const int pl = (p < 0) ? a.G + (int)p : (p >= a.nloc ? (int)(p - a.nloc) : (int)p);
          mbar_expect_tx(fb, bytes);
          tma_load_3d(st, &maps.m[g + 0], fb, kb, jtop, pl);
          tma_load_3d(st + C::BODY_OFF, &maps.m[g + 1], fb, kb, j0, pl);
          tma_load_3d(st + C::BOT_OFF, &maps.m[g + 0], fb, kb, jbot, pl);
          if (first_k) {
            const uint32_t wl = st + C::WL_OFF;
            tma_load_3d(wl, &maps.m[g + 2], fb, (int)a.n2 - 2, jtop, pl);
            tma_load_3d(wl + C::WBODY_OFF, &maps.m[g + 3], fb, (int)a.n2 - 2, j0, pl);
            tma_load_3d(wl + C::WBOT_OFF, &maps.m[g + 2], fb, (int)a.n2 - 2, jbot, pl);
          }
          if (last_k) {
            const uint32_t wr = st + C::WR_OFF;
            tma_load_3d(wr, &maps.m[g + 2], fb, 0, jtop, pl);
            tma_load_3d(wr + C::WBODY_OFF, &maps.m[g + 3], fb, 0, j0, pl);
            tma_load_3d(wr + C::WBOT_OFF, &maps.m[g + 2], fb, 0, jbot, pl);
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
  const int q0 = ty * C::R;  // first level-1 row of this thread (level-1 row q = global row j0-1+q)
  const int lane = tid & 31;
  int stage = 0;
  uint32_t phase = 0;
  uint32_t xsel = 0;
  const double w0 = a.w[0], w1 = a.w[1], w2 = a.w[2], w3 = a.w[3], w4 = a.w[4], w5 = a.w[5], w6 = a.w[6];
  const uint32_t cb = tx * 16;  // byte offset of this thread's pair in a tile row
  // in-row neighbours; the outermost columns of level 1 are never used, their out-of-tile
  // neighbours are clamped to something readable
  const uint32_t kmb = (tx == 0) ? cb : cb - 8;
  const uint32_t kpb = (tx == C::TX - 1) ? cb + 8 : cb + 16;
  // exchange-tile addresses (level 1): own rows, the rows above / below (clamped at the tile edge,
  // where the result is never used), the cells left / right
  const uint32_t x_own = q0 * C::XP + cb;
  const uint32_t x_up = (q0 == 0 ? 0 : q0 - 1) * C::XP + cb;
  const uint32_t x_dn = (q0 + C::R >= C::CJ ? C::CJ - 1 : q0 + C::R) * C::XP + cb;
  const uint32_t x_km = q0 * C::XP + kmb;
  const uint32_t x_kp = q0 * C::XP + kpb;
  // level-0 stage row s = level-1 row q + 1
  auto main_row = [](int s) -> uint32_t { return s * C::ROW_BYTES + (s >= 2 ? C::ROW_SKEW : 0); };
  auto wrap_row = [](int s) -> uint32_t { return s * 16 + (s >= 2 ? C::WSKEW : 0); };

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = w / ((int64_t)a.nkt * a.njt);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - 2 + 2 * tx;  // global column of this thread's first cell
    const int64_t j = (int64_t)jt * C::BJ - 1 + q0;      // global row of this thread's first row
    const bool first_k = (kt == 0), last_k = (kt == a.nkt - 1);
    // where level 0 of this thread's pair / left cell / right cell lives: tile rows or a wrap area
    const bool own_wl = first_k && tx == 0, own_wr = last_k && tx == C::TX - 1;
    const bool km_wl = first_k && tx == 1, kp_wr = last_k && tx == C::TX - 2;
    const bool store_cols = worker && tx >= 1 && tx <= C::TX - 2;

    double2 below0[C::R], part1[C::R], below1[C::R], part2[C::R];
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      below0[r] = make_double2(0.0, 0.0);
      part1[r] = make_double2(0.0, 0.0);
      below1[r] = make_double2(0.0, 0.0);
      part2[r] = make_double2(0.0, 0.0);
    }

    for (int64_t p = i0 - 2; p <= i1 + 1; ++p) {
      mbar_wait(full + 8 * stage, phase);
      const uint32_t st = smem + stage * C::STAGE_BYTES;
      double2 c[C::R], up, dn;
      double km[C::R], kp[C::R];
      {
        auto pair_at = [&](int s) -> double2 {
          const uint32_t ad = own_wl ? st + C::WL_OFF + wrap_row(s)
                                     : (own_wr ? st + C::WR_OFF + wrap_row(s) : st + main_row(s) + cb);
          return lds_v2(ad);
        };
        up = pair_at(q0);
        dn = pair_at(q0 + C::R + 1);
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const int s = q0 + 1 + r;
          c[r] = pair_at(s);
          km[r] = lds_f64(km_wl ? st + C::WL_OFF + wrap_row(s) + 8 : st + main_row(s) + kmb);
          kp[r] = lds_f64(kp_wr ? st + C::WR_OFF + wrap_row(s) : st + main_row(s) + kpb);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 8 * stage);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }

      // level 1 of plane p-1: the (+1,0,0) branch is this plane's centre
      double2 l1[C::R];
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        l1[r].x = lf_acc(part1[r].x, w6, c[r].x);
        l1[r].y = lf_acc(part1[r].y, w6, c[r].y);
      }
      // level 2 of plane p-2: the (+1,0,0) branch is level 1 of plane p-1 -- done, store it
      if (p >= i0 + 2 && store_cols) {
        double* orow = a.out + ((p - 2) * a.n1 + j) * a.n2 + k;
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const int q = q0 + r;
          if (q >= 1 && q <= C::CJ - 2)
            st_global_v2(orow + (int64_t)r * a.n2, lf_acc(part2[r].x, w6, l1[r].x),
                         lf_acc(part2[r].y, w6, l1[r].y));
        }
      }
      // hand level 1 of plane p-1 to the neighbours
      const uint32_t xb = xbuf + xsel * C::X_BYTES;
      xsel ^= 1;
      if (worker) {
#pragma unroll
        for (int r = 0; r < C::R; ++r) sts_v2(xb + x_own + r * C::XP, l1[r].x, l1[r].y);
      }
      // first six branches of level 1 of plane p (keeps the FP64 pipe busy while warps gather)
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        const double2 jm = (r == 0) ? up : c[r - 1];
        const double2 jp = (r == C::R - 1) ? dn : c[r + 1];
        double x = 0.0, y = 0.0;
        x = lf_acc(x, w0, below0[r].x);  y = lf_acc(y, w0, below0[r].y);
        x = lf_acc(x, w1, jm.x);         y = lf_acc(y, w1, jm.y);
        x = lf_acc(x, w2, km[r]);        y = lf_acc(y, w2, c[r].x);
        x = lf_acc(x, w3, c[r].x);       y = lf_acc(y, w3, c[r].y);
        x = lf_acc(x, w4, c[r].y);       y = lf_acc(y, w4, kp[r]);
        x = lf_acc(x, w5, jp.x);         y = lf_acc(y, w5, jp.y);
        part1[r] = make_double2(x, y);
        below0[r] = c[r];
      }
      named_bar_sync(1, C::CONSUMERS);
      // first six branches of level 2 of plane p-1
      {
        const double2 up1 = lds_v2(xb + x_up);
        const double2 dn1 = lds_v2(xb + x_dn);
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          km[r] = lds_f64(xb + x_km + r * C::XP);
          kp[r] = lds_f64(xb + x_kp + r * C::XP);
        }
#pragma unroll
        for (int r = 0; r < C::R; ++r) {
          const double2 jm = (r == 0) ? up1 : l1[r - 1];
          const double2 jp = (r == C::R - 1) ? dn1 : l1[r + 1];
          double x = 0.0, y = 0.0;
          x = lf_acc(x, w0, below1[r].x);  y = lf_acc(y, w0, below1[r].y);
          x = lf_acc(x, w1, jm.x);         y = lf_acc(y, w1, jm.y);
          x = lf_acc(x, w2, km[r]);        y = lf_acc(y, w2, l1[r].x);
          x = lf_acc(x, w3, l1[r].x);      y = lf_acc(y, w3, l1[r].y);
          x = lf_acc(x, w4, l1[r].y);      y = lf_acc(y, w4, kp[r]);
          x = lf_acc(x, w5, jp.x);         y = lf_acc(y, w5, jp.y);
          part2[r] = make_double2(x, y);
          below1[r] = l1[r];
        }
      }
    }
  }
}

// ---- configurations ---------------------------------------------------------------------
typedef void (*LapFusedKernel)(const LapFusedMaps, const LapFusedArgs);
struct LapFusedConfig {
  int BJ, BK, CK, threads, smem;
  LapFusedKernel kernel;
  const char* name;
};
template <class C>
constexpr LapFusedConfig make_lapf(const char* name) {
  return LapFusedConfig{C::BJ, C::BK, C::CK, C::THREADS, C::SMEM_BYTES, lap7_fused2_kernel<C>, name};
}
// index 0 is the default; the rest are tuning alternatives (env FDB_LAPF_CFG)
const LapFusedConfig kLapFused[] = {
    make_lapf<LapFusedCfg<16, 3, 4>>("bj16_r3_s4"),
    make_lapf<LapFusedCfg<16, 3, 3>>("bj16_r3_s3"),
    make_lapf<LapFusedCfg<16, 6, 4>>("bj16_r6_s4"),
    make_lapf<LapFusedCfg<16, 2, 4>>("bj16_r2_s4"),
    make_lapf<LapFusedCfg<8, 5, 4, 2>>("bj8_r5_s4_2cta"),
    make_lapf<LapFusedCfg<8, 2, 6>>("bj8_r2_s6"),
    make_lapf<LapFusedCfg<16, 3, 6>>("bj16_r3_s6"),
    make_lapf<LapFusedCfg<16, 6, 6>>("bj16_r6_s6"),
    make_lapf<LapFusedCfg<8, 5, 8>>("bj8_r5_s8"),
    make_lapf<LapFusedCfg<16, 3, 8>>("bj16_r3_s8"),
};
constexpr int kNumLapFused = sizeof(kLapFused) / sizeof(kLapFused[0]);

int lf_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// the first configuration (from the forced or default one on) whose tile divides the plane
const LapFusedConfig* lapf_pick(const Field& f) {
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  int first = lf_env_int("FDB_LAPF_CFG", 0);
  if (first < 0 || first >= kNumLapFused) first = 0;
  for (int t = 0; t < kNumLapFused; ++t) {
    const LapFusedConfig& C = kLapFused[(first + t) % kNumLapFused];
    if (n1 % C.BJ == 0 && n2 % C.BK == 0) return &C;
  }
  return nullptr;
}

struct LapFusedAttr {
  const LapFusedConfig* cfg = nullptr;
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
  const int boxes[4][2] = {{C.CK, 2}, {C.CK, C.BJ}, {2, 2}, {2, C.BJ}};
  for (int t = 0; t < 3; ++t)
    for (int b = 0; b < 4; ++b)
      FDB_TRY(encode_tensor_map_3d(&out->m[4 * t + b], base[t], n2, n1, planes[t], boxes[b][0], boxes[b][1]));
  return FDB_OK;
}

int launch_stencil_lap7_fused(Field& f, int d, int X, int64_t ibeg, int64_t iend, const StencilBranches& b,
                              cudaStream_t s) {
  if (iend <= ibeg) return FDB_OK;
  Slab& sl = f.slabs[d];
  const LapFusedConfig* C = lapf_pick(f);
  if (!C) return set_error(FDB_E_INVALID, "no fused 7-point tile divides a %lld x %lld plane",
                           (long long)f.geo.n[1], (long long)f.geo.n[2]);
  LapFusedAttr& at = g_lapf_attr[sl.device & 15];
  if (at.cfg != C) {
    FDB_CUDA(cudaFuncSetAttribute(C->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C->smem));
    int nb = 0;
    FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, C->kernel, C->threads, C->smem));
    if (nb < 1) return set_error(FDB_E_CUDA, "fused 7-point kernel %s does not fit on an SM", C->name);
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, sl.device));
    at.ctas_per_sm = nb;
    at.sms = prop.multiProcessorCount;
    at.cfg = C;
  }
  if (sl.lapf_cfg != (const void*)C) {
    for (int p = 0; p < 2; ++p) FDB_TRY(lapf_maps(f, d, *C, p, reinterpret_cast<LapFusedMaps*>(sl.lapf_maps[p])));
    sl.lapf_cfg = (const void*)C;
  }
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
    // every work item warms up on four extra planes: favour long chunks
    ci = 128;
    while (ci > 8 && tiles * ((planes + ci - 1) / ci) < 2 * grid_max) ci /= 2;
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nwork = tiles * ((planes + ci - 1) / ci);
  const int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  C->kernel<<<(unsigned)grid, C->threads, C->smem, s>>>(*reinterpret_cast<const LapFusedMaps*>(sl.lapf_maps[X]), a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

const char* stencil_lap7_fused_name(const Field& f) {
  const LapFusedConfig* C = lapf_pick(f);
  return C ? C->name : "";
}

}  // namespace fdb
