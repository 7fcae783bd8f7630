// kernels_fused.cu -- temporal blocking for the 3-D upwind step: T time steps per sweep.
//
// The single-step kernel (kernels_tma.cu) is pinned to the HBM roofline: 16 B of DRAM
// traffic per cell-update.  The upwind stencil only looks at LOWER indices (i-1, j-1, k-1),
// so T steps can be fused with one-sided halos and no look-ahead along the marching axis:
// when plane i of the input (level 0) lands in shared memory, level 1 of plane i follows from
// level 0 of planes i and i-1, level 2 from level 1 of planes i and i-1, ... up to level T,
// which is the only one written back.  DRAM traffic per cell-update drops to ~16/T bytes.
//
//   * each consumer thread owns R rows x 2 cells of a CJ x CK compute tile and keeps plane
//     i-1 of every level 0..T-1 in registers (T*R double2);
//   * in-plane neighbours (j-1, k-1) of levels >= 1 travel through two ping-pong exchange
//     tiles in shared memory, one named barrier among the consumer warps per level;
//   * the tile computes T-1 halo rows/columns redundantly on its low sides (level s is valid
//     from compute row/column s-1 on); only the BJ x BK interior of level T is stored, so
//     output tiles stay 128-byte aligned along k;
//   * a work item warms up on T extra planes below its chunk (levels become valid one plane
//     after the other), nothing is stored for them; planes below the slab come from the ghost
//     tensor (depth T), which aliases the far planes on a single device.
//
// Shared-memory layout (the second one of round 1; the first addressed rows through per-row
// selects and cost ~20 % more instructions, profiles/r01l_*):
//   * ONE row pitch for the whole stage.  Rows are BKP = BK + 8 cells wide (global columns
//     k0-8 .. k0+BK-1) so that two rows are a multiple of 128 bytes; the halo box holds HR rows
//     (T rounded up to even) and the tile rows follow it directly.  Every shared-memory access of a
//     consumer thread is `base + immediate`: no per-row selects, no per-plane address arithmetic.
//   * the periodic wrap columns of the first k-tile (an 8-cell box at column N2-8) are copied into
//     the zero-filled cells of the tile rows by the LOADER warp, which issues the TMA boxes and,
//     a few planes later, hands each stage to the consumers; the consumers never see the wrap.
//   * the exchange tiles use the same pitch and column offsets, with a spare row on top, so the
//     stage and the exchange tile are addressed from the same thread base.
//   * output addresses advance by one plane per iteration instead of being rebuilt.
//
// Arithmetic per level is exactly the single step's (separately rounded, reference order,
// ref: upwind/cxx/upwind.cxx:72-80), so T fused steps are bit-identical to T single steps.
// tests/host_model_fused.py restates the tile pipeline with numpy and is pinned to the oracle.
#include <algorithm>

#include "fdb_internal.h"
#include "tma_ptx.cuh"

namespace fdb {

namespace {

using namespace ptx;

constexpr int fz_align128(int x) { return (x + 127) / 128 * 128; }

template <int T_, int CJ_, int R_, int STAGES_, int BK_ = 128, int MINB_ = 1>
struct FusedCfg {
  static constexpr int T = T_, CJ = CJ_, R = R_, STAGES = STAGES_, BK = BK_, MINB = MINB_;
  static constexpr int HKC = 2 * (T / 2);          // redundant compute columns on the left (even, >= T-1)
  static constexpr int CK = BK + HKC;              // compute columns: global k0-HKC .. k0+BK-1
  static constexpr int TX = CK / 2;                // threads per row (2 cells each)
  static constexpr int TY = CJ / R;
  static constexpr int WORKERS = TX * TY;
  static constexpr int CONSUMERS = (WORKERS + 31) / 32 * 32;
  static constexpr int CONSUMER_WARPS = CONSUMERS / 32;
  static constexpr int THREADS = CONSUMERS + 32;   // + the loader warp
  static constexpr int BJ = CJ - (T - 1);          // output rows per tile: global j0 .. j0+BJ-1
  static constexpr int HR = (T + 1) / 2 * 2;       // halo rows loaded above the tile: global j0-HR .. j0-1
  static constexpr int IN_ROWS = HR + BJ;          // stage rows; compute row q is stage row q + HR - T + 1
  static constexpr int BKP = BK + 8;               // stage columns: global k0-8 .. k0+BK-1
  static constexpr int LEFT = 8 - HKC;             // stage column of the first compute column
  static constexpr int PITCH = BKP * 8;
  static constexpr int BODY_OFF = HR * PITCH;
  static constexpr int MAIN_BYTES = IN_ROWS * PITCH;
  static constexpr int WPITCH = 64;                // wrap box rows: 8 cells, global columns N2-8 .. N2-1
  static constexpr int W_OFF = fz_align128(MAIN_BYTES);
  static constexpr int STAGE_BYTES = fz_align128(W_OFF + IN_ROWS * WPITCH);
  static constexpr int TX_MAIN = IN_ROWS * PITCH;
  static constexpr int TX_WRAP = IN_ROWS * WPITCH;
  static constexpr int X_BYTES = fz_align128((CJ + 1) * PITCH);  // exchange tile: compute row q at row q + 1
  static constexpr int NX = (T > 1) ? 2 : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NX * X_BYTES + 3 * STAGES * 8 + 128;
  static constexpr int LAG = STAGES - 2;
  static_assert(CJ % R == 0 && T >= 2 && T <= 4 && HKC >= T - 1 && CJ > T, "bad fused tile");
  static_assert((HR * PITCH) % 128 == 0 && (HR * WPITCH) % 128 == 0, "misaligned TMA destination");
  static_assert(BKP <= 256 && BJ <= 256, "TMA box limit");
  static_assert(IN_ROWS <= 32, "one loader lane per stage row");
  static_assert(STAGES >= 3, "the loader needs two stages of slack");
  static_assert(THREADS <= 1024, "too many threads");
  static_assert(SMEM_BYTES <= 227 * 1024, "more shared memory than a CTA can have");
};

struct FusedArgs {
  double* out;
  int64_t n1, n2;
  int64_t ibeg, iend;
  int ci, njt, nkt;
  int64_t nwork;
  int G;  // planes in the ghost tensor; local plane p < 0 is its plane G + p
  double c0, c1, c2;
  double* peer_out;   // planes p >= peer_from are also stored here (the next slab's ghost planes), or null
  int64_t peer_from;
  int64_t plane;      // n1 * n2
  // single-launch sweeps on a slab ring (runtime.cu: sweep_device_direct): the chunks are walked top chunk first, the
  // work items that hold pushed planes wait for the neighbour's ACK before their first peer store, and the last of
  // them to finish raises the neighbour's ghost flag -- no separate boundary launch, no host or stream in between
  int nchunks, reverse;           // chunks along the marching axis; reverse: chunk nchunks-1 first
  unsigned int* done_ctr;         // items with pushed planes finished so far (device memory of this slab), or null
  unsigned int done_target;
  unsigned long long* nbr_flag;   // the next slab's F_GHOST_LO counter (peer-mapped) <- flag_value when all are done
  unsigned long long flag_value;
  const unsigned long long* ack_flag;  // this slab's F_ACK_NEXT counter: wait for >= ack_value before the first peer store
  unsigned long long ack_value;
  const unsigned long long* ghost_flag;  // this slab's F_GHOST_LO counter: the loader waits for >= ghost_value before
  unsigned long long ghost_value;        // its first load from the ghost planes (null: the stream waited already)
  int64_t peer_delta;  // byte distance from a stored cell of `out` to the same cell in the next slab's ghost planes
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// chunk of work item w along the marching axis (top chunk first on single-launch ring sweeps)
__device__ __forceinline__ int64_t fz_chunk_of(const FusedArgs& a, int64_t w) {
  const int64_t ic = w / ((int64_t)a.nkt * a.njt);
  return a.reverse ? (int64_t)a.nchunks - 1 - ic : ic;
}

// Tensor maps: m[0..3] over the local planes, m[4..7] over the ghost planes below; box shapes
//   0/4: {BKP, HR} halo rows   1/5: {BKP, BJ} tile rows   2/6: {8, HR} wrap corner   3/7: {8, BJ} wrap columns
struct FusedMaps {
  CUtensorMap m[8];
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

// position of the loader in the sequence of (work item, plane) pairs of this CTA
template <class C>
struct FusedCursor {
  int64_t w, p, i1;
  int kt, jt;
  __device__ __forceinline__ void open(const FusedArgs& a) {
    kt = (int)(w % a.nkt);
    jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = fz_chunk_of(a, w);
    const int64_t i0 = a.ibeg + ic * a.ci;
    i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    p = i0 - C::T;
  }
  __device__ __forceinline__ bool valid(const FusedArgs& a) const { return w < a.nwork; }
  __device__ __forceinline__ void next(const FusedArgs& a) {
    if (++p >= i1) {
      w += gridDim.x;
      if (w < a.nwork) open(a);
    }
  }
};

// ---- loader warp (shared by both consumer formulations) --------------------------------------------
// Lane 0 issues the TMA boxes of plane n; LAG planes later the warp waits for a plane to land, one lane per
// stage row patches the periodic wrap columns of the first k-tile, and the stage is handed to the consumers.
template <class C>
__device__ __forceinline__ void fused_loader_warp(const FusedMaps& maps, const FusedArgs& a, uint32_t smem, uint32_t landed,
                                                  uint32_t full, uint32_t empty, int lane) {
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < 8; ++m) prefetch_tmap(&maps.m[m]);
    }
    FusedCursor<C> ci, cf;  // issue / hand-over
    ci.w = blockIdx.x;
    if (ci.valid(a)) ci.open(a);
    cf = ci;
    int si = 0, sf = 0;
    uint32_t phi = 0, phf = 0;
    int ahead = 0;
    bool ghosts_seen = false;
    while (cf.valid(a)) {
      if (ci.valid(a)) {
        mbar_wait(empty + 8 * si, phi ^ 1);
        if (lane == 0) {
          if (ci.p < 0 && a.ghost_flag != nullptr && !ghosts_seen) {
            // single-launch ring sweep: the ghost planes are the previous slab's peer stores; its kernel raises this
            // counter behind the last of them (the bottom chunk comes last here, so this rarely spins)
            while (ld_acquire_sys_u64(a.ghost_flag) < a.ghost_value) {}
            fence_proxy_async_all();  // ... and the TMA reads below go through the async proxy
            ghosts_seen = true;
          }
          const uint32_t st = smem + si * C::STAGE_BYTES;
          const uint32_t lb = landed + 8 * si;
          const int kb = ci.kt * C::BK - 8;  // first stage column (negative for the first k-tile: zero fill)
          const int j0 = ci.jt * C::BJ;
          const int jh = (j0 - C::HR < 0) ? j0 - C::HR + (int)a.n1 : j0 - C::HR;  // periodic halo rows
          const bool first_k = (ci.kt == 0);
          const int64_t p = ci.p;
          const int g = (p < 0) ? 4 : 0;  // ghost tensor below the slab
          const int pl = (p < 0) ? a.G + (int)p : (int)p;
          mbar_expect_tx(lb, C::TX_MAIN + (first_k ? C::TX_WRAP : 0));
          tma_load_3d(st, &maps.m[g + 0], lb, kb, jh, pl);
          tma_load_3d(st + C::BODY_OFF, &maps.m[g + 1], lb, kb, j0, pl);
          if (first_k) {
            tma_load_3d(st + C::W_OFF, &maps.m[g + 2], lb, (int)a.n2 - 8, jh, pl);
            tma_load_3d(st + C::W_OFF + C::HR * C::WPITCH, &maps.m[g + 3], lb, (int)a.n2 - 8, j0, pl);
          }
        }
        ci.next(a);
        if (++si == C::STAGES) { si = 0; phi ^= 1; }
        ++ahead;
      }
      if (ahead > C::LAG || !ci.valid(a)) {
        mbar_wait(landed + 8 * sf, phf);
        if (cf.kt == 0 && lane < C::IN_ROWS) {
          // columns -6 .. -1 of the tile rows <- columns N2-6 .. N2-1 (a sweep of T <= 4 steps reads
          // back to column -(HKC + 1) >= -5)
          const uint32_t st = smem + sf * C::STAGE_BYTES;
          const double2 v0 = lds_v2(st + C::W_OFF + lane * C::WPITCH + 16);
          const double2 v1 = lds_v2(st + C::W_OFF + lane * C::WPITCH + 32);
          const double2 v2 = lds_v2(st + C::W_OFF + lane * C::WPITCH + 48);
          sts_v2(st + lane * C::PITCH + 16, v0.x, v0.y);
          sts_v2(st + lane * C::PITCH + 32, v1.x, v1.y);
          sts_v2(st + lane * C::PITCH + 48, v2.x, v2.y);
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

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    upwind3d_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t landed = xbuf + C::NX * C::X_BYTES;  // TMA bytes of the stage have arrived
  const uint32_t full = landed + C::STAGES * 8;        // ... and its wrap columns are in place
  const uint32_t empty = full + C::STAGES * 8;         // every consumer warp has read the stage

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
    fused_loader_warp<C>(maps, a, smem, landed, full, empty, lane);
    return;
  }

  // ===================== consumer warps =====================
  const bool worker = tid < C::WORKERS;  // threads past the tile only keep the barriers company
  const int wid = worker ? tid : 0;
  const int tx = wid % C::TX;
  const int ty = wid / C::TX;
  const int q0 = ty * C::R;  // first compute row of this thread (compute row q = global row j0-(T-1)+q)
  int stage = 0;
  uint32_t phase = 0;
  uint32_t xsel = 0;
  const double c0 = a.c0, c1 = a.c1, c2 = a.c2;
  constexpr uint32_t P = C::PITCH;
  // thread base: the row above the thread's first row, its own pair.  Exchange tile: compute row q
  // at tile row q + 1; stage: compute row q at stage row q + HR - T + 1.
  const uint32_t xt = q0 * P + (C::LEFT + 2 * tx) * 8;
  const uint32_t tb = xt + (C::HR - C::T) * P;
  // rows of this thread that belong to the output tile (compute rows T-1 .. CJ-1)
  uint32_t rowmask = 0;
#pragma unroll
  for (int r = 0; r < C::R; ++r)
    if (q0 + r >= C::T - 1) rowmask |= 1u << r;
  const int64_t plane = a.n1 * a.n2;

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = fz_chunk_of(a, w);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - C::HKC + 2 * tx;   // global column of this thread's first cell
    const int64_t j = (int64_t)jt * C::BJ - (C::T - 1) + q0;   // global row of this thread's first row
    const bool store_cols = worker && (2 * tx >= C::HKC) && (k < a.n2);
    // rows of the output tile that exist (the last j-tile may be ragged)
    uint32_t rmask = rowmask;
#pragma unroll
    for (int r = 0; r < C::R; ++r)
      if (j + r >= a.n1) rmask &= ~(1u << r);
    // offset of (plane p, row j, column k), advanced by one plane per iteration
    int64_t ooff = ((i0 - C::T - 1) * a.n1 + j) * a.n2 + k;

    double2 carry[C::T][C::R];  // plane i-1 of levels 0..T-1
#pragma unroll
    for (int s = 0; s < C::T; ++s)
#pragma unroll
      for (int r = 0; r < C::R; ++r) carry[s][r] = make_double2(0.0, 0.0);

    for (int64_t p = i0 - C::T; p < i1; ++p) {
      ooff += plane;
      mbar_wait(full + 8 * stage, phase);  // the loader's hand-over (see Lean::step)
      const uint32_t sb = smem + stage * C::STAGE_BYTES + tb;
      double2 v[C::R];
      double km[C::R];
      double2 up = lds_v2(sb);
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
        v[r] = lds_v2(sb + (1 + r) * P);
        km[r] = lds_f64(sb + (1 + r) * P - 8);
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
            double* orow = a.out + ooff;
            const bool push = (a.peer_out != nullptr) && (p >= a.peer_from);
            double* prow = a.peer_out + (ooff - a.peer_from * plane);
#pragma unroll
            for (int r = 0; r < C::R; ++r) {
              if ((rmask >> r) & 1u) {
                st_global_v2(orow + (int64_t)r * a.n2, nv[r].x, nv[r].y);
                if (push) st_global_v2(prow + (int64_t)r * a.n2, nv[r].x, nv[r].y);
              }
            }
          }
        } else {
          // hand level s+1 of this plane to the neighbours through the exchange tile
          const uint32_t xb = xbuf + xsel * C::X_BYTES + xt;
          xsel ^= 1;
          if (worker) {
#pragma unroll
            for (int r = 0; r < C::R; ++r) sts_v2(xb + (1 + r) * P, nv[r].x, nv[r].y);
          }
          named_bar_sync(1, C::CONSUMERS);
          up = lds_v2(xb);
#pragma unroll
          for (int r = 0; r < C::R; ++r) {
            km[r] = lds_f64(xb + (1 + r) * P - 8);
            v[r] = nv[r];
          }
        }
      }
    }
  }
}

// ---- second consumer formulation ("lean"): same tile pipeline, same arithmetic, fewer instructions --------
// Measured on B200 (profiles/r02e_ubench_fp64_issue.txt): an FP64 instruction holds the issue port of its SM
// sub-partition for two cycles and nothing else issues beside it, so the kernel's time is the SUM of
// 2 x (FP64 instructions) + (everything else), and the first formulation spent 150 other instructions per 162
// FP64 ones per warp and plane (44 of them register moves that shift plane i into the "plane i-1" registers).
// Here
//   * the plane loop is unrolled by two and the two register sets swap roles (previous plane / this plane):
//     no moves;
//   * plane counters are 32-bit, the stage/barrier addresses advance by constants, output rows are R pointers that
//     advance by one plane, idle threads of the last warp duplicate thread 0 instead of being predicated off, and
//     the peer stores of the halo push live in their own instantiation (PUSH).
//   * SPLIT: the exchange tile keeps the even cells (x) and the odd cells (y) of a row in two contiguous halves and
//     only what a neighbour reads is stored (every row's y for the right-hand neighbour's k-1 operand, the last
//     row's x for the j-1 operand of the thread below): the 8-byte accesses are conflict-free (2 shared-memory
//     wavefronts per warp instead of the 4 of a 16-byte lane stride) -- 18 instead of 28 wavefronts per exchange.
template <class C, bool PUSH, bool SPLIT>
struct Lean {
  static constexpr int T = C::T, R = C::R;
  static constexpr uint32_t P = C::PITCH;
  static constexpr uint32_t YOFF = C::PITCH / 2;   // y half of an exchange-tile row (TX <= PITCH / 16)
  static_assert(!SPLIT || (C::TX + 1) * 8 <= C::PITCH / 2, "exchange row halves too narrow");

  struct State {
    uint32_t st;      // this thread's base inside the current stage
    uint32_t bar;     // `full` barrier of the current stage (empty = bar + 8*STAGES)
    uint32_t par;     // phase parity of the current round over the stages
    uint32_t st0, bar0, bar_end;
    uint32_t xt0, xt1;  // this thread's base inside the two exchange tiles
    double c0, c1, c2;
    double* orow[C::R];   // output rows of the plane being computed
    int64_t plane_elems;
    int64_t peer_delta;   // PUSH: byte distance to the same cell in the next slab's ghost planes
    uint32_t smask;       // rows x columns this thread stores
    int lane;
  };

  // one plane: level 0 comes from the stage, level s+1 from level s of this plane (X) and of the previous one (Pv)
  template <int PARITY>
  static __device__ __forceinline__ void step(State& z, double2 (&Pv)[C::T][C::R], double2 (&X)[C::T][C::R], bool store,
                                              bool push) {
    // the loader warp hands the stage over once its TMA bytes have landed (it waits on `landed`, an acquire) and the wrap
    // columns are patched: its arrive on `full` (release) / this wait (acquire) order the TMA writes before the loads
    // below.  The consumers used to wait on `landed` as well; the hand-over already implies it, and compute-sanitizer's
    // synccheck reports a transaction barrier that two groups of threads wait on ("Missing init", profiles/r02s_*).
    mbar_wait(z.bar, z.par);
    double km[R];
    double2 up = lds_v2(z.st);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      X[0][r] = lds_v2(z.st + (1 + r) * P);
      km[r] = lds_f64(z.st + (1 + r) * P - 8);
    }
    __syncwarp();
    if (z.lane == 0) mbar_arrive(z.bar + 8 * C::STAGES);
    z.st += C::STAGE_BYTES;
    z.bar += 8;
    if (z.bar == z.bar_end) { z.st = z.st0; z.bar = z.bar0; z.par ^= 1; }
    const uint32_t m = store ? z.smask : 0u;
#pragma unroll
    for (int s = 0; s < T; ++s) {
      double2 out[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double2 jm = (r == 0) ? up : X[s][r - 1];
        double2 n;
        n.x = upwind_cell(X[s][r].x, Pv[s][r].x, jm.x, km[r], z.c0, z.c1, z.c2);
        n.y = upwind_cell(X[s][r].y, Pv[s][r].y, jm.y, X[s][r].x, z.c0, z.c1, z.c2);
        if (s == T - 1) out[r] = n; else X[s + 1][r] = n;
      }
      if (s == T - 1) {
        if (PUSH && push) {  // uniform over the CTA: the planes the next slab reads as ghosts
#pragma unroll
          for (int r = 0; r < R; ++r)
            if ((m >> r) & 1u)
              st_global_v2(reinterpret_cast<double*>(reinterpret_cast<char*>(z.orow[r]) + z.peer_delta), out[r].x, out[r].y);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((m >> r) & 1u) st_global_v2(z.orow[r], out[r].x, out[r].y);
          z.orow[r] += z.plane_elems;
        }
      } else {
        // hand level s+1 of this plane to the neighbours; the tiles alternate with every exchange
        const uint32_t xb = (((PARITY * (T - 1) + s) & 1) == 0) ? z.xt0 : z.xt1;
        if (SPLIT) {
          // z.xt0/z.xt1: row above the thread's first row, x half, this thread's slot (8 bytes per thread)
#pragma unroll
          for (int r = 0; r < R; ++r) sts_f64(xb + (1 + r) * P + YOFF, X[s + 1][r].y);
          sts_f64(xb + R * P, X[s + 1][R - 1].x);
          named_bar_sync(1, C::CONSUMERS);
          up.x = lds_f64(xb);
          up.y = lds_f64(xb + YOFF);
#pragma unroll
          for (int r = 0; r < R; ++r) km[r] = lds_f64(xb + (1 + r) * P + YOFF - 8);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) sts_v2(xb + (1 + r) * P, X[s + 1][r].x, X[s + 1][r].y);
          named_bar_sync(1, C::CONSUMERS);
          up = lds_v2(xb);
#pragma unroll
          for (int r = 0; r < R; ++r) km[r] = lds_f64(xb + (1 + r) * P - 8);
        }
      }
    }
  }
};

template <class C, bool PUSH, bool SPLIT>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
    upwind3d_fused_lean_kernel(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t xbuf = smem + C::STAGES * C::STAGE_BYTES;
  const uint32_t landed = xbuf + C::NX * C::X_BYTES;
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
    fused_loader_warp<C>(maps, a, smem, landed, full, empty, lane);
    return;
  }

  using L = Lean<C, PUSH, SPLIT>;
  // threads past the tile repeat thread 0's work (same values to the same shared-memory cells) and store nothing
  const bool worker = tid < C::WORKERS;
  const int wid = worker ? tid : 0;
  const int tx = wid % C::TX;
  const int ty = wid / C::TX;
  const int q0 = ty * C::R;
  const uint32_t xt = q0 * C::PITCH + (C::LEFT + 2 * tx) * 8;
  typename L::State z;
  z.st0 = smem + xt + (C::HR - C::T) * C::PITCH;
  z.bar0 = full;
  z.bar_end = full + 8 * C::STAGES;
  z.st = z.st0;
  z.bar = z.bar0;
  z.par = 0;
  // exchange tiles: interleaved rows (same offsets as the stage) or split halves (8 bytes per thread from the row start)
  const uint32_t xx = SPLIT ? (uint32_t)(q0 * C::PITCH + 8 + tx * 8) : xt;
  z.xt0 = xbuf + xx;
  z.xt1 = xbuf + C::X_BYTES + xx;
  z.c0 = a.c0; z.c1 = a.c1; z.c2 = a.c2;
  z.plane_elems = a.n1 * a.n2;
  z.peer_delta = PUSH ? a.peer_delta : 0;
  z.lane = lane;
  uint32_t rowmask = 0;
#pragma unroll
  for (int r = 0; r < C::R; ++r)
    if (worker && q0 + r >= C::T - 1) rowmask |= 1u << r;

  for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
    const int kt = (int)(w % a.nkt);
    const int jt = (int)((w / a.nkt) % a.njt);
    const int64_t ic = fz_chunk_of(a, w);
    const int64_t i0 = a.ibeg + ic * a.ci;
    const int64_t i1 = (i0 + a.ci < a.iend) ? i0 + a.ci : a.iend;
    const int64_t k = (int64_t)kt * C::BK - C::HKC + 2 * tx;
    const int64_t j = (int64_t)jt * C::BJ - (C::T - 1) + q0;
    z.smask = (2 * tx >= C::HKC && k < a.n2) ? rowmask : 0u;
#pragma unroll
    for (int r = 0; r < C::R; ++r) {
      if (j + r >= a.n1) z.smask &= ~(1u << r);
      // rows of plane i0 - T (the first warm-up plane: never dereferenced before plane i0)
      z.orow[r] = a.out + ((i0 - C::T) * a.n1 + j + r) * a.n2 + k;
    }
    double2 A[C::T][C::R], B[C::T][C::R];  // levels 0..T-1 of the previous plane / of this plane, swapping roles
#pragma unroll
    for (int s = 0; s < C::T; ++s)
#pragma unroll
      for (int r = 0; r < C::R; ++r) A[s][r] = B[s][r] = make_double2(0.0, 0.0);

    // T warm-up planes below the chunk, then the chunk; `q` counts planes from i0 - T
    const int np = (int)(i1 - i0) + C::T;
    const int push_from = PUSH ? (int)(a.peer_from - (i0 - C::T)) : 0;
    // single-launch ring sweep: this item holds planes the next slab reads as ghosts
    const bool signals = PUSH && a.done_ctr != nullptr && i1 > a.peer_from;
    const int ack_at = push_from > C::T ? push_from : C::T;  // first plane-step that stores to the peer
    int q = 0;
    for (; q + 2 <= np; q += 2) {
      if (signals && (q == ack_at || q + 1 == ack_at)) {
        // WAR on the neighbour's ghost planes: it has finished the sweep that read them (its ACK, written by a stream
        // memory operation behind that sweep, lands in this slab's memory)
        if (tid == 0)
          while (ld_acquire_sys_u64(a.ack_flag) < a.ack_value) {}
        __syncwarp();
        named_bar_sync(1, C::CONSUMERS);
      }
      L::template step<0>(z, A, B, q >= C::T, q >= push_from);
      L::template step<1>(z, B, A, q + 1 >= C::T, q + 1 >= push_from);
    }
    if (q < np) {
      if (signals && q == ack_at) {
        if (tid == 0)
          while (ld_acquire_sys_u64(a.ack_flag) < a.ack_value) {}
        __syncwarp();
        named_bar_sync(1, C::CONSUMERS);
      }
      // odd plane count: one more plane in the even roles (the next item starts from zeroed sets anyway)
      L::template step<0>(z, A, B, q >= C::T, q >= push_from);
      // with an odd number of exchanges per plane the next item would write the exchange tile this plane just read
      if ((C::T - 1) & 1) named_bar_sync(1, C::CONSUMERS);
    }
    if (signals) {
      // every peer store of this item is issued; the last item to get here publishes them to the neighbour
      named_bar_sync(1, C::CONSUMERS);
      if (tid == 0) {
        __threadfence_system();
        const unsigned int done = atomicAdd(a.done_ctr, 1u) + 1u;
        if (done == a.done_target) {
          __threadfence_system();
          *a.done_ctr = 0u;  // the next sweep's kernel is ordered behind this one
          st_release_sys_u64(a.nbr_flag, a.flag_value);
        }
      }
      __syncwarp();
    }
  }
}

// ---- configurations ---------------------------------------------------------------------
typedef void (*FusedKernel)(const FusedMaps, const FusedArgs);
struct FusedConfig {
  int T, CJ, BJ, BK, BKP, HR, threads, smem;
  FusedKernel kernel;                 // first consumer formulation (FDB_FUSED_IMPL=1)
  FusedKernel lean, lean_push;        // lean formulation, without / with the peer stores of the halo push
  FusedKernel split, split_push;      // lean formulation with the split exchange-tile layout
  const char* name;
};
template <class C>
constexpr FusedConfig make_fused(const char* name) {
  return FusedConfig{C::T, C::CJ, C::BJ, C::BK, C::BKP, C::HR, C::THREADS, C::SMEM_BYTES,
                      upwind3d_fused_kernel<C>, upwind3d_fused_lean_kernel<C, false, false>, upwind3d_fused_lean_kernel<C, true, false>,
                      upwind3d_fused_lean_kernel<C, false, true>, upwind3d_fused_lean_kernel<C, true, true>, name};
}
// per T: index 0 is the default, the rest are tuning alternatives (env FDB_FUSED_CFG)
// per T: index 0 is the default, the rest are tuning alternatives (env FDB_FUSED_CFG).  Round-1/2 sweeps of tiles that
// were dropped from the build since: six / seven rows per thread (7 warps), five rows, four rows (13 / 11 warps), 64-cell
// tiles at two CTAs per SM, deeper stage rings -- all slower (profiles/r01c_*, r01l_*, r02i_sweep.txt).
const FusedConfig kFused2[] = {
    make_fused<FusedCfg<2, 21, 3, 5>>("t2_cj21_r3_s5"),
    make_fused<FusedCfg<2, 16, 2, 4>>("t2_cj16_r2_s4"),
};
const FusedConfig kFused3[] = {
    make_fused<FusedCfg<3, 21, 3, 4>>("t3_cj21_r3_s4"),  // 15 consumer warps x 3 rows
    make_fused<FusedCfg<3, 18, 2, 4>>("t3_cj18_r2_s4"),  // 19 consumer warps x 2 rows
    make_fused<FusedCfg<3, 24, 3, 4>>("t3_cj24_r3_s4"),  // 17 consumer warps x 3 rows, 10.8 % redundant halo
    make_fused<FusedCfg<3, 18, 3, 4>>("t3_cj18_r3_s4"),  // 13 consumer warps x 3 rows
};
const FusedConfig kFused4[] = {
    make_fused<FusedCfg<4, 21, 3, 4>>("t4_cj21_r3_s4"),
    make_fused<FusedCfg<4, 18, 2, 4>>("t4_cj18_r2_s4"),
};

const FusedConfig* fz_table(int T, int* count) {
  switch (T) {
    case 2: *count = sizeof(kFused2) / sizeof(kFused2[0]); return kFused2;
    case 3: *count = sizeof(kFused3) / sizeof(kFused3[0]); return kFused3;
    case 4: *count = sizeof(kFused4) / sizeof(kFused4[0]); return kFused4;
  }
  *count = 0;
  return nullptr;
}

int fz_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

const FusedConfig* fz_pick(int T) {
  int n = 0;
  const FusedConfig* tab = fz_table(T, &n);
  if (!tab) return nullptr;
  int c = fz_env_int("FDB_FUSED_CFG", 0);
  if (c < 0 || c >= n) c = 0;
  return &tab[c];
}

// measured on B200 at 512^3 (profiles/r02f_*, r02h_*, r02j_*), first / lean / lean + split exchange tile:
// T=3 835 / 915 / 973 GCUPS, T=4 768 / 850 / 939, T=2 747 / 720 / 715
constexpr int fz_default_impl(int T) { return T >= 3 ? 4 : 1; }

struct FusedAttr {
  const FusedConfig* cfg = nullptr;
  int ctas_per_sm = 1;
  int sms = 148;
};
FusedAttr g_fz_attr[16][kMaxFuse + 1];

int fz_maps(const Field& f, int d, const FusedConfig& C, int p, FusedMaps* out) {
  const Slab& s = f.slabs[d];
  const int64_t n1 = f.geo.n[1], n2 = f.geo.n[2];
  const double* base[2] = {f.body(d, p), f.ghost_lo(d, p)};
  const int64_t planes[2] = {s.nloc(), f.G};
  const int boxes[4][2] = {{C.BKP, C.HR}, {C.BKP, C.BJ}, {8, C.HR}, {8, C.BJ}};
  for (int t = 0; t < 2; ++t)
    for (int b = 0; b < 4; ++b)
      FDB_TRY(encode_tensor_map_3d(&out->m[4 * t + b], base[t], n2, n1, planes[t], boxes[b][0], boxes[b][1]));
  return FDB_OK;
}

}  // namespace

int upwind_fused_max() { return kMaxFuse; }

bool upwind_fused_supported(const Field& f, const UpwindCoeffs& k, int T) {
  if (T < 2 || T > kMaxFuse) return false;
  if (!upwind_tma_supported(f, k)) return false;
  if (f.G < T) return false;
  if (f.geo.n[1] < 8 || f.geo.n[2] < 16) return false;
  for (const Slab& s : f.slabs)
    if (s.nloc() < T) return false;
  return fz_pick(T) != nullptr;
}

const char* upwind_fused_name(int T) {
  const FusedConfig* C = fz_pick(T);
  return C ? C->name : "";
}

int upwind_fused_prepare(Field& f, int d, int T) {
  Slab& sl = f.slabs[d];
  FusedAttr& at = g_fz_attr[sl.device & 15][T];
  const FusedConfig* C = fz_pick(T);
  if (!C) return set_error(FDB_E_INVALID, "no fused kernel for %d steps per sweep", T);
  if (at.cfg != C) {
    for (FusedKernel kf : {C->kernel, C->lean, C->lean_push, C->split, C->split_push})
      FDB_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, C->smem));
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
    for (int p = 0; p < 2; ++p) FDB_TRY(fz_maps(f, d, *C, p, reinterpret_cast<FusedMaps*>(sl.fused_maps[p])));
    sl.fused_T = T;
    sl.fused_cfg = (const void*)C;
  }
  return FDB_OK;
}

// the __global__ function a sweep of T steps launches (what ncu lists)
const char* upwind_fused_kernel_name(int T) {
  return fz_env_int("FDB_FUSED_IMPL", fz_default_impl(T)) == 1 ? "upwind3d_fused_kernel" : "upwind3d_fused_lean_kernel";
}

bool upwind_fused_can_signal(int T) { return fz_env_int("FDB_FUSED_IMPL", fz_default_impl(T)) != 1; }

int launch_upwind_fused(Field& f, int d, int X, int T, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                         cudaStream_t s, double* peer_out, int64_t peer_from, const HaloSignal* sig) {
  if (iend <= ibeg) return FDB_OK;
  FDB_TRY(upwind_fused_prepare(f, d, T));
  Slab& sl = f.slabs[d];
  FusedAttr& at = g_fz_attr[sl.device & 15][T];
  const FusedConfig* C = fz_pick(T);
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
  a.plane = a.n1 * a.n2;
  const int reserve = f.single() ? 0 : fz_env_int("FDB_COMM_SMS", 0);  // SMs left free for NCCL halo kernels
  int64_t grid_max = (int64_t)at.ctas_per_sm * (at.sms - reserve);
  if (grid_max < 1) grid_max = 1;
  const int64_t tiles = (int64_t)a.njt * a.nkt;
  const int64_t planes = iend - ibeg;
  int64_t ci = fz_env_int("FDB_TMA_CI", 0);
  if (ci <= 0) {
    // static round-robin: ceil(items / CTAs) rounds of (chunk + T warm-up planes) plane-steps
    static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 24, 16, 8};
    int64_t best_cost = -1;
    for (int c : cand) {
      const int64_t cc = c < planes ? c : planes;
      const int64_t items = tiles * ((planes + cc - 1) / cc);
      const int64_t cost = ((items + grid_max - 1) / grid_max) * (cc + T);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; ci = cc; }
    }
  }
  if (ci > planes) ci = planes;
  a.ci = (int)ci;
  a.nchunks = (int)((planes + ci - 1) / ci);
  a.nwork = tiles * a.nchunks;
  a.reverse = 0;
  a.done_ctr = nullptr;
  a.done_target = 0;
  a.nbr_flag = nullptr;
  a.flag_value = 0;
  a.ack_flag = nullptr;
  a.ack_value = 0;
  a.ghost_flag = nullptr;
  a.ghost_value = 0;
  a.peer_delta = peer_out ? (int64_t)((char*)peer_out - (char*)a.out) - peer_from * a.plane * (int64_t)sizeof(double) : 0;
  if (sig) {
    if (!peer_out || !upwind_fused_can_signal(T))
      return set_error(FDB_E_STATE, "this fused kernel formulation cannot signal its halo push");
    int64_t pushing = 0;  // chunks that hold planes >= peer_from
    for (int64_t c = 0; c < a.nchunks; ++c) {
      const int64_t c1 = std::min<int64_t>(ibeg + (c + 1) * ci, iend);
      if (c1 > peer_from) ++pushing;
    }
    a.reverse = 1;
    a.done_ctr = sig->done_ctr;
    a.done_target = (unsigned int)(tiles * pushing);
    a.nbr_flag = sig->nbr_flag;
    a.flag_value = sig->flag_value;
    a.ack_flag = sig->ack_flag;
    a.ack_value = sig->ack_value;
    a.ghost_flag = sig->ghost_flag;
    a.ghost_value = sig->ghost_value;
  }
  int64_t grid = a.nwork < grid_max ? a.nwork : grid_max;
  // FDB_MAX_CTAS (tests): fewer CTAs than the device holds, so every CTA walks many work items
  if (const int cap = fz_env_int("FDB_MAX_CTAS", 0); cap > 0 && grid > cap) grid = cap;
  // FDB_FUSED_IMPL: 1 = first consumer formulation, 2 = lean, 4 = lean + split exchange tile
  const int impl = fz_env_int("FDB_FUSED_IMPL", fz_default_impl(T));
  const FusedKernel kf = (impl == 1)   ? C->kernel
                         : (impl == 4) ? (peer_out ? C->split_push : C->split)
                                       : (peer_out ? C->lean_push : C->lean);
  kf<<<(unsigned)grid, C->threads, C->smem, s>>>(*reinterpret_cast<const FusedMaps*>(sl.fused_maps[X]), a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

}  // namespace fdb
