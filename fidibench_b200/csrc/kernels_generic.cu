// kernels_generic.cu -- any-shape global-memory kernels: the N-D upwind step, the
// generic offset->weight stencil (Filter), fills, permutes and reductions.
//
// Arithmetic contract: IEEE binary64, round to nearest,
// every multiply and add/subtract rounded separately (no FMA) and in the
// reference's order.  The __d*_rn intrinsics are never contracted by nvcc.
#include "fdb_internal.h"

namespace fdb {

namespace {

constexpr int kThreads = 256;

// Source pointer of the plane `ii` (local index, may be -G..-1 or nloc..nloc+G-1).
__device__ __forceinline__ const double* plane_ptr(const double* __restrict__ body,
                                                   const double* __restrict__ glo,
                                                   const double* __restrict__ ghi, int64_t ii,
                                                   int64_t nloc, int G, int64_t plane) {
  if (ii < 0) return glo + (ii + G) * plane;
  if (ii >= nloc) return ghi + (ii - nloc) * plane;
  return body + ii * plane;
}

struct UpwindArgs {
  const double* body;
  const double* glo;
  const double* ghi;
  double* out;
  int64_t nloc, n1, n2, ibeg, iend;
  int G;
  double c0, c1, c2;
  int up0, up1, up2;
  int act0, act1, act2;
};

// ref: Upwind<NDIMS>::advect inner loop, upwind/cxx/upwind.cxx:64-84.
// One thread per cell; axes are visited in the reference's order 0,1,2 and
// inactive (non-existent) axes are skipped as the loop over j < NDIMS does.
__global__ void __launch_bounds__(kThreads) upwind_generic_kernel(UpwindArgs a) {
  const int64_t plane = a.n1 * a.n2;
  const int64_t total = (a.iend - a.ibeg) * plane;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * kThreads) {
    const int64_t i = a.ibeg + t / plane;
    const int64_t r = t % plane;
    const int64_t j = r / a.n2;
    const int64_t k = r % a.n2;
    const double ctr = a.body[i * plane + r];
    double v = ctr;
    if (a.act0) {
      const double* p = plane_ptr(a.body, a.glo, a.ghi, i + a.up0, a.nloc, a.G, plane);
      v = __dsub_rn(v, __dmul_rn(a.c0, __dsub_rn(p[r], ctr)));
    }
    if (a.act1) {
      int64_t jj = j + a.up1;
      jj = jj < 0 ? jj + a.n1 : (jj >= a.n1 ? jj - a.n1 : jj);
      v = __dsub_rn(v, __dmul_rn(a.c1, __dsub_rn(a.body[i * plane + jj * a.n2 + k], ctr)));
    }
    if (a.act2) {
      int64_t kk = k + a.up2;
      kk = kk < 0 ? kk + a.n2 : (kk >= a.n2 ? kk - a.n2 : kk);
      v = __dsub_rn(v, __dmul_rn(a.c2, __dsub_rn(a.body[i * plane + j * a.n2 + kk], ctr)));
    }
    a.out[i * plane + r] = v;
  }
}

struct StencilArgs {
  const double* body;
  const double* glo;
  const double* ghi;
  double* out;
  int64_t nloc, n1, n2, ibeg, iend;
  int G;
  int nbranch;
  int ref_wrap;  // reproduce Filter.cpp:240's (int %= size_t) wrap instead of the periodic one (single slab only)
  int off[32][3];
  double w[32];
};

__device__ __forceinline__ int64_t wrap(int64_t v, int64_t n) {
  v %= n;
  return v < 0 ? v + n : v;
}

// ref: Filter.cpp:237-240 -- `indOffset[j] = inds[j] + offset[j]` is an int, `indOffset[j] %= globalDims[j]`
// converts it to size_t first: -1 becomes (2^64 - 1) mod n, which is n - 1 only when n divides 2^64
// (SURVEY.md H2).  Opt-in compatibility mode, see fdb_stencil_set_ref_wrap.
__device__ __forceinline__ int64_t wrap_ref(int64_t ind, int off, int64_t n) {
  const int v = (int)ind + off;
  const unsigned long long u = (unsigned long long)(long long)v;
  return (int64_t)(int)(u % (unsigned long long)n);
}

// ref: Filter::applyFilter, cxx/Filter.cpp:191-263.  acc starts at 0.0 and takes
// one separately rounded multiply and add per branch, in std::map order.
__global__ void __launch_bounds__(kThreads) stencil_generic_kernel(StencilArgs a) {
  const int64_t plane = a.n1 * a.n2;
  const int64_t total = (a.iend - a.ibeg) * plane;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * kThreads) {
    const int64_t i = a.ibeg + t / plane;
    const int64_t r = t % plane;
    const int64_t j = r / a.n2;
    const int64_t k = r % a.n2;
    double acc = 0.0;
    for (int b = 0; b < a.nbranch; ++b) {
      // axis 0 is not wrapped here: planes outside [0,nloc) live in the ghosts
      const double* p = a.ref_wrap ? a.body + wrap_ref(i, a.off[b][0], a.nloc) * plane
                                   : plane_ptr(a.body, a.glo, a.ghi, i + a.off[b][0], a.nloc, a.G, plane);
      const int64_t jj = a.ref_wrap ? wrap_ref(j, a.off[b][1], a.n1) : wrap(j + a.off[b][1], a.n1);
      const int64_t kk = a.ref_wrap ? wrap_ref(k, a.off[b][2], a.n2) : wrap(k + a.off[b][2], a.n2);
      acc = __dadd_rn(acc, __dmul_rn(a.w[b], p[jj * a.n2 + kk]));
    }
    a.out[i * plane + r] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) fill_kernel(double* p, int64_t n, double v) {
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * kThreads)
    p[t] = v;
}

// Device-side synthetic input (bench/parity fields without a host copy): cell g of the GLOBAL row-major field
// gets u(seed, g) = (splitmix64(seed + (g+1)*0x9E3779B97F4A7C15) >> 11) * 2^-53, a pure function of the global
// index, so the field does not depend on how planes are spread over devices (tests restate it with numpy).
__device__ __forceinline__ double hash_u01(uint64_t seed, uint64_t g) {
  uint64_t z = seed + (g + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * 0x1.0p-53;  // exact: 53 bits
}
__global__ void __launch_bounds__(kThreads) fill_random_kernel(double* p, int64_t n, uint64_t seed, int64_t g0) {
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < n;
       t += (int64_t)gridDim.x * kThreads)
    p[t] = hash_u01(seed, (uint64_t)(g0 + t));
}

// Separable input: cell (i, j, k) = ((1 * xa[.]) * xb[.]) * xc[.] in the REFERENCE's axis order -- the product
// Filter::setInData accumulates for laplacian.cxx's func (ref: laplacian.cxx:22-28, Filter.cpp:103-112), with the
// 1-D factors sin(2 pi x) evaluated on the host (same libm, same bits) and only 8 B x (d0+d1+d2) uploaded.
// ax[] = internal axis of reference axis 0..nd-1.
struct SeparableArgs {
  double* out;
  int64_t i_lo, nplanes, n1, n2;
  const double* x[3];
  int ax[3];
  int nd;
};
__global__ void __launch_bounds__(kThreads) separable_kernel(SeparableArgs a) {
  const int64_t plane = a.n1 * a.n2;
  const int64_t total = a.nplanes * plane;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * kThreads) {
    int64_t idx[3];
    idx[0] = a.i_lo + t / plane;
    idx[1] = (t % plane) / a.n2;
    idx[2] = t % a.n2;
    double res = 1.0;
    for (int j = 0; j < a.nd; ++j) res = __dmul_rn(res, a.x[j][idx[a.ax[j]]]);
    a.out[t] = res;
  }
}

// out[k][j][i] (row-major n2 x n1 x n0) = in[i][j][k] (row-major n0 x n1 x n2):
// converts between row-major and Filter's column-major storage (Filter.cpp:50).
__global__ void __launch_bounds__(kThreads) permute210_kernel(const double* __restrict__ in,
                                                              double* __restrict__ out, int64_t n0,
                                                              int64_t n1, int64_t n2) {
  __shared__ double tile[32][33];
  // grid: x over k tiles, y over i tiles, z over j
  const int64_t j = blockIdx.z;
  const int64_t k0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = i0 + r, k = k0 + tx;
    if (i < n0 && k < n2) tile[r][tx] = in[(i * n1 + j) * n2 + k];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t k = k0 + r, i = i0 + tx;
    if (i < n0 && k < n2) out[(k * n1 + j) * n0 + i] = tile[tx][r];
  }
}

// out[i][j][k] = in[i'][j'][k'] with each primed index mirrored (n-1-x) where its flag is set: turns a
// negative velocity along an axis into a positive one on the mirrored grid (capi.cu: upwind flips)
__global__ void __launch_bounds__(kThreads) mirror_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                          int64_t n0, int64_t n1, int64_t n2, int f0, int f1, int f2) {
  const int64_t total = n0 * n1 * n2;
  for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * kThreads) {
    const int64_t k = t % n2;
    const int64_t j = (t / n2) % n1;
    const int64_t i = t / (n2 * n1);
    const int64_t ii = f0 ? n0 - 1 - i : i, jj = f1 ? n1 - 1 - j : j, kk = f2 ? n2 - 1 - k : k;
    out[t] = in[(ii * n1 + jj) * n2 + kk];
  }
}

// ---- reductions ---------------------------------------------------------------
// Stage 1: block (c, i) reduces chunk c of plane i in a fixed order -> partial.
// Stage 2: block i reduces its plane's partials in a fixed order -> plane_sums[i].
// The shape of both trees depends only on the plane size, never on how planes
// are spread over devices, so the final sum is bitwise partition-invariant.
constexpr int64_t kChunk = 8192;  // cells per stage-1 block

__device__ __forceinline__ double block_tree_sum(double v) {
  __shared__ double sh[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, o));
  }
  __syncthreads();
  return s;  // valid in thread 0
}

// mode 0: sum f ; mode 1: sum (f - mean)^2   (ref: upwind.cxx:91-103)
__global__ void __launch_bounds__(kThreads) plane_partial_kernel(const double* __restrict__ body,
                                                                 int64_t plane, int64_t nchunk,
                                                                 int mode, double mean,
                                                                 double* __restrict__ partial) {
  const int64_t c = blockIdx.x, i = blockIdx.y;
  const double* p = body + i * plane;
  const int64_t beg = c * kChunk;
  const int64_t end = beg + kChunk < plane ? beg + kChunk : plane;
  double acc = 0.0;
  for (int64_t t = beg + threadIdx.x; t < end; t += kThreads) {
    double v = p[t];
    if (mode == 1) {
      const double d = __dsub_rn(v, mean);
      v = __dmul_rn(d, d);
    }
    acc = __dadd_rn(acc, v);
  }
  const double s = block_tree_sum(acc);
  if (threadIdx.x == 0) partial[i * nchunk + c] = s;
}

__global__ void __launch_bounds__(kThreads) plane_final_kernel(const double* __restrict__ partial,
                                                               int64_t nchunk,
                                                               double* __restrict__ plane_sums) {
  const int64_t i = blockIdx.x;
  double acc = 0.0;
  for (int64_t t = threadIdx.x; t < nchunk; t += kThreads)
    acc = __dadd_rn(acc, partial[i * nchunk + t]);
  const double s = block_tree_sum(acc);
  if (threadIdx.x == 0) plane_sums[i] = s;
}

int grid_for(int64_t work_items) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = 148 * 16;  // grid-stride beyond a few waves
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

int launch_upwind_generic(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                          cudaStream_t s) {
  if (iend <= ibeg) return FDB_OK;
  const Slab& sl = f.slabs[d];
  UpwindArgs a;
  a.body = f.body(d, X);
  a.glo = f.ghost_lo(d, X);
  a.ghi = f.ghost_hi(d, X);
  a.out = f.body(d, 1 - X);
  a.nloc = sl.nloc();
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.ibeg = ibeg;
  a.iend = iend;
  a.G = f.G;
  a.c0 = k.c[0]; a.c1 = k.c[1]; a.c2 = k.c[2];
  a.up0 = k.up[0]; a.up1 = k.up[1]; a.up2 = k.up[2];
  a.act0 = k.active[0]; a.act1 = k.active[1]; a.act2 = k.active[2];
  upwind_generic_kernel<<<grid_for((iend - ibeg) * f.geo.plane()), kThreads, 0, s>>>(a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_stencil_generic(const Field& f, int d, int X, int64_t ibeg, int64_t iend,
                           const StencilBranches& b, cudaStream_t s) {
  if (iend <= ibeg) return FDB_OK;
  const Slab& sl = f.slabs[d];
  StencilArgs a;
  a.body = f.body(d, X);
  a.glo = f.ghost_lo(d, X);
  a.ghi = f.ghost_hi(d, X);
  a.out = f.body(d, 1 - X);
  a.nloc = sl.nloc();
  a.n1 = f.geo.n[1];
  a.n2 = f.geo.n[2];
  a.ibeg = ibeg;
  a.iend = iend;
  a.G = f.G;
  a.nbranch = b.nbranch;
  a.ref_wrap = (b.ref_wrap && f.single()) ? 1 : 0;
  for (int i = 0; i < b.nbranch; ++i) {
    a.off[i][0] = b.off[i][0]; a.off[i][1] = b.off[i][1]; a.off[i][2] = b.off[i][2];
    a.w[i] = b.w[i];
  }
  stencil_generic_kernel<<<grid_for((iend - ibeg) * f.geo.plane()), kThreads, 0, s>>>(a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

// forces the (lazily loaded) generic kernels onto the current device; see SweepLauncher::prepare
int generic_kernels_prepare() {
  cudaFuncAttributes fa;
  FDB_CUDA(cudaFuncGetAttributes(&fa, upwind_generic_kernel));
  FDB_CUDA(cudaFuncGetAttributes(&fa, stencil_generic_kernel));
  FDB_CUDA(cudaFuncGetAttributes(&fa, fill_kernel));
  FDB_CUDA(cudaFuncGetAttributes(&fa, mirror_kernel));
  FDB_CUDA(cudaFuncGetAttributes(&fa, plane_partial_kernel));
  FDB_CUDA(cudaFuncGetAttributes(&fa, plane_final_kernel));
  return FDB_OK;
}

int64_t reduce_partials_per_plane(int64_t plane) { return (plane + kChunk - 1) / kChunk; }

int launch_plane_sums(const double* body, int64_t nloc, int64_t plane, int mode, double mean,
                      double* partial, double* plane_sums, cudaStream_t s) {
  const int64_t nchunk = reduce_partials_per_plane(plane);
  // gridDim.y is limited to 65535 planes per launch
  for (int64_t i0 = 0; i0 < nloc; i0 += 32768) {
    const int64_t ni = nloc - i0 < 32768 ? nloc - i0 : 32768;
    dim3 grid((unsigned)nchunk, (unsigned)ni);
    plane_partial_kernel<<<grid, kThreads, 0, s>>>(body + i0 * plane, plane, nchunk, mode, mean,
                                                   partial + i0 * nchunk);
    count_launch();
    FDB_CUDA(cudaGetLastError());
  }
  plane_final_kernel<<<(unsigned)nloc, kThreads, 0, s>>>(partial, nchunk, plane_sums);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_fill(double* p, int64_t n, double v, cudaStream_t s) {
  if (n <= 0) return FDB_OK;
  fill_kernel<<<grid_for(n), kThreads, 0, s>>>(p, n, v);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_fill_random(double* p, int64_t n, uint64_t seed, int64_t g0, cudaStream_t s) {
  if (n <= 0) return FDB_OK;
  fill_random_kernel<<<grid_for(n), kThreads, 0, s>>>(p, n, seed, g0);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_separable(double* out, int64_t i_lo, int64_t nplanes, int64_t n1, int64_t n2, int nd,
                     const double* const* x, const int* ax, cudaStream_t s) {
  if (nplanes <= 0) return FDB_OK;
  SeparableArgs a;
  a.out = out; a.i_lo = i_lo; a.nplanes = nplanes; a.n1 = n1; a.n2 = n2; a.nd = nd;
  for (int j = 0; j < 3; ++j) { a.x[j] = j < nd ? x[j] : nullptr; a.ax[j] = j < nd ? ax[j] : 0; }
  separable_kernel<<<grid_for(nplanes * n1 * n2), kThreads, 0, s>>>(a);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_mirror(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2, const bool* flip,
                  cudaStream_t s) {
  const int64_t total = n0 * n1 * n2;
  if (total <= 0) return FDB_OK;
  mirror_kernel<<<grid_for(total), kThreads, 0, s>>>(in, out, n0, n1, n2, flip[0], flip[1], flip[2]);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

int launch_permute(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2,
                   cudaStream_t s) {
  if (n1 > 65535) return set_error(FDB_E_INVALID, "permute: middle extent %lld too large", (long long)n1);
  dim3 grid((unsigned)((n2 + 31) / 32), (unsigned)((n0 + 31) / 32), (unsigned)n1);
  if (grid.y > 65535) return set_error(FDB_E_INVALID, "permute: first extent too large");
  permute210_kernel<<<grid, kThreads, 0, s>>>(in, out, n0, n1, n2);
  count_launch();
  FDB_CUDA(cudaGetLastError());
  return FDB_OK;
}

}  // namespace fdb
