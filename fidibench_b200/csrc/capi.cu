// capi.cu -- the extern "C" surface declared in include/fidib200.h.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "fdb_internal.h"

using namespace fdb;

// ---- handles ----------------------------------------------------------------------
struct fdb_upwind {
  Field field;
  double velocity[3] = {0, 0, 0};  // reference axis order
  double lengths[3] = {1, 1, 1};
  int64_t num_cells[3] = {1, 1, 1};
  int kernel = FDB_KERNEL_AUTO;
  int fuse = 0;  // time steps per sweep; 0 = auto (kAutoFuse when the fused kernel can run)
  // Axes (internal) along which the device holds the field MIRRORED: a negative velocity along an axis is
  // a positive one on the mirrored grid, with the same coefficient bits ((dt*-v)*-1 == (dt*v)*+1), so the
  // tiled kernels (written for the upwind direction -1) run every sign.  Single-slab handles only.
  bool flip[3] = {false, false, false};
  bool any_flip = false;
};

struct fdb_stencil {
  Field field;
  StencilBranches br;
  int ndims = 3;
  int64_t dims[3] = {1, 1, 1};
  bool out_valid = false;  // buf[1-cur] holds the last apply's output
  int kernel = FDB_KERNEL_AUTO;
  int reach = 1;           // ghost planes one apply reads: max |offset| along the slab axis
  int fuse = 0;            // applies per sweep in iterate(): 0 = auto, 1, or 2 (fused 7-point kernel)
};

namespace {

struct UpwindSweep : SweepLauncher {
  UpwindCoeffs k;
  bool tma = false;
  // depth = time steps this sweep advances (> 1: the fused temporal-blocking kernel)
  int launch(Field* f, int d, int X, int depth, int64_t ibeg, int64_t iend, cudaStream_t s) override {
    if (depth > 1) return launch_upwind_fused(*f, d, X, depth, ibeg, iend, k, s);
    return tma ? launch_upwind_tma(*f, d, X, ibeg, iend, k, s)
               : launch_upwind_generic(*f, d, X, ibeg, iend, k, s);
  }
  int prepare(Field* f, int d, int depth) override {
    FDB_TRY(generic_kernels_prepare());
    if (depth > 1) return upwind_fused_prepare(*f, d, depth);
    return tma ? upwind_tma_prepare(*f, d) : FDB_OK;
  }
  // the TMA kernels can mirror their top planes into the next slab's ghost planes themselves
  bool can_push(const Field*, int) const override { return tma; }
  bool can_push_single(const Field*, int depth) const override { return tma && depth > 1 && upwind_fused_can_signal(depth); }
  int launch_single(Field* f, int d, int X, int depth, cudaStream_t s, double* peer_out, int64_t peer_from,
                    const HaloSignal& sig) override {
    if (!can_push_single(f, depth)) return FDB_E_STATE;
    return launch_upwind_fused(*f, d, X, depth, 0, f->slabs[d].nloc(), k, s, peer_out, peer_from, &sig);
  }
  uint64_t key() const override {
    uint64_t h = tma ? 0x51ull : 0x77ull;
    for (int a = 0; a < 3; ++a) {
      uint64_t bits;
      memcpy(&bits, &k.c[a], sizeof(bits));
      h = (h ^ bits) * 0x100000001B3ull;
      h = (h ^ (uint64_t)(k.up[a] + 2) ^ ((uint64_t)k.active[a] << 8)) * 0x100000001B3ull;
    }
    return h | 1ull;
  }
  int launch_push(Field* f, int d, int X, int depth, int64_t ibeg, int64_t iend, cudaStream_t s, double* peer_out,
                  int64_t peer_from) override {
    if (!tma) return FDB_E_STATE;
    if (depth > 1) return launch_upwind_fused(*f, d, X, depth, ibeg, iend, k, s, peer_out, peer_from);
    return launch_upwind_tma(*f, d, X, ibeg, iend, k, s, peer_out, peer_from);
  }
};

struct StencilSweep : SweepLauncher {
  const StencilBranches* b = nullptr;
  bool fast = false;
  int fused_depth = 0;  // sweeps of this ghost depth run the two-applies-per-sweep kernel (0 = never)
  int prepare(Field* f, int d, int depth) override {
    FDB_TRY(generic_kernels_prepare());
    if (fused_depth > 0 && depth == fused_depth) return stencil_lap7_fused_prepare(*f, d, *b);
    return fast ? stencil_lap7_prepare(*f, d) : FDB_OK;
  }
  uint64_t key() const override {  // the branches live in the handle and never change after creation
    return (0xF17ull * 0x100000001B3ull) ^ ((uint64_t)fast << 1) ^ ((uint64_t)fused_depth << 4) ^ ((uint64_t)b->ref_wrap << 9) | 1ull;
  }
  int launch(Field* f, int d, int X, int depth, int64_t ibeg, int64_t iend, cudaStream_t s) override {
    if (fused_depth > 0 && depth == fused_depth) return launch_stencil_lap7_fused(*f, d, X, ibeg, iend, *b, s);
    return fast ? launch_stencil_lap7(*f, d, X, ibeg, iend, *b, s)
                : launch_stencil_generic(*f, d, X, ibeg, iend, *b, s);
  }
};

// Sustained runs (bench.py, power-capped; profiles/r02zz_*): T = 4 sweeps are 2.6-3.4 % ahead of T = 3 at 1 and 2 GPUs
// (25 % less DRAM traffic per step, and 100 steps are 25 x 4 with no remainder); short bursts at 512^3 favour T = 3.
constexpr int kAutoFuse = 4;

// ref: upwind.cxx:34-36,72 -- upDirection, deltas and the per-axis coefficient
// ((deltaTime * v[j]) * upDirection[j]) / deltas[j], evaluated left to right.
void upwind_coeffs(const fdb_upwind* h, double dt, UpwindCoeffs* k) {
  const Geometry& g = h->field.geo;
  for (int a = 0; a < 3; ++a) { k->c[a] = 0.0; k->up[a] = -1; k->active[a] = false; }
  for (int j = 0; j < g.ndims; ++j) {
    const int a = g.axis_of[j];
    const double v = h->flip[a] ? -h->velocity[j] : h->velocity[j];  // the velocity on the (mirrored) device grid
    const int up = (v < 0.) ? +1 : -1;
    const double delta = h->lengths[j] / (double)(size_t)h->num_cells[j];
    k->c[a] = dt * v * up / delta;
    k->up[a] = up;
    k->active[a] = true;
  }
}

int timing_begin(Field* f) {
  f->last_halo_bytes = 0;
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaEventRecord(s.ev_t0, s.s_main));
  }
  return FDB_OK;
}
int timing_end(Field* f) {
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaEventRecord(s.ev_t1, s.s_main));
  }
  return FDB_OK;
}
int timing_collect(Field* f) {
  double ms = 0;
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaEventSynchronize(s.ev_t1));
    float t = 0;
    if (cudaEventElapsedTime(&t, s.ev_t0, s.ev_t1) == cudaSuccess) ms = std::max(ms, (double)t);
    (void)cudaGetLastError();
  }
  f->last_ms = ms;
  return FDB_OK;
}

// where the ctor's cell 0 (upwind.cxx:48) sits on the device grid
int64_t upwind_mirrored_cell0(const fdb_upwind* h, const Geometry& g) {
  int64_t cell = 0;
  const int64_t stride[3] = {g.n[1] * g.n[2], g.n[2], 1};
  // axis 0 is mirrored inside each slab: global plane 0 is the LAST device plane of the slab that owns it
  const int64_t ext[3] = {g.n[0] / h->field.nparts, g.n[1], g.n[2]};
  for (int a = 0; a < 3; ++a)
    if (h->flip[a]) cell += (ext[a] - 1) * stride[a];
  return cell;
}

// host data lands in the spare buffer and is mirrored into the current one (and back for downloads)
int upwind_upload(fdb_upwind* h, const double* host_global, const double* host_slab, bool async) {
  Field* f = &h->field;
  if (!h->any_flip) return field_upload(f, f->cur, host_global, host_slab, async);
  const int cur = f->cur;
  FDB_TRY(field_upload(f, 1 - cur, host_global, host_slab, async));  // (publishes the spare buffer for a moment)
  return field_mirror(f, 1 - cur, cur, h->flip, /*publish=*/true);
}

int upwind_download(fdb_upwind* h, double* host_global, double* host_slab) {
  Field* f = &h->field;
  if (!h->any_flip) return field_download(f, f->cur, host_global, host_slab);
  FDB_TRY(field_mirror(f, f->cur, 1 - f->cur, h->flip, /*publish=*/false));  // the spare buffer holds no live data
  return field_download(f, 1 - f->cur, host_global, host_slab);
}

int upwind_common_create(int ndims, const int64_t* numCells, const double* velocity,
                         const double* lengths, int ngpus, fdb_comm* comm, fdb_upwind** out) {
  if (!out) return set_error(FDB_E_INVALID, "null output handle");
  *out = nullptr;
  if (!numCells || !velocity || !lengths) return set_error(FDB_E_INVALID, "null argument");
  Geometry geo;
  FDB_TRY(make_geometry(ndims, numCells, &geo));
  for (int j = 0; j < ndims; ++j)
    if (!(lengths[j] > 0.0)) return set_error(FDB_E_INVALID, "lengths[%d] must be positive", j);
  fdb_upwind* h = new (std::nothrow) fdb_upwind();
  if (!h) return set_error(FDB_E_OOM, "out of host memory");
  for (int j = 0; j < ndims; ++j) {
    h->velocity[j] = velocity[j];
    h->lengths[j] = lengths[j];
    h->num_cells[j] = numCells[j];
  }
  // ghost depth: as many planes as the fused kernel may advance per sweep, slab permitting
  const int nparts = comm ? comm->nranks : (ngpus > 0 ? ngpus : 1);
  const int64_t nloc = geo.n[0] / nparts;
  const int G = (int)std::max<int64_t>(1, std::min<int64_t>(kMaxFuse, nloc));
  // Mirror the axes with a negative velocity when that lets the tiled kernels run (3-D, the kernels' shape
  // requirements); FDB_NO_FLIP=1 keeps the generic kernel for them.  On several slabs every slab is mirrored in
  // place: each device keeps its own global planes, reversed, and along axis 0 the ring then runs backwards.
  {
    const char* nf = getenv("FDB_NO_FLIP");
    const bool allow = !(nf && *nf && atoi(nf) != 0) && ndims == 3 && geo.n[2] % 2 == 0 && geo.n[2] >= 4 && geo.n[1] >= 2;
    for (int j = 0; j < ndims && allow; ++j)
      if (velocity[j] < 0. && geo.n[geo.axis_of[j]] > 1) {
        h->flip[geo.axis_of[j]] = true;
        h->any_flip = true;
      }
  }
  bool need_lo = false, need_hi = false;
  for (int j = 0; j < ndims; ++j)
    if (geo.axis_of[j] == 0 && geo.n[0] > 1) {
      if (velocity[j] < 0. && !h->flip[0]) need_hi = true; else need_lo = true;
    }
  const bool reversed = h->flip[0] && nparts > 1;
  int rc = field_create(&h->field, geo, G, need_lo, need_hi, ngpus, comm, /*want_tma=*/1, reversed);
  h->field.planes_mirrored = h->flip[0];  // reductions report planes in global order
  if (rc == FDB_OK) rc = field_fill_delta(&h->field, 0, upwind_mirrored_cell0(h, geo));
  if (rc == FDB_OK) rc = field_sync(&h->field);
  if (rc != FDB_OK) {
    field_destroy(&h->field);
    delete h;
    return rc;
  }
  *out = h;
  return FDB_OK;
}

// std::map<std::vector<int>, double> order (Filter.cpp:202)
bool lex_less(const int* a, const int* b, int nd) {
  for (int j = 0; j < nd; ++j) {
    if (a[j] < b[j]) return true;
    if (a[j] > b[j]) return false;
  }
  return false;
}

int stencil_common_create(int ndims, const int64_t* dims, int nbranch, const int* offsets,
                          const double* weights, int ngpus, fdb_comm* comm, fdb_stencil** out) {
  if (!out) return set_error(FDB_E_INVALID, "null output handle");
  *out = nullptr;
  if (!dims || !offsets || !weights) return set_error(FDB_E_INVALID, "null argument");
  if (nbranch < 1 || nbranch > 32)
    return set_error(FDB_E_INVALID, "nbranch must be in 1..32 (got %d)", nbranch);
  Geometry geo;
  // a 2-D problem on one device is carried as a single plane so that the tiled kernel runs
  // (the reference's default laplacian case is 2-D, ref: laplacian.cxx:41-42)
  FDB_TRY(make_geometry(ndims, dims, &geo, /*plane2d=*/ndims == 2 && !comm && ngpus == 1));
  fdb_stencil* h = new (std::nothrow) fdb_stencil();
  if (!h) return set_error(FDB_E_OOM, "out of host memory");
  h->ndims = ndims;
  for (int j = 0; j < ndims; ++j) h->dims[j] = dims[j];
  // sort the branches the way std::map iterates them
  std::vector<int> order(nbranch);
  for (int b = 0; b < nbranch; ++b) order[b] = b;
  std::sort(order.begin(), order.end(), [&](int x, int y) {
    return lex_less(offsets + x * ndims, offsets + y * ndims, ndims);
  });
  for (int b = 1; b < nbranch; ++b)
    if (!lex_less(offsets + order[b - 1] * ndims, offsets + order[b] * ndims, ndims)) {
      delete h;
      return set_error(FDB_E_INVALID, "duplicate stencil offset (a std::map key can appear once)");
    }
  int G = 0;
  bool need_lo = false, need_hi = false;
  h->br.nbranch = nbranch;
  for (int b = 0; b < nbranch; ++b) {
    const int* o = offsets + order[b] * ndims;
    int io[3] = {0, 0, 0};
    for (int j = 0; j < ndims; ++j) io[geo.axis_of[j]] = o[j];
    for (int a = 0; a < 3; ++a) h->br.off[b][a] = io[a];
    h->br.w[b] = weights[order[b]];
    if (geo.n[0] > 1) {
      if (io[0] < 0) need_lo = true;
      if (io[0] > 0) need_hi = true;
      G = std::max(G, std::abs(io[0]));
    } else {
      h->br.off[b][0] = 0;  // a single plane is its own periodic neighbour
    }
  }
  if (G == 0) G = 1;
  h->reach = G;
  {
    // the 3-D 7-point set can run two applies per sweep, which reads two ghost planes per side
    const int nparts = comm ? comm->nranks : (ngpus > 0 ? ngpus : 1);
    const int64_t nloc = geo.n[0] / nparts;
    bool full7 = (ndims == 3 && nbranch == 7 && G == 1 && nloc >= 2);
    for (int b = 0; b < nbranch && full7; ++b) {
      const int* o = h->br.off[b];
      full7 = (std::abs(o[0]) + std::abs(o[1]) + std::abs(o[2]) <= 1);  // distinct offsets: exactly the 7
    }
    if (full7) G = 2;
  }
  int rc = field_create(&h->field, geo, G, need_lo, need_hi, ngpus, comm, /*want_tma=*/2);
  if (rc == FDB_OK) rc = field_sync(&h->field);
  if (rc == FDB_OK) {  // FDB_REF_WRAP=1: handles start in the reference-wrap compatibility mode where it applies
    const char* rw = getenv("FDB_REF_WRAP");
    if (rw && *rw && atoi(rw) != 0 && h->field.single()) h->br.ref_wrap = true;
  }
  if (rc != FDB_OK) {
    field_destroy(&h->field);
    delete h;
    return rc;
  }
  *out = h;
  return FDB_OK;
}

}  // namespace

#define FDB_GUARD_BEGIN try {
#define FDB_GUARD_END                                                    \
  }                                                                      \
  catch (const std::bad_alloc&) {                                        \
    return set_error(FDB_E_OOM, "out of host memory");                   \
  }                                                                      \
  catch (...) {                                                          \
    return set_error(FDB_E_INVALID, "unexpected C++ exception");         \
  }

extern "C" {

const char* fdb_last_error(void) { return last_error(); }

int fdb_version(int* major, int* minor) {
  if (major) *major = FDB_VERSION_MAJOR;
  if (minor) *minor = FDB_VERSION_MINOR;
  return FDB_OK;
}

int fdb_device_count(int* count) {
  if (!count) return set_error(FDB_E_INVALID, "null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    (void)cudaGetLastError();
    return set_error(FDB_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return FDB_OK;
}

int fdb_launch_count(int64_t* count) {
  if (!count) return set_error(FDB_E_INVALID, "null argument");
  *count = launch_count();
  return FDB_OK;
}

int fdb_slab_partition(int64_t n0, int nparts, int part, int64_t* lo, int64_t* hi) {
  if (!lo || !hi || n0 < 1 || nparts < 1 || part < 0 || part >= nparts)
    return set_error(FDB_E_INVALID, "bad slab partition request (n0=%lld nparts=%d part=%d)",
                     (long long)n0, nparts, part);
  if (n0 % nparts != 0)
    return set_error(FDB_E_DECOMP,
                     "No valid domain decomposition: %d slab(s) do not divide the %lld planes of axis 0",
                     nparts, (long long)n0);
  const int64_t nloc = n0 / nparts;
  *lo = part * nloc;
  *hi = *lo + nloc;
  return FDB_OK;
}

// ---- communicator ---------------------------------------------------------------------
int fdb_comm_unique_id(void* id_bytes) {
  if (!id_bytes) return set_error(FDB_E_INVALID, "null argument");
  static_assert(sizeof(ncclUniqueId) <= FDB_COMM_ID_BYTES, "id size");
  ncclUniqueId id;
  FDB_NCCL(ncclGetUniqueId(&id));
  memset(id_bytes, 0, FDB_COMM_ID_BYTES);
  memcpy(id_bytes, &id, sizeof(id));
  return FDB_OK;
}

int fdb_comm_create(int rank, int nranks, const void* id_bytes, int device, fdb_comm** out) {
  FDB_GUARD_BEGIN
  if (!out) return set_error(FDB_E_INVALID, "null output handle");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks || (!id_bytes && nranks > 1))
    return set_error(FDB_E_INVALID, "bad communicator request (rank=%d nranks=%d)", rank, nranks);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    (void)cudaGetLastError();
    return set_error(FDB_E_CUDA, "device %d not available (%d visible); no CPU fallback", device, ndev);
  }
  fdb_comm* c = new fdb_comm();
  c->rank = rank;
  c->nranks = nranks;
  c->device = device;
  auto init = [&]() -> int {
    FDB_CUDA(cudaSetDevice(device));
    FDB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    FDB_CUDA(cudaMalloc(&c->scratch, 64 * sizeof(double)));
    c->scratch_doubles = 64;
    if (nranks > 1) {
      ncclUniqueId id;
      memcpy(&id, id_bytes, sizeof(id));
      FDB_NCCL(ncclCommInitRank(&c->nccl, nranks, id, rank));
    }
    return FDB_OK;
  };
  const int rc = init();
  if (rc != FDB_OK) {
    fdb_comm_destroy(c);  // releases whatever was created
    return rc;
  }
  *out = c;
  return FDB_OK;
  FDB_GUARD_END
}

int fdb_comm_rank(const fdb_comm* c, int* rank, int* nranks) {
  if (!c) return set_error(FDB_E_INVALID, "null communicator");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return FDB_OK;
}

int fdb_comm_max(fdb_comm* c, double* value) {
  if (!c || !value) return set_error(FDB_E_INVALID, "null argument");
  if (c->nranks == 1) return FDB_OK;
  FDB_CUDA(cudaSetDevice(c->device));
  FDB_CUDA(cudaMemcpyAsync(c->scratch, value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  FDB_NCCL(ncclAllReduce(c->scratch, c->scratch + 1, 1, ncclDouble, ncclMax, c->nccl, c->stream));
  count_launch();
  FDB_CUDA(cudaMemcpyAsync(value, c->scratch + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FDB_CUDA(cudaStreamSynchronize(c->stream));
  return FDB_OK;
}

int fdb_comm_barrier(fdb_comm* c) {
  double v = 0.0;
  return fdb_comm_max(c, &v);
}

int fdb_comm_destroy(fdb_comm* c) {
  if (!c) return FDB_OK;
  if (c->users > 0)
    return set_error(FDB_E_STATE, "%d engine handle(s) still use this communicator: destroy them first (their teardown "
                     "is a collective over it)", c->users);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->nccl) ncclCommDestroy(c->nccl);
  if (c->scratch) cudaFree(c->scratch);
  if (c->stream) cudaStreamDestroy(c->stream);
  (void)cudaGetLastError();
  delete c;
  return FDB_OK;
}

// ---- upwind -----------------------------------------------------------------------------
int fdb_upwind_create(int ndims, const int64_t* numCells, const double* velocity,
                      const double* lengths, int ngpus, fdb_upwind** out) {
  FDB_GUARD_BEGIN
  return upwind_common_create(ndims, numCells, velocity, lengths, ngpus, nullptr, out);
  FDB_GUARD_END
}

int fdb_upwind_create_dist(int ndims, const int64_t* numCells, const double* velocity,
                           const double* lengths, fdb_comm* comm, fdb_upwind** out) {
  FDB_GUARD_BEGIN
  if (!comm) return set_error(FDB_E_INVALID, "null communicator");
  return upwind_common_create(ndims, numCells, velocity, lengths, 1, comm, out);
  FDB_GUARD_END
}

int fdb_upwind_local_range(const fdb_upwind* h, int64_t* lo, int64_t* hi) {
  if (!h || !lo || !hi) return set_error(FDB_E_INVALID, "null argument");
  *lo = h->field.slabs.front().lo;
  *hi = h->field.slabs.back().hi;
  return FDB_OK;
}

int fdb_upwind_set_field(fdb_upwind* h, const double* host_field) {
  FDB_GUARD_BEGIN
  if (!h || !host_field) return set_error(FDB_E_INVALID, "null argument");
  FDB_TRY(upwind_upload(h, host_field, nullptr, false));
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_upwind_set_slab(fdb_upwind* h, const double* host_slab) {
  FDB_GUARD_BEGIN
  if (!h || !host_slab) return set_error(FDB_E_INVALID, "null argument");
  FDB_TRY(upwind_upload(h, nullptr, host_slab, false));
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_upwind_set_slab_async(fdb_upwind* h, const double* host_slab) {
  FDB_GUARD_BEGIN
  if (!h || !host_slab) return set_error(FDB_E_INVALID, "null argument");
  return upwind_upload(h, nullptr, host_slab, /*async=*/true);
  FDB_GUARD_END
}

int fdb_upwind_reset(fdb_upwind* h) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  FDB_TRY(field_fill_delta(&h->field, h->field.cur, upwind_mirrored_cell0(h, h->field.geo)));
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_upwind_fill_random(fdb_upwind* h, uint64_t seed) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  Field* f = &h->field;
  if (!h->any_flip) {
    FDB_TRY(field_fill_random(f, f->cur, seed, /*publish=*/true));
  } else {  // the hash is a function of the LOGICAL cell index: fill the spare buffer, mirror it into place
    FDB_TRY(field_fill_random(f, 1 - f->cur, seed, /*publish=*/false));
    FDB_TRY(field_mirror(f, 1 - f->cur, f->cur, h->flip, /*publish=*/true));
  }
  return field_sync(f);
  FDB_GUARD_END
}

int fdb_upwind_default_dt(const fdb_upwind* h, double* dt) {
  if (!h || !dt) return set_error(FDB_E_INVALID, "null argument");
  // ref: upwind.cxx:186-192
  const double courant = 0.1;
  double best = DBL_MAX;
  for (int j = 0; j < h->field.geo.ndims; ++j) {
    const double dx = h->lengths[j] / (double)(size_t)h->num_cells[j];
    const double val = courant * dx / fabs(h->velocity[j]);  // |v|: bit-identical for the reference's v > 0
    best = (val < best ? val : best);
  }
  *dt = best;
  return FDB_OK;
}

int fdb_upwind_get_kernel(const fdb_upwind* h, int* kernel) {
  if (!h || !kernel) return set_error(FDB_E_INVALID, "null argument");
  UpwindCoeffs k;
  upwind_coeffs(h, 1.0, &k);
  const bool can = upwind_tma_supported(h->field, k);
  *kernel = (h->kernel == FDB_KERNEL_GENERIC || !can) ? FDB_KERNEL_GENERIC : FDB_KERNEL_TMA;
  return FDB_OK;
}

int fdb_upwind_describe(const fdb_upwind* h, char* text, size_t capacity) {
  if (!h || !text || capacity == 0) return set_error(FDB_E_INVALID, "null argument");
  const Field& f = h->field;
  UpwindCoeffs k;
  upwind_coeffs(h, 1.0, &k);
  int kern = FDB_KERNEL_GENERIC;
  FDB_TRY(fdb_upwind_get_kernel(h, &kern));
  const char* halo = f.single() ? "periodic alias (one slab)"
                     : !f.direct ? (f.comm ? "NCCL send/recv ring" : "event-ordered peer copies")
                     : f.push_stores ? "peer stores from the boundary kernel + stream counters" : "copy engines + stream counters";
  if (kern == FDB_KERNEL_TMA) {
    const int want = (h->fuse == 0) ? kAutoFuse : h->fuse;
    const int fuse = (want > 1 && upwind_fused_supported(f, k, want)) ? want : 1;
    if (fuse > 1)
      snprintf(text, capacity, "%s<T=%d> (tile %s)%s; %d slab(s), halo: %s", upwind_fused_kernel_name(fuse), fuse, upwind_fused_name(fuse),
               h->any_flip ? ", field held mirrored along the axes with a negative velocity" : "", f.nparts, halo);
    else
      snprintf(text, capacity, "upwind3d_tma_kernel%s; %d slab(s), halo: %s", h->any_flip ? " (mirrored axes)" : "", f.nparts, halo);
    return FDB_OK;
  }
  const char* why = h->kernel == FDB_KERNEL_GENERIC ? "asked for with fdb_upwind_set_kernel"
                    : f.geo.ndims != 3              ? "the tiled kernels are 3-D"
                    : (f.geo.n[2] % 2 != 0 || f.geo.n[2] < 4) ? "odd (or < 4) last extent: rows are not 16-byte aligned for TMA"
                    : f.geo.n[1] < 2                ? "a single row per plane"
                    : (k.up[0] != -1 || k.up[1] != -1 || k.up[2] != -1) ? "a negative velocity that was not mirrored (FDB_NO_FLIP)"
                                                    : "no tiled configuration fits";
  snprintf(text, capacity, "upwind_generic_kernel (one thread per cell, several times slower than the tiled kernels): %s; %d slab(s), halo: %s",
           why, f.nparts, halo);
  return FDB_OK;
}

int fdb_upwind_set_kernel(fdb_upwind* h, int kernel) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (kernel != FDB_KERNEL_AUTO && kernel != FDB_KERNEL_GENERIC && kernel != FDB_KERNEL_TMA)
    return set_error(FDB_E_INVALID, "unknown kernel id %d", kernel);
  if (kernel == FDB_KERNEL_TMA) {
    UpwindCoeffs k;
    upwind_coeffs(h, 1.0, &k);
    if (!upwind_tma_supported(h->field, k))
      return set_error(FDB_E_INVALID,
                       "the TMA kernel needs a 3-D grid and an even last extent");
  }
  h->kernel = kernel;
  return FDB_OK;
}

int fdb_upwind_set_fuse(fdb_upwind* h, int steps_per_sweep) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (steps_per_sweep < 0 || steps_per_sweep > kMaxFuse)
    return set_error(FDB_E_INVALID, "steps per sweep must be in 1..%d, or 0 for auto (got %d)", kMaxFuse,
                     steps_per_sweep);
  if (steps_per_sweep > 1) {
    UpwindCoeffs k;
    upwind_coeffs(h, 1.0, &k);
    if (!upwind_fused_supported(h->field, k, steps_per_sweep))
      return set_error(FDB_E_INVALID,
                       "the fused kernel needs what the TMA kernel needs, slabs of at least %d planes and a "
                       "plane of at least 8 x 16 cells", steps_per_sweep);
  }
  h->fuse = steps_per_sweep;
  return FDB_OK;
}

int fdb_upwind_set_stream(fdb_upwind* h, void* cuda_stream) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  return field_set_stream(&h->field, cuda_stream);
  FDB_GUARD_END
}

int fdb_upwind_advect_async(fdb_upwind* h, int64_t numTimeSteps, double deltaTime) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (numTimeSteps < 0) return set_error(FDB_E_INVALID, "negative step count");
  Field* f = &h->field;
  UpwindSweep sw;
  upwind_coeffs(h, deltaTime, &sw.k);
  int kern = FDB_KERNEL_GENERIC;
  FDB_TRY(fdb_upwind_get_kernel(h, &kern));
  sw.tma = (kern == FDB_KERNEL_TMA);
  const int want = (h->fuse == 0) ? kAutoFuse : h->fuse;
  const int fuse = (sw.tma && want > 1 && upwind_fused_supported(*f, sw.k, want)) ? want : 1;
  // plan: sweeps of `fuse` time steps, then the remainder (depths never increase along a plan)
  std::vector<int> depths;
  for (int64_t done = 0; done < numTimeSteps;) {
    const int64_t left = numTimeSteps - done;
    int t = (int)(left < fuse ? left : fuse);
    if (t > 1 && !upwind_fused_supported(*f, sw.k, t)) t = 1;
    depths.push_back(t);
    done += t;
  }
  // a plan ending in [3, 1] runs as [2, 2], one ending in [4, 1] as [3, 2]: the single-step kernel is the slowest per
  // time step (depths still never increase along the plan)
  if (depths.size() >= 2 && depths.back() == 1) {
    const int before = depths[depths.size() - 2];
    if (before == 3 && upwind_fused_supported(*f, sw.k, 2)) {
      depths[depths.size() - 2] = 2;
      depths.back() = 2;
    } else if (before == 4 && upwind_fused_supported(*f, sw.k, 3) && upwind_fused_supported(*f, sw.k, 2)) {
      depths[depths.size() - 2] = 3;
      depths.back() = 2;
    }
  }
  FDB_TRY(timing_begin(f));
  FDB_TRY(field_run_sweeps(f, &sw, depths.data(), (int)depths.size()));
  FDB_TRY(timing_end(f));
  f->last_updates = (double)numTimeSteps * (double)f->geo.total();
  return FDB_OK;
  FDB_GUARD_END
}

int fdb_upwind_sync(fdb_upwind* h) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  FDB_TRY(field_sync(&h->field));
  return timing_collect(&h->field);
}

int fdb_upwind_advect(fdb_upwind* h, int64_t numTimeSteps, double deltaTime) {
  FDB_TRY(fdb_upwind_advect_async(h, numTimeSteps, deltaTime));
  return fdb_upwind_sync(h);
}

int fdb_upwind_checksum(fdb_upwind* h, double* sum) {
  FDB_GUARD_BEGIN
  if (!h || !sum) return set_error(FDB_E_INVALID, "null argument");
  return field_sum(&h->field, h->field.cur, sum);
  FDB_GUARD_END
}

int fdb_upwind_plane_sums(fdb_upwind* h, double* sums, int64_t capacity, int64_t* count) {
  FDB_GUARD_BEGIN
  if (!h || !count) return set_error(FDB_E_INVALID, "null argument");
  const Geometry& g = h->field.geo;
  *count = g.n[0];
  if (!sums) return FDB_OK;  // size query
  if (capacity < g.n[0]) return set_error(FDB_E_INVALID, "room for %lld plane sums, %lld needed", (long long)capacity, (long long)g.n[0]);
  return field_plane_sums(&h->field, h->field.cur, sums);  // already in global plane order
  FDB_GUARD_END
}

int fdb_upwind_std(fdb_upwind* h, double* stddev) {
  FDB_GUARD_BEGIN
  if (!h || !stddev) return set_error(FDB_E_INVALID, "null argument");
  // ref: upwind.cxx:95-103
  double sum = 0, sq = 0;
  FDB_TRY(field_sum(&h->field, h->field.cur, &sum));
  const double ntot = (double)(size_t)h->field.geo.total();
  const double mean = sum / ntot;
  FDB_TRY(field_sqdev(&h->field, h->field.cur, mean, &sq));
  *stddev = sqrt(sq / ntot);
  return FDB_OK;
  FDB_GUARD_END
}

int fdb_upwind_get_field(fdb_upwind* h, double* host_field) {
  FDB_GUARD_BEGIN
  if (!h || !host_field) return set_error(FDB_E_INVALID, "null argument");
  return upwind_download(h, host_field, nullptr);
  FDB_GUARD_END
}

int fdb_upwind_get_slab(fdb_upwind* h, double* host_slab) {
  FDB_GUARD_BEGIN
  if (!h || !host_slab) return set_error(FDB_E_INVALID, "null argument");
  return upwind_download(h, nullptr, host_slab);
  FDB_GUARD_END
}

int fdb_upwind_last_timing(const fdb_upwind* h, double* gpu_ms, double* cell_updates,
                           double* halo_bytes) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (gpu_ms) *gpu_ms = h->field.last_ms;
  if (cell_updates) *cell_updates = h->field.last_updates;
  if (halo_bytes) *halo_bytes = h->field.last_halo_bytes / (double)h->field.ngpus;
  return FDB_OK;
}

int fdb_upwind_destroy(fdb_upwind* h) {
  if (!h) return FDB_OK;
  field_destroy(&h->field);
  delete h;
  return FDB_OK;
}

// ---- stencil (Filter) -----------------------------------------------------------------------
int fdb_stencil_create(int ndims, const int64_t* globalDims, int nbranch, const int* offsets,
                       const double* weights, int ngpus, fdb_stencil** out) {
  FDB_GUARD_BEGIN
  return stencil_common_create(ndims, globalDims, nbranch, offsets, weights, ngpus, nullptr, out);
  FDB_GUARD_END
}

int fdb_stencil_create_dist(int ndims, const int64_t* globalDims, int nbranch, const int* offsets,
                            const double* weights, fdb_comm* comm, fdb_stencil** out) {
  FDB_GUARD_BEGIN
  if (!comm) return set_error(FDB_E_INVALID, "null communicator");
  return stencil_common_create(ndims, globalDims, nbranch, offsets, weights, 1, comm, out);
  FDB_GUARD_END
}

int fdb_stencil_local_range(const fdb_stencil* h, int64_t* lo, int64_t* hi) {
  if (!h || !lo || !hi) return set_error(FDB_E_INVALID, "null argument");
  *lo = h->field.slabs.front().lo;
  *hi = h->field.slabs.back().hi;
  if (h->ndims == 2 && h->field.geo.n[0] == 1) *hi = h->dims[0];  // a 2-D problem carried as one plane: all of axis 0
  return FDB_OK;
}

// Column-major (first reference axis fastest) host data of extents d[0..nd) is the
// row-major array of the reversed extents: permute on the device.
static int stencil_colmajor_io(fdb_stencil* h, int which_buf, double* host, bool upload) {
  Field* f = &h->field;
  if (f->nparts != 1)
    return set_error(FDB_E_STATE, "column-major transfers need a single-device handle");
  if (h->ndims == 1) {
    return upload ? field_upload(f, which_buf, host, nullptr) : field_download(f, which_buf, host, nullptr);
  }
  Slab& s = f->slabs[0];
  FDB_CUDA(cudaSetDevice(s.device));
  const int64_t total = f->geo.total();
  double* tmp = nullptr;
  FDB_CUDA(cudaMalloc(&tmp, (size_t)total * sizeof(double)));
  // reference extents (e0,e1,e2) padded on the left so that permute210 applies
  int64_t e0 = 1, e1 = 1, e2 = 1;
  if (h->ndims == 3) { e0 = h->dims[0]; e1 = h->dims[1]; e2 = h->dims[2]; }
  else { e0 = h->dims[0]; e1 = 1; e2 = h->dims[1]; }
  int rc = FDB_OK;
  if (upload) {
    // host col-major == row-major (e2, e1, e0) -> device row-major (e0, e1, e2)
    FDB_TRY(field_sync(f));
    cudaError_t ce = cudaMemcpyAsync(tmp, host, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, s.s_main);
    if (ce != cudaSuccess) { cudaFree(tmp); return set_error(FDB_E_CUDA, "H2D copy: %s", cudaGetErrorString(ce)); }
    rc = launch_permute(tmp, f->body(0, which_buf), e2, e1, e0, s.s_main);
  } else {
    rc = launch_permute(f->body(0, which_buf), tmp, e0, e1, e2, s.s_main);
    if (rc == FDB_OK) {
      cudaError_t ce = cudaMemcpyAsync(host, tmp, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s.s_main);
      if (ce != cudaSuccess) rc = set_error(FDB_E_CUDA, "D2H copy: %s", cudaGetErrorString(ce));
    }
  }
  cudaStreamSynchronize(s.s_main);
  cudaFree(tmp);
  return rc;
}

int fdb_stencil_set_input(fdb_stencil* h, const double* host_field, int layout) {
  FDB_GUARD_BEGIN
  if (!h || !host_field) return set_error(FDB_E_INVALID, "null argument");
  Field* f = &h->field;
  if (layout == FDB_ROW_MAJOR) {
    FDB_TRY(field_upload(f, f->cur, host_field, nullptr));
  } else if (layout == FDB_COL_MAJOR) {
    FDB_TRY(stencil_colmajor_io(h, f->cur, const_cast<double*>(host_field), true));
    // publish (records events, refreshes ghosts)
    for (auto& s : f->slabs) {
      FDB_CUDA(cudaSetDevice(s.device));
      FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
    }
  } else {
    return set_error(FDB_E_INVALID, "unknown layout %d", layout);
  }
  h->out_valid = false;
  return field_sync(f);
  FDB_GUARD_END
}

int fdb_stencil_set_input_slab(fdb_stencil* h, const double* host_slab) {
  FDB_GUARD_BEGIN
  if (!h || !host_slab) return set_error(FDB_E_INVALID, "null argument");
  FDB_TRY(field_upload(&h->field, h->field.cur, nullptr, host_slab));
  h->out_valid = false;
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_stencil_fill_random(fdb_stencil* h, uint64_t seed) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  // the hash runs over the device's row-major cell order; a single-plane 2-D handle has the same order
  FDB_TRY(field_fill_random(&h->field, h->field.cur, seed, /*publish=*/true));
  h->out_valid = false;
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_stencil_set_input_separable(fdb_stencil* h, const double* const* factors) {
  FDB_GUARD_BEGIN
  if (!h || !factors) return set_error(FDB_E_INVALID, "null argument");
  for (int j = 0; j < h->ndims; ++j)
    if (!factors[j]) return set_error(FDB_E_INVALID, "null factor array for axis %d", j);
  FDB_TRY(field_fill_separable(&h->field, h->field.cur, h->ndims, factors, h->dims));
  h->out_valid = false;
  return field_sync(&h->field);
  FDB_GUARD_END
}

int fdb_stencil_set_ref_wrap(fdb_stencil* h, int on) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (on && !h->field.single())
    return set_error(FDB_E_STATE, "the reference's wrap is only reproduced on a single slab (its off-rank indices go "
                     "through MPI windows, not through the modulo)");
  h->br.ref_wrap = on != 0;
  return FDB_OK;
}

int fdb_stencil_get_kernel(const fdb_stencil* h, int* kernel) {
  if (!h || !kernel) return set_error(FDB_E_INVALID, "null argument");
  const bool can = !h->br.ref_wrap && stencil_lap7_supported(h->field, h->br);
  *kernel = (h->kernel == FDB_KERNEL_GENERIC || !can) ? FDB_KERNEL_GENERIC : FDB_KERNEL_TMA;
  return FDB_OK;
}

int fdb_stencil_describe(const fdb_stencil* h, char* text, size_t capacity) {
  if (!h || !text || capacity == 0) return set_error(FDB_E_INVALID, "null argument");
  const Field& f = h->field;
  int kern = FDB_KERNEL_GENERIC, fuse = 1;
  FDB_TRY(fdb_stencil_get_kernel(h, &kern));
  FDB_TRY(fdb_stencil_get_fuse(h, &fuse));
  if (kern == FDB_KERNEL_TMA) {
    if (fuse == 2)
      snprintf(text, capacity, "iterate: %s (two applies per sweep, tile %s), apply: lap7_tma_kernel; %d slab(s)",
               stencil_lap7_fused_kernel_name(), stencil_lap7_fused_name(f), f.nparts);
    else
      snprintf(text, capacity, "lap7_tma_kernel (one apply per sweep%s); %d slab(s)",
               f.geo.ndims == 2 ? ", 2-D problem carried as one plane" : "", f.nparts);
    return FDB_OK;
  }
  bool seven = h->br.nbranch <= 7;
  for (int b = 0; b < h->br.nbranch && seven; ++b)
    seven = std::abs(h->br.off[b][0]) + std::abs(h->br.off[b][1]) + std::abs(h->br.off[b][2]) <= 1;
  const char* why = h->kernel == FDB_KERNEL_GENERIC ? "asked for with fdb_stencil_set_kernel"
                    : h->br.ref_wrap                ? "reference index-wrap compatibility mode"
                    : !seven                        ? "offsets beyond the radius-1 axis-aligned (7-point) set"
                    : (f.geo.ndims == 1 || (f.geo.ndims == 2 && f.geo.n[0] != 1))
                        ? "1-D, or a 2-D problem on several slabs: the tiled kernel needs whole planes"
                        : "odd (or < 4) last extent, or a single row per plane";
  snprintf(text, capacity, "stencil_generic_kernel (one thread per cell, several times slower than the tiled kernels): %s; %d slab(s)",
           why, f.nparts);
  return FDB_OK;
}

int fdb_stencil_set_kernel(fdb_stencil* h, int kernel) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (kernel != FDB_KERNEL_AUTO && kernel != FDB_KERNEL_GENERIC && kernel != FDB_KERNEL_TMA)
    return set_error(FDB_E_INVALID, "unknown kernel id %d", kernel);
  if (kernel == FDB_KERNEL_TMA && (h->br.ref_wrap || !stencil_lap7_supported(h->field, h->br)))
    return set_error(FDB_E_INVALID, "the TMA kernel only runs the 3-D 7-point stencil with an even last extent");
  h->kernel = kernel;
  return FDB_OK;
}

int fdb_stencil_set_fuse(fdb_stencil* h, int applies_per_sweep) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (applies_per_sweep < 0 || applies_per_sweep > 2)
    return set_error(FDB_E_INVALID, "applies per sweep must be 1 or 2, or 0 for auto (got %d)", applies_per_sweep);
  if (applies_per_sweep == 2 && !stencil_lap7_fused_supported(h->field, h->br))
    return set_error(FDB_E_INVALID,
                     "the fused kernel runs the 3-D 7-point stencil on planes of 8k x 128m cells and slabs of at "
                     "least two planes");
  h->fuse = applies_per_sweep;
  return FDB_OK;
}

int fdb_stencil_get_fuse(const fdb_stencil* h, int* applies_per_sweep) {
  if (!h || !applies_per_sweep) return set_error(FDB_E_INVALID, "null argument");
  int kern = FDB_KERNEL_GENERIC;
  FDB_TRY(fdb_stencil_get_kernel(h, &kern));
  *applies_per_sweep =
      (kern == FDB_KERNEL_TMA && h->fuse != 1 && stencil_lap7_fused_supported(h->field, h->br)) ? 2 : 1;
  return FDB_OK;
}

int fdb_stencil_set_stream(fdb_stencil* h, void* cuda_stream) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  return field_set_stream(&h->field, cuda_stream);
  FDB_GUARD_END
}

static int stencil_apply_async(fdb_stencil* h) {
  Field* f = &h->field;
  StencilSweep sw;
  sw.b = &h->br;
  int kern = FDB_KERNEL_GENERIC;
  FDB_TRY(fdb_stencil_get_kernel(h, &kern));
  sw.fast = (kern == FDB_KERNEL_TMA);
  const int depth = h->reach;
  FDB_TRY(field_run_sweeps(f, &sw, &depth, 1));
  f->cur = 1 - f->cur;  // an apply leaves `cur` on the input; the swap flips it (copyOutToIn)
  h->out_valid = true;
  return FDB_OK;
}

// The sweep also exchanged the ghosts of the output buffer, so the O(1) swap
// leaves a fully valid input field (ref: copyOutToIn, Filter.cpp:440-463, which
// copies the block and repacks every window).
static int stencil_swap(fdb_stencil* h) {
  if (!h->out_valid) return FDB_OK;  // in == out already (nothing applied since the last swap)
  h->field.cur = 1 - h->field.cur;
  h->out_valid = false;
  return FDB_OK;
}

int fdb_stencil_apply(fdb_stencil* h) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  Field* f = &h->field;
  FDB_TRY(timing_begin(f));
  FDB_TRY(stencil_apply_async(h));
  FDB_TRY(timing_end(f));
  f->last_updates = (double)f->geo.total();
  FDB_TRY(field_sync(f));
  return timing_collect(f);
  FDB_GUARD_END
}

int fdb_stencil_swap(fdb_stencil* h) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  return stencil_swap(h);
}

int fdb_stencil_iterate(fdb_stencil* h, int64_t niter) {
  FDB_GUARD_BEGIN
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (niter < 0) return set_error(FDB_E_INVALID, "negative iteration count");
  Field* f = &h->field;
  StencilSweep sw;
  sw.b = &h->br;
  int kern = FDB_KERNEL_GENERIC;
  FDB_TRY(fdb_stencil_get_kernel(h, &kern));
  sw.fast = (kern == FDB_KERNEL_TMA);
  const bool fuse = sw.fast && h->fuse != 1 && stencil_lap7_fused_supported(*f, h->br);
  if (h->fuse == 2 && !fuse)
    return set_error(FDB_E_STATE, "two applies per sweep were requested but the fused kernel cannot run this problem");
  sw.fused_depth = fuse ? 2 * h->reach : 0;
  // plan: pairs of applies in one sweep each, then the odd one
  std::vector<int> depths;
  for (int64_t done = 0; done < niter;) {
    const int t = (fuse && niter - done >= 2) ? 2 : 1;
    depths.push_back(t * h->reach);
    done += t;
  }
  FDB_TRY(timing_begin(f));
  // every sweep flips `cur` onto its output, which is exactly apply + swap
  FDB_TRY(field_run_sweeps(f, &sw, depths.data(), (int)depths.size()));
  if (niter > 0) h->out_valid = false;
  FDB_TRY(timing_end(f));
  f->last_updates = (double)niter * (double)f->geo.total();
  FDB_TRY(field_sync(f));
  return timing_collect(f);
  FDB_GUARD_END
}

// after a swap "output" reads as the same data as "input" (copyOutToIn semantics)
static int stencil_buffer(const fdb_stencil* h, int which, int* p) {
  if (which != FDB_INPUT && which != FDB_OUTPUT) return set_error(FDB_E_INVALID, "which must be FDB_INPUT or FDB_OUTPUT");
  *p = (which == FDB_OUTPUT && h->out_valid) ? 1 - h->field.cur : h->field.cur;
  return FDB_OK;
}

int fdb_stencil_checksum(fdb_stencil* h, int which, double* sum) {
  FDB_GUARD_BEGIN
  if (!h || !sum) return set_error(FDB_E_INVALID, "null argument");
  int p = 0;
  FDB_TRY(stencil_buffer(h, which, &p));
  return field_sum(&h->field, p, sum);
  FDB_GUARD_END
}

int fdb_stencil_sumsq(fdb_stencil* h, int which, double* sumsq) {
  FDB_GUARD_BEGIN
  if (!h || !sumsq) return set_error(FDB_E_INVALID, "null argument");
  int p = 0;
  FDB_TRY(stencil_buffer(h, which, &p));
  return field_sqdev(&h->field, p, 0.0, sumsq);
  FDB_GUARD_END
}

int fdb_stencil_get(fdb_stencil* h, int which, double* host_field, int layout) {
  FDB_GUARD_BEGIN
  if (!h || !host_field) return set_error(FDB_E_INVALID, "null argument");
  int p = 0;
  FDB_TRY(stencil_buffer(h, which, &p));
  if (layout == FDB_ROW_MAJOR) return field_download(&h->field, p, host_field, nullptr);
  if (layout == FDB_COL_MAJOR) return stencil_colmajor_io(h, p, host_field, false);
  return set_error(FDB_E_INVALID, "unknown layout %d", layout);
  FDB_GUARD_END
}

int fdb_stencil_get_slab(fdb_stencil* h, int which, double* host_slab) {
  FDB_GUARD_BEGIN
  if (!h || !host_slab) return set_error(FDB_E_INVALID, "null argument");
  int p = 0;
  FDB_TRY(stencil_buffer(h, which, &p));
  return field_download(&h->field, p, nullptr, host_slab);
  FDB_GUARD_END
}

int fdb_stencil_last_timing(const fdb_stencil* h, double* gpu_ms, double* cell_updates,
                            double* halo_bytes) {
  if (!h) return set_error(FDB_E_INVALID, "null handle");
  if (gpu_ms) *gpu_ms = h->field.last_ms;
  if (cell_updates) *cell_updates = h->field.last_updates;
  if (halo_bytes) *halo_bytes = h->field.last_halo_bytes / (double)h->field.ngpus;
  return FDB_OK;
}

int fdb_stencil_destroy(fdb_stencil* h) {
  if (!h) return FDB_OK;
  field_destroy(&h->field);
  delete h;
  return FDB_OK;
}

}  // extern "C"
