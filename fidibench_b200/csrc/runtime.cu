// runtime.cu -- device context of the fidib200 engines: slab-decomposed fields,
// streams/events, halo exchange (peer copies in one process, NCCL send/recv
// between processes), uploads/downloads and the host side of the reductions.
//
// Replaces, B200-first, what the reference spreads over CubeDecomp
// (cxx/CubeDecomp.cpp:11-131), Filter's MPI-3 RMA windows (cxx/Filter.cpp:114-129,
// :320-340) and copyOutToIn's window repacking (:440-463): slabs along axis 0 make
// every halo a contiguous run of planes, received straight into the ghost planes
// of the field allocation.
#include "fdb_internal.h"

#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

namespace fdb {

// ---- errors / counters ----------------------------------------------------------
static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

static int64_t g_launches = 0;
void count_launch(int64_t n) { __atomic_fetch_add(&g_launches, n, __ATOMIC_RELAXED); }
int64_t launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---- geometry ---------------------------------------------------------------------
int make_geometry(int ndims, const int64_t* dims, Geometry* g, bool plane2d) {
  if (ndims < 1 || ndims > 3) return set_error(FDB_E_INVALID, "ndims must be 1, 2 or 3 (got %d)", ndims);
  if (!dims) return set_error(FDB_E_INVALID, "null dimensions");
  for (int j = 0; j < ndims; ++j)
    if (dims[j] < 1) return set_error(FDB_E_INVALID, "extent %d is %lld", j, (long long)dims[j]);
  *g = Geometry();
  g->ndims = ndims;
  if (ndims == 3) {
    for (int j = 0; j < 3; ++j) { g->n[j] = dims[j]; g->axis_of[j] = j; }
  } else if (ndims == 2 && plane2d) {
    // one plane: both axes are in-plane, so the tiled (TMA) kernels apply; no slab axis
    g->n[0] = 1; g->n[1] = dims[0]; g->n[2] = dims[1];
    g->axis_of[0] = 1; g->axis_of[1] = 2;
    g->active[0] = false;
  } else if (ndims == 2) {
    g->n[0] = dims[0]; g->n[1] = 1; g->n[2] = dims[1];
    g->axis_of[0] = 0; g->axis_of[1] = 2;
    g->active[1] = false;
  } else {
    g->n[0] = 1; g->n[1] = 1; g->n[2] = dims[0];
    g->axis_of[0] = 2;
    g->active[0] = g->active[1] = false;
  }
  return FDB_OK;
}

// ---- tensor maps ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 3-D FP64 tensor (n0 planes x n1 rows x n2 cells, n2 contiguous), box {box2, box1, 1}
int encode_tensor_map_3d(CUtensorMap* tm, const double* base, int64_t n2, int64_t n1, int64_t n0,
                         int box2, int box1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(FDB_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0};
  cuuint64_t strides[2] = {(cuuint64_t)n2 * 8, (cuuint64_t)n2 * (cuuint64_t)n1 * 8};
  cuuint32_t box[3] = {(cuuint32_t)box2, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  // L2 promotion of the TMA requests: 256 B unless FDB_TMA_L2PROMO says 0 (none), 64 or 128 (tuning)
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (const char* pv = getenv("FDB_TMA_L2PROMO")) {
    const int v = atoi(pv);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                   : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                              : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FDB_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) dims=%lldx%lldx%lld box=%dx%d",
                     (int)r, (long long)n0, (long long)n1, (long long)n2, box1, box2);
  return FDB_OK;
}

// ---- stream memory operations (driver API, resolved at run time) -------------------------
enum { F_GHOST_LO = 0, F_GHOST_HI = 1, F_ACK_NEXT = 2, F_ACK_PREV = 3, F_DONE_CTR = 6 /* local: single-launch sweeps */, F_COUNT = 8 };
enum { NBR_PREV = 0, NBR_NEXT = 1 };

typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
static StreamValueFn g_wait_value = nullptr, g_write_value = nullptr;

static bool stream_memops_available() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_wait_value = reinterpret_cast<StreamValueFn>(p);
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_write_value = reinterpret_cast<StreamValueFn>(p);
    (void)cudaGetLastError();
  });
  return g_wait_value && g_write_value;
}

// stream `s` proceeds once *addr >= v
static int stream_wait_geq(cudaStream_t s, uint64_t* addr, uint64_t v) {
  CUresult r = g_wait_value((CUstream)s, (CUdeviceptr)(uintptr_t)addr, v, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return set_error(FDB_E_CUDA, "cuStreamWaitValue64 failed (CUresult %d)", (int)r);
  return FDB_OK;
}
// *addr = v once everything before it in `s` is done and visible
static int stream_write(cudaStream_t s, uint64_t* addr, uint64_t v) {
  CUresult r = g_write_value((CUstream)s, (CUdeviceptr)(uintptr_t)addr, v, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) return set_error(FDB_E_CUDA, "cuStreamWriteValue64 failed (CUresult %d)", (int)r);
  return FDB_OK;
}

// Map the neighbours' buffers and counters of a one-slab-per-process field into this
// process (CUDA IPC); the handles travel through one NCCL all-gather.
static int field_open_neighbours_ipc(Field* f) {
  Slab& s = f->slabs[0];
  fdb_comm* c = f->comm;
  const int n = c->nranks;
  struct Pack { cudaIpcMemHandle_t h[3]; double ok; };
  static_assert(sizeof(Pack) % sizeof(double) == 0, "pack size");
  const size_t words = sizeof(Pack) / sizeof(double);
  // Every rank runs the SAME sequence of collectives whatever fails locally (a rank that skipped the
  // all-gather would leave the others hanging in it): local failures travel inside the pack.
  Pack mine;
  memset(&mine, 0, sizeof(mine));
  int local = FDB_OK;
  {
    cudaError_t e = cudaIpcGetMemHandle(&mine.h[0], s.buf[0]);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.h[1], s.buf[1]);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.h[2], s.flags);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      local = set_error(FDB_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
  }
  mine.ok = (local == FDB_OK) ? 1.0 : 0.0;
  std::vector<Pack> all((size_t)n);
  double *dsend = nullptr, *drecv = nullptr;
  auto gather = [&]() -> int {
    FDB_CUDA(cudaMalloc(&dsend, sizeof(Pack)));
    FDB_CUDA(cudaMalloc(&drecv, sizeof(Pack) * (size_t)n));
    FDB_CUDA(cudaMemcpy(dsend, &mine, sizeof(Pack), cudaMemcpyHostToDevice));
    FDB_NCCL(ncclAllGather(dsend, drecv, words, ncclDouble, c->nccl, c->stream));
    FDB_CUDA(cudaStreamSynchronize(c->stream));
    FDB_CUDA(cudaMemcpy(all.data(), drecv, sizeof(Pack) * (size_t)n, cudaMemcpyDeviceToHost));
    return FDB_OK;
  };
  const int grc = gather();
  if (dsend) cudaFree(dsend);
  if (drecv) cudaFree(drecv);
  FDB_TRY(grc);
  if (local != FDB_OK) return local;
  for (int r = 0; r < n; ++r)
    if (all[(size_t)r].ok != 1.0) return set_error(FDB_E_CUDA, "rank %d could not export its buffers over CUDA IPC", r);
  const int prev = f->prev_of(c->rank), next = f->next_of(c->rank);
  int opened = 0;
  auto open3 = [&](int rank, int side) -> int {
    void* p[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; ++k) {
      cudaError_t e = cudaIpcOpenMemHandle(&p[k], all[(size_t)rank].h[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return set_error(FDB_E_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", rank, cudaGetErrorString(e));
      }
      s.ipc_opened[opened++] = p[k];
    }
    s.nbr_buf[side][0] = static_cast<double*>(p[0]);
    s.nbr_buf[side][1] = static_cast<double*>(p[1]);
    s.nbr_flags[side] = static_cast<uint64_t*>(p[2]);
    return FDB_OK;
  };
  FDB_TRY(open3(next, NBR_NEXT));
  if (prev == next) {  // two ranks: one neighbour on both sides, one mapping
    s.nbr_buf[NBR_PREV][0] = s.nbr_buf[NBR_NEXT][0];
    s.nbr_buf[NBR_PREV][1] = s.nbr_buf[NBR_NEXT][1];
    s.nbr_flags[NBR_PREV] = s.nbr_flags[NBR_NEXT];
  } else {
    FDB_TRY(open3(prev, NBR_PREV));
  }
  return FDB_OK;
}

// ---- field --------------------------------------------------------------------------
double* Field::ghost_lo(int d, int p) const {
  if (single()) return body(d, p) + (slabs[d].nloc() - G) * geo.plane();  // periodic alias
  return slabs[d].buf[p];
}
double* Field::ghost_hi(int d, int p) const {
  if (single()) return body(d, p);  // periodic alias
  return body(d, p) + slabs[d].nloc() * geo.plane();
}

int tma_encode_slab(Field* f, int d);       // kernels_tma.cu (upwind box shapes)
int tma_encode_slab_lap7(Field* f, int d);  // kernels_tma.cu (7-point stencil box shapes)

int field_create(Field* f, const Geometry& geo, int G, bool need_lo, bool need_hi, int ngpus,
                 fdb_comm* comm, int want_tma, bool ring_reversed) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return set_error(FDB_E_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  f->geo = geo;
  f->G = G;
  f->need_lo = need_lo;
  f->need_hi = need_hi;
  f->comm = comm;
  f->cur = 0;
  f->ring_reversed = ring_reversed;
  f->planes_mirrored = ring_reversed;
  if (comm) {
    ++comm->users;  // fdb_comm_destroy refuses while handles still use the communicator
    f->ngpus = 1;
    f->nparts = comm->nranks;
  } else {
    if (ngpus < 1) return set_error(FDB_E_INVALID, "ngpus must be >= 1 (got %d)", ngpus);
    if (ngpus > ndev)
      return set_error(FDB_E_INVALID, "ngpus=%d but only %d CUDA device(s) visible", ngpus, ndev);
    f->ngpus = ngpus;
    f->nparts = ngpus;
  }
  if (geo.n[0] % f->nparts != 0)
    return set_error(FDB_E_DECOMP,
                     "No valid domain decomposition: %d slab(s) do not divide the %lld planes of axis 0",
                     f->nparts, (long long)geo.n[0]);
  const int64_t nloc = geo.n[0] / f->nparts;
  if (nloc < G)
    return set_error(FDB_E_DECOMP, "slabs of %lld plane(s) are thinner than the %d ghost plane(s)",
                     (long long)nloc, G);
  const int64_t plane = geo.plane();
  f->slabs.assign(f->ngpus, Slab());
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    const int part = comm ? comm->rank : d;
    s.device = comm ? comm->device : d;
    s.lo = part * nloc;
    s.hi = s.lo + nloc;
    FDB_CUDA(cudaSetDevice(s.device));
    const size_t bytes = (size_t)(nloc + 2 * G) * (size_t)plane * sizeof(double);
    for (int p = 0; p < 2; ++p) {
      FDB_CUDA(cudaMalloc(&s.buf[p], bytes));
      FDB_CUDA(cudaMemset(s.buf[p], 0, bytes));
    }
    const int64_t nchunk = reduce_partials_per_plane(plane);
    FDB_CUDA(cudaMalloc(&s.partial, (size_t)(nloc * nchunk) * sizeof(double)));
    FDB_CUDA(cudaMalloc(&s.plane_sums, (size_t)nloc * sizeof(double)));
    FDB_CUDA(cudaMalloc(&s.flags, F_COUNT * sizeof(uint64_t)));
    FDB_CUDA(cudaMemset(s.flags, 0, F_COUNT * sizeof(uint64_t)));
    int lo_prio = 0, hi_prio = 0;
    FDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    FDB_CUDA(cudaStreamCreateWithPriority(&s.s_main, cudaStreamNonBlocking, lo_prio));
    FDB_CUDA(cudaStreamCreateWithPriority(&s.s_bnd, cudaStreamNonBlocking, hi_prio));
    s.own_main = true;
    FDB_CUDA(cudaEventCreateWithFlags(&s.ev_local_done, cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreateWithFlags(&s.ev_bnd_done, cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreateWithFlags(&s.ev_xchg_done, cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreateWithFlags(&s.ev_ghost_ready[0], cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreateWithFlags(&s.ev_ghost_ready[1], cudaEventDisableTiming));
    FDB_CUDA(cudaEventCreate(&s.ev_t0));
    FDB_CUDA(cudaEventCreate(&s.ev_t1));
    FDB_CUDA(cudaDeviceSynchronize());  // memsets done before anything else touches the buffers
    FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
    FDB_CUDA(cudaEventRecord(s.ev_ghost_ready[0], s.s_bnd));
    FDB_CUDA(cudaEventRecord(s.ev_ghost_ready[1], s.s_bnd));
    FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_bnd));
  }
  // peer access between neighbouring devices of this process.  The direct transport (raw neighbour
  // pointers, kernel peer stores, stream memory operations on the neighbour's counters) needs it on
  // EVERY neighbour pair; without it the event-ordered cudaMemcpyPeerAsync transport is used, which
  // stages through the host when it has to.  FDB_NO_PEER=1 pretends there is none (tests).
  bool all_peer = true;
  if (!comm && f->ngpus > 1) {
    const char* np = getenv("FDB_NO_PEER");
    const bool pretend_none = np && *np && atoi(np) != 0;
    for (int d = 0; d < f->ngpus; ++d) {
      FDB_CUDA(cudaSetDevice(f->slabs[d].device));
      for (int o : {f->next_of(d), f->prev_of(d)}) {
        if (o == d) continue;
        int can = 0;
        FDB_CUDA(cudaDeviceCanAccessPeer(&can, f->slabs[d].device, f->slabs[o].device));
        if (can && !pretend_none) {
          cudaError_t pe = cudaDeviceEnablePeerAccess(f->slabs[o].device, 0);
          if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) all_peer = false;
          (void)cudaGetLastError();
        } else {
          all_peer = false;
        }
      }
    }
  }
  // halo transport
  const char* halo_env = getenv("FDB_HALO");
  f->direct = !(halo_env && strcmp(halo_env, "nccl") == 0) && stream_memops_available();
  f->push_stores = !(halo_env && strcmp(halo_env, "copy") == 0);  // FDB_HALO=copy: copy engines only
  if (!all_peer) { f->direct = false; f->push_stores = false; }
  if (f->nparts > 1 && f->direct) {
    if (comm) {
      int rc = field_open_neighbours_ipc(f);
      int ok = (rc == FDB_OK) ? 1 : 0, all_ok = ok;
      {  // every rank must take the same transport
        int* dflag = reinterpret_cast<int*>(comm->scratch);
        FDB_CUDA(cudaMemcpy(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice));
        FDB_NCCL(ncclAllReduce(dflag, dflag + 1, 1, ncclInt, ncclMin, comm->nccl, comm->stream));
        FDB_CUDA(cudaStreamSynchronize(comm->stream));
        FDB_CUDA(cudaMemcpy(&all_ok, dflag + 1, sizeof(int), cudaMemcpyDeviceToHost));
      }
      if (!all_ok) f->direct = false;  // e.g. no IPC between these processes: NCCL send/recv instead
    } else {
      const int g = f->ngpus;
      for (int d = 0; d < g; ++d) {
        Slab& s = f->slabs[d];
        const Slab& pv = f->slabs[f->prev_of(d)];
        const Slab& nx = f->slabs[f->next_of(d)];
        for (int p = 0; p < 2; ++p) { s.nbr_buf[NBR_PREV][p] = pv.buf[p]; s.nbr_buf[NBR_NEXT][p] = nx.buf[p]; }
        s.nbr_flags[NBR_PREV] = pv.flags;
        s.nbr_flags[NBR_NEXT] = nx.flags;
      }
    }
  }
  for (int d = 0; d < f->ngpus; ++d) {
    if (want_tma == 1) FDB_TRY(tma_encode_slab(f, d));
    if (want_tma == 2) FDB_TRY(tma_encode_slab_lap7(f, d));
  }
  return FDB_OK;
}

void field_destroy(Field* f) {
  for (auto& s : f->slabs) {
    cudaSetDevice(s.device);
    cudaDeviceSynchronize();
  }
  if (f->comm && f->comm->nranks > 1 && f->direct && !f->slabs.empty()) {
    // neighbours write into this slab's memory: nobody frees before everybody is done
    double v = 0.0;
    fdb_comm_max(f->comm, &v);
  }
  for (auto& s : f->slabs) {
    cudaSetDevice(s.device);
    for (void*& p : s.ipc_opened)
      if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    if (s.flags) cudaFree(s.flags);
    for (int p = 0; p < 2; ++p) if (s.buf[p]) cudaFree(s.buf[p]);
    if (s.partial) cudaFree(s.partial);
    if (s.plane_sums) cudaFree(s.plane_sums);
    if (s.s_main && s.own_main) cudaStreamDestroy(s.s_main);
    if (s.s_bnd) cudaStreamDestroy(s.s_bnd);
    for (cudaEvent_t ev : {s.ev_local_done, s.ev_bnd_done, s.ev_xchg_done, s.ev_ghost_ready[0], s.ev_ghost_ready[1],
                           s.ev_t0, s.ev_t1})
      if (ev) cudaEventDestroy(ev);
  }
  f->slabs.clear();
  for (auto& g : f->plan_graphs) cudaGraphExecDestroy(g.exec);
  f->plan_graphs.clear();
  if (f->s_capture) { cudaStreamDestroy(f->s_capture); f->s_capture = nullptr; }
  if (f->comm) {
    --f->comm->users;
    f->comm = nullptr;
  }
  (void)cudaGetLastError();
}

int field_set_stream(Field* f, void* stream) {
  if (f->ngpus != 1)
    return set_error(FDB_E_STATE, "a caller stream can only drive a single-device handle");
  Slab& s = f->slabs[0];
  FDB_CUDA(cudaSetDevice(s.device));
  FDB_CUDA(cudaStreamSynchronize(s.s_main));
  FDB_CUDA(cudaStreamSynchronize(s.s_bnd));
  if (s.own_main) FDB_CUDA(cudaStreamDestroy(s.s_main));
  if (stream) {
    s.s_main = static_cast<cudaStream_t>(stream);
    s.own_main = false;
  } else {
    int lo_prio = 0, hi_prio = 0;
    FDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    FDB_CUDA(cudaStreamCreateWithPriority(&s.s_main, cudaStreamNonBlocking, lo_prio));
    s.own_main = true;
  }
  FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
  return FDB_OK;
}

int field_sync(Field* f) {
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaStreamSynchronize(s.s_main));
    FDB_CUDA(cudaStreamSynchronize(s.s_bnd));
  }
  return FDB_OK;
}

// direct transport: tell the producers of this slab's ghosts that the round of exchange
// `xseq` is complete here (everything enqueued on s_main so far), so they may overwrite the
// ghost planes that round was reading
static int field_ack(Field* f) {
  if (f->single() || !f->direct) return FDB_OK;
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    if (f->need_lo) FDB_TRY(stream_write(s.s_main, &s.nbr_flags[NBR_PREV][F_ACK_NEXT], f->xseq));
    if (f->need_hi) FDB_TRY(stream_write(s.s_main, &s.nbr_flags[NBR_NEXT][F_ACK_PREV], f->xseq));
  }
  return FDB_OK;
}

// After a host-driven overwrite of buf[p] (upload, reset): publish it.
static int field_publish(Field* f, int p) {
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
  }
  f->cur = p;
  FDB_TRY(field_exchange(f, p, /*after_bnd=*/false, f->G));
  return field_ack(f);
}

int field_upload(Field* f, int p, const double* host_global, const double* host_slab, bool async) {
  if (!async) FDB_TRY(field_sync(f));
  const int64_t plane = f->geo.plane();
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    // the last halo exchange may still be reading this buffer's boundary planes on s_bnd
    FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_xchg_done, 0));
    const double* src = host_global ? host_global + s.lo * plane
                                    : host_slab + (s.lo - f->slabs[0].lo) * plane;
    FDB_CUDA(cudaMemcpyAsync(f->body(d, p), src, (size_t)(s.nloc() * plane) * sizeof(double),
                             cudaMemcpyHostToDevice, s.s_main));
  }
  return field_publish(f, p);
}

int field_download(Field* f, int p, double* host_global, double* host_slab) {
  const int64_t plane = f->geo.plane();
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    double* dst = host_global ? host_global + s.lo * plane
                              : host_slab + (s.lo - f->slabs[0].lo) * plane;
    FDB_CUDA(cudaMemcpyAsync(dst, f->body(d, p), (size_t)(s.nloc() * plane) * sizeof(double),
                             cudaMemcpyDeviceToHost, s.s_main));
  }
  for (auto& s : f->slabs) {
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_CUDA(cudaStreamSynchronize(s.s_main));
  }
  return FDB_OK;
}

// ref: Upwind ctor, upwind.cxx:45-48 -- zero field, cell 0 = 1
int field_fill_delta(Field* f, int p, int64_t cell) {
  FDB_TRY(field_sync(f));
  const int64_t plane = f->geo.plane();
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_TRY(launch_fill(f->body(d, p), s.nloc() * plane, 0.0, s.s_main));
    if (cell >= s.lo * plane && cell < s.hi * plane)
      FDB_TRY(launch_fill(f->body(d, p) + (cell - s.lo * plane), 1, 1.0, s.s_main));
  }
  return field_publish(f, p);
}

int field_fill_random(Field* f, int p, uint64_t seed, bool publish) {
  FDB_TRY(field_sync(f));
  const int64_t plane = f->geo.plane();
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_TRY(launch_fill_random(f->body(d, p), s.nloc() * plane, seed, s.lo * plane, s.s_main));
  }
  return publish ? field_publish(f, p) : FDB_OK;
}

int field_fill_separable(Field* f, int p, int nd, const double* const* x_host, const int64_t* extents) {
  FDB_TRY(field_sync(f));
  int64_t total = 0;
  for (int j = 0; j < nd; ++j) total += extents[j];
  std::vector<double> packed((size_t)total);
  for (int j = 0, o = 0; j < nd; o += (int)extents[j], ++j) memcpy(packed.data() + o, x_host[j], (size_t)extents[j] * sizeof(double));
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    double* dx = nullptr;
    FDB_CUDA(cudaMalloc(&dx, (size_t)total * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(dx, packed.data(), (size_t)total * sizeof(double), cudaMemcpyHostToDevice, s.s_main);
    const double* x[3] = {nullptr, nullptr, nullptr};
    for (int j = 0, o = 0; j < nd; o += (int)extents[j], ++j) x[j] = dx + o;
    int rc = (e == cudaSuccess) ? launch_separable(f->body(d, p), s.lo, s.nloc(), f->geo.n[1], f->geo.n[2], nd, x,
                                                   f->geo.axis_of, s.s_main)
                                : set_error(FDB_E_CUDA, "H2D copy: %s", cudaGetErrorString(e));
    cudaStreamSynchronize(s.s_main);  // `packed` and dx are released below
    cudaFree(dx);
    FDB_TRY(rc);
  }
  return field_publish(f, p);
}

int field_mirror(Field* f, int src, int dst, const bool* flip, bool publish) {
  if (src == dst) return set_error(FDB_E_STATE, "mirroring needs two buffers");
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    // a host-driven overwrite of buf[dst]: the last halo exchange may still be reading its boundary planes
    FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_xchg_done, 0));
    FDB_TRY(launch_mirror(f->body(d, src), f->body(d, dst), s.nloc(), f->geo.n[1], f->geo.n[2], flip, s.s_main));
  }
  return publish ? field_publish(f, dst) : FDB_OK;
}

// Fill the ghosts of buf[p].  Pull model: the copy runs on the RECEIVING device's
// boundary stream, so every event is recorded on the device that owns it.
int field_exchange(Field* f, int p, bool after_bnd, int depth) {
  if (f->single()) return FDB_OK;  // ghosts alias the far planes of the same buffer
  f->ghost_depth[p] = depth;
  const int64_t plane = f->geo.plane();
  const size_t bytes = (size_t)depth * (size_t)plane * sizeof(double);
  const int64_t gcount = (int64_t)depth * plane;
  const int64_t lo_skip = (int64_t)(f->G - depth) * plane;  // the ghost planes nearest the body
  if (f->direct) {
    // Push model on copy engines: each slab copies its boundary planes straight into the
    // neighbour's ghost planes (no SM is needed, so the transfer overlaps a persistent interior
    // kernel that owns every SM) and bumps the neighbour's sequence counter behind the copy.
    const uint64_t e = ++f->xseq;
    for (int d = 0; d < f->ngpus; ++d) {
      Slab& s = f->slabs[d];
      FDB_CUDA(cudaSetDevice(s.device));
      if (!after_bnd) FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_local_done, 0));
      const int64_t nloc = s.nloc();
      if (f->need_lo) {
        // WAR on the neighbour's ghost planes: it has finished the round before this one
        FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_ACK_NEXT], e - 1));
        FDB_CUDA(cudaMemcpyAsync(s.nbr_buf[NBR_NEXT][p] + lo_skip, f->body(d, p) + (nloc - depth) * plane, bytes,
                                 cudaMemcpyDeviceToDevice, s.s_bnd));
        FDB_TRY(stream_write(s.s_bnd, &s.nbr_flags[NBR_NEXT][F_GHOST_LO], e));
      }
      if (f->need_hi) {
        FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_ACK_PREV], e - 1));
        FDB_CUDA(cudaMemcpyAsync(s.nbr_buf[NBR_PREV][p] + (int64_t)(f->G + nloc) * plane, f->body(d, p), bytes,
                                 cudaMemcpyDeviceToDevice, s.s_bnd));
        FDB_TRY(stream_write(s.s_bnd, &s.nbr_flags[NBR_PREV][F_GHOST_HI], e));
      }
      FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_bnd));
      f->last_halo_bytes += (double)bytes * ((f->need_lo ? 1 : 0) + (f->need_hi ? 1 : 0));
    }
    f->ghost_seq[p] = e;
    return FDB_OK;
  }
  if (f->comm) {
    Slab& s = f->slabs[0];
    fdb_comm* c = f->comm;
    FDB_CUDA(cudaSetDevice(s.device));
    // the planes to send are produced on s_bnd (after_bnd) or were published on s_main
    if (!after_bnd) FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_local_done, 0));
    const int next = f->next_of(c->rank), prev = f->prev_of(c->rank);
    double* top = f->body(0, p) + (s.nloc() - depth) * plane;
    double* bottom = f->body(0, p);
    FDB_NCCL(ncclGroupStart());
    if (f->need_lo) {
      FDB_NCCL(ncclSend(top, (size_t)gcount, ncclDouble, next, c->nccl, s.s_bnd));
      FDB_NCCL(ncclRecv(f->ghost_lo(0, p) + lo_skip, (size_t)gcount, ncclDouble, prev, c->nccl, s.s_bnd));
    }
    if (f->need_hi) {
      FDB_NCCL(ncclSend(bottom, (size_t)gcount, ncclDouble, prev, c->nccl, s.s_bnd));
      FDB_NCCL(ncclRecv(f->ghost_hi(0, p), (size_t)gcount, ncclDouble, next, c->nccl, s.s_bnd));
    }
    FDB_NCCL(ncclGroupEnd());
    count_launch();
    FDB_CUDA(cudaEventRecord(s.ev_ghost_ready[p], s.s_bnd));
    FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_bnd));
    f->last_halo_bytes += (double)bytes * ((f->need_lo ? 1 : 0) + (f->need_hi ? 1 : 0));
    return FDB_OK;
  }
  const int g = f->ngpus;
  for (int d = 0; d < g; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    // WAR: this device's previous readers of ghost buffer p are done once its
    // last sweep is (ev_local_done as recorded at the end of that sweep)
    FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_local_done, 0));
    if (f->need_lo) {
      const int src = f->prev_of(d);
      Slab& o = f->slabs[src];
      FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, after_bnd ? o.ev_bnd_done : o.ev_local_done, 0));
      FDB_CUDA(cudaMemcpyPeerAsync(f->ghost_lo(d, p) + lo_skip, s.device,
                                   f->body(src, p) + (o.nloc() - depth) * plane, o.device, bytes,
                                   s.s_bnd));
    }
    if (f->need_hi) {
      const int src = f->next_of(d);
      Slab& o = f->slabs[src];
      FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, after_bnd ? o.ev_bnd_done : o.ev_local_done, 0));
      FDB_CUDA(cudaMemcpyPeerAsync(f->ghost_hi(d, p), s.device, f->body(src, p), o.device, bytes,
                                   s.s_bnd));
    }
    FDB_CUDA(cudaEventRecord(s.ev_ghost_ready[p], s.s_bnd));
    FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_bnd));
    f->last_halo_bytes += (double)bytes * ((f->need_lo ? 1 : 0) + (f->need_hi ? 1 : 0));
  }
  return FDB_OK;
}

// ---- sweeps ----------------------------------------------------------------------------------
// Boundary/interior split of a slab for ghost depth `depth`.
static void split_slab(const Field* f, int64_t nloc, int depth, int64_t* b_end, int64_t* t_beg) {
  *b_end = f->need_hi ? (depth < nloc ? depth : nloc) : 0;
  *t_beg = f->need_lo ? (nloc - depth > *b_end ? nloc - depth : *b_end) : nloc;
}

// Direct transport, one slab, one sweep: everything this device has to enqueue for the sweep
// that reads buf[X] (ghosts from exchange `gseq`) and performs exchange `e` into the
// neighbours' ghosts of buf[Y].  Touches no shared host state, so the devices of one process
// can be driven by independent host threads.
static int sweep_device_direct(Field* f, int d, SweepLauncher* L, int depth, int X, uint64_t gseq, uint64_t e,
                               double* halo_bytes) {
  const int Y = 1 - X;
  Slab& s = f->slabs[d];
  const int64_t plane = f->geo.plane();
  const int64_t nloc = s.nloc();
  const size_t bytes = (size_t)depth * (size_t)plane * sizeof(double);
  const int64_t lo_skip = (int64_t)(f->G - depth) * plane;
  int64_t b_end, t_beg;
  split_slab(f, nloc, depth, &b_end, &t_beg);
  // 0. Single-launch sweep (opt-in: FDB_HALO=single; one-sided rings, kernels that run the HaloSignal protocol): ONE
  //    persistent kernel over the whole slab on the main stream.  It walks the top chunk first, stores the planes the
  //    next slab needs straight into that slab's ghost planes (peer stores over NVLink), waits on the device for the
  //    neighbour's ACK before the first of those stores and for its own ghost planes before the loader reads them, and
  //    raises the neighbour's ghost flag behind the last store.  Measured on B200 (profiles/r02n_*): bit-exact, and
  //    1.3 % SLOWER than the two-launch form at 8 GPUs (6372 vs 6455 GCUPS; 1.0 % at 2), so it is not the default.
  static const bool single_ok = [] { const char* h = getenv("FDB_HALO"); return h && strcmp(h, "single") == 0; }();
  if (single_ok && f->need_lo && !f->need_hi && f->push_stores && t_beg == nloc - depth && t_beg > 0 &&
      L->can_push_single(f, depth)) {
    FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_xchg_done, 0));  // boundary-stream work of an earlier two-launch sweep
    HaloSignal sig;
    sig.ghost_flag = reinterpret_cast<const unsigned long long*>(&s.flags[F_GHOST_LO]);  // waited for on the device
    sig.ghost_value = gseq;
    sig.done_ctr = reinterpret_cast<unsigned int*>(&s.flags[F_DONE_CTR]);
    sig.nbr_flag = reinterpret_cast<unsigned long long*>(&s.nbr_flags[NBR_NEXT][F_GHOST_LO]);
    sig.flag_value = e;
    sig.ack_flag = reinterpret_cast<const unsigned long long*>(&s.flags[F_ACK_NEXT]);
    sig.ack_value = e - 1;
    FDB_TRY(L->launch_single(f, d, X, depth, s.s_main, s.nbr_buf[NBR_NEXT][Y] + lo_skip, t_beg, sig));
    FDB_CUDA(cudaEventRecord(s.ev_bnd_done, s.s_main));
    FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_main));
    FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
    FDB_TRY(stream_write(s.s_main, &s.nbr_flags[NBR_PREV][F_ACK_NEXT], e));
    *halo_bytes += (double)bytes;
    return FDB_OK;
  }
  // 1. boundary planes on the high-priority stream
  FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_local_done, 0));
  if (f->need_lo) FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_GHOST_LO], gseq));
  if (f->need_hi) FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_GHOST_HI], gseq));
  FDB_TRY(L->launch(f, d, X, depth, 0, b_end, s.s_bnd));
  // Fused compute + halo push: when the kernel can, the top planes are stored straight into the
  // next slab's ghost planes over NVLink by the boundary kernel itself (peer stores), tile by
  // tile, instead of a copy afterwards.  The WAR wait on the neighbour's ACK then precedes it.
  const bool push = f->need_lo && f->push_stores && t_beg == nloc - depth && L->can_push(f, depth);
  if (push) {
    FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_ACK_NEXT], e - 1));
    FDB_TRY(L->launch_push(f, d, X, depth, t_beg, nloc, s.s_bnd, s.nbr_buf[NBR_NEXT][Y] + lo_skip, t_beg));
  } else {
    FDB_TRY(L->launch(f, d, X, depth, t_beg, nloc, s.s_bnd));
  }
  FDB_CUDA(cudaEventRecord(s.ev_bnd_done, s.s_bnd));
  // 2. hand the boundary planes to the neighbours and bump their counters
  if (f->need_lo) {
    if (!push) {
      FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_ACK_NEXT], e - 1));
      FDB_CUDA(cudaMemcpyAsync(s.nbr_buf[NBR_NEXT][Y] + lo_skip, f->body(d, Y) + (nloc - depth) * plane, bytes,
                               cudaMemcpyDeviceToDevice, s.s_bnd));
    }
    FDB_TRY(stream_write(s.s_bnd, &s.nbr_flags[NBR_NEXT][F_GHOST_LO], e));
    *halo_bytes += (double)bytes;
  }
  if (f->need_hi) {
    FDB_TRY(stream_wait_geq(s.s_bnd, &s.flags[F_ACK_PREV], e - 1));
    FDB_CUDA(cudaMemcpyAsync(s.nbr_buf[NBR_PREV][Y] + (int64_t)(f->G + nloc) * plane, f->body(d, Y), bytes,
                             cudaMemcpyDeviceToDevice, s.s_bnd));
    FDB_TRY(stream_write(s.s_bnd, &s.nbr_flags[NBR_PREV][F_GHOST_HI], e));
    *halo_bytes += (double)bytes;
  }
  FDB_CUDA(cudaEventRecord(s.ev_xchg_done, s.s_bnd));
  // 3. interior on the main stream, overlapped with the transfer
  if (f->need_lo) FDB_TRY(stream_wait_geq(s.s_main, &s.flags[F_GHOST_LO], gseq));
  if (f->need_hi) FDB_TRY(stream_wait_geq(s.s_main, &s.flags[F_GHOST_HI], gseq));
  FDB_TRY(L->launch(f, d, X, depth, b_end, t_beg, s.s_main));
  FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_bnd_done, 0));
  FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
  // 4. this round is complete here: the producers of this slab's ghosts may overwrite them
  if (f->need_lo) FDB_TRY(stream_write(s.s_main, &s.nbr_flags[NBR_PREV][F_ACK_NEXT], e));
  if (f->need_hi) FDB_TRY(stream_write(s.s_main, &s.nbr_flags[NBR_NEXT][F_ACK_PREV], e));
  return FDB_OK;
}

// Legacy transports (NCCL send/recv between processes, event-ordered peer copies in one
// process): one sweep over every slab, driven by the calling thread.
static int field_sweep_legacy(Field* f, SweepLauncher* L, int depth) {
  const int X = f->cur, Y = 1 - X;
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    int64_t b_end, t_beg;
    split_slab(f, s.nloc(), depth, &b_end, &t_beg);
    FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_local_done, 0));
    FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, s.ev_ghost_ready[X], 0));
    if (!f->comm) {
      // WAR on the planes about to be rewritten: the neighbours' last pull of them (two
      // sweeps ago, same buffer) must have finished.  NCCL sends are ordered by s_bnd itself.
      const int g = f->ngpus;
      if (f->need_lo) FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, f->slabs[f->next_of(d)].ev_ghost_ready[Y], 0));
      if (f->need_hi) FDB_CUDA(cudaStreamWaitEvent(s.s_bnd, f->slabs[f->prev_of(d)].ev_ghost_ready[Y], 0));
    }
    FDB_TRY(L->launch(f, d, X, depth, 0, b_end, s.s_bnd));
    FDB_TRY(L->launch(f, d, X, depth, t_beg, s.nloc(), s.s_bnd));
    FDB_CUDA(cudaEventRecord(s.ev_bnd_done, s.s_bnd));
  }
  FDB_TRY(field_exchange(f, Y, /*after_bnd=*/true, depth));
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    int64_t b_end, t_beg;
    split_slab(f, s.nloc(), depth, &b_end, &t_beg);
    FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_ghost_ready[X], 0));
    FDB_TRY(L->launch(f, d, X, depth, b_end, t_beg, s.s_main));
    FDB_CUDA(cudaStreamWaitEvent(s.s_main, s.ev_bnd_done, 0));
    FDB_CUDA(cudaEventRecord(s.ev_local_done, s.s_main));
  }
  return FDB_OK;
}

int field_run_sweeps(Field* f, SweepLauncher* L, const int* depths, int n) {
  if (n <= 0) return FDB_OK;
  if (!f->single()) {
    // A sweep of depth T reads T ghost planes and refreshes T ghost planes of its output, so a plan
    // must not deepen on the way; and if the exchange that filled the current buffer's ghosts was
    // shallower than the first sweep (the previous plan ended on a remainder sweep), they are
    // refreshed to full depth first.
    for (int i = 1; i < n; ++i)
      if (depths[i] > depths[i - 1])
        return set_error(FDB_E_STATE, "sweep plan deepens from %d to %d planes", depths[i - 1], depths[i]);
    // all one-time kernel set-up before anything waits on a neighbour (see SweepLauncher::prepare)
    for (int d = 0; d < f->ngpus; ++d) {
      FDB_CUDA(cudaSetDevice(f->slabs[d].device));
      int seen = 0;
      for (int i = 0; i < n; ++i) {
        if (depths[i] < 31 && ((seen >> depths[i]) & 1)) continue;
        if (depths[i] < 31) seen |= 1 << depths[i];
        FDB_TRY(L->prepare(f, d, depths[i]));
      }
    }
    if (depths[0] > f->ghost_depth[f->cur]) FDB_TRY(field_publish(f, f->cur));
  }
  if (f->single()) {
    Slab& s = f->slabs[0];
    FDB_CUDA(cudaSetDevice(s.device));
    // Launch-bound grids: a plan of several sweeps is captured once into a CUDA graph and
    // replayed -- one launch call per advect()/iterate() instead of one per sweep.  Grids whose sweeps run for
    // hundreds of microseconds gain nothing, so only small fields take this path (FDB_GRAPH=0 never, =1 always).
    static const int graph_mode = [] { const char* g = getenv("FDB_GRAPH"); return (g && *g) ? atoi(g) : -1; }();
    const bool want_graph = n >= 2 && L->key() != 0 &&
                            (graph_mode > 0 || (graph_mode < 0 && f->geo.total() <= (int64_t)(1 << 24)));
    if (want_graph) {
      uint64_t key = L->key() * 0x9E3779B97F4A7C15ull + (uint64_t)f->cur;
      // the launchers read their tuning knobs from the environment at every launch: part of what is baked in
      for (const char* knob : {"FDB_FUSED_CFG", "FDB_TMA_CFG", "FDB_TMA_CI", "FDB_MAX_CTAS", "FDB_LAPF_CFG", "FDB_LAP_CFG",
                               "FDB_LAPF_GENERAL", "FDB_FUSED_IMPL"}) {
        const char* v = getenv(knob);
        for (const char* c = (v ? v : ""); *c; ++c) key = (key ^ (uint64_t)(unsigned char)*c) * 0x100000001B3ull;
        key = (key ^ 0xFFull) * 0x100000001B3ull;
      }
      for (int i = 0; i < n; ++i) key = (key ^ (uint64_t)depths[i]) * 0x100000001B3ull;
      key ^= (uint64_t)n << 56;
      Field::PlanGraph* hit = nullptr;
      for (auto& g : f->plan_graphs)
        if (g.key == key) hit = &g;
      if (!hit) {
        if (!f->s_capture) FDB_CUDA(cudaStreamCreateWithFlags(&f->s_capture, cudaStreamNonBlocking));
        const int64_t before = launch_count();
        FDB_CUDA(cudaStreamBeginCapture(f->s_capture, cudaStreamCaptureModeRelaxed));  // first-use setup calls are not stream work
        int rc = FDB_OK, X = f->cur;
        for (int i = 0; i < n && rc == FDB_OK; ++i, X = 1 - X) rc = L->launch(f, 0, X, depths[i], 0, s.nloc(), f->s_capture);
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(f->s_capture, &graph);
        const int launches = (int)(launch_count() - before);
        count_launch(-launches);  // captured, not launched
        if (rc != FDB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return set_error(FDB_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
        cudaGraphExec_t exec = nullptr;
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return set_error(FDB_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
        if (f->plan_graphs.size() >= 16) {  // plans vary with (steps, dt): keep the cache small
          FDB_CUDA(cudaStreamSynchronize(s.s_main));  // the evicted graph may still be running (rare path)
          cudaGraphExecDestroy(f->plan_graphs.front().exec);
          f->plan_graphs.erase(f->plan_graphs.begin());
        }
        f->plan_graphs.push_back({key, exec, launches});
        hit = &f->plan_graphs.back();
      }
      FDB_CUDA(cudaGraphLaunch(hit->exec, s.s_main));
      count_launch(hit->launches);
      f->cur = (f->cur + n) & 1;
      return FDB_OK;
    }
    for (int i = 0; i < n; ++i) {
      FDB_TRY(L->launch(f, 0, f->cur, depths[i], 0, s.nloc(), s.s_main));
      f->cur = 1 - f->cur;
    }
    return FDB_OK;
  }
  if (!f->direct) {
    for (int i = 0; i < n; ++i) {
      FDB_TRY(field_sweep_legacy(f, L, depths[i]));
      f->cur = 1 - f->cur;
    }
    return FDB_OK;
  }
  // direct transport: the per-device enqueue sequences are independent of one another
  const int g = f->ngpus;
  const int X0 = f->cur;
  const uint64_t e0 = f->xseq;
  const uint64_t gs0 = f->ghost_seq[X0];
  std::vector<int> rc((size_t)g, FDB_OK);
  std::vector<double> halo((size_t)g, 0.0);
  std::vector<std::string> err((size_t)g);
  auto drive = [&](int d) {
    if (cudaSetDevice(f->slabs[d].device) != cudaSuccess) {
      rc[(size_t)d] = set_error(FDB_E_CUDA, "cudaSetDevice(%d) failed", f->slabs[d].device);
      err[(size_t)d] = last_error();
      return;
    }
    for (int i = 0; i < n; ++i) {
      const int X = (X0 + i) & 1;
      // sweep i reads the ghosts of exchange e0+i (the first sweep: whatever filled buf[X0]) and
      // performs exchange e0+i+1
      const uint64_t gseq = (i == 0) ? gs0 : e0 + (uint64_t)i;
      const int r = sweep_device_direct(f, d, L, depths[i], X, gseq, e0 + (uint64_t)i + 1, &halo[(size_t)d]);
      if (r != FDB_OK) {
        rc[(size_t)d] = r;
        err[(size_t)d] = last_error();  // thread-local text, carried back to the caller's thread
        return;
      }
    }
  };
  if (g == 1) {
    drive(0);
  } else {
    std::vector<std::thread> threads;
    for (int d = 0; d < g; ++d) threads.emplace_back(drive, d);
    for (auto& t : threads) t.join();
  }
  for (int d = 0; d < g; ++d)
    if (rc[(size_t)d] != FDB_OK) return set_error(rc[(size_t)d], "%s", err[(size_t)d].c_str());
  for (int d = 0; d < g; ++d) f->last_halo_bytes += halo[(size_t)d];
  f->xseq = e0 + (uint64_t)n;
  f->cur = (X0 + n) & 1;
  f->ghost_seq[f->cur] = f->xseq;                 // the last exchange filled the new current buffer
  f->ghost_seq[1 - f->cur] = f->xseq - (n >= 1 ? 1 : 0);
  for (int i = 0; i < n; ++i) f->ghost_depth[(X0 + i + 1) & 1] = depths[i];
  return FDB_OK;
}

// ---- reductions -------------------------------------------------------------------------
static int field_reduce(Field* f, int p, int mode, double mean, double* out, double* planes_out = nullptr) {
  const int64_t plane = f->geo.plane();
  const int64_t n0 = f->geo.n[0];
  std::vector<double> sums((size_t)n0, 0.0);
  for (int d = 0; d < f->ngpus; ++d) {
    Slab& s = f->slabs[d];
    FDB_CUDA(cudaSetDevice(s.device));
    FDB_TRY(launch_plane_sums(f->body(d, p), s.nloc(), plane, mode, mean, s.partial, s.plane_sums,
                              s.s_main));
  }
  if (f->comm && f->comm->nranks > 1) {
    fdb_comm* c = f->comm;
    Slab& s = f->slabs[0];
    if (c->scratch_doubles < (size_t)n0) {
      if (c->scratch) FDB_CUDA(cudaFree(c->scratch));
      FDB_CUDA(cudaMalloc(&c->scratch, (size_t)n0 * sizeof(double)));
      c->scratch_doubles = (size_t)n0;
    }
    FDB_NCCL(ncclAllGather(s.plane_sums, c->scratch, (size_t)s.nloc(), ncclDouble, c->nccl, s.s_main));
    count_launch();
    FDB_CUDA(cudaMemcpyAsync(sums.data(), c->scratch, (size_t)n0 * sizeof(double),
                             cudaMemcpyDeviceToHost, s.s_main));
    FDB_CUDA(cudaStreamSynchronize(s.s_main));
  } else {
    for (int d = 0; d < f->ngpus; ++d) {
      Slab& s = f->slabs[d];
      FDB_CUDA(cudaSetDevice(s.device));
      FDB_CUDA(cudaMemcpyAsync(sums.data() + s.lo, s.plane_sums, (size_t)s.nloc() * sizeof(double),
                               cudaMemcpyDeviceToHost, s.s_main));
    }
    for (auto& s : f->slabs) {
      FDB_CUDA(cudaSetDevice(s.device));
      FDB_CUDA(cudaStreamSynchronize(s.s_main));
    }
  }
  if (f->planes_mirrored) {  // device order inside a slab is the reverse of the global plane order
    const int64_t nl = n0 / f->nparts;
    for (int64_t b = 0; b < n0; b += nl) std::reverse(sums.begin() + b, sums.begin() + b + nl);
  }
  double acc = 0.0;
  for (int64_t i = 0; i < n0; ++i) acc += sums[(size_t)i];  // global plane order
  if (out) *out = acc;
  if (planes_out) memcpy(planes_out, sums.data(), (size_t)n0 * sizeof(double));
  return FDB_OK;
}

int field_plane_sums(Field* f, int p, double* planes_out) { return field_reduce(f, p, 0, 0.0, nullptr, planes_out); }

int field_sum(Field* f, int p, double* out) { return field_reduce(f, p, 0, 0.0, out); }
int field_sqdev(Field* f, int p, double mean, double* out) { return field_reduce(f, p, 1, mean, out); }

}  // namespace fdb
