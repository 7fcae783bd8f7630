// fdb_internal.h -- shared declarations of the fidib200 runtime (not part of the ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fidib200.h"

namespace fdb {

constexpr int kMaxFuse = 4;  // most time steps one sweep of the fused upwind kernel advances

// ---- errors -----------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
const char* last_error();

#define FDB_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::fdb::set_error(_e == cudaErrorMemoryAllocation ? FDB_E_OOM : FDB_E_CUDA,    \
                              "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
  } while (0)

#define FDB_NCCL(expr)                                                                    \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess)                                                                \
      return ::fdb::set_error(FDB_E_NCCL, "%s failed: %s (%s:%d)", #expr,                 \
                              ncclGetErrorString(_r), __FILE__, __LINE__);                \
  } while (0)

#define FDB_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != FDB_OK) return _rc; \
  } while (0)

void count_launch(int64_t n = 1);
int64_t launch_count();

// ---- geometry of one engine -------------------------------------------------
// Every problem is carried as a 3-D row-major box (n0, n1, n2), n2 fastest.
// ndims == 2 maps (d0, d1) -> (d0, 1, d1) (or (1, d0, d1) on a single device, see plane2d)
// and ndims == 1 maps (d0) -> (1, 1, d0),
// so the reference's last axis is always the contiguous one and its first axis
// (when it exists beside another) is the slab axis.  `axis_of[j]` is the
// internal axis of reference axis j.
struct Geometry {
  int ndims = 3;
  int64_t n[3] = {1, 1, 1};
  int axis_of[3] = {0, 1, 2};
  bool active[3] = {true, true, true};
  int64_t plane() const { return n[1] * n[2]; }
  int64_t total() const { return n[0] * n[1] * n[2]; }
};
// plane2d: carry a 2-D problem as ONE plane (1, d0, d1) instead of d0 planes of one row
int make_geometry(int ndims, const int64_t* dims, Geometry* g, bool plane2d = false);

// ---- communicator (one process per GPU) -------------------------------------
}  // namespace fdb

struct fdb_comm {
  int rank = 0, nranks = 1, device = 0;
  ncclComm_t nccl = nullptr;
  cudaStream_t stream = nullptr;  // small collectives (checksum gather, barrier)
  double* scratch = nullptr;      // device scratch for those collectives
  size_t scratch_doubles = 0;
  int users = 0;                  // engine handles created on this communicator and not yet destroyed
};

namespace fdb {

// ---- one device's share of a field -------------------------------------------
// Buffer layout (doubles): [G ghost planes below][nloc planes][G ghost planes above]
struct Slab {
  int device = 0;
  int64_t lo = 0, hi = 0;  // global plane range on axis 0
  int64_t nloc() const { return hi - lo; }
  double* buf[2] = {nullptr, nullptr};  // ping-pong, each (nloc + 2G) planes
  double* partial = nullptr;            // reduction scratch
  double* plane_sums = nullptr;         // nloc doubles
  cudaStream_t s_main = nullptr, s_bnd = nullptr;
  bool own_main = true;
  cudaEvent_t ev_local_done = nullptr;              // all of the current field written
  cudaEvent_t ev_bnd_done = nullptr;                // boundary planes of the next field written
  cudaEvent_t ev_xchg_done = nullptr;               // last halo exchange issued on s_bnd has drained
  cudaEvent_t ev_ghost_ready[2] = {nullptr, nullptr};  // ghosts of buf[p] filled
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;     // timing
  // direct halo transport: 64-bit sequence counters in THIS slab's memory, written by the
  // neighbours (stream memory operations), and the neighbours' buffers/counters mapped here
  // (plain peer pointers in one process, CUDA IPC mappings between processes)
  uint64_t* flags = nullptr;
  double* nbr_buf[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [0 = prev, 1 = next][parity]
  uint64_t* nbr_flags[2] = {nullptr, nullptr};
  void* ipc_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // TMA descriptors, index = buffer parity
  // (several box shapes over the same tensors, see kernels_tma.cu)
  CUtensorMap tm_body[2];  // local planes, box = tile rows
  CUtensorMap tm_row[2];   // local planes, box = one (halo) row
  CUtensorMap tm_col[2];   // local planes, box = 2 cells x tile rows (k wrap)
  CUtensorMap tm_glo[2];   // the G planes below local plane 0 (ghost, or wrap for 1 device)
  CUtensorMap tm_ghi[2];   // the G planes above local plane nloc-1
  bool have_tma = false;
  int tma_cfg = 0;         // index into the TMA kernel configuration table
  bool lap_ragged = false; // 7-point single-apply kernel: the plane is not a whole number of tiles
  // fused (temporal blocking) kernel: 8 tensor maps per buffer parity, encoded on first use
  alignas(64) unsigned char fused_maps[2][8 * sizeof(CUtensorMap)];
  int fused_T = 0;
  const void* fused_cfg = nullptr;
  // fused two-apply 7-point kernel: 12 tensor maps per buffer parity, encoded on first use
  alignas(64) unsigned char lapf_maps[2][12 * sizeof(CUtensorMap)];
  const void* lapf_cfg = nullptr;
};

// A field decomposed in slabs over the devices this process drives, plus the
// halo plumbing between them (peer copies in-process, NCCL between processes).
struct Field {
  Geometry geo;
  int G = 1;                // ghost planes per side
  bool need_lo = true;      // someone reads plane i-1.. (ghost below)
  bool need_hi = false;     // someone reads plane i+1.. (ghost above)
  int ngpus = 1;            // devices in this process
  fdb_comm* comm = nullptr; // non-null: one slab here, neighbours are other ranks
  int nparts = 1;           // total slabs in the ring
  std::vector<Slab> slabs;
  int cur = 0;              // buf[cur] holds the current field
  // The ring can run BACKWARDS along axis 0: a field held mirrored along that axis inside every slab (negative
  // velocity along the slab axis, capi.cu) keeps each slab on the device that owns its global planes, so in device
  // space the slab "after" part p is part p-1.  next_of / prev_of are the only places that know.
  bool ring_reversed = false;
  int next_of(int part) const { return ring_reversed ? (part + nparts - 1) % nparts : (part + 1) % nparts; }
  int prev_of(int part) const { return ring_reversed ? (part + 1) % nparts : (part + nparts - 1) % nparts; }
  bool planes_mirrored = false;  // device plane order inside a slab is the reverse of the global one (reductions)
  bool direct = true;       // halo transport: peer copies + stream flags (else NCCL / event-ordered copies)
  bool push_stores = true;  // direct transport: boundary kernels store into the neighbour's ghosts themselves
  uint64_t xseq = 0;        // halo exchanges issued so far (same on every rank)
  uint64_t ghost_seq[2] = {0, 0};  // the exchange that filled the ghosts of buf[p]
  int ghost_depth[2] = {0, 0};     // how many ghost planes (nearest the body) that exchange refreshed
  bool ghosts_valid = false;
  double last_ms = 0, last_updates = 0, last_halo_bytes = 0;
  // CUDA graphs of whole sweep plans on single-slab fields (launch-bound grids), keyed by launcher + parity + depths
  struct PlanGraph { uint64_t key; cudaGraphExec_t exec; int launches; };
  std::vector<PlanGraph> plan_graphs;
  cudaStream_t s_capture = nullptr;

  double* body(int d, int p) const { return slabs[d].buf[p] + (int64_t)G * geo.plane(); }
  double* ghost_lo(int d, int p) const;  // G planes below local plane 0 (never null)
  double* ghost_hi(int d, int p) const;  // G planes above local plane nloc-1
  bool single() const { return nparts == 1; }
};

int field_create(Field* f, const Geometry& geo, int G, bool need_lo, bool need_hi, int ngpus,
                 fdb_comm* comm, int want_tma,  // want_tma: 0 none, 1 upwind, 2 7-point stencil
                 bool ring_reversed = false);
void field_destroy(Field* f);
int field_set_stream(Field* f, void* stream);
// async: no host synchronisation at all, the copy is ordered behind the handle's earlier work
int field_upload(Field* f, int p, const double* host_global, const double* host_slab, bool async = false);
int field_download(Field* f, int p, double* host_global, double* host_slab);
int field_fill_delta(Field* f, int p, int64_t cell = 0);  // zero field, global cell `cell` = 1
// device-side synthetic inputs (no host copy): hash of the global cell index / separable product of 1-D factors
// (x[j] = HOST array of the extent of reference axis j); both publish buf[p] as the current field when `publish`
int field_fill_random(Field* f, int p, uint64_t seed, bool publish);
int field_fill_separable(Field* f, int p, int nd, const double* const* x_host, const int64_t* extents);
// buf[dst] = buf[src] mirrored along the flagged axes INSIDE every slab, on the main streams; buf[dst] becomes the
// current field when `publish`
int field_mirror(Field* f, int src, int dst, const bool* flip, bool publish);
// (re)fill the ghosts of buf[p] from the neighbours' boundary planes; enqueued on
// the boundary streams, ev_ghost_ready[p] recorded.  `after_bnd` = wait for
// ev_bnd_done (the planes were just produced by a boundary kernel) instead of
// ev_local_done.
int field_exchange(Field* f, int p, bool after_bnd, int depth);
int field_sync(Field* f);
// deterministic, partition-invariant reductions: per-plane tree sums on the
// device, then a sequential sum over planes in global order (collective in
// dist mode, every rank gets the result)
int field_sum(Field* f, int p, double* out);
int field_sqdev(Field* f, int p, double mean, double* out);
int field_plane_sums(Field* f, int p, double* planes_out);  // geo.n[0] per-plane sums, global plane order

// one sweep of a stencil kernel over every slab: boundary planes first (their halos start
// travelling while the interior is computed), ghosts of the new field exchanged, buffers not
// swapped.  `launch(f, d, X, ibeg, iend, stream)` must enqueue the kernel computing local planes
// [ibeg,iend) of buf[1-X] from buf[X] on slab d; `depth` is set per sweep by the plan.
// Device-side completion protocol of a single-launch ring sweep (kernels_fused.cu): the kernel waits for
// *ack_flag >= ack_value before its first peer store and sets *nbr_flag = flag_value behind its last one.
struct HaloSignal {
  unsigned int* done_ctr = nullptr;
  unsigned long long* nbr_flag = nullptr;
  unsigned long long flag_value = 0;
  const unsigned long long* ack_flag = nullptr;
  unsigned long long ack_value = 0;
  // the kernel's loader waits for *ghost_flag >= ghost_value before it reads this slab's ghost planes
  const unsigned long long* ghost_flag = nullptr;
  unsigned long long ghost_value = 0;
};

struct SweepLauncher {
  virtual int launch(Field* f, int d, int X, int depth, int64_t ibeg, int64_t iend, cudaStream_t s) = 0;
  // Fused halo push: same as launch(), and the kernel also stores local planes >= peer_from through
  // `peer_out` (the next slab's ghost planes, peer-mapped).  Launchers that cannot do it return
  // FDB_E_STATE without launching and the runtime falls back to a copy-engine transfer.
  virtual int launch_push(Field*, int, int, int, int64_t, int64_t, cudaStream_t, double* /*peer_out*/,
                          int64_t /*peer_from*/) {
    return FDB_E_STATE;
  }
  virtual bool can_push(const Field*, int /*depth*/) const { return false; }
  // One-time set-up of whatever kernel a sweep of this depth launches on slab d (function attributes, occupancy,
  // tensor maps, module load).  Such calls may synchronise the device or take driver-wide locks, so multi-device
  // fields make them all up front: once a stream waits on another device's counter, a host thread that blocks in one
  // of them while holding a driver lock can keep the thread that would raise that counter from enqueuing its work.
  virtual int prepare(Field*, int /*d*/, int /*depth*/) { return FDB_OK; }
  // Single-launch ring sweep: ONE kernel over the whole slab that walks the top chunk first, stores the planes the next
  // slab reads through `peer_out` and runs the HaloSignal protocol itself.  FDB_E_STATE when the launcher cannot.
  virtual bool can_push_single(const Field*, int /*depth*/) const { return false; }
  virtual int launch_single(Field*, int /*d*/, int /*X*/, int /*depth*/, cudaStream_t, double* /*peer_out*/,
                            int64_t /*peer_from*/, const HaloSignal&) {
    return FDB_E_STATE;
  }
  // Everything the launches bake in besides (field, parity, depths): two launchers with the same non-zero key
  // enqueue identical kernels, so a captured CUDA graph of a sweep plan can be replayed.  0 = do not cache.
  virtual uint64_t key() const { return 0; }
  virtual ~SweepLauncher() {}
};
// runs the sweeps depths[0..n) back to back (each advances the field once; cur flips after
// each).  Devices of one process are driven by one host thread each when the transport allows.
int field_run_sweeps(Field* f, SweepLauncher* L, const int* depths, int n);

// ---- kernels ----------------------------------------------------------------
struct UpwindCoeffs {
  double c[3];   // ((dt*v)*up)/delta per internal axis (0 where inactive)
  int up[3];     // -1 / +1 upwind direction per internal axis
  bool active[3];
};

int launch_upwind_generic(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                          cudaStream_t s);
bool upwind_tma_supported(const Field& f, const UpwindCoeffs& k);
bool upwind_fused_supported(const Field& f, const UpwindCoeffs& k, int T);
int launch_upwind_fused(Field& f, int d, int X, int T, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                        cudaStream_t s, double* peer_out = nullptr, int64_t peer_from = 0, const HaloSignal* sig = nullptr);
bool upwind_fused_can_signal(int T);
const char* upwind_fused_kernel_name(int T);
int upwind_fused_prepare(Field& f, int d, int T);
int upwind_tma_prepare(const Field& f, int d);
int stencil_lap7_prepare(const Field& f, int d);
int generic_kernels_prepare();  // the selected consumer formulation runs the HaloSignal protocol
const char* upwind_fused_name(int T);
// peer_out/peer_from: planes >= peer_from are also stored into the next slab's ghost planes
int launch_upwind_tma(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const UpwindCoeffs& k,
                      cudaStream_t s, double* peer_out = nullptr, int64_t peer_from = 0);

struct StencilBranches {
  int nbranch = 0;
  int off[32][3];   // internal-axis offsets, already in application order
  double w[32];
  bool ref_wrap = false;  // the reference's non-periodic wrap for extents that do not divide 2^64 (generic kernel)
};
int launch_stencil_generic(const Field& f, int d, int X, int64_t ibeg, int64_t iend,
                           const StencilBranches& b, cudaStream_t s);
bool stencil_lap7_supported(const Field& f, const StencilBranches& b);
int launch_stencil_lap7(const Field& f, int d, int X, int64_t ibeg, int64_t iend, const StencilBranches& b,
                        cudaStream_t s);
// two applies per sweep (kernels_lapfused.cu): buf[1-X] = stencil(stencil(buf[X])) on planes [ibeg,iend)
bool stencil_lap7_fused_supported(const Field& f, const StencilBranches& b);
int launch_stencil_lap7_fused(Field& f, int d, int X, int64_t ibeg, int64_t iend, const StencilBranches& b,
                              cudaStream_t s);
const char* stencil_lap7_fused_name(const Field& f);
const char* stencil_lap7_fused_kernel_name();
int stencil_lap7_fused_prepare(Field& f, int d, const StencilBranches& b);

int launch_plane_sums(const double* body, int64_t nloc, int64_t plane, int mode, double mean,
                      double* partial, double* plane_sums, cudaStream_t s);
int64_t reduce_partials_per_plane(int64_t plane);
int launch_fill(double* p, int64_t n, double v, cudaStream_t s);
int launch_fill_random(double* p, int64_t n, uint64_t seed, int64_t g0, cudaStream_t s);
int launch_separable(double* out, int64_t i_lo, int64_t nplanes, int64_t n1, int64_t n2, int nd,
                     const double* const* x, const int* ax, cudaStream_t s);
int launch_permute(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2, cudaStream_t s);
int launch_mirror(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2, const bool* flip, cudaStream_t s);

int encode_tensor_map_3d(CUtensorMap* tm, const double* base, int64_t n2, int64_t n1, int64_t n0,
                         int box2, int box1);

}  // namespace fdb
