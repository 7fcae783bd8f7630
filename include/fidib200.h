/*
 * fidib200.h -- C ABI of the B200-native FiDiBench finite-difference engines.
 *
 * This is the drop-in boundary: the reference (pletzer/fidibench) has no FFI
 * layer, its engines are two C++ classes that each driver constructs directly,
 * so the ABI below is the public surface of those classes restated with plain
 * pointers and sizes.  "ref:" citations are relative to the reference tree.
 *
 *   fdb_upwind_*   <->  template<size_t NDIMS> class Upwind   ref: upwind/cxx/upwind.cxx:19-135
 *   fdb_stencil_*  <->  class Filter                          ref: cxx/Filter.h:36-170, cxx/Filter.cpp
 *   fdb_comm_*, fdb_slab_partition  <->  MPI_COMM_WORLD + CubeDecomp   ref: cxx/CubeDecomp.cpp:11-131
 *
 * Conventions
 *   - every entry point returns 0 (FDB_OK) or a negative FDB_E_* code; the text
 *     of the last failure on the calling thread is fdb_last_error().  No C++
 *     exception crosses this boundary.
 *   - there is no CPU fallback: without a usable CUDA device every create call
 *     fails with FDB_E_CUDA.
 *   - fields are FP64.  Host fields are row-major (last axis fastest, as
 *     upwind.cxx:39-44) unless a layout argument says FDB_COL_MAJOR (first axis
 *     fastest, Filter's local storage, Filter.cpp:50).
 *   - a handle owns all device memory, streams and events it needs; host
 *     buffers are caller-owned and only touched during the call.
 *   - a handle is driven by one caller thread at a time.  Calls are synchronous
 *     at return unless the name ends in _async.
 *   - domains are periodic in every axis and partitioned in slabs along axis 0.
 *     Two ways to use several GPUs:
 *       (1) one process, `ngpus` devices 0..ngpus-1 (peer-to-peer halo stores);
 *       (2) one process per GPU: each rank builds an fdb_comm (NCCL over
 *           NVLink) and creates its engines with the *_dist constructors; it
 *           then owns planes [lo,hi) of axis 0 only.
 */
#ifndef FIDIB200_H
#define FIDIB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDB_VERSION_MAJOR 0
#define FDB_VERSION_MINOR 1

/* ---- status ------------------------------------------------------------ */
enum {
  FDB_OK = 0,
  FDB_E_INVALID = -1, /* bad argument                                            */
  FDB_E_CUDA = -2,    /* CUDA runtime/driver failure, or no device               */
  FDB_E_NCCL = -3,    /* NCCL failure                                            */
  FDB_E_OOM = -4,     /* device or host allocation failed                        */
  FDB_E_DECOMP = -5,  /* no valid slab decomposition (ref: Filter.cpp:27-34)     */
  FDB_E_STATE = -6    /* call not valid in the handle's current state            */
};
enum { FDB_ROW_MAJOR = 0, FDB_COL_MAJOR = 1 };
enum { FDB_INPUT = 0, FDB_OUTPUT = 1 }; /* ref: Filter::computeCheckSum("input"|"output") */

/* kernel selection for fdb_upwind_set_kernel / fdb_stencil_set_kernel */
enum {
  FDB_KERNEL_AUTO = 0,    /* fastest kernel that supports the problem            */
  FDB_KERNEL_GENERIC = 1, /* one-thread-per-cell global-memory kernel, any shape */
  FDB_KERNEL_TMA = 2      /* TMA-staged shared-memory tile pipeline (3-D)        */
};

const char *fdb_last_error(void);
int fdb_version(int *major, int *minor);
int fdb_device_count(int *count);
/* number of CUDA kernels this library has launched in this process so far */
int fdb_launch_count(int64_t *count);

/* ---- partition: replaces CubeDecomp (ref: cxx/CubeDecomp.cpp:88-109) ----- */
/* planes [lo,hi) of axis 0 owned by `part` of `nparts`; FDB_E_DECOMP unless
 * nparts divides n0 (the reference has the same requirement, CubeDecomp.cpp:26-28). */
int fdb_slab_partition(int64_t n0, int nparts, int part, int64_t *lo, int64_t *hi);

/* The reference's own process-grid chooser, restated on the host (ref: CubeDecomp::build /
 * computeOptimalDecomp / getBegIndices / getEndIndices / getNeighborRank, cxx/CubeDecomp.cpp:11-131), for
 * callers that want its block decomposition (e.g. to compare with an MPI run): ndims in 1..3.
 * grid[ndims] <- processes per axis; FDB_E_DECOMP when no divisor tuple multiplies to nprocs
 * (ref: Filter.cpp:27-34).  Ranks are row-major over the grid; neighbours are periodic. */
int fdb_cube_decomp(int nprocs, int ndims, const int64_t *dims, int64_t *grid);
int fdb_cube_block(int nprocs, int ndims, const int64_t *dims, int rank, int64_t *lo, int64_t *hi);
int fdb_cube_neighbor(int nprocs, int ndims, const int64_t *dims, int rank, const int *dir, int *neighbor);

/* ---- communicator for one-process-per-GPU runs --------------------------- */
typedef struct fdb_comm fdb_comm;
#define FDB_COMM_ID_BYTES 128
/* rank 0 calls this and ships the 128 bytes to every rank (any transport) */
int fdb_comm_unique_id(void *id_bytes);
/* collective over all ranks; `device` is the CUDA device this rank drives */
int fdb_comm_create(int rank, int nranks, const void *id_bytes, int device, fdb_comm **out);
int fdb_comm_rank(const fdb_comm *c, int *rank, int *nranks);
int fdb_comm_barrier(fdb_comm *c);
/* max over ranks of a host double (timing), collective */
int fdb_comm_max(fdb_comm *c, double *value);
/* FDB_E_STATE while engine handles created on the communicator still exist: destroy those first (on every
 * rank, in the same order -- their teardown is a collective) */
int fdb_comm_destroy(fdb_comm *c);

/* ---- Upwind engine -------------------------------------------------------- */
typedef struct fdb_upwind fdb_upwind;

/* ref: Upwind<NDIMS>::Upwind(velocity, lengths, numCells), upwind.cxx:23-49.
 * ndims in 1..3.  The field starts as the ctor's: all zero, cell 0 = 1.
 * ngpus >= 1 devices of this process; must divide numCells[0]. */
int fdb_upwind_create(int ndims, const int64_t *numCells, const double *velocity,
                      const double *lengths, int ngpus, fdb_upwind **out);
/* same engine, this process owning one slab of the comm's ranks */
int fdb_upwind_create_dist(int ndims, const int64_t *numCells, const double *velocity,
                           const double *lengths, fdb_comm *comm, fdb_upwind **out);
/* planes [lo,hi) of axis 0 held by this handle (whole domain when not dist) */
int fdb_upwind_local_range(const fdb_upwind *h, int64_t *lo, int64_t *hi);

/* overwrite the field from a host array holding the WHOLE domain (row-major);
 * in dist mode each rank takes its own planes out of it */
int fdb_upwind_set_field(fdb_upwind *h, const double *host_field);
/* overwrite only this handle's planes [lo,hi) from a host array of that size */
int fdb_upwind_set_slab(fdb_upwind *h, const double *host_slab);
/* same, enqueue only: the copy is ordered behind the handle's earlier work and ahead of
 * its later work; host_slab (pinned memory for a truly asynchronous copy) must stay valid
 * and unchanged until the next fdb_upwind_sync/checksum/get on this handle.  Lets a caller
 * overlap the upload of one handle with the advect of another. */
int fdb_upwind_set_slab_async(fdb_upwind *h, const double *host_slab);
/* reset to the ctor's initial condition (delta at cell 0), upwind.cxx:45-48 */
int fdb_upwind_reset(fdb_upwind *h);
/* synthetic input generated ON the device (no host copy; benches and parity checks at sizes where a host
 * field is unwieldy): cell g of the global row-major field <- (splitmix64(seed + (g+1)*0x9E3779B97F4A7C15)
 * >> 11) * 2^-53, uniform in [0,1) and independent of the slab partition. */
int fdb_upwind_fill_random(fdb_upwind *h, uint64_t seed);

/* ref: Upwind::advect(numTimeSteps, deltaTime), upwind.cxx:51-86 */
int fdb_upwind_advect(fdb_upwind *h, int64_t numTimeSteps, double deltaTime);
/* enqueue only; pair with fdb_upwind_sync */
int fdb_upwind_advect_async(fdb_upwind *h, int64_t numTimeSteps, double deltaTime);
int fdb_upwind_sync(fdb_upwind *h);
/* ref: main()'s dt = min_j 0.1*dx_j/v_j, upwind.cxx:186-192, with |v_j| so that the negative velocities the
 * class supports still give a positive, stable step (identical bits for the reference's v > 0) */
int fdb_upwind_default_dt(const fdb_upwind *h, double *dt);

/* ref: Upwind::checksum(), upwind.cxx:91-93 (collective in dist mode; every rank gets the sum) */
int fdb_upwind_checksum(fdb_upwind *h, double *sum);
/* the per-plane sums behind the checksum: sums[i] = sum of plane i of axis 0 (fixed-shape tree on the device, the
 * checksum is their sequential sum); *count <- planes (numCells[0] for 3-D, 1 for a 1-D field); sums == NULL only
 * queries the count.  Collective in dist mode, every rank gets all planes.  A cheap size-independent parity
 * probe for fields too large to copy back. */
int fdb_upwind_plane_sums(fdb_upwind *h, double *sums, int64_t capacity, int64_t *count);
/* ref: Upwind::std(), upwind.cxx:95-103 */
int fdb_upwind_std(fdb_upwind *h, double *stddev);
/* whole-domain copy-out, row-major (feeds the host-side saveVTK/print);
 * in dist mode only planes [lo,hi) of host_field are written */
int fdb_upwind_get_field(fdb_upwind *h, double *host_field);
int fdb_upwind_get_slab(fdb_upwind *h, double *host_slab);

/* FDB_KERNEL_*; FDB_E_INVALID if the kernel cannot run this problem */
int fdb_upwind_set_kernel(fdb_upwind *h, int kernel);
/* which kernel the next advect will use (FDB_KERNEL_GENERIC or FDB_KERNEL_TMA) */
int fdb_upwind_get_kernel(const fdb_upwind *h, int *kernel);
/* one line of text: which kernel the next advect launches (and, for the slow generic kernel, WHY the tiled ones do
 * not apply), how many slabs, which halo transport */
int fdb_upwind_describe(const fdb_upwind *h, char *text, size_t capacity);
/* time steps fused per sweep by the TMA kernel (temporal blocking): 1..4, or 0 = auto
 * (the default: the fastest setting the problem supports).  Results do not depend on it. */
int fdb_upwind_set_fuse(fdb_upwind *h, int steps_per_sweep);
/* run the handle's work on a caller-owned cudaStream_t (NULL = the handle's own) */
int fdb_upwind_set_stream(fdb_upwind *h, void *cuda_stream);
/* device time of the last advect (CUDA events on the handle's stream, max over
 * this process's devices), cell-updates done, and halo bytes sent per device */
int fdb_upwind_last_timing(const fdb_upwind *h, double *gpu_ms, double *cell_updates,
                           double *halo_bytes);
int fdb_upwind_destroy(fdb_upwind *h);

/* ---- generic stencil engine (Filter) -------------------------------------- */
typedef struct fdb_stencil fdb_stencil;

/* ref: Filter::Filter(globalDims, xmins, xmaxs, stencil), Filter.cpp:11-78.
 * offsets is nbranch x ndims (row-major), weights nbranch; the branches are
 * applied in std::map order (lexicographic on the offset vector,
 * Filter.cpp:202) whatever order they are passed in; duplicates are rejected.
 * out = sum_b w_b * in[(idx + offset_b) mod dims], accumulated from 0.0 with a
 * separately rounded multiply and add per branch, exactly as Filter.cpp:247-251.
 * ngpus must divide globalDims[0]. */
int fdb_stencil_create(int ndims, const int64_t *globalDims, int nbranch, const int *offsets,
                       const double *weights, int ngpus, fdb_stencil **out);
int fdb_stencil_create_dist(int ndims, const int64_t *globalDims, int nbranch, const int *offsets,
                            const double *weights, fdb_comm *comm, fdb_stencil **out);
int fdb_stencil_local_range(const fdb_stencil *h, int64_t *lo, int64_t *hi);
/* ref: Filter::setInData / setInDataByIndices (Filter.cpp:131-188): the driver
 * evaluates its callback on the host and hands the array over (whole domain) */
int fdb_stencil_set_input(fdb_stencil *h, const double *host_field, int layout);
int fdb_stencil_set_input_slab(fdb_stencil *h, const double *host_slab);
/* Separable input evaluated on the device: cell (i_0..i_{nd-1}) <- ((1 * factors[0][i_0]) * factors[1][i_1]) * ...
 * in the reference's axis order, i.e. what Filter::setInData accumulates for laplacian.cxx's
 * func = prod_j sin(2 pi x_j) (ref: laplacian.cxx:22-28, Filter.cpp:103-112,131-160) when the driver evaluates
 * the 1-D factors sin(2 pi x_j(i)) on the host: the same bits as the host-evaluated field, with
 * 8 B x (d_0 + ... + d_{nd-1}) uploaded instead of 8 B x d_0 ... d_{nd-1}.  factors[j] holds globalDims[j] doubles. */
int fdb_stencil_set_input_separable(fdb_stencil *h, const double *const *factors);
/* same synthetic field as fdb_upwind_fill_random, over the handle's row-major cell order */
int fdb_stencil_fill_random(fdb_stencil *h, uint64_t seed);
/* ref: Filter::applyFilter(), Filter.cpp:191-263 */
int fdb_stencil_apply(fdb_stencil *h);
/* ref: Filter::copyOutToIn(), Filter.cpp:440-463 -- O(1): the buffers swap
 * roles and "output" keeps reading as the same data until the next apply */
int fdb_stencil_swap(fdb_stencil *h);
/* niter x { apply; swap } as laplacian.cxx:86-90, enqueued back to back.  The 3-D 7-point
 * stencil runs two applies per sweep (temporal blocking, half the DRAM traffic) when the
 * plane extents allow; the result is bit-identical either way. */
int fdb_stencil_iterate(fdb_stencil *h, int64_t niter);
/* applies fused per sweep by fdb_stencil_iterate: 1, 2, or 0 = auto (the default: 2 when the
 * fused kernel supports the problem).  FDB_E_INVALID if 2 is asked for and cannot run. */
int fdb_stencil_set_fuse(fdb_stencil *h, int applies_per_sweep);
/* what the next fdb_stencil_iterate will use (1 or 2) */
int fdb_stencil_get_fuse(const fdb_stencil *h, int *applies_per_sweep);
/* ref: Filter::computeCheckSum("input"|"output"), Filter.cpp:465-485 */
int fdb_stencil_checksum(fdb_stencil *h, int which, double *sum);
/* sum of squares of the input or output data, same deterministic reduction (a norm for parity probes) */
int fdb_stencil_sumsq(fdb_stencil *h, int which, double *sumsq);
int fdb_stencil_get(fdb_stencil *h, int which, double *host_field, int layout);
int fdb_stencil_get_slab(fdb_stencil *h, int which, double *host_slab);
/* Compatibility with the reference's index wrap.  Filter.cpp:237-240 reduces `int(index) + offset` with
 * `%= size_t(extent)`: a negative index is converted to size_t first, so -1 wraps to (2^64 - 1) mod N -- the
 * periodic N - 1 only when N is a power of two (for the reference's default `laplacian -numCells 8000`, 7615).
 * This library wraps periodically by default; on = 1 reproduces the reference's arithmetic bit for bit on a
 * single-slab handle (generic kernel; FDB_E_STATE on several slabs).  Environment FDB_REF_WRAP=1 turns it on
 * at creation. */
int fdb_stencil_set_ref_wrap(fdb_stencil *h, int on);
int fdb_stencil_set_kernel(fdb_stencil *h, int kernel);
int fdb_stencil_get_kernel(const fdb_stencil *h, int *kernel);
/* one line of text: the kernels apply / iterate launch, or why only the generic kernel applies */
int fdb_stencil_describe(const fdb_stencil *h, char *text, size_t capacity);
int fdb_stencil_set_stream(fdb_stencil *h, void *cuda_stream);
int fdb_stencil_last_timing(const fdb_stencil *h, double *gpu_ms, double *cell_updates,
                            double *halo_bytes);
int fdb_stencil_destroy(fdb_stencil *h);

#ifdef __cplusplus
}
#endif
#endif /* FIDIB200_H */
