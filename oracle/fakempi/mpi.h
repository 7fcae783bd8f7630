/*
 * Single-rank stand-in for <mpi.h>: just enough of MPI-3 for the reference's
 * cxx/Filter.cpp, upwind/cxx/upwindMpi.cxx and laplacian/cxx/laplacian.cxx to
 * compile UNMODIFIED and run as one rank (there is no MPI in this image).
 * TEST INFRASTRUCTURE ONLY -- used by oracle/Makefile to build oracle/_ref/.
 * One rank means: a window's only target is this process, MPI_Get is a memcpy
 * out of the exposed buffer, fences do nothing, reductions/gathers copy.
 */
#ifndef FDB_FAKE_MPI_H
#define FDB_FAKE_MPI_H
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
struct fdb_fake_win { char *base; MPI_Aint bytes; int disp_unit; };
typedef fdb_fake_win *MPI_Win;

enum { MPI_COMM_WORLD = 0, MPI_INFO_NULL = 0, MPI_SUCCESS = 0 };
enum { MPI_DOUBLE = 8 };                       /* value = element size */
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };
enum { MPI_THREAD_FUNNELED = 1 };
enum { MPI_MODE_NOPUT = 1, MPI_MODE_NOPRECEDE = 2, MPI_MODE_NOSUCCEED = 4 };

inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
inline int MPI_Init_thread(int *, char ***, int req, int *prov) { *prov = req; return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int *rk) { *rk = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *sz) { *sz = 1; return MPI_SUCCESS; }
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Alloc_mem(MPI_Aint size, MPI_Info, void *baseptr) {
  *(void **)baseptr = std::malloc((size_t)(size > 0 ? size : 1)); return MPI_SUCCESS;
}
inline int MPI_Free_mem(void *base) { std::free(base); return MPI_SUCCESS; }
inline int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info, MPI_Comm, MPI_Win *win) {
  *win = new fdb_fake_win{(char *)base, size, disp_unit}; return MPI_SUCCESS;
}
inline int MPI_Win_free(MPI_Win *win) { delete *win; *win = 0; return MPI_SUCCESS; }
inline int MPI_Win_fence(int, MPI_Win) { return MPI_SUCCESS; }
inline int MPI_Get(void *origin, int count, MPI_Datatype dt, int /*target rank 0*/, MPI_Aint disp,
                   int, MPI_Datatype, MPI_Win win) {
  std::memcpy(origin, win->base + disp * win->disp_unit, (size_t)count * (size_t)dt); return MPI_SUCCESS;
}
inline int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op, int, MPI_Comm) {
  std::memcpy(r, s, (size_t)count * (size_t)dt); return MPI_SUCCESS;
}
inline int MPI_Gather(const void *s, int count, MPI_Datatype dt, void *r, int, MPI_Datatype, int, MPI_Comm) {
  std::memcpy(r, s, (size_t)count * (size_t)dt); return MPI_SUCCESS;
}
#endif
