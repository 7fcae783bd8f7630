/*
 * fdb_oracle.h -- CPU restatement of FiDiBench's finite-difference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fidibench_b200/,
 * drivers/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against
 * the untouched reference sources compiled into oracle/_ref/ (see
 * oracle/Makefile, oracle/ref_harness_*.cxx) and against the golden fixtures
 * in tests/golden/ that were generated from those reference builds
 * (tests/golden/make_golden.py).
 *
 * Citations are relative to the reference tree (pletzer/fidibench).
 */
#ifndef FDB_ORACLE_H
#define FDB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* upwind/cxx/upwind.cxx:186-192 -- dt = min_j courant*dx_j/v_j, courant = 0.1 */
double fdb_oracle_upwind_dt(int ndims, const int64_t *numCells,
                            const double *velocity, const double *lengths);

/* upwind/cxx/upwind.cxx:51-86 -- Upwind<NDIMS>::advect restated for ndims in
 * 1..3, row-major field (last axis fastest), periodic, in place on f.
 * scratch must hold prod(numCells) doubles (the reference's oldF). */
void fdb_oracle_upwind_advect(int ndims, const int64_t *numCells,
                              const double *velocity, const double *lengths,
                              double *f, double *scratch,
                              int64_t numSteps, double dt);

/* upwind/cxx/upwind.cxx:91-93 -- sequential std::accumulate */
double fdb_oracle_checksum(const double *f, int64_t n);

/* upwind/cxx/upwind.cxx:95-103 -- population standard deviation */
double fdb_oracle_std(const double *f, int64_t n);

/* cxx/Filter.cpp:191-263 -- Filter::applyFilter restated for one domain:
 * out[c] = 0; for each branch IN THE ORDER GIVEN: out[c] += w * in[wrap(c+off)].
 * Fields are row-major (last axis fastest).  The caller passes branches in the
 * reference's std::map order (lexicographic on the offset vector); see
 * fdb_oracle_sort_branches.  If ref_wrap_quirk != 0 the negative-index wrap is
 * computed exactly as Filter.cpp:237-240 does it ((size_t)(int) % size_t, which
 * is only a true periodic wrap for power-of-two extents). */
void fdb_oracle_stencil_apply(int ndims, const int64_t *dims, int nbranch,
                              const int *offsets, const double *weights,
                              const double *in, double *out,
                              int ref_wrap_quirk);

/* std::map<std::vector<int>,double> iteration order (Filter.cpp:202): sorts
 * the branches lexicographically by offset vector, in place. */
void fdb_oracle_sort_branches(int ndims, int nbranch, int *offsets,
                              double *weights);

/* laplacian/cxx/laplacian.cxx:22-28 + cxx/Filter.cpp:103-112 -- cell-centred
 * prod_j sin(2*pi*x_j) input on [xmin,xmax]^ndims, row-major. */
void fdb_oracle_laplacian_input(int ndims, const int64_t *dims,
                                const double *xmins, const double *xmaxs,
                                double *out);

/* number of OpenMP threads the oracle loops will use (1 when built without) */
int fdb_oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
