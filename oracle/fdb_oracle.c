/*
 * fdb_oracle.c -- CPU restatement of FiDiBench's finite-difference hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see fdb_oracle.h).  Plain C, IEEE-754 binary64,
 * round-to-nearest, built with -ffp-contract=off so no FMA is ever formed:
 * the reference's default x86-64 build has none either.
 *
 * Parity status: PINNED against oracle/_ref (the untouched reference sources)
 * by tests/test_oracle.py and against the tests/golden fixtures.
 */
#include "fdb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int fdb_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* upwind/cxx/upwind.cxx:186-192 */
double fdb_oracle_upwind_dt(int ndims, const int64_t *numCells,
                            const double *velocity, const double *lengths) {
  const double courant = 0.1;
  double dt = DBL_MAX;
  for (int j = 0; j < ndims; ++j) {
    double dx = lengths[j] / (double)(size_t)numCells[j];
    double val = courant * dx / velocity[j];
    dt = (val < dt ? val : dt);
  }
  return dt;
}

/* upwind/cxx/upwind.cxx:51-86.  The reference rebuilds the index set with
 * div/mod per cell (:122-129); here the three loops are explicit, which
 * changes nothing numerically: each cell does, for j = 0,1,2 in that order,
 *     f[i] = f[i] - coeff_j * (old[up_j(i)] - old[i])
 * with coeff_j = ((dt * v_j) * up_j) / delta_j evaluated left to right (:72). */
void fdb_oracle_upwind_advect(int ndims, const int64_t *numCells,
                              const double *velocity, const double *lengths,
                              double *f, double *scratch,
                              int64_t numSteps, double dt) {
  int64_t n[3] = {1, 1, 1};
  double coeff[3] = {0, 0, 0};
  int up[3] = {0, 0, 0};
  /* right-align the axes so that the last reference axis is always n[2] */
  const int shift = 3 - ndims;
  for (int j = 0; j < ndims; ++j) {
    n[j + shift] = numCells[j];
    up[j + shift] = (velocity[j] < 0.) ? +1 : -1;            /* :34-35 */
    double delta = lengths[j] / (double)(size_t)numCells[j]; /* :36    */
    coeff[j + shift] = dt * velocity[j] * up[j + shift] / delta; /* :72 */
  }
  const int64_t n0 = n[0], n1 = n[1], n2 = n[2];
  const int64_t ntot = n0 * n1 * n2;
  double *old = scratch;

  for (int64_t step = 0; step < numSteps; ++step) {
    memcpy(old, f, (size_t)ntot * sizeof(double)); /* :59-62 */
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (int64_t i0 = 0; i0 < n0; ++i0) {
      for (int64_t i1 = 0; i1 < n1; ++i1) {
        const int64_t u0 = (i0 + up[0] + n0) % n0; /* :75-76 */
        const int64_t u1 = (i1 + up[1] + n1) % n1;
        const double *c = old + (i0 * n1 + i1) * n2;
        const double *p0 = old + (u0 * n1 + i1) * n2;
        const double *p1 = old + (i0 * n1 + u1) * n2;
        double *o = f + (i0 * n1 + i1) * n2;
        for (int64_t i2 = 0; i2 < n2; ++i2) {
          const int64_t u2 = (i2 + up[2] + n2) % n2;
          double t = c[i2];
          /* axes that do not exist in an ndims<3 run are skipped, exactly
           * as the reference's loop over j < NDIMS does */
          if (shift < 1) t = t - coeff[0] * (p0[i2] - c[i2]); /* :80 */
          if (shift < 2) t = t - coeff[1] * (p1[i2] - c[i2]);
          t = t - coeff[2] * (c[u2] - c[i2]);
          o[i2] = t;
        }
      }
    }
  }
}

/* upwind/cxx/upwind.cxx:91-93 */
double fdb_oracle_checksum(const double *f, int64_t n) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += f[i];
  return s;
}

/* upwind/cxx/upwind.cxx:95-103 */
double fdb_oracle_std(const double *f, int64_t n) {
  double mean = fdb_oracle_checksum(f, n) / (double)(size_t)n;
  double res = 0;
  for (int64_t i = 0; i < n; ++i) {
    double d = f[i] - mean;
    res += d * d;
  }
  return sqrt(res / (double)(size_t)n);
}

/* std::map<std::vector<int>, double> ordering (Filter.cpp:202) */
static int lex_less(int ndims, const int *a, const int *b) {
  for (int j = 0; j < ndims; ++j) {
    if (a[j] < b[j]) return 1;
    if (a[j] > b[j]) return 0;
  }
  return 0;
}

void fdb_oracle_sort_branches(int ndims, int nbranch, int *offsets,
                              double *weights) {
  for (int a = 1; a < nbranch; ++a) { /* insertion sort, nbranch is tiny */
    int key[8];
    double w = weights[a];
    memcpy(key, offsets + a * ndims, sizeof(int) * (size_t)ndims);
    int b = a - 1;
    while (b >= 0 && lex_less(ndims, key, offsets + b * ndims)) {
      memcpy(offsets + (b + 1) * ndims, offsets + b * ndims,
             sizeof(int) * (size_t)ndims);
      weights[b + 1] = weights[b];
      --b;
    }
    memcpy(offsets + (b + 1) * ndims, key, sizeof(int) * (size_t)ndims);
    weights[b + 1] = w;
  }
}

/* Filter.cpp:237-240: indOffset[j] = (int) inds[j] + offset[j];
 *                     indOffset[j] %= this->globalDims[j];   (int %= size_t) */
static inline int64_t wrap_index(int64_t ind, int off, int64_t n, int quirk) {
  if (quirk) {
    int v = (int)ind + off;
    size_t u = (size_t)v; /* the usual arithmetic conversion of int to size_t */
    return (int64_t)(int)(u % (size_t)n);
  }
  int64_t v = (ind + off) % n;
  return v < 0 ? v + n : v;
}

/* cxx/Filter.cpp:191-263, single domain (every neighbour is "inside"). */
void fdb_oracle_stencil_apply(int ndims, const int64_t *dims, int nbranch,
                              const int *offsets, const double *weights,
                              const double *in, double *out,
                              int ref_wrap_quirk) {
  int64_t n[3] = {1, 1, 1};
  const int shift = 3 - ndims;
  for (int j = 0; j < ndims; ++j) n[j + shift] = dims[j];
  const int64_t n0 = n[0], n1 = n[1], n2 = n[2];
  const int64_t ntot = n0 * n1 * n2;

  for (int64_t c = 0; c < ntot; ++c) out[c] = 0; /* :195-199 */

  for (int b = 0; b < nbranch; ++b) { /* :202, one full sweep per branch */
    int off[3] = {0, 0, 0};
    for (int j = 0; j < ndims; ++j) off[j + shift] = offsets[b * ndims + j];
    const double val = weights[b];
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (int64_t i0 = 0; i0 < n0; ++i0) {
      for (int64_t i1 = 0; i1 < n1; ++i1) {
        const int64_t s0 = wrap_index(i0, off[0], n0, ref_wrap_quirk);
        const int64_t s1 = wrap_index(i1, off[1], n1, ref_wrap_quirk);
        const double *src = in + (s0 * n1 + s1) * n2;
        double *dst = out + (i0 * n1 + i1) * n2;
        for (int64_t i2 = 0; i2 < n2; ++i2) {
          const int64_t s2 = wrap_index(i2, off[2], n2, ref_wrap_quirk);
          dst[i2] += val * src[s2]; /* :247-251 */
        }
      }
    }
  }
}

/* laplacian/cxx/laplacian.cxx:22-28 evaluated at Filter::getPosition
 * (cxx/Filter.cpp:103-112): pos = xmin + (ind + 0.5) * ((xmax-xmin)/double(N)) */
void fdb_oracle_laplacian_input(int ndims, const int64_t *dims,
                                const double *xmins, const double *xmaxs,
                                double *out) {
  int64_t ntot = 1;
  for (int j = 0; j < ndims; ++j) ntot *= dims[j];
  int64_t stride[3];
  stride[ndims - 1] = 1;
  for (int j = ndims - 2; j >= 0; --j) stride[j] = stride[j + 1] * dims[j + 1];
  for (int64_t c = 0; c < ntot; ++c) {
    double res = 1;
    for (int j = 0; j < ndims; ++j) {
      int64_t ind = c / stride[j] % dims[j];
      double delta = (xmaxs[j] - xmins[j]) / (double)(size_t)dims[j];
      double pos = xmins[j] + ((size_t)ind + 0.5) * delta;
      res *= sin(2.0 * M_PI * pos);
    }
    out[c] = res;
  }
}
