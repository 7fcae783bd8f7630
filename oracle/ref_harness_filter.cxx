/*
 * ref_harness_filter.cxx -- C entry points around the UNTOUCHED reference
 * class Filter (cxx/Filter.{h,cpp} + CubeDecomp, MultiArrayIter, writeVTK),
 * compiled where they lie against the single-rank oracle/fakempi/mpi.h.
 *
 * TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/libref_filter.so.
 * Filter stores its local block column-major (Filter.cpp:50); everything that
 * crosses this harness is row-major (last axis fastest), converted by index.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include <mpi.h>

#define private public
#include "Filter.h"
#undef private

namespace {

const double *g_in = nullptr;
std::vector<size_t> g_dims;

size_t rowMajorIndex(const std::vector<size_t> &inds) {
  size_t idx = 0;
  for (size_t j = 0; j < g_dims.size(); ++j) idx = idx * g_dims[j] + inds[j];
  return idx;
}

double pickInput(const std::vector<size_t> &inds) { return g_in[rowMajorIndex(inds)]; }

/* laplacian/cxx/laplacian.cxx:22-28, restated because that file's func() lives
 * next to its main(); checked against the driver through fdb_ref_laplacian_cli */
double sinProduct(const std::vector<double> &pos) {
  double res = 1;
  for (size_t i = 0; i < pos.size(); ++i) res *= sin(2.0 * M_PI * pos[i]);
  return res;
}

void toRowMajor(Filter &f, const std::vector<double> &colMajor, double *out) {
  f.mit.begin();
  for (size_t i = 0; i < f.mit.getNumberOfTerms(); ++i) {
    out[rowMajorIndex(f.mit.getIndices())] = colMajor[i];
    f.mit.next();
  }
}

}  // namespace

extern "C" {

/* Builds Filter(dims, 0, 1, stencil), loads `in` (row-major) through
 * setInDataByIndices (or, if in == NULL, the laplacian driver's sin-product
 * through setInData), then runs niter x { applyFilter(); copyOutToIn(); }
 * exactly as laplacian.cxx:86-90 / upwindMpi.cxx:115-126 do.
 * outField  <- outData after the last apply (row-major), may be NULL
 * inField   <- inData as initialised, before any apply (row-major), may be NULL
 * sums[0..1] <- computeCheckSum("input"), ("output") after the loop
 * Returns wall seconds of the apply loop, or -1 if the decomposition failed. */
double fdb_ref_filter_run(int ndims, const long long *dims, int nbranch,
                          const int *offsets, const double *weights,
                          const double *in, int niter, double *outField,
                          double *inField, double *sums) {
  std::vector<size_t> gd(dims, dims + ndims);
  std::vector<double> xmins(ndims, 0.0), xmaxs(ndims, 1.0);
  std::map<std::vector<int>, double> stencil;
  for (int b = 0; b < nbranch; ++b) {
    std::vector<int> off(offsets + b * ndims, offsets + (b + 1) * ndims);
    stencil.insert(std::pair<std::vector<int>, double>(off, weights[b]));
  }
  std::streambuf *saved = std::cout.rdbuf(nullptr); /* silence the ctor banner */
  Filter fltr(gd, xmins, xmaxs, stencil);
  std::cout.rdbuf(saved);
  if (!fltr.isDecompValid()) return -1.0;
  g_dims = gd;
  if (in) {
    g_in = in;
    fltr.setInDataByIndices(pickInput);
  } else {
    fltr.setInData(sinProduct);
  }
  if (inField) toRowMajor(fltr, fltr.inData, inField);
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < niter; ++it) {
    fltr.applyFilter();
    fltr.copyOutToIn();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (outField) toRowMajor(fltr, fltr.outData, outField);
  if (sums) {
    sums[0] = fltr.computeCheckSum("input");
    sums[1] = fltr.computeCheckSum("output");
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

/* The stencil's branches in the order Filter iterates them (std::map order,
 * Filter.cpp:202).  offsets/weights are rewritten in that order. */
void fdb_ref_filter_branch_order(int ndims, int nbranch, int *offsets, double *weights) {
  std::map<std::vector<int>, double> stencil;
  for (int b = 0; b < nbranch; ++b) {
    std::vector<int> off(offsets + b * ndims, offsets + (b + 1) * ndims);
    stencil.insert(std::pair<std::vector<int>, double>(off, weights[b]));
  }
  int b = 0;
  for (std::map<std::vector<int>, double>::const_iterator it = stencil.begin(); it != stencil.end(); ++it, ++b) {
    std::copy(it->first.begin(), it->first.end(), offsets + b * ndims);
    weights[b] = it->second;
  }
}

/* CubeDecomp::build + getDecomp (CubeDecomp.cpp:11-86) for nprocs ranks:
 * decomp[ndims] <- chosen process grid; returns 1 if valid else 0. */
int fdb_ref_cubedecomp(int nprocs, int ndims, const long long *dims, long long *decomp) {
  std::vector<size_t> gd(dims, dims + ndims);
  CubeDecomp d;
  if (!d.build(nprocs, gd)) return 0;
  std::vector<size_t> res = d.getDecomp();
  for (int j = 0; j < ndims; ++j) decomp[j] = (long long)res[j];
  return 1;
}

/* the rest of CubeDecomp's surface for one rank: getBegIndices / getEndIndices (CubeDecomp.cpp:88-109) and
 * getNeighborRank for every direction in dirs (ndir x ndims, CubeDecomp.cpp:111-131). Returns 1 if valid. */
int fdb_ref_cubedecomp_rank(int nprocs, int ndims, const long long *dims, int rank, long long *lo, long long *hi,
                            int ndir, const int *dirs, int *nbr) {
  std::vector<size_t> gd(dims, dims + ndims);
  CubeDecomp d;
  if (!d.build(nprocs, gd)) return 0;
  std::vector<size_t> b = d.getBegIndices(rank), e = d.getEndIndices(rank);
  for (int j = 0; j < ndims; ++j) { lo[j] = (long long)b[j]; hi[j] = (long long)e[j]; }
  for (int k = 0; k < ndir; ++k) {
    std::vector<int> dir(dirs + k * ndims, dirs + (k + 1) * ndims);
    nbr[k] = d.getNeighborRank(rank, dir);
  }
  return 1;
}

int fdb_ref_laplacian_main(int, char **);
int fdb_ref_upwindmpi_main(int, char **);
int fdb_ref_laplacian_cli(int argc, char **argv) { return fdb_ref_laplacian_main(argc, argv); }
int fdb_ref_upwindmpi_cli(int argc, char **argv) { return fdb_ref_upwindmpi_main(argc, argv); }

}  // extern "C"
