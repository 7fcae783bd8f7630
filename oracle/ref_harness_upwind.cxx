/*
 * ref_harness_upwind.cxx -- C entry points around the UNTOUCHED reference
 * translation unit upwind/cxx/upwind.cxx, compiled where it lies (the Makefile
 * passes -I$(REF)/cxx -I$(REF)/upwind/cxx; nothing is copied into this repo).
 *
 * TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/libref_upwind.so, used
 * by tests/ (to pin oracle/fdb_oracle.c and to make tests/golden/), by
 * bench.py --impl reference and by bench.py's cpu_baseline leg.
 *
 * The reference keeps its fields private and has its own main(); both are
 * opened up with the preprocessor (std headers first: libstdc++ does not
 * survive "#define private public").
 */
#include <algorithm>
#include <array>
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
#include <sstream>
#include <string>
#include <vector>
#ifdef HAVE_OPENMP
#include <omp.h>
#endif

#define private public
#define main fdb_ref_upwind_main
#include "upwind.cxx" /* resolved through -I$(REF)/upwind/cxx */
#undef main
#undef private

namespace {

template <size_t ND>
double run(const long long *numCells, const double *velocity,
           const double *lengths, const double *init, int numSteps, double dt,
           double *outField, double *checksum, double *stddev) {
  std::vector<size_t> nc(numCells, numCells + ND);
  std::vector<double> v(velocity, velocity + ND);
  std::vector<double> len(lengths, lengths + ND);
  Upwind<ND> up(v, len, nc);
  if (init) std::copy(init, init + up.ntot, up.f.begin());
  auto t0 = std::chrono::steady_clock::now();
  up.advect(numSteps, dt);
  auto t1 = std::chrono::steady_clock::now();
  if (outField) std::copy(up.f.begin(), up.f.end(), outField);
  if (checksum) *checksum = up.checksum();
  if (stddev) *stddev = up.std();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

/* Runs Upwind<ndims>(velocity, lengths, numCells); optional initial field
 * (row-major, NULL = the ctor's delta at cell 0); advect(numSteps, dt).
 * Returns the wall seconds spent inside advect(), or -1 on bad ndims. */
double fdb_ref_upwind_run(int ndims, const long long *numCells,
                          const double *velocity, const double *lengths,
                          const double *init, int numSteps, double dt,
                          double *outField, double *checksum, double *stddev) {
  switch (ndims) {
    case 1: return run<1>(numCells, velocity, lengths, init, numSteps, dt, outField, checksum, stddev);
    case 2: return run<2>(numCells, velocity, lengths, init, numSteps, dt, outField, checksum, stddev);
    case 3: return run<3>(numCells, velocity, lengths, init, numSteps, dt, outField, checksum, stddev);
  }
  return -1.0;
}

/* the reference's own command line, e.g. {"upwindCxx","-numCells","32"} */
int fdb_ref_upwind_cli(int argc, char **argv) {
  return fdb_ref_upwind_main(argc, argv);
}

int fdb_ref_upwind_threads(void) {
#ifdef HAVE_OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void fdb_ref_upwind_set_threads(int n) {
#ifdef HAVE_OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
