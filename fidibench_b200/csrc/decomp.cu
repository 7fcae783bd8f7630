// decomp.cu -- host-side restatement of the reference's process-grid chooser (CubeDecomp), kept for
// callers that want the reference's block decomposition (more than 8 GPUs, or parity with an MPI run).
// The CUDA engines themselves partition in slabs along axis 0 (runtime.cu); nothing here touches a GPU.
//
// ref: cxx/CubeDecomp.cpp:11-170, cxx/MultiArrayIter.h:51-57,92-99.  The reference's quirks are part of
// the contract and reproduced on purpose:
//   * the candidate extents per axis are ALL divisors of the axis length (getPrimeFactors returns
//     1, every k <= n/2 dividing n, and n -- so an axis of length 1 lists 1 twice);
//   * candidates are enumerated with the FIRST axis running fastest (column-major MultiArrayIter);
//   * the first valid candidate is kept only if it is the only one: the cost scan starts at the second
//     (CubeDecomp.cpp:73-80), ties keep the earlier candidate;
//   * cost = surface / volume of the local block with truncating integer division of the extents;
//   * ranks are laid out row-major over the process grid (last axis fastest), neighbours are periodic.
#include <cfloat>
#include <vector>

#include "fdb_internal.h"

namespace {

std::vector<int64_t> divisors(int64_t n) {
  std::vector<int64_t> d(1, 1);
  for (int64_t k = 2; k <= n / 2; ++k)
    if ((n / k) * k == n) d.push_back(k);
  d.push_back(n);
  return d;
}

double block_cost(int nd, const int64_t* dims, const int64_t* grid) {
  int64_t loc[3];
  for (int j = 0; j < nd; ++j) loc[j] = dims[j] / grid[j];
  uint64_t volume = 1, surface = 0;
  for (int i = 0; i < nd; ++i) {
    volume *= (uint64_t)loc[i];
    uint64_t area = 1;
    for (int j = 0; j < nd; ++j)
      if (j != i) area *= (uint64_t)loc[j];
    surface += area;
  }
  return (double)surface / (double)volume;
}

int choose_grid(int nprocs, int nd, const int64_t* dims, int64_t* grid) {
  if (nprocs < 1 || nd < 1 || nd > 3 || !dims || !grid)
    return fdb::set_error(FDB_E_INVALID, "bad cube decomposition request (nprocs=%d ndims=%d)", nprocs, nd);
  std::vector<int64_t> cand[3];
  int64_t total = 1;
  for (int j = 0; j < nd; ++j) {
    if (dims[j] < 1) return fdb::set_error(FDB_E_INVALID, "extent %d is %lld", j, (long long)dims[j]);
    cand[j] = divisors(dims[j]);
    total *= (int64_t)cand[j].size();
  }
  bool have_first = false, have_best = false;
  int64_t first[3] = {1, 1, 1}, best[3] = {1, 1, 1};
  double min_cost = DBL_MAX;
  int64_t nvalid = 0;
  for (int64_t t = 0; t < total; ++t) {
    int64_t g[3] = {1, 1, 1}, rest = t, prod = 1;
    for (int j = 0; j < nd; ++j) {  // first axis fastest
      g[j] = cand[j][(size_t)(rest % (int64_t)cand[j].size())];
      rest /= (int64_t)cand[j].size();
      prod *= g[j];
    }
    if (prod != nprocs) continue;
    ++nvalid;
    if (!have_first) {
      for (int j = 0; j < nd; ++j) first[j] = g[j];
      have_first = true;
      continue;  // never costed (CubeDecomp.cpp:73)
    }
    const double c = block_cost(nd, dims, g);
    if (c < min_cost) {
      min_cost = c;
      for (int j = 0; j < nd; ++j) best[j] = g[j];
      have_best = true;
    }
  }
  if (nvalid == 0)
    return fdb::set_error(FDB_E_DECOMP, "No valid domain decomposition: %d process(es) do not tile the grid", nprocs);
  for (int j = 0; j < nd; ++j) grid[j] = have_best ? best[j] : first[j];
  return FDB_OK;
}

}  // namespace

extern "C" {

int fdb_cube_decomp(int nprocs, int ndims, const int64_t* dims, int64_t* grid) {
  try {
    return choose_grid(nprocs, ndims, dims, grid);
  } catch (...) {
    return fdb::set_error(FDB_E_OOM, "out of host memory");
  }
}

int fdb_cube_block(int nprocs, int ndims, const int64_t* dims, int rank, int64_t* lo, int64_t* hi) {
  int64_t grid[3] = {1, 1, 1};
  if (!lo || !hi || rank < 0 || rank >= nprocs) return fdb::set_error(FDB_E_INVALID, "bad cube block request");
  FDB_TRY(fdb_cube_decomp(nprocs, ndims, dims, grid));
  int64_t rest = rank;
  for (int j = ndims - 1; j >= 0; --j) {  // row-major rank layout: last axis fastest
    const int64_t idx = rest % grid[j];
    rest /= grid[j];
    const int64_t loc = dims[j] / grid[j];
    lo[j] = idx * loc;
    hi[j] = (idx + 1) * loc;
  }
  return FDB_OK;
}

int fdb_cube_neighbor(int nprocs, int ndims, const int64_t* dims, int rank, const int* dir, int* neighbor) {
  int64_t grid[3] = {1, 1, 1};
  if (!dir || !neighbor || rank < 0 || rank >= nprocs)
    return fdb::set_error(FDB_E_INVALID, "bad cube neighbour request");
  FDB_TRY(fdb_cube_decomp(nprocs, ndims, dims, grid));
  int64_t idx[3] = {0, 0, 0}, rest = rank;
  for (int j = ndims - 1; j >= 0; --j) {
    idx[j] = rest % grid[j];
    rest /= grid[j];
  }
  int64_t nb = 0;
  for (int j = 0; j < ndims; ++j) {
    int64_t v = idx[j] + dir[j];  // one periodic image either way, as the reference (CubeDecomp.cpp:119-126)
    if (v < 0) v += grid[j];
    else if (v >= grid[j]) v -= grid[j];
    nb = nb * grid[j] + v;
  }
  *neighbor = (int)nb;
  return FDB_OK;
}

}  // extern "C"
