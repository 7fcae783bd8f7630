// CubeDecomp.hpp -- C++ front with the interface of the reference's `class CubeDecomp`
// (ref: cxx/CubeDecomp.h, cxx/CubeDecomp.cpp:11-131) over fdb_cube_* (include/fidib200.h): the process grid
// the reference would choose, each rank's block and its periodic neighbours.  Host-only; the CUDA engines
// partition in slabs, this serves callers that need the reference's decomposition (e.g. to match an MPI run).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "fidib200.h"

namespace fidib200 {

class CubeDecomp {
 public:
  CubeDecomp() : nprocs_(0) {}

  // ref: CubeDecomp::build, CubeDecomp.cpp:11-31 -- false when no divisor tuple multiplies to nprocs
  bool build(int nprocs, const std::vector<size_t>& dims) {
    nprocs_ = nprocs;
    dims_.assign(dims.begin(), dims.end());
    std::vector<int64_t> grid(dims.size(), 0);
    if (fdb_cube_decomp(nprocs, (int)dims.size(), dims_.data(), grid.data()) != FDB_OK) {
      decomp_.clear();
      return false;
    }
    decomp_.assign(grid.begin(), grid.end());
    return true;
  }
  std::vector<size_t> getDecomp() const { return decomp_; }
  std::vector<size_t> getBegIndices(int rk) const { return block(rk, true); }
  std::vector<size_t> getEndIndices(int rk) const { return block(rk, false); }
  int getNeighborRank(int rk, const std::vector<int>& dir) const {
    int nb = -1;
    fdb_cube_neighbor(nprocs_, (int)dims_.size(), dims_.data(), rk, dir.data(), &nb);
    return nb;
  }

 private:
  std::vector<size_t> block(int rk, bool begin) const {
    std::vector<int64_t> lo(dims_.size(), 0), hi(dims_.size(), 0);
    fdb_cube_block(nprocs_, (int)dims_.size(), dims_.data(), rk, lo.data(), hi.data());
    const std::vector<int64_t>& v = begin ? lo : hi;
    return std::vector<size_t>(v.begin(), v.end());
  }
  int nprocs_;
  std::vector<int64_t> dims_;
  std::vector<size_t> decomp_;
};

}  // namespace fidib200
