// Upwind.hpp -- the reference's `template <size_t NDIMS> class Upwind`
// (ref: upwind/cxx/upwind.cxx:19-135) as a header-only C++ front of the C ABI.
// Same constructor and methods; the field lives in GPU memory inside the handle and
// every method forwards to libfidib200.so.  There is no CPU path: a failed call
// throws std::runtime_error carrying fdb_last_error().
#pragma once

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fidib200.h"

namespace fidib200 {

inline void check(int rc) {
  if (rc != FDB_OK) throw std::runtime_error(fdb_last_error());
}

template <size_t NDIMS>
class Upwind {
 public:
  // ref: upwind.cxx:23-49.  `ngpus` devices share the domain in slabs along axis 0.
  Upwind(const std::vector<double>& velocity, const std::vector<double>& lengths,
         const std::vector<size_t>& numCells, int ngpus = 1)
      : numCells_(numCells), deltas_(NDIMS), ntot_(1), h_(0) {
    std::vector<int64_t> nc(NDIMS);
    for (size_t j = 0; j < NDIMS; ++j) {
      nc[j] = static_cast<int64_t>(numCells[j]);
      deltas_[j] = lengths[j] / numCells[j];
      ntot_ *= numCells[j];
    }
    check(fdb_upwind_create(static_cast<int>(NDIMS), nc.data(), velocity.data(), lengths.data(), ngpus, &h_));
  }
  ~Upwind() { fdb_upwind_destroy(h_); }
  Upwind(const Upwind&) = delete;
  Upwind& operator=(const Upwind&) = delete;

  // ref: upwind.cxx:51-86
  void advect(int numTimeSteps, double deltaTime) { check(fdb_upwind_advect(h_, numTimeSteps, deltaTime)); }

  // ref: upwind.cxx:91-93
  double checksum() const {
    double s = 0;
    check(fdb_upwind_checksum(h_, &s));
    return s;
  }

  // ref: upwind.cxx:95-103
  double std() const {
    double s = 0;
    check(fdb_upwind_std(h_, &s));
    return s;
  }

  // ref: upwind/cxx/saveVTK.h -- same ASCII rectilinear-grid file, written from a host copy
  void saveVTK(const std::string& filename) const {
    const std::vector<double> f = field();
    std::ofstream file(filename.c_str());
    file << "# vtk DataFile Version 2.0\nupwind.cxx\nASCII\nDATASET RECTILINEAR_GRID\nDIMENSIONS";
    // VTK's first dimension varies fastest: axes go out in reverse order
    for (int a = 2; a >= 0; --a) {
      if (a == 0 || static_cast<size_t>(a) < NDIMS) file << ' ' << numCells_[a] + 1;
      else file << " 1";
    }
    static const char* names[3] = {"X", "Y", "Z"};
    for (int v = 0; v < 3; ++v) {
      const int a = 2 - v;  // X <- axis 2, Y <- axis 1, Z <- axis 0
      file << '\n' << names[v] << "_COORDINATES ";
      if (a == 0 || static_cast<size_t>(a) < NDIMS) {
        file << numCells_[a] + 1 << " double\n";
        for (size_t i = 0; i < numCells_[a] + 1; ++i) file << ' ' << 0.0 + deltas_[a] * i;
      } else {
        file << "1 double\n0.0\n";
      }
    }
    file << "\nCELL_DATA " << ntot_ << "\nSCALARS f double 1\nLOOKUP_TABLE default\n";
    for (size_t i = 0; i < ntot_; ++i) {
      file << f[i] << " ";
      if ((i + 1) % 10 == 0) file << '\n';
    }
    file << '\n';
  }

  // ref: upwind.cxx:105-109
  void print() const {
    const std::vector<double> f = field();
    for (size_t i = 0; i < f.size(); ++i) std::cout << i << " " << f[i] << '\n';
  }

  std::string describe() const {
    char buf[512];
    check(fdb_upwind_describe(h_, buf, sizeof(buf)));
    return buf;
  }
  // binary dump of the field: row-major FP64, native byte order, no header
  void saveRaw(const std::string& filename) const {
    const std::vector<double> f = field();
    std::ofstream file(filename.c_str(), std::ios::binary);
    file.write(reinterpret_cast<const char*>(f.data()), (std::streamsize)(f.size() * sizeof(double)));
  }

  // beyond the reference: host access to the device field, dt helper, timing
  std::vector<double> field() const {
    std::vector<double> f(ntot_);
    check(fdb_upwind_get_field(h_, f.data()));
    return f;
  }
  void setField(const std::vector<double>& f) {
    if (f.size() != ntot_) throw std::runtime_error("setField: wrong number of cells");
    check(fdb_upwind_set_field(h_, f.data()));
  }
  double defaultDt() const {
    double dt = 0;
    check(fdb_upwind_default_dt(h_, &dt));
    return dt;
  }
  double lastGpuMilliseconds() const {
    double ms = 0;
    check(fdb_upwind_last_timing(h_, &ms, 0, 0));
    return ms;
  }
  void setKernel(int kernel) { check(fdb_upwind_set_kernel(h_, kernel)); }
  fdb_upwind* handle() { return h_; }

 private:
  std::vector<size_t> numCells_;
  std::vector<double> deltas_;
  size_t ntot_;
  fdb_upwind* h_;
};

}  // namespace fidib200
