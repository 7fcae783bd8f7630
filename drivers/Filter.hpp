// Filter.hpp -- the reference's `class Filter` (ref: cxx/Filter.h:36-170) as a header-only
// C++ front of the C ABI: an offset->weight stencil applied to a periodic field that is
// slab-decomposed over the GPUs of one box.  Same public methods as the reference for the
// ones its drivers call (laplacian.cxx:75-114, upwindMpi.cxx:101-150, testStencil2d.cxx:85-96).
// MPI ranks become GPUs of this process: getRank() is 0 and getNumProcs() the GPU count.
#pragma once

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "fidib200.h"

namespace fidib200 {

class Filter {
 public:
  // ref: Filter.cpp:11-78
  Filter(const std::vector<size_t>& globalDims, const std::vector<double>& xmins,
         const std::vector<double>& xmaxs, const std::map<std::vector<int>, double>& stencil, int ngpus = 1)
      : globalDims_(globalDims), xmins_(xmins), xmaxs_(xmaxs), ngpus_(ngpus), valid_(false), h_(0) {
    const int nd = static_cast<int>(globalDims.size());
    std::vector<int64_t> dims(globalDims.begin(), globalDims.end());
    std::vector<int> offs;
    std::vector<double> w;
    for (std::map<std::vector<int>, double>::const_iterator it = stencil.begin(); it != stencil.end(); ++it) {
      offs.insert(offs.end(), it->first.begin(), it->first.end());
      w.push_back(it->second);
    }
    const int rc = fdb_stencil_create(nd, dims.data(), static_cast<int>(w.size()), offs.data(), w.data(), ngpus, &h_);
    if (rc == FDB_E_DECOMP) {  // ref: Filter.cpp:28-34 -- report, stay alive, let the driver check
      std::cerr << "ERROR: No valid domain decomposition could be found. Adjust the number\n";
      std::cerr << "of processes and/or the domain dimensions.\n";
      return;
    }
    if (rc != FDB_OK) throw std::runtime_error(fdb_last_error());
    valid_ = true;
    std::cout << "Number of procs: " << ngpus << "\nglobal dimensions: ";
    for (int i = 0; i < nd; ++i) std::cout << globalDims[i] << ' ';
    std::cout << "\nDomain decomp ";  // slabs along axis 0 instead of CubeDecomp's process grid
    for (int i = 0; i < nd; ++i) std::cout << (i == 0 ? ngpus : 1) << ' ';
    std::cout << '\n';
  }
  ~Filter() { fdb_stencil_destroy(h_); }
  Filter(const Filter&) = delete;
  Filter& operator=(const Filter&) = delete;

  int getRank() const { return 0; }
  int getNumProcs() const { return ngpus_; }
  bool isDecompValid() const { return valid_; }

  // ref: Filter.cpp:103-112
  std::vector<double> getPosition(const std::vector<size_t>& globalInds) const {
    std::vector<double> pos(globalDims_.size());
    for (size_t i = 0; i < pos.size(); ++i) {
      const double delta = (xmaxs_[i] - xmins_[i]) / double(globalDims_[i]);
      pos[i] = xmins_[i] + (globalInds[i] + 0.5) * delta;
    }
    return pos;
  }

  // ref: Filter.cpp:131-159 -- the callback runs on the host (same libm, same bits)
  void setInData(double (*f)(const std::vector<double>&)) {
    std::vector<double> a(total());
    std::vector<size_t> inds(globalDims_.size(), 0);
    for (size_t c = 0; c < a.size(); ++c) {
      a[c] = f(getPosition(inds));
      next(inds);
    }
    check(fdb_stencil_set_input(h_, a.data(), FDB_ROW_MAJOR));
  }

  // ref: Filter.cpp:161-188
  void setInDataByIndices(double (*f)(const std::vector<size_t>&)) {
    std::vector<double> a(total());
    std::vector<size_t> inds(globalDims_.size(), 0);
    for (size_t c = 0; c < a.size(); ++c) {
      a[c] = f(inds);
      next(inds);
    }
    check(fdb_stencil_set_input(h_, a.data(), FDB_ROW_MAJOR));
  }

  // setInData for a callback that is a product of one 1-D function over the axes (laplacian.cxx:22-28's func is
  // prod_j sin(2 pi x_j)): the 1-D factors are evaluated here with the host's libm at Filter::getPosition's
  // positions and multiplied out on the device in setInData's order -- the same bits as setInData(func), with
  // 8 B x (d_0 + ... + d_{n-1}) crossing PCIe instead of the whole field.
  void setInDataProduct(double (*f1)(double)) {
    const size_t nd = globalDims_.size();
    std::vector<std::vector<double> > fac(nd);
    std::vector<const double*> ptr(nd);
    for (size_t j = 0; j < nd; ++j) {
      fac[j].resize(globalDims_[j]);
      const double delta = (xmaxs_[j] - xmins_[j]) / double(globalDims_[j]);
      for (size_t i = 0; i < globalDims_[j]; ++i) fac[j][i] = f1(xmins_[j] + (i + 0.5) * delta);
      ptr[j] = fac[j].data();
    }
    check(fdb_stencil_set_input_separable(h_, ptr.data()));
  }
  // reproduce Filter.cpp:240's (int %= size_t) index wrap (non-periodic unless the extent is a power of two)
  void setRefWrap(bool on) { check(fdb_stencil_set_ref_wrap(h_, on ? 1 : 0)); }
  // binary dump of the output data: row-major FP64, native byte order, no header (full precision; the ASCII
  // VTK file keeps six digits)
  void saveRaw(const std::string& filename) {
    const std::vector<double> f = getData(FDB_OUTPUT, FDB_ROW_MAJOR);
    std::ofstream file(filename.c_str(), std::ios::binary);
    file.write(reinterpret_cast<const char*>(f.data()), (std::streamsize)(f.size() * sizeof(double)));
  }

  void applyFilter() { check(fdb_stencil_apply(h_)); }  // ref: Filter.cpp:191-263
  void copyOutToIn() { check(fdb_stencil_swap(h_)); }   // ref: Filter.cpp:440-463
  // niter x { applyFilter(); copyOutToIn(); } (ref: laplacian.cxx:86-90) in one call: the 3-D 7-point stencil
  // then runs two applies per sweep (same bits); setFuse(1) turns that off, fuse() tells what will run
  void iterate(long niter) { check(fdb_stencil_iterate(h_, niter)); }
  void setFuse(int appliesPerSweep) { check(fdb_stencil_set_fuse(h_, appliesPerSweep)); }
  int fuse() const {
    int n = 1;
    check(fdb_stencil_get_fuse(h_, &n));
    return n;
  }

  // ref: Filter.cpp:465-485
  double computeCheckSum(const std::string& inOrOut) {
    double s = 0;
    check(fdb_stencil_checksum(h_, inOrOut == "input" ? FDB_INPUT : FDB_OUTPUT, &s));
    return s;
  }

  std::vector<double> getData(int which, int layout = FDB_ROW_MAJOR) {
    std::vector<double> a(total());
    check(fdb_stencil_get(h_, which, a.data(), layout));
    return a;
  }

  // ref: Filter.cpp:231-261 (printOutData) -- "[rk] inds = .. outData = .." in the
  // reference's column-major visiting order
  void printOutData() {
    const std::vector<double> a = getData(FDB_OUTPUT, FDB_ROW_MAJOR);
    const size_t nd = globalDims_.size();
    std::cerr << "[0] outData: \n";
    std::vector<size_t> inds(nd, 0);
    for (size_t c = 0; c < a.size(); ++c) {
      size_t flat = 0;
      for (size_t j = 0; j < nd; ++j) flat = flat * globalDims_[j] + inds[j];
      std::cerr << "\t[0] inds = ";
      for (size_t j = 0; j < nd; ++j) std::cerr << inds[j] << ' ';
      std::cerr << " outData = " << a[flat] << '\n';
      for (size_t j = 0; j < nd; ++j) {  // first axis fastest
        if (++inds[j] < globalDims_[j]) break;
        inds[j] = 0;
      }
    }
  }

  // ref: Filter.cpp:487-538 + cxx/writeVTK.cpp:12-95 -- same ASCII structured-grid file
  void saveVTK(const std::string& filename) {
    const size_t nd = globalDims_.size();
    if (nd > 3) {
      std::cerr << "WARNING: writeVTK does not support more than 3 dimensions\n";
      return;
    }
    const std::vector<double> col = getData(FDB_OUTPUT, nd > 1 && ngpus_ == 1 ? FDB_COL_MAJOR : FDB_ROW_MAJOR);
    std::vector<double> field = col;
    if (nd > 1 && ngpus_ != 1) {  // multi-GPU handles hand out row-major only: reorder on the host
      std::vector<size_t> inds(nd, 0);
      for (size_t c = 0; c < field.size(); ++c) {
        size_t flat = 0;
        for (size_t j = 0; j < nd; ++j) flat = flat * globalDims_[j] + inds[j];
        field[c] = col[flat];
        for (size_t j = 0; j < nd; ++j) {
          if (++inds[j] < globalDims_[j]) break;
          inds[j] = 0;
        }
      }
    }
    size_t cells[3] = {1, 1, 1};
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    for (size_t j = 0; j < nd; ++j) {
      cells[j] = globalDims_[j];
      lo[j] = xmins_[j];
      hi[j] = xmaxs_[j];
    }
    const size_t nodes[3] = {nd > 0 ? cells[0] + 1 : 2, nd > 1 ? cells[1] + 1 : 2, nd > 2 ? cells[2] + 1 : 2};
    std::ofstream file(filename.c_str());
    file << "# vtk DataFile Version 2.0\nproduced by laplacian\nASCII\nDATASET STRUCTURED_GRID\n";
    file << "DIMENSIONS " << nodes[0] << ' ' << nodes[1] << ' ' << nodes[2] << '\n';
    file << "POINTS " << nodes[0] * nodes[1] * nodes[2] << " float\n";
    for (size_t k = 0; k < nodes[2]; ++k)
      for (size_t j = 0; j < nodes[1]; ++j)
        for (size_t i = 0; i < nodes[0]; ++i)
          file << lo[0] + (hi[0] - lo[0]) * i / double(cells[0]) << ' '
               << lo[1] + (hi[1] - lo[1]) * j / double(cells[1]) << ' '
               << lo[2] + (hi[2] - lo[2]) * k / double(cells[2]) << '\n';
    file << "CELL_DATA " << cells[0] * cells[1] * cells[2] << '\n';
    file << "SCALARS outData float\nLOOKUP_TABLE default\n";
    for (size_t c = 0; c < field.size(); ++c) file << field[c] << '\n';
  }

  double lastGpuMilliseconds() const {
    double ms = 0;
    check(fdb_stencil_last_timing(h_, &ms, 0, 0));
    return ms;
  }
  fdb_stencil* handle() { return h_; }

 private:
  static void check(int rc) {
    if (rc != FDB_OK) throw std::runtime_error(fdb_last_error());
  }
  size_t total() const {
    size_t n = 1;
    for (size_t j = 0; j < globalDims_.size(); ++j) n *= globalDims_[j];
    return n;
  }
  void next(std::vector<size_t>& inds) const {  // row-major successor (last axis fastest)
    for (size_t j = inds.size(); j-- > 0;) {
      if (++inds[j] < globalDims_[j]) return;
      inds[j] = 0;
    }
  }

  std::vector<size_t> globalDims_;
  std::vector<double> xmins_, xmaxs_;
  int ngpus_;
  bool valid_;
  fdb_stencil* h_;
};

}  // namespace fidib200
