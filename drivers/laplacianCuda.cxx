// laplacianCuda -- the reference's laplacian driver (ref: laplacian/cxx/laplacian.cxx:30-129)
// on the B200 backend: -numCells (8000) -numDims (2) -vtk, 10 x { applyFilter; copyOutToIn },
// "Laplace times min/max/avg:" and "Check sums: input = .. output = .." lines.
// MPI ranks are replaced by the GPUs of one box (-ngpus).
#include <chrono>
#include <cmath>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "Filter.hpp"
#include "cmdline.hpp"

// ref: laplacian.cxx:22-28
static double func(const std::vector<double>& pos) {
  double res = 1;
  for (size_t i = 0; i < pos.size(); ++i) res *= sin(2.0 * M_PI * pos[i]);
  return res;
}

// func's 1-D factor: func(pos) = ((1 * sin1(pos[0])) * sin1(pos[1])) * ...
static double sin1(double x) { return sin(2.0 * M_PI * x); }

// ref: laplacian.cxx:55-65 -- 2*numDims+1 points, diagonal -2*numDims, neighbours +1, no 1/h^2 scaling
static std::map<std::vector<int>, double> laplacianStencil(size_t numDims) {
  std::map<std::vector<int>, double> st;
  const std::vector<int> centre(numDims, 0);
  st[centre] = -2.0 * numDims;
  for (size_t axis = 0; axis < numDims; ++axis)
    for (int side = -1; side <= 1; side += 2) {
      std::vector<int> o(centre);
      o[axis] = side;
      st[o] = 1.0;
    }
  return st;
}

int main(int argc, char** argv) {
  CmdLineArgParser args;
  args.setPurpose("Purpose: benchmark finite difference operations.");
  args.set("-numCells", 8000, "Number of cells along each axis");
  args.set("-numDims", 2, "Number of dimensions");
  args.set("-vtk", false, "Write output to VTK file");
  args.set("-ngpus", 1, "Number of GPUs of this box sharing the domain (slabs along axis 0)");
  args.set("-numIter", 10, "Number of apply/copy iterations (the reference hard-codes 10)");
  args.set("-raw", std::string(""), "Also dump the output data to this file (row-major FP64, no header)");
  args.set("-refwrap", false, "Wrap indices as the reference does (int %= size_t: not periodic unless numCells is a power of two)");
  args.set("-hostinit", false, "Evaluate the input function cell by cell on the host, as the reference does (same bits, slower)");

  const bool success = args.parse(argc, argv);
  const bool help = args.get<bool>("-h");

  if (success && !help) {
    const size_t numCells = (size_t)args.get<int>("-numCells");
    const size_t numDims = (size_t)args.get<int>("-numDims");
    const bool writeVTK = args.get<bool>("-vtk");

    const std::map<std::vector<int>, double> stencil = laplacianStencil(numDims);

    std::vector<size_t> globalDims(numDims, numCells);
    std::vector<double> xmins(numDims, 0.0), xmaxs(numDims, 1.0);

    try {
      fidib200::Filter fltr(globalDims, xmins, xmaxs, stencil, args.get<int>("-ngpus"));
      if (!fltr.isDecompValid()) std::cerr << "Decomposition is invalid\n";
      if (fltr.isDecompValid()) {
        if (args.get<bool>("-refwrap")) fltr.setRefWrap(true);
        if (args.get<bool>("-hostinit"))
          fltr.setInData(func);
        else
          fltr.setInDataProduct(sin1);  // same bits as setInData(func), the product runs on the device
        const auto tic = std::chrono::steady_clock::now();
        // repeat to improve statistics
        const size_t numIter = (size_t)args.get<int>("-numIter");
        // numIter x { applyFilter(); copyOutToIn(); } (ref: laplacian.cxx:86-90) handed over as one call,
        // so that the library may run two applies per sweep (same bits, half the DRAM traffic)
        fltr.iterate((long)numIter);
        const double walltime = std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count();

        if (writeVTK) {
          std::cout << "Data will be written to file laplacian.vtk\n";
          fltr.saveVTK("laplacian.vtk");
        }
        const std::string raw = args.get<std::string>("-raw");
        if (!raw.empty()) fltr.saveRaw(raw);
        const double inSum = fltr.computeCheckSum("input");
        const double outSum = fltr.computeCheckSum("output");
        std::cout << "Laplace times min/max/avg: " << walltime << '/' << walltime << '/' << walltime << " [seconds]\n";
        std::cout << "Check sums: input = " << inSum << " output = " << outSum << '\n';
      }
    } catch (const std::exception& e) {
      std::cerr << "ERROR: " << e.what() << '\n';
      return 1;
    }
  } else {
    if (!success) std::cerr << "ERROR when parsing command line arguments\n";
    args.help();
  }
  return 0;
}
