// upwindMpiCuda -- the reference's distributed upwind driver (ref: upwind/cxx/upwindMpi.cxx:30-166)
// on the B200 backend: the upwind step expressed as a 4-branch Filter stencil, one
// applyFilter + copyOutToIn per step, the per-iteration "iter i check sum  in/out" lines
// (CHECK_NAN is defined in the reference), " times min/max/avg:" and "Check sum:".
#include <chrono>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "Filter.hpp"
#include "cmdline.hpp"

// ref: upwindMpi.cxx:21-27 -- one in the cell whose indices are all zero
static double initialCondition(const std::vector<size_t>& inds) {
  return std::accumulate(inds.begin(), inds.end(), size_t(0)) == 0 ? 1.0 : 0.0;
}

// ref: upwindMpi.cxx:62-92 -- dt from the Courant number 0.1, then the upwind step as a stencil:
// centre 1 - sum_i s_i dt v_i / dx_i, and s_i dt v_i / dx_i at offset -s_i along axis i
static std::map<std::vector<int>, double> upwindStencil(size_t numCells, const std::vector<double>& v,
                                                        const std::vector<double>& lengths) {
  const size_t nd = v.size();
  const double courant = 0.1;
  std::vector<double> dx(nd);
  std::vector<int> sign(nd);
  double dt = std::numeric_limits<double>::max();
  for (size_t j = 0; j < nd; ++j) {
    dx[j] = lengths[j] / (double)numCells;
    const double val = courant * dx[j] / v[j];
    dt = (val < dt ? val : dt);
    sign[j] = (v[j] > 0 ? 1 : -1);
  }
  std::map<std::vector<int>, double> st;
  std::vector<int> o(nd, 0);
  double diag = 1.0;
  for (size_t i = 0; i < nd; ++i) diag -= sign[i] * dt * v[i] / dx[i];
  st[o] = diag;
  for (size_t i = 0; i < nd; ++i) {
    o[i] = -sign[i];
    st[o] = sign[i] * dt * v[i] / dx[i];
    o[i] = 0;
  }
  return st;
}

int main(int argc, char** argv) {
  CmdLineArgParser args;
  args.setPurpose("Purpose: benchmark finite difference operations.");
  args.set("-numCells", 128, "Number of cells along each axis");
  args.set("-numSteps", 10, "Number of time steps");
  args.set("-vtk", false, "Write output to VTK file");
  args.set("-ngpus", 1, "Number of GPUs of this box sharing the domain (slabs along axis 0)");
  args.set("-raw", std::string(""), "Also dump the output data to this file (row-major FP64, no header)");

  const bool success = args.parse(argc, argv);
  const bool help = args.get<bool>("-h");

  if (success && !help) {
    const size_t numDims = 3;
    const size_t numCells = (size_t)args.get<int>("-numCells");
    const size_t numSteps = (size_t)args.get<int>("-numSteps");
    const bool writeVTK = args.get<bool>("-vtk");

    const std::vector<double> velocities(numDims, 1.), lengths(numDims, 1.);
    const std::map<std::vector<int>, double> stencil = upwindStencil(numCells, velocities, lengths);

    std::vector<size_t> globalDims(numDims, numCells);
    std::vector<double> xmins(numDims, 0.0);

    try {
      fidib200::Filter fltr(globalDims, xmins, lengths, stencil, args.get<int>("-ngpus"));
      if (!fltr.isDecompValid()) std::cerr << "Decomposition is invalid\n";
      if (fltr.isDecompValid()) {
        const auto tic = std::chrono::steady_clock::now();
        fltr.setInDataByIndices(initialCondition);
        for (size_t i = 0; i < numSteps; ++i) {
          fltr.applyFilter();
          const double inSum = fltr.computeCheckSum("input");
          const double outSum = fltr.computeCheckSum("output");
          std::cout << "iter " << i << " check sum  in/out = " << inSum << " / " << outSum << '\n';
          fltr.copyOutToIn();
        }
        const double walltime = std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count();
        if (writeVTK) {
          std::cout << "Data will be written to file upMpi.vtk\n";
          fltr.saveVTK("upMpi.vtk");
        }
        const std::string raw = args.get<std::string>("-raw");
        if (!raw.empty()) fltr.saveRaw(raw);
        const double outSum = fltr.computeCheckSum("output");
        std::cout << " times min/max/avg: " << walltime << '/' << walltime << '/' << walltime << " [seconds]\n";
        std::cout << "Check sum: " << outSum << '\n';
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
