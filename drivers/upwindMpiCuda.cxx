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

int main(int argc, char** argv) {
  CmdLineArgParser args;
  args.setPurpose("Purpose: benchmark finite difference operations.");
  args.set("-numCells", 128, "Number of cells along each axis");
  args.set("-numSteps", 10, "Number of time steps");
  args.set("-vtk", false, "Write output to VTK file");
  args.set("-ngpus", 1, "Number of GPUs of this box sharing the domain (slabs along axis 0)");

  const bool success = args.parse(argc, argv);
  const bool help = args.get<bool>("-h");

  if (success && !help) {
    const size_t numDims = 3;
    const size_t numCells = (size_t)args.get<int>("-numCells");
    const size_t numSteps = (size_t)args.get<int>("-numSteps");
    const bool writeVTK = args.get<bool>("-vtk");

    const std::vector<double> velocities(numDims, 1.);
    const std::vector<double> lengths(numDims, 1.);
    // time step and stencil weights, ref: upwindMpi.cxx:62-92
    const double courant = 0.1;
    std::vector<double> deltas(numDims);
    std::vector<int> signs(numDims);
    double dt = std::numeric_limits<double>::max();
    for (size_t j = 0; j < numDims; ++j) {
      const double dx = lengths[j] / (double)numCells;
      deltas[j] = dx;
      const double val = courant * dx / velocities[j];
      dt = (val < dt ? val : dt);
      signs[j] = (velocities[j] > 0 ? 1 : -1);
    }
    std::map<std::vector<int>, double> stencil;
    std::vector<int> offset(numDims, 0);
    double diag = 1.0;
    for (size_t i = 0; i < numDims; ++i) diag -= signs[i] * dt * velocities[i] / deltas[i];
    stencil[offset] = diag;
    for (size_t i = 0; i < numDims; ++i) {
      offset[i] = -signs[i];
      stencil[offset] = signs[i] * dt * velocities[i] / deltas[i];
      offset[i] = 0;
    }

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
