// testStencil2dCuda -- the reference's 2-D stencil check (ref: laplacian/cxx/testStencil2d.cxx:42-118)
// on the B200 backend: an 8 x 8 field that is 1 wherever an index is 0, one applyFilter of the
// stencil {(0,0): 0, (1,0): +1, (0,-1): -1}, outData printed through printOutData.
#include <cstddef>
#include <iostream>
#include <map>
#include <vector>

#include "Filter.hpp"
#include "cmdline.hpp"

// ref: testStencil2d.cxx:28-37 -- one wherever some index is zero
static double func(const std::vector<size_t>& inds) {
  size_t prod = 1;
  for (size_t i = 0; i < inds.size(); ++i) prod *= inds[i];
  return prod == 0 ? 1.0 : 0.0;
}

int main(int argc, char** argv) {
  CmdLineArgParser args;
  args.setPurpose("Purpose: benchmark finite difference operations.");
  args.set("-numCells", 8, "Number of cells along each axis");
  args.set("-vtk", false, "Write output to VTK file");
  args.set("-ngpus", 1, "Number of GPUs of this box sharing the domain (slabs along axis 0)");

  const bool success = args.parse(argc, argv);
  const bool help = args.get<bool>("-h");

  if (success && !help) {
    const size_t numCells = (size_t)args.get<int>("-numCells");
    const size_t numDims = 2;
    const bool writeVTK = args.get<bool>("-vtk");

    // ref: testStencil2d.cxx:63-75
    std::map<std::vector<int>, double> stencil;
    std::vector<int> offset(numDims, 0);
    stencil[offset] = 0.0;
    offset[0] = 1;
    stencil[offset] = 1.0;
    offset[0] = 0;
    offset[1] = -1;
    stencil[offset] = -1.0;
    offset[1] = 0;

    std::vector<size_t> globalDims(numDims, numCells);
    std::vector<double> xmins(numDims, 0.0), xmaxs(numDims, 1.0);

    try {
      fidib200::Filter fltr(globalDims, xmins, xmaxs, stencil, args.get<int>("-ngpus"));
      if (!fltr.isDecompValid()) std::cerr << "Decomposition is invalid\n";
      if (fltr.isDecompValid()) {
        fltr.setInDataByIndices(func);
        fltr.applyFilter();
        fltr.printOutData();
        if (writeVTK) {
          std::cout << "Data will be written to file stencil2d.vtk\n";
          fltr.saveVTK("stencil2d.vtk");
        }
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
