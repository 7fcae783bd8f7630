// upwindCuda -- the reference's upwindCxx driver (ref: upwind/cxx/upwind.cxx:139-217) on the
// B200 backend.  Same flags (-numCells -numSteps -vtk -std -h), same stdout lines
// ("number of cells:", "number of time steps:", "check sum:", "std      :"), same exit code;
// new, additive options select GPUs / velocity / timing output and default to the reference.
#include <cmath>
#include <iomanip>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "Upwind.hpp"
#include "cmdline.hpp"

int main(int argc, char** argv) {
  const int ndims = 3;

  CmdLineArgParser args;
  args.setPurpose("Purpose: benchmark finite difference operations.");
  args.set("-numCells", 128, "Number of cells along each axis");
  args.set("-numSteps", 10, "Number of time steps");
  args.set("-vtk", false, "Write output to VTK file");
  args.set("-std", false, "Print out spread of solution");
  // additions (defaults reproduce the reference run)
  args.set("-ngpus", 1, "Number of GPUs of this box sharing the domain (slabs along axis 0)");
  args.set("-vx", 1.0, "Velocity along axis 0");
  args.set("-vy", 1.0, "Velocity along axis 1");
  args.set("-vz", 1.0, "Velocity along axis 2");
  args.set("-kernel", std::string("auto"), "auto | tma | generic");
  args.set("-timing", false, "Print GPU time, GCUPS and full-precision sums");
  args.set("-raw", std::string(""), "Also dump the final field to this file (row-major FP64, no header)");
  // the class supports anisotropic grids, the reference's main() does not expose them (upwind.cxx:174,182-183)
  args.set("-nx", 0, "Cells along axis 0 (0 = -numCells)");
  args.set("-ny", 0, "Cells along axis 1 (0 = -numCells)");
  args.set("-nz", 0, "Cells along axis 2 (0 = -numCells)");
  args.set("-lx", 1.0, "Domain length along axis 0");
  args.set("-ly", 1.0, "Domain length along axis 1");
  args.set("-lz", 1.0, "Domain length along axis 2");

  const bool success = args.parse(argc, argv);
  const bool help = args.get<bool>("-h");

  if (success && !help) {
    const int numTimeSteps = args.get<int>("-numSteps");
    const bool doVtk = args.get<bool>("-vtk");
    const bool doStd = args.get<bool>("-std");

    // same resolution in each direction
    std::vector<size_t> numCells(ndims, args.get<int>("-numCells"));
    const char* perAxis[3] = {"-nx", "-ny", "-nz"};
    for (int j = 0; j < ndims; ++j)
      if (args.get<int>(perAxis[j]) > 0) numCells[j] = (size_t)args.get<int>(perAxis[j]);
    std::cout << "number of cells: ";
    for (size_t i = 0; i < numCells.size(); ++i) std::cout << ' ' << numCells[i];
    std::cout << '\n';
    std::cout << "number of time steps: " << numTimeSteps << '\n';

    std::vector<double> velocity(ndims);
    velocity[0] = args.get<double>("-vx");
    velocity[1] = args.get<double>("-vy");
    velocity[2] = args.get<double>("-vz");
    std::vector<double> lengths(ndims, 1.0);
    lengths[0] = args.get<double>("-lx");
    lengths[1] = args.get<double>("-ly");
    lengths[2] = args.get<double>("-lz");

    // dt from the Courant number, ref: upwind.cxx:186-192 (|v| so that a negative
    // velocity, which the class supports, still gives a positive step)
    const double courant = 0.1;
    double dt = std::numeric_limits<double>::max();
    for (size_t j = 0; j < velocity.size(); ++j) {
      const double dx = lengths[j] / numCells[j];
      const double val = courant * dx / std::fabs(velocity[j]);
      dt = (val < dt ? val : dt);
    }

    try {
      if (args.get<bool>("-timing")) {
        // the first launch of a CUDA kernel loads its module: do that on a throw-away 32^3 problem so
        // that "gpu time" below is the stepping loop, not one-off initialisation
        fidib200::Upwind<ndims> scratch(velocity, lengths, std::vector<size_t>(ndims, 32), 1);
        scratch.advect(7, 1e-3);
      }
      fidib200::Upwind<ndims> up(velocity, lengths, numCells, args.get<int>("-ngpus"));
      const std::string kernel = args.get<std::string>("-kernel");
      if (kernel == "tma") up.setKernel(FDB_KERNEL_TMA);
      if (kernel == "generic") up.setKernel(FDB_KERNEL_GENERIC);
      if (doVtk) up.saveVTK("up0.vtk");
      up.advect(numTimeSteps, dt);
      const double sum = up.checksum();
      std::cout << "check sum: " << sum << '\n';
      double sd = 0;
      if (doStd) {
        sd = up.std();
        std::cout << "std      : " << sd << '\n';
      }
      if (doVtk) up.saveVTK("up1.vtk");
      if (!args.get<std::string>("-raw").empty()) up.saveRaw(args.get<std::string>("-raw"));
      if (args.get<bool>("-timing")) {
        const double ms = up.lastGpuMilliseconds();
        const double updates = double(numCells[0]) * numCells[1] * numCells[2] * numTimeSteps;
        std::cout << "kernel: " << up.describe() << '\n';
        std::cout << std::setprecision(17) << "check sum (17 digits): " << sum << '\n';
        if (doStd) std::cout << "std (17 digits): " << sd << '\n';
        std::cout << std::setprecision(6) << "gpu time [ms]: " << ms << "  GCUPS: " << updates / ms / 1e6
                  << "  HBM roofline GB/s (16 B/update): " << updates * 16 / ms / 1e6 << '\n';
      }
    } catch (const std::exception& e) {
      std::cerr << "ERROR: " << e.what() << '\n';
      return 1;
    }
  } else {
    // error when parsing command line arguments
    if (!success) std::cerr << "ERROR when parsing command line arguments\n";
    args.help();
  }
  return 0;
}
