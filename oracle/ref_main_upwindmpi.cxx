/* TEST INFRASTRUCTURE ONLY: the reference's MPI upwind driver, untouched, with
 * its main() renamed so it can live in oracle/_ref/libref_filter.so. */
#include <numeric> /* upwindMpi.cxx uses std::accumulate without including it */
#define main fdb_ref_upwindmpi_main_cxx
#include "upwindMpi.cxx" /* -I$(REF)/upwind/cxx */
#undef main
extern "C" int fdb_ref_upwindmpi_main(int argc, char **argv) { return fdb_ref_upwindmpi_main_cxx(argc, argv); }
