/* TEST INFRASTRUCTURE ONLY: the reference's laplacian driver, untouched, with
 * its main() renamed so it can live in oracle/_ref/libref_filter.so. */
#define main fdb_ref_laplacian_main_cxx
#include "laplacian.cxx" /* -I$(REF)/laplacian/cxx */
#undef main
extern "C" int fdb_ref_laplacian_main(int argc, char **argv) { return fdb_ref_laplacian_main_cxx(argc, argv); }
