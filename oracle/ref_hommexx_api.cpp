// Section B of include/hommexx_b200.h for the REFERENCE library: the few build-information symbols the host driver
// (hommexx_b200/driver) asks every dycore library for, so that oracle/_ref/libref_hommexx_<PLEV>_<QSIZE_D>.so — the
// reference's own src/share/cxx sources compiled against the serial Kokkos stand-in of oracle/ref_shim — binds to the
// same driver as the product and the oracle. Everything in section A comes from the reference's own
// cxx_f90_interface.cpp / prim_driver.cpp / mpi_cxx_f90_interface.cpp. TEST INFRASTRUCTURE.
#include <cstdint>
extern "C" {
int hommexx_b200_nlev(void) { return PLEV; }
int hommexx_b200_qsize_d(void) { return QSIZE_D; }
const char* hommexx_b200_backend(void) { return "reference-serial"; }
void hommexx_b200_set_comm(int, int, int, const void*) {}
int64_t hommexx_b200_launch_count(void) { return 0; }
void hommexx_b200_sync(void) {}
}
