// C ABI around the reference's remap_Q_ppm (src/preqx/unit_tests/remap.cpp), compiled from
// the reference tree by oracle/Makefile into oracle/_ref/. Test infrastructure only.
#include "remap.hpp"
extern "C" {
int ref_remap_nlev() { return NLEV; }
// Qdp [qsize][NLEV][NP][NP], dp1/dp2 [NLEV][NP][NP]
void ref_remap_Q_ppm(double* Qdp, int qsize, const double* dp1, const double* dp2, int alg) {
  remap_Q_ppm(reinterpret_cast<Real(*)[NLEV][NP][NP]>(Qdp), qsize,
              reinterpret_cast<const Real(*)[NP][NP]>(dp1), reinterpret_cast<const Real(*)[NP][NP]>(dp2), alg);
}
}
