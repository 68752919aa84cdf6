// Quasi-monotone limiters on one level (16 GLL points) held in registers — thread-local.
// Restates the limiter shell and limiter_optim_iter_full / limiter_clip_and_sum of
// EulerStepFunctorImpl.hpp:693-884 (serial reduction order k = 0..15, as SerialLimiter).
#pragma once
#include "hxx.cuh"

namespace hxx {

// x = ptens/dpmass, c = spheremp*dpmass (both prepared by the caller). minp/maxp are the
// level's qlim entries and are updated as the reference updates them. Returns false when the
// level is skipped (sum of weights <= 0), in which case x must not be written back.
// limiter_level_w: the same with the weight sum sumc = c[0] + ... + c[15] (> 0) supplied by the caller,
// who computes it once for all tracers of a level.
// The weights come through an accessor: a register array (RegPlane) or per-thread shared-memory slots
// (SlotPlane: point k at base[k * stride]).
struct RegPlane {
  const double (&c)[NPSQ];
  __device__ __forceinline__ double operator[](int k) const { return c[k]; }
};
struct SlotPlane {
  const double* base;
  int stride;
  __device__ __forceinline__ double operator[](int k) const { return base[k * stride]; }
};
template <class W>
__device__ __forceinline__ void limiter_level_w(int limiter_option, const W c, double sumc, double (&x)[NPSQ],
                                                double& qmin, double& qmax) {
  double mass = x[0] * c[0];
#pragma unroll
  for (int k = 1; k < NPSQ; ++k) mass += x[k] * c[k];
  double minp = qmin, maxp = qmax;
  if (minp < 0) minp = qmin = 0.0;
  if (mass < minp * sumc) minp = qmin = mass / sumc;
  if (mass > maxp * sumc) maxp = qmax = mass / sumc;

  if (limiter_option == 8) {  // :766-823
    const int maxiter = NP * NP - 1;
    const double tol_limiter = 5e-14;
    for (int iter = 0; iter < maxiter; ++iter) {
      double addmass = 0.0;
#pragma unroll
      for (int k = 0; k < NPSQ; ++k) {
        double delta = 0.0;
        if (x[k] > maxp) { delta = x[k] - maxp; x[k] = maxp; }
        else if (x[k] < minp) { delta = x[k] - minp; x[k] = minp; }
        addmass += delta * c[k];
      }
      if (fabs(addmass) <= tol_limiter * fabs(mass)) break;
      if (addmass > 0) {
        double weightssum = 0.0;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) weightssum += (x[k] < maxp) ? c[k] : 0.0;
        const double adw = addmass / weightssum;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) x[k] += (x[k] < maxp) ? adw : 0.0;
      } else {
        double weightssum = 0.0;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) weightssum += (x[k] > minp) ? c[k] : 0.0;
        const double adw = addmass / weightssum;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) x[k] += (x[k] > minp) ? adw : 0.0;
      }
    }
  } else {  // limiter 9, :826-884
    double addmass = 0.0;
#pragma unroll
    for (int k = 0; k < NPSQ; ++k) {
      double delta = 0.0;
      if (x[k] > maxp) { delta = x[k] - maxp; x[k] = maxp; }
      else if (x[k] < minp) { delta = x[k] - minp; x[k] = minp; }
      addmass += delta * c[k];
    }
    if (addmass != 0) {
      if (addmass > 0) {
        double fac = 0.0;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) fac += c[k] * (maxp - x[k]);
        if (fac > 0) {
          fac = addmass / fac;
#pragma unroll
          for (int k = 0; k < NPSQ; ++k) x[k] += fac * (maxp - x[k]);
        }
      } else {
        double fac = 0.0;
#pragma unroll
        for (int k = 0; k < NPSQ; ++k) fac += c[k] * (x[k] - minp);
        if (fac > 0) {
          fac = addmass / fac;
#pragma unroll
          for (int k = 0; k < NPSQ; ++k) x[k] += fac * (x[k] - minp);
        }
      }
    }
  }
}

__device__ __forceinline__ bool limiter_level(int limiter_option, const double (&c)[NPSQ], double (&x)[NPSQ],
                                              double& qmin, double& qmax) {
  double sumc = c[0];
#pragma unroll
  for (int k = 1; k < NPSQ; ++k) sumc += c[k];
  if (sumc <= 0) return false;
  limiter_level_w(limiter_option, RegPlane{c}, sumc, x, qmin, qmax);
  return true;
}

}  // namespace hxx
