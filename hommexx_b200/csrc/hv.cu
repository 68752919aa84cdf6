// Hyperviscosity — replaces HyperviscosityFunctor{,Impl}.{hpp,cpp} of the reference:
// hypervis_subcycle x [ first weak Laplacian (TagFirstLaplaceHV, .hpp:72-87) -> DSS*rspheremp ->
// second Laplacian (TagSecondLaplaceConstHV/TensorHV :92-131) fused with TagHyperPreExchange
// (:161-257) -> DSS -> TagUpdateStates (:134-158) ].
// One thread per (element, level), the level's 4x4 planes in registers, no shared memory.
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

#include "hxx_sphere.cuh"

namespace hxx {

struct HvArgs {
  const double *geo, *metinv, *tensorvisc, *vec_sph2cart;
  double *v, *t, *dp3d, *vtens, *ttens, *dptens, *dpdiss_ave, *dpdiss_biharmonic;
  int nelem, np1;
  double dt, eta_ave_w, nu, nu_s, nu_p, nu_top, nu_ratio1, nu_ratio2;
  int hypervis_subcycle, consthv;
};

__device__ __forceinline__ bool map_thread(int nelem, int& ie, int& k) {
  // flat mapping: consecutive threads walk the levels of consecutive elements, so a block need
  // not hold whole elements and its size is free (4 warps = one per SM sub-partition)
  const long long g = (long long)blockIdx.x * TPB + threadIdx.x;
  ie = (int)(g / NLEV);
  k = (int)(g % NLEV);
  return ie < nelem;
}

// biharmonic_wk_dp3d first pass; interior points already carry the rspheremp of the DSS that follows
__global__ void __launch_bounds__(TPB, 2) hv_first_laplace_kernel(const HvArgs a) {
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  const double* __restrict__ g = a.geo + (size_t)ie * NPSQ * GEO_N;
  const double* __restrict__ mi = a.metinv + (size_t)ie * 4 * NPSQ;
  double s[NPSQ], lap[NPSQ];
  plane_load(a.t + off_s(ie, a.np1) + k, s);
  laplace_simple(g, s, lap);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p)
    if (is_interior_pt(p)) lap[p] *= geo_ld(g, p, G_RSPHEREMP);
  plane_store(a.ttens + off_f(ie) + k, lap);
  plane_load(a.dp3d + off_s(ie, a.np1) + k, s);
  laplace_simple(g, s, lap);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p)
    if (is_interior_pt(p)) lap[p] *= geo_ld(g, p, G_RSPHEREMP);
  plane_store(a.dptens + off_f(ie) + k, lap);
  double v1[NPSQ], l1[NPSQ];
  plane_load(a.v + off_v(ie, a.np1, 0) + k, s);
  plane_load(a.v + off_v(ie, a.np1, 1) + k, v1);
  vlaplace_sphere_wk_contra(g, mi, a.nu_ratio1, s, v1, lap, l1);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p)
    if (is_interior_pt(p)) {
      const double rs = geo_ld(g, p, G_RSPHEREMP);
      lap[p] *= rs;
      l1[p] *= rs;
    }
  plane_store(a.vtens + ((size_t)ie * 2 + 0) * NLF + k, lap);
  plane_store(a.vtens + ((size_t)ie * 2 + 1) * NLF + k, l1);
}

// second Laplacian + TagHyperPreExchange
__global__ void __launch_bounds__(TPB, 2) hv_second_laplace_pre_exchange_kernel(const HvArgs a) {
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  const double* __restrict__ g = a.geo + (size_t)ie * NPSQ * GEO_N;
  const double* __restrict__ mi = a.metinv + (size_t)ie * 4 * NPSQ;
  const double* __restrict__ tv = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const double* __restrict__ vs = a.consthv ? nullptr : a.vec_sph2cart + (size_t)ie * 6 * NPSQ;
  const double nst = (k == 0 ? 4.0 : k == 1 ? 2.0 : 1.0) * a.nu_top;  // HyperviscosityFunctorImpl.cpp:24-38
  const bool sponge = a.nu_top > 0 && k < 3;                           // NUM_BIHARMONIC_LEV
  double s[NPSQ], lap[NPSQ], top[NPSQ];
  {  // T
    double* tt = a.ttens + off_f(ie) + k;
    plane_load(tt, s);
    if (a.consthv) laplace_simple(g, s, lap); else laplace_tensor(g, tv, s, lap);
    if (sponge) {
      plane_load(a.t + off_s(ie, a.np1) + k, s);
      laplace_simple(g, s, top);
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      lap[p] *= -a.nu_s;
      if (sponge) lap[p] += nst * top[p];
    }
    plane_store(tt, lap);
  }
  {  // dp3d
    double* dt_ = a.dptens + off_f(ie) + k;
    double dp[NPSQ];
    plane_load(dt_, s);
    if (a.consthv) laplace_simple(g, s, lap); else laplace_tensor(g, tv, s, lap);
    plane_load(a.dp3d + off_s(ie, a.np1) + k, dp);
    double* dave = a.dpdiss_ave + off_f(ie) + k;
    double* dbih = a.dpdiss_biharmonic + off_f(ie) + k;
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      dave[p * NLEV] += a.eta_ave_w * dp[p] / a.hypervis_subcycle;
      dbih[p * NLEV] += a.eta_ave_w * lap[p] / a.hypervis_subcycle;
    }
    if (sponge) laplace_simple(g, dp, top);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      lap[p] *= -a.nu_p;
      if (sponge) lap[p] += nst * top[p];
      lap[p] *= a.dt;
      lap[p] += dp[p] * geo_ld(g, p, G_SPHEREMP);
    }
    plane_store(dt_, lap);
  }
  {  // v
    double* vt0 = a.vtens + ((size_t)ie * 2 + 0) * NLF + k;
    double* vt1 = a.vtens + ((size_t)ie * 2 + 1) * NLF + k;
    double s1[NPSQ], l1[NPSQ], top1[NPSQ];
    plane_load(vt0, s);
    plane_load(vt1, s1);
    if (a.consthv) vlaplace_sphere_wk_contra(g, mi, a.nu_ratio2, s, s1, lap, l1);
    else vlaplace_sphere_wk_cartesian(g, tv, vs, s, s1, lap, l1);
    if (sponge) {
      plane_load(a.v + off_v(ie, a.np1, 0) + k, s);
      plane_load(a.v + off_v(ie, a.np1, 1) + k, s1);
      vlaplace_sphere_wk_contra(g, mi, 1.0, s, s1, top, top1);
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      lap[p] *= -a.nu;
      l1[p] *= -a.nu;
      if (sponge) {
        lap[p] += nst * top[p];
        l1[p] += nst * top1[p];
      }
    }
    plane_store(vt0, lap);
    plane_store(vt1, l1);
  }
}

// TagUpdateStates .hpp:134-158
__global__ void hv_update_states_kernel(const HvArgs a) {
  const int ie = blockIdx.x;
  const double* __restrict__ g = a.geo + (size_t)ie * NPSQ * GEO_N;
  const double* vt0 = a.vtens + ((size_t)ie * 2 + 0) * NLF;
  const double* vt1 = a.vtens + ((size_t)ie * 2 + 1) * NLF;
  const double* tt = a.ttens + off_f(ie);
  const double* dpt = a.dptens + off_f(ie);
  double* v0 = a.v + off_v(ie, a.np1, 0);
  double* v1 = a.v + off_v(ie, a.np1, 1);
  double* t = a.t + off_s(ie, a.np1);
  double* dp = a.dp3d + off_s(ie, a.np1);
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) {
    const int p = i / NLEV;
    const double rs = geo_ld(g, p, G_RSPHEREMP);
    const double a0 = a.dt * vt0[i] * rs, a1 = a.dt * vt1[i] * rs;
    const double n0 = v0[i] + a0, n1 = v1[i] + a1;
    v0[i] = n0;
    v1[i] = n1;
    const double th = a.dt * tt[i] * rs;
    const double heating = a0 * n0 + a1 * n1;
    t[i] = t[i] + th - heating / cp;
    dp[i] = dpt[i] * rs;
  }
}

void hypervis_run(int np1, double dt_in, double eta_ave_w) {
  if (!S.nelemd) return;
  const Params& p = S.p;
  HvArgs a{S.geo, S.metinv, S.tensorvisc, S.vec_sph2cart, S.v, S.t, S.dp3d, S.vtens, S.ttens, S.dptens,
           S.dpdiss_ave, S.dpdiss_biharmonic, S.nelemd, np1, dt_in / p.hypervis_subcycle, eta_ave_w, p.nu, p.nu_s,
           p.nu_p, p.nu_top, p.nu_ratio1, p.nu_ratio2, p.hypervis_subcycle, p.consthv ? 1 : 0};
  const int nb = nblocks_flat(S.nelemd);
  for (int icycle = 0; icycle < p.hypervis_subcycle; ++icycle) {
    PROBE(K_HV_FIRST);
    hv_first_laplace_kernel<<<nb, TPB, 0, S.stream>>>(a);
    KERNEL_LAUNCHED(K_HV_FIRST);
    dss_exchange(fields_hv(), true);
    PROBE(K_HV_SECOND);
    hv_second_laplace_pre_exchange_kernel<<<nb, TPB, 0, S.stream>>>(a);
    KERNEL_LAUNCHED(K_HV_SECOND);
    dss_exchange(fields_hv(), false);
    PROBE(K_HV_UPDATE);
    hv_update_states_kernel<<<S.nelemd, 288, 0, S.stream>>>(a);
    KERNEL_LAUNCHED(K_HV_UPDATE);
  }
}

}  // namespace hxx
