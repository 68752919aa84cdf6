// Hyperviscosity — replaces HyperviscosityFunctor{,Impl}.{hpp,cpp} of the reference:
// hypervis_subcycle x [ first weak Laplacian (TagFirstLaplaceHV, .hpp:72-87) -> DSS*rspheremp ->
// second Laplacian (TagSecondLaplaceConstHV/TensorHV :92-131) fused with TagHyperPreExchange
// (:161-257) -> DSS -> TagUpdateStates (:134-158) ].
// One thread per (element, level), the level's 4x4 planes in registers; operators are finished one point
// at a time (hxx_sphere.cuh), the block's geometry / metinv records and the vector Laplacian's parked
// weak gradient live in shared memory.
#include <type_traits>

#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

#include "hxx_sphere.cuh"

namespace hxx {

#ifndef HXX_HV_MINB
#define HXX_HV_MINB 2
#endif
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

// Cooperative copy of a per-element [N]-double record (metinv: 64) for the elements a block touches.
template <int NE, int N, int NT>
__device__ __forceinline__ void stage_records(double* s_dst, const double* __restrict__ src, int e_first, int nelem) {
  for (int i = threadIdx.x; i < NE * N; i += NT) {
    const int e = e_first + i / N;
    if (e < nelem) s_dst[i] = __ldg(src + (size_t)e_first * N + i);
  }
}

// All three kernels below finish their operators one point at a time (hxx_sphere.cuh): a result goes to
// HBM as soon as it exists, every plane a kernel reads is in registers before its first store, and the
// weak gradient inside the vector Laplacian waits in per-thread shared-memory slots [32][TPB].

// biharmonic_wk_dp3d first pass; interior points already carry the rspheremp of the DSS that follows
__global__ void __launch_bounds__(TPB, HXX_HV_MINB) hv_first_laplace_kernel(const HvArgs a) {
  extern __shared__ double s_park[];
  __shared__ double s_geo[geo_span(TPB) * NPSQ * GEO_N];
  __shared__ double s_mi[geo_span(TPB) * 4 * NPSQ];
  const int e_first = (int)(((long long)blockIdx.x * TPB) / NLEV);
  stage_records<geo_span(TPB), 4 * NPSQ, TPB>(s_mi, a.metinv, e_first, a.nelem);
  stage_geo<geo_span(TPB), TPB>(s_geo, a.geo, e_first, a.nelem);
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const GeoShared mi{s_mi + (ie - e_first) * 4 * NPSQ};
  {
    double s[NPSQ];
    plane_load(a.t + off_s(ie, a.np1) + k, s);
    double* out = a.ttens + off_f(ie) + k;
    laplace_points<false>(g, nullptr, s, [&](int p, double lap) {
      if (is_interior_pt(p)) lap *= geo_ld(g, p, G_RSPHEREMP);
      out[p * NLEV] = lap;
    });
  }
  phase_fence();
  {
    double s[NPSQ];
    plane_load(a.dp3d + off_s(ie, a.np1) + k, s);
    double* out = a.dptens + off_f(ie) + k;
    laplace_points<false>(g, nullptr, s, [&](int p, double lap) {
      if (is_interior_pt(p)) lap *= geo_ld(g, p, G_RSPHEREMP);
      out[p * NLEV] = lap;
    });
  }
  phase_fence();
  double* o0 = a.vtens + ((size_t)ie * 2 + 0) * NLF + k;
  double* o1 = a.vtens + ((size_t)ie * 2 + 1) * NLF + k;
  vlaplace_contra_points(g, mi, a.nu_ratio1, a.v + off_v(ie, a.np1, 0) + k, a.v + off_v(ie, a.np1, 1) + k,
                         s_park + threadIdx.x, s_park + NPSQ * TPB + threadIdx.x, TPB, [&](int p, double l0, double l1) {
                           if (is_interior_pt(p)) {
                             const double rs = geo_ld(g, p, G_RSPHEREMP);
                             l0 *= rs;
                             l1 *= rs;
                           }
                           o0[p * NLEV] = l0;
                           o1[p * NLEV] = l1;
                         });
}

// second Laplacian + TagHyperPreExchange. The scalar fields (T, dp3d) and the vector field are separate
// kernels, and the nu_top sponge layer — extra Laplacians of the state on levels 0..2 only
// (NUM_BIHARMONIC_LEV) — is a template flag: the SPONGE=true instantiation runs over just those
// three levels of every element, the SPONGE=false one over the remaining levels, so 69 of 72
// levels run code that never allocates registers for the sponge terms.
constexpr int HV2_SPAN = (TPB + (NLEV > 3 ? NLEV - 3 : 1) - 2) / (NLEV > 3 ? NLEV - 3 : 1) + 1;

template <bool SPONGE>
__device__ __forceinline__ bool map_thread_hv2(int nelem, int nsponge, int& ie, int& k) {
  const long long g = (long long)blockIdx.x * TPB + threadIdx.x;
  if (SPONGE) {
    ie = (int)(g / nsponge);
    k = (int)(g % nsponge);
  } else {
    const int nl = NLEV - nsponge;
    ie = (int)(g / nl);
    k = (int)(g % nl) + nsponge;
  }
  return ie < nelem;
}

template <bool SPONGE>
__global__ void __launch_bounds__(TPB, SPONGE ? 2 : 3) hv_second_scalar_kernel(const HvArgs a, int nsponge) {
  // the main (non-sponge) instantiation walks >= NLEV - 3 levels per element, so a block spans
  // at most HV2_SPAN elements and reads their geometry from shared memory
  __shared__ double s_geo[SPONGE ? 1 : HV2_SPAN * NPSQ * GEO_N];
  const int e_first = SPONGE ? 0 : (int)(((long long)blockIdx.x * TPB) / (NLEV - nsponge));
  if (!SPONGE) stage_geo<HV2_SPAN, TPB>(s_geo, a.geo, e_first, a.nelem);
  int ie, k;
  if (!map_thread_hv2<SPONGE>(a.nelem, nsponge, ie, k)) return;
  using Geo = typename std::conditional<SPONGE, GeoGlobal, GeoShared>::type;
  const Geo g{SPONGE ? a.geo + (size_t)ie * NPSQ * GEO_N : s_geo + (ie - e_first) * NPSQ * GEO_N};
  const double* __restrict__ tv = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const double nst = (k == 0 ? 4.0 : k == 1 ? 2.0 : 1.0) * a.nu_top;  // HyperviscosityFunctorImpl.cpp:24-38
  {  // T
    double* tt = a.ttens + off_f(ie) + k;
    double s[NPSQ], top[SPONGE ? NPSQ : 1];
    plane_load(tt, s);
    if constexpr (SPONGE) {
      double t[NPSQ];
      plane_load(a.t + off_s(ie, a.np1) + k, t);
      laplace_simple(g, t, top);
    }
    auto emit = [&](int p, double lap) {
      lap *= -a.nu_s;
      if constexpr (SPONGE) lap += nst * top[p];
      tt[p * NLEV] = lap;
    };
    if (a.consthv) laplace_points<false>(g, tv, s, emit); else laplace_points<true>(g, tv, s, emit);
  }
  phase_fence();
  {  // dp3d
    double* dt_ = a.dptens + off_f(ie) + k;
    double* dbih = a.dpdiss_biharmonic + off_f(ie) + k;
    double dp[NPSQ], top[SPONGE ? NPSQ : 1];
    plane_load(a.dp3d + off_s(ie, a.np1) + k, dp);
    const double hs = (double)a.hypervis_subcycle, rhs = 1.0 / hs;  // the subcycle count divides through its reciprocal
    {
      double* dave = a.dpdiss_ave + off_f(ie) + k;
      double r0[NPSQ];
      plane_load(dave, r0);
      double t[NPSQ];
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) t[p] = a.eta_ave_w * dp[p];
      div_rcp_plane(t, [&](int) { return hs; }, [&](int) { return rhs; });
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) r0[p] += t[p];
      plane_store(dave, r0);
    }
    if constexpr (SPONGE) laplace_simple(g, dp, top);
    double s[NPSQ], r1[NPSQ];
    plane_load(dt_, s);
    plane_load(dbih, r1);
    phase_fence();
    auto emit = [&](int p, double lap) {
      dbih[p * NLEV] = r1[p] + div_rcp(a.eta_ave_w * lap, hs, rhs);
      lap *= -a.nu_p;
      if constexpr (SPONGE) lap += nst * top[p];
      lap *= a.dt;
      lap += dp[p] * geo_ld(g, p, G_SPHEREMP);
      dt_[p * NLEV] = lap;
    };
    if (a.consthv) laplace_points<false>(g, tv, s, emit); else laplace_points<true>(g, tv, s, emit);
  }
}

template <bool SPONGE, bool CONSTHV>
__global__ void __launch_bounds__(TPB, SPONGE ? 2 : HXX_HV_MINB) hv_second_vector_kernel(const HvArgs a, int nsponge) {
  extern __shared__ double s_park[];
  __shared__ double s_geo[SPONGE ? 1 : HV2_SPAN * NPSQ * GEO_N];
  __shared__ double s_mi[SPONGE ? 1 : HV2_SPAN * 4 * NPSQ];
  const int e_first = SPONGE ? 0 : (int)(((long long)blockIdx.x * TPB) / (NLEV - nsponge));
  if (!SPONGE) {
    stage_records<HV2_SPAN, 4 * NPSQ, TPB>(s_mi, a.metinv, e_first, a.nelem);
    stage_geo<HV2_SPAN, TPB>(s_geo, a.geo, e_first, a.nelem);
  }
  int ie, k;
  if (!map_thread_hv2<SPONGE>(a.nelem, nsponge, ie, k)) return;
  using Geo = typename std::conditional<SPONGE, GeoGlobal, GeoShared>::type;
  const Geo g{SPONGE ? a.geo + (size_t)ie * NPSQ * GEO_N : s_geo + (ie - e_first) * NPSQ * GEO_N};
  const Geo mig{SPONGE ? a.metinv + (size_t)ie * 4 * NPSQ : s_mi + (ie - e_first) * 4 * NPSQ};
  const double* __restrict__ mi = a.metinv + (size_t)ie * 4 * NPSQ;
  const double* __restrict__ tv = CONSTHV ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const double* __restrict__ vs = CONSTHV ? nullptr : a.vec_sph2cart + (size_t)ie * 6 * NPSQ;
  const double nst = (k == 0 ? 4.0 : k == 1 ? 2.0 : 1.0) * a.nu_top;
  double* vt0 = a.vtens + ((size_t)ie * 2 + 0) * NLF + k;
  double* vt1 = a.vtens + ((size_t)ie * 2 + 1) * NLF + k;
  if constexpr (CONSTHV) {
    double* const park0 = s_park + threadIdx.x;
    double* const park1 = s_park + NPSQ * TPB + threadIdx.x;
    double top0[SPONGE ? NPSQ : 1], top1[SPONGE ? NPSQ : 1];
    if constexpr (SPONGE) {
      vlaplace_contra_points(g, mig, 1.0, a.v + off_v(ie, a.np1, 0) + k, a.v + off_v(ie, a.np1, 1) + k, park0, park1,
                             TPB, [&](int p, double l0, double l1) { top0[p] = l0; top1[p] = l1; });
      phase_fence();
    }
    // in place: the operator's last reads of vtens are whole planes taken before its first emit
    vlaplace_contra_points(g, mig, a.nu_ratio2, vt0, vt1, park0, park1, TPB, [&](int p, double l0, double l1) {
      l0 *= -a.nu;
      l1 *= -a.nu;
      if constexpr (SPONGE) {
        l0 += nst * top0[p];
        l1 += nst * top1[p];
      }
      vt0[p * NLEV] = l0;
      vt1[p * NLEV] = l1;
    });
  } else {
  double lap[NPSQ], l1[NPSQ];
  if constexpr (CONSTHV) vlaplace_sphere_wk_contra_mem(g, mi, a.nu_ratio2, vt0, vt1, lap, l1);
  else {
    double s[NPSQ], s1[NPSQ];
    plane_load(vt0, s);
    plane_load(vt1, s1);
    vlaplace_sphere_wk_cartesian(g, tv, vs, s, s1, lap, l1);
  }
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    lap[p] *= -a.nu;
    l1[p] *= -a.nu;
  }
  if (SPONGE) {
    double top[NPSQ], top1[NPSQ];
    vlaplace_sphere_wk_contra_mem(g, mi, 1.0, a.v + off_v(ie, a.np1, 0) + k, a.v + off_v(ie, a.np1, 1) + k, top, top1);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
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
  constexpr size_t park_bytes = 2 * (size_t)NPSQ * TPB * sizeof(double);  // two parked planes per thread
  if (HXX_ONCE_PER_SESSION()) {  // static + dynamic shared memory passes 48 KB at small NLEV (more elements per block)
    const int pb = (int)park_bytes;
    CUDA_OK(cudaFuncSetAttribute(hv_first_laplace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pb));
    CUDA_OK(cudaFuncSetAttribute(hv_second_vector_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb));
    CUDA_OK(cudaFuncSetAttribute(hv_second_vector_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb));
    CUDA_OK(cudaFuncSetAttribute(hv_second_vector_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pb));
  }
  for (int icycle = 0; icycle < p.hypervis_subcycle; ++icycle) {
    {
      HXX_TIMER("hvf-bhwk");
      PROBE(K_HV_FIRST);
      hv_first_laplace_kernel<<<nb, TPB, park_bytes, S.stream>>>(a);
      KERNEL_LAUNCHED(K_HV_FIRST);
    }
    {
      HXX_TIMER("hvf-bexch");
      dss_exchange(fields_hv(), true);
    }
    {
      const int nsp = p.nu_top > 0 ? (NLEV < 3 ? NLEV : 3) : 0;  // NUM_BIHARMONIC_LEV
      const int nb_main = (int)(((long long)S.nelemd * (NLEV - nsp) + TPB - 1) / TPB);
      const int nb_sp = (int)(((long long)S.nelemd * nsp + TPB - 1) / TPB);
      PROBE(K_HV_SECOND);
      if (nb_main) hv_second_scalar_kernel<false><<<nb_main, TPB, 0, S.stream>>>(a, nsp);
      if (nb_main) {
        if (p.consthv) hv_second_vector_kernel<false, true><<<nb_main, TPB, park_bytes, S.stream>>>(a, nsp);
        else hv_second_vector_kernel<false, false><<<nb_main, TPB, park_bytes, S.stream>>>(a, nsp);
      }
      if (nb_sp) hv_second_scalar_kernel<true><<<nb_sp, TPB, 0, S.stream>>>(a, nsp);
      if (nb_sp) {
        if (p.consthv) hv_second_vector_kernel<true, true><<<nb_sp, TPB, park_bytes, S.stream>>>(a, nsp);
        else hv_second_vector_kernel<true, false><<<nb_sp, TPB, 0, S.stream>>>(a, nsp);
      }
      KERNEL_LAUNCHED(K_HV_SECOND);
      S.launches += 3;
    }
    {
      HXX_TIMER("hvf-bexch");
      dss_exchange(fields_hv(), false);
    }
    PROBE(K_HV_UPDATE);
    hv_update_states_kernel<<<S.nelemd, 288, 0, S.stream>>>(a);
    KERNEL_LAUNCHED(K_HV_UPDATE);
  }
}

}  // namespace hxx
