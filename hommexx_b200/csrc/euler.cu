// Tracer advection — replaces EulerStepFunctor{,Impl}.hpp of the reference: one stage of the
// 3-stage SSP-RK2 scheme per euler_step() call (EulerStepFunctorImpl.hpp:514-561).
//
// One thread per (element, level) walks a chunk of tracers; the level's 4x4 plane is in
// registers, so the divergence, the weak Laplacians, the per-level min/max AND the
// quasi-monotone limiter (whose reductions run over the 16 points of one level) are all
// thread-local — no team reductions, no shared-memory exchange. The reference's per-element
// set-up kernels (compute_dp :406-434, compute_2d_advection_step :585-626) are recomputed in
// registers by each tracer chunk instead of round-tripping vstar/dpdissk/dp_star through HBM,
// and the second biharmonic Laplacian (:216-231) is applied on the fly by the advection kernel.
// Algorithmic HBM traffic per (element, tracer, stage): min/max pass 1 tile read; advection
// 1 read + 1 write (+1 read +1 write of qtens_biharmonic on the stage with hyperviscosity).
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

#include "hxx_limiter.cuh"
#include "hxx_sphere.cuh"

namespace hxx {

struct EulerArgs {
  const double* geo;
  const double* tensorvisc;
  double *qdp, *qtens_biharmonic, *qlim;
  const double *derived_dp, *divdp_proj, *divdp, *derived_vn0, *dpdiss_ave, *dpdiss_biharmonic;
  double* f_dss;
  int nelem, qsize, qchunk, n0_qdp, np1_qdp;
  double dt, rhsmdt, nu_p, nu_q, rhs_viss;
  int rhs_mode;  // 0, 1, 2 = rhs_multiplier
  int tavg_n0;   // >= 0: fuse qdp_time_avg with this qdp time level (interior points here, boundary in the DSS)
  int limiter_option, consthv;
};

__device__ __forceinline__ bool map_thread(int nelem, int& ie, int& k) {
  // flat mapping: consecutive threads walk the levels of consecutive elements, so a block need
  // not hold whole elements and its size is free (4 warps = one per SM sub-partition)
  const long long g = (long long)blockIdx.x * TPB + threadIdx.x;
  ie = (int)(g / NLEV);
  k = (int)(g % NLEV);
  return ie < nelem;
}

// precompute_divdp :348-377
__global__ void __launch_bounds__(TPB, 2)
    euler_divdp_kernel(const double* __restrict__ geo, const double* __restrict__ vn0, double* __restrict__ divdp,
                       double* __restrict__ divdp_proj, int nelem) {
  int ie, k;
  if (!map_thread(nelem, ie, k)) return;
  const double* g = geo + (size_t)ie * NPSQ * GEO_N;
  double v0[NPSQ], v1[NPSQ], div[NPSQ];
  plane_load(vn0 + ((size_t)ie * 2 + 0) * NLF + k, v0);
  plane_load(vn0 + ((size_t)ie * 2 + 1) * NLF + k, v1);
  divergence_sphere(g, v0, v1, div);
  plane_store(divdp + off_f(ie) + k, div);
  plane_store(divdp_proj + off_f(ie) + k, div);
}

// compute_dp + compute_qmin_qmax (:406-485) and, on the hyperviscosity stage,
// compute_biharmonic_pre (:196-214, dpdiss_adjustment :251-267): Q -> laplace(Q * dpdiss_ave / dp0).
// The tracer planes are staged two tracers ahead with cp.async, as in the advection kernel.
template <bool BIH>
__global__ void __launch_bounds__(TPB, BIH ? 2 : 4) euler_qminmax_kernel(const EulerArgs a) {
  extern __shared__ double s_all[];
  __shared__ double s_geo[BIH ? geo_span(TPB) * NPSQ * GEO_N : 1];
  const int e_first = (int)(((long long)blockIdx.x * TPB) / NLEV);
  if (BIH) stage_geo<geo_span(TPB), TPB>(s_geo, a.geo, e_first, a.nelem);
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  double* const s_q = s_all + threadIdx.x;  // [2][16][TPB]
  const GeoShared g{s_geo + (BIH ? (ie - e_first) * NPSQ * GEO_N : 0)};
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  const double* const qin = a.qdp + off_q(ie, a.n0_qdp, 0) + k;
  auto prefetch = [&](int q, int buf) {
    if (q < q1) {
      const double* src = qin + (size_t)q * NLF;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) cp_async8(s_q + (buf * NPSQ + p) * TPB, src + p * NLEV);
    }
    cp_async_commit();
  };
  prefetch(q0, 0);
  prefetch(q0 + 1, 1);
  // dp of compute_dp, its reciprocal and (hyperviscosity stage) dpdiss_ave are the same for every
  // tracer: the BIH instantiation parks them in per-thread shared-memory slots [32 + 16][TPB] behind
  // the staging buffers so that its Laplacian has the registers, the other one keeps them in registers
  double* const s_dp = s_q + 2 * NPSQ * TPB;
  double* const s_rdp = s_dp + NPSQ * TPB;
  double* const s_dave = s_rdp + NPSQ * TPB;
  double dps[BIH ? 1 : NPSQ];
  const bool scale = BIH && a.nu_p > 0;
  {
    double r0[NPSQ], r1[NPSQ];
    plane_load(a.derived_dp + off_f(ie) + k, r0);
    plane_load(a.divdp_proj + off_f(ie) + k, r1);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const double d = r0[p] - a.rhsmdt * r1[p];
      if constexpr (BIH) { s_dp[p * TPB] = d; s_rdp[p * TPB] = 1.0 / d; }
      else dps[p] = d;
    }
    if (scale) {
      plane_load(a.dpdiss_ave + off_f(ie) + k, r0);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) s_dave[p * TPB] = r0[p];
    }
  }
  const double dp0k = dc.dp0[k], rdp0k = 1.0 / dp0k;
  double* ql = a.qlim + ((size_t)ie * QSIZE_D + q0) * 2 * NLEV + k;
  double* qtb = a.qtens_biharmonic + ((size_t)ie * QSIZE_D + q0) * NLF + k;
  double mn_n = 0.0, mx_n = 0.0;
  if (a.rhs_mode == 1) { mn_n = ql[0]; mx_n = ql[NLEV]; }
  for (int q = q0; q < q1; ++q, ql += 2 * NLEV, qtb += NLF) {
    const int buf = (q - q0) & 1;
    double mn = mn_n, mx = mx_n;
    if (a.rhs_mode == 1 && q + 1 < q1) { mn_n = ql[2 * NLEV]; mx_n = ql[3 * NLEV]; }
    cp_async_wait<1>();
    double Q[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) Q[p] = s_q[(buf * NPSQ + p) * TPB];
    prefetch(q + 2, buf);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      if constexpr (BIH) Q[p] = div_rcp(Q[p], s_dp[p * TPB], s_rdp[p * TPB]);
      else Q[p] = Q[p] / dps[p];
    }
    if (a.rhs_mode != 1) { mn = Q[0]; mx = Q[0]; }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) { mn = fmin(mn, Q[p]); mx = fmax(mx, Q[p]); }
    ql[0] = mn;
    ql[NLEV] = mx;
    if (BIH) {
      double lap[NPSQ];
      if (scale) {
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) Q[p] = div_rcp(Q[p] * s_dave[p * TPB], dp0k, rdp0k);
      }
      laplace_simple(g, Q, lap);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p)
        if (is_interior_pt(p)) lap[p] *= geo_ld(g, p, G_RSPHEREMP);  // rspheremp of the DSS that follows
      plane_store(qtb, lap);
    }
  }
  cp_async_wait<0>();
}

// shared-memory doubles per thread: vstar (2x16), dpdissk (16), 2 staged qdp planes (2x16) and,
// on the hyperviscosity stage, 2 staged qtens_biharmonic planes (2x16); slot s of thread t lives
// at [s][t], so a warp's access to one slot is 256 contiguous bytes (conflict-free).
#ifndef HXX_ADV_STAGES
#define HXX_ADV_STAGES 1
#endif
#ifndef HXX_ADV_MINB
#define HXX_ADV_MINB 3
#endif
#ifndef HXX_ADV_MINB_HV
#define HXX_ADV_MINB_HV 2
#endif
constexpr int ADV_NST = HXX_ADV_STAGES;  // staged tracers in flight per thread
// HV mode of the advection kernel: 0 = no hyperviscosity term; 1 = second Laplacian applied on
// the fly (one pass less over qtens_biharmonic, but the heaviest register footprint); 2 = the term
// was prepared in place by euler_hvpost_kernel and is only added here.
// shared-memory slots per thread: vstar (2x16) and dpdissk (16), then per staged tracer the qdp
// plane (16), the qtens_biharmonic / prepared-term plane when HV != 0 (16), the two qlim rows and
// the four interior time-average partners when TAVG
template <int HV, bool TAVG>
__host__ __device__ constexpr int advect_stage_slots() { return NPSQ + (HV ? NPSQ : 0) + 2 + (TAVG ? 4 : 0); }
template <int HV, bool TAVG>
__host__ __device__ constexpr int advect_slots() { return 48 + ADV_NST * advect_stage_slots<HV, TAVG>(); }
#ifndef HXX_ADV_MINB_HV2
#define HXX_ADV_MINB_HV2 2
#endif

// compute_biharmonic_post :216-231 with rhsviss_adjustment :293-310, in place:
// qtens_biharmonic <- (-rhs_viss dt nu_q) dp0 laplace(qtens_biharmonic) / spheremp
__global__ void __launch_bounds__(TPB, 3) euler_hvpost_kernel(const EulerArgs a) {
  extern __shared__ double s_all[];
  __shared__ double s_geo[geo_span(TPB) * NPSQ * GEO_N];
  const int e_first = (int)(((long long)blockIdx.x * TPB) / NLEV);
  stage_geo<geo_span(TPB), TPB>(s_geo, a.geo, e_first, a.nelem);
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  double* const s_q = s_all + threadIdx.x;  // [2][16][TPB]
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const double* __restrict__ tv = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  double* const qtb = a.qtens_biharmonic + (size_t)ie * QSIZE_D * NLF + k;
  auto prefetch = [&](int q, int buf) {
    if (q < q1) {
      const double* src = qtb + (size_t)q * NLF;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) cp_async8(s_q + (buf * NPSQ + p) * TPB, src + p * NLEV);
    }
    cp_async_commit();
  };
  prefetch(q0, 0);
  prefetch(q0 + 1, 1);
  const double dp0k = dc.dp0[k];
  const double bfac = -a.rhs_viss * a.dt * a.nu_q;
  for (int q = q0; q < q1; ++q) {
    const int buf = (q - q0) & 1;
    cp_async_wait<1>();
    double s[NPSQ], lap[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) s[p] = s_q[(buf * NPSQ + p) * TPB];
    prefetch(q + 2, buf);
    if (a.consthv) laplace_simple(g, s, lap); else laplace_tensor(g, tv, s, lap);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p)
      lap[p] = div_rcp(bfac * dp0k * lap[p], geo_ld(g, p, G_SPHEREMP), geo_ld(g, p, G_INV_SPHEREMP));
    plane_store(qtb + (size_t)q * NLF, lap);
  }
  cp_async_wait<0>();
}

// advect_and_limit :317-332 = compute_2d_advection_step (:585-626) + run_tracer_phase (:571-582),
// with compute_biharmonic_post (:216-231, rhsviss_adjustment :293-310) applied on the fly.
template <int HV, bool TAVG>
__global__ void __launch_bounds__(TPB, HV == 1 ? HXX_ADV_MINB_HV : HV == 2 ? HXX_ADV_MINB_HV2 : HXX_ADV_MINB)
    euler_advect_kernel(const EulerArgs a) {
  extern __shared__ double s_all[];
  __shared__ double s_geo[geo_span(TPB) * NPSQ * GEO_N];
  const int e_first = (int)(((long long)blockIdx.x * TPB) / NLEV);
  stage_geo<geo_span(TPB), TPB>(s_geo, a.geo, e_first, a.nelem);
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;  // no block-wide barrier below: early exit is safe
  const int tid = threadIdx.x;
  double* const s_vs0 = s_all + tid;
  double* const s_vs1 = s_all + 16 * TPB + tid;
  double* const s_dpk = s_all + 32 * TPB + tid;
  constexpr int SS = advect_stage_slots<HV, TAVG>();
  double* const s_q = s_all + 48 * TPB + tid;               // stage i: slots [i*SS, (i+1)*SS)
  double* const s_b = s_q + NPSQ * TPB;                     // second plane (HV != 0)
  double* const s_l = s_q + (HV ? 2 : 1) * NPSQ * TPB;      // qlim rows
  double* const s_a = s_l + 2 * TPB;                        // time-average partners (TAVG)
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const double* __restrict__ tv = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  const double* const qin = a.qdp + off_q(ie, a.n0_qdp, 0) + k;
  const double* const qtb = a.qtens_biharmonic + (size_t)ie * QSIZE_D * NLF + k;
  const double* const qlim_in = a.qlim + (size_t)ie * QSIZE_D * 2 * NLEV + k;
  const double* const qavg = TAVG ? a.qdp + off_q(ie, a.tavg_n0, 0) + k : nullptr;
  auto prefetch = [&](int q, int buf) {
    if (q < q1) {
      const int o = buf * SS * TPB;
      const double* src = qin + (size_t)q * NLF;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) cp_async8(s_q + o + p * TPB, src + p * NLEV);
      if (HV) {
        const double* sb = qtb + (size_t)q * NLF;
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) cp_async8(s_b + o + p * TPB, sb + p * NLEV);
      }
      cp_async8(s_l + o, qlim_in + (size_t)q * 2 * NLEV);
      cp_async8(s_l + o + TPB, qlim_in + (size_t)q * 2 * NLEV + NLEV);
      if (TAVG) {  // qdp_time_avg :379-403 partner values of the 4 interior points
        const double* pa = qavg + (size_t)q * NLF;
        cp_async8(s_a + o, pa + 5 * NLEV);
        cp_async8(s_a + o + TPB, pa + 6 * NLEV);
        cp_async8(s_a + o + 2 * TPB, pa + 9 * NLEV);
        cp_async8(s_a + o + 3 * TPB, pa + 10 * NLEV);
      }
    }
    cp_async_commit();  // possibly empty: keeps one group per loop iteration
  };
  HXX_UNROLL
  for (int i = 0; i < ADV_NST; ++i) prefetch(q0 + i, i);

  const bool add_ps_diss = a.nu_p > 0 && HV != 0;
  const double diss_fac = add_ps_diss ? -a.rhs_viss * a.dt * a.nu_q : 0.0;
  double c[NPSQ];
  {
    const double* dd = a.derived_dp + off_f(ie) + k;
    const double* dj = a.divdp_proj + off_f(ie) + k;
    const double* dv = a.divdp + off_f(ie) + k;
    const double* n0 = a.derived_vn0 + ((size_t)ie * 2 + 0) * NLF + k;
    const double* n1 = a.derived_vn0 + ((size_t)ie * 2 + 1) * NLF + k;
    const double* db = a.dpdiss_biharmonic + off_f(ie) + k;
    double r0[NPSQ], r1[NPSQ], r2[NPSQ], r3[NPSQ], r4[NPSQ], r5[NPSQ];
    plane_load(dd, r0);
    plane_load(dj, r1);
    plane_load(dv, r2);
    plane_load(n0, r3);
    plane_load(n1, r4);
    if (add_ps_diss) plane_load(db, r5);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const double sm_ = geo_ld(g, p, G_SPHEREMP);
      const double dp = r0[p] - a.rhsmdt * r1[p];
      s_vs0[p * TPB] = r3[p] / dp;
      s_vs1[p * TPB] = r4[p] / dp;
      double d = dp - a.dt * r2[p];
      if (add_ps_diss) d += diss_fac * r5[p] / sm_;
      s_dpk[p * TPB] = d;
      c[p] = sm_ * d;
    }
  }
  if (blockIdx.y == 0 && a.f_dss) {  // f_dss *= spheremp (and the interior part of the DSS rspheremp)
    double* f = a.f_dss + off_f(ie) + k;
    double r[NPSQ];
    plane_load(f, r);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      r[p] = r[p] * geo_ld(g, p, G_SPHEREMP);
      if (is_interior_pt(p)) r[p] *= geo_ld(g, p, G_RSPHEREMP);
    }
    plane_store(f, r);
  }
  // the limiter's weight sum does not depend on the tracer (serial order k = 0..15, as the reference)
  double sumc = c[0];
  HXX_UNROLL
  for (int p = 1; p < NPSQ; ++p) sumc += c[p];
  const bool skip = sumc <= 0;
  const double dp0k = dc.dp0[k];
  const double bfac = -a.rhs_viss * a.dt * a.nu_q;
  const double alpha = -a.dt;
  double* qlp = a.qlim + ((size_t)ie * QSIZE_D + q0) * 2 * NLEV + k;
  double* out = a.qdp + off_q(ie, a.np1_qdp, q0) + k;
  for (int q = q0; q < q1; ++q, qlp += 2 * NLEV, out += NLF) {
    const int buf = (q - q0) % ADV_NST;
    const int o = buf * SS * TPB;
    cp_async_wait<ADV_NST - 1>();  // this thread's copies of tracer q have landed
    const double qmin0 = s_l[o], qmax0 = s_l[o + TPB];
    double qa[4] = {0.0, 0.0, 0.0, 0.0};
    if (TAVG) { qa[0] = s_a[o]; qa[1] = s_a[o + TPB]; qa[2] = s_a[o + 2 * TPB]; qa[3] = s_a[o + 3 * TPB]; }
    // The qdp plane becomes the advected value in place, one point at a time: with the limiter's
    // weights that makes four live planes (c, x, gv0, gv1) at the widest spot.
    double x[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) x[p] = s_q[o + p * TPB];
    {
      // divergence_sphere_update, SphereOperators.hpp:398-444
      double gv0[NPSQ], gv1[NPSQ];
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        const double u = s_vs0[p * TPB] * x[p];
        const double v = s_vs1[p * TPB] * x[p];
        const double md = geo_ld(g, p, G_METDET);
        gv0[p] = (geo_ld(g, p, G_DINV00) * u + geo_ld(g, p, G_DINV10) * v) * md;
        gv1[p] = (geo_ld(g, p, G_DINV01) * u + geo_ld(g, p, G_DINV11) * v) * md;
      }
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        double dx, dy;
        deriv_point(gv0, gv1, p / NP, p % NP, dx, dy);
        x[p] = x[p] + alpha * ((dx + dy) * geo_ld(g, p, G_RMETDET_R));
        if (HV == 2) x[p] += s_b[o + p * TPB];  // the prepared hyperviscosity term
      }
    }
    if (HV == 1) {
      // x is parked in the (already consumed) qdp staging slot while the Laplacian needs registers
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) s_q[o + p * TPB] = x[p];
      double s[NPSQ], lap[NPSQ];
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) s[p] = s_b[o + p * TPB];
      if (a.consthv) laplace_simple(g, s, lap); else laplace_tensor(g, tv, s, lap);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p)
        x[p] = s_q[o + p * TPB] + div_rcp(bfac * dp0k * lap[p], geo_ld(g, p, G_SPHEREMP), geo_ld(g, p, G_INV_SPHEREMP));
    }
    prefetch(q + ADV_NST, buf);  // the staged planes of tracer q are in registers now: refill the slot
    // limiter shell :693-761; a level whose weights do not sum to a positive number is left alone
    if (!skip) {
      double qmin = qmin0, qmax = qmax0;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = x[p] / s_dpk[p * TPB];
      limiter_level_w(a.limiter_option, c, sumc, x, qmin, qmax);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = x[p] * s_dpk[p * TPB];
      if (qmin != qmin0) qlp[0] = qmin;
      if (qmax != qmax0) qlp[NLEV] = qmax;
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {  // apply_spheremp :672-687
      double r = geo_ld(g, p, G_SPHEREMP) * x[p];
      if (is_interior_pt(p)) {
        r *= geo_ld(g, p, G_RSPHEREMP);
        if (TAVG) r = (qa[p == 5 ? 0 : p == 6 ? 1 : p == 9 ? 2 : 3] + 2.0 * r) / 3.0;
      }
      out[p * NLEV] = r;
    }
  }
  cp_async_wait<0>();
}

// f_dss *= spheremp on its own, for the one case where the advection kernel still reads it
__global__ void euler_fdss_kernel(double* __restrict__ f_dss, const double* __restrict__ geo) {
  const int ie = blockIdx.x;
  const double* g = geo + (size_t)ie * NPSQ * GEO_N;
  double* f = f_dss + off_f(ie);
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) {
    const int p = i / NLEV;
    double r = f[i] * geo_ld(g, p, G_SPHEREMP);
    if (is_interior_pt(p)) r *= geo_ld(g, p, G_RSPHEREMP);
    f[i] = r;
  }
}

// qdp_time_avg :379-403
__global__ void euler_time_avg_kernel(double* __restrict__ qdp, int n0_qdp, int np1_qdp) {
  const int ie = blockIdx.x, q = blockIdx.y;
  const double* a = qdp + off_q(ie, n0_qdp, q);
  double* b = qdp + off_q(ie, np1_qdp, q);
  const double rkstage = 3.0;
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) b[i] = (a[i] + (rkstage - 1) * b[i]) / rkstage;
}

void euler_precompute_divdp() {
  if (!S.nelemd) return;
  PROBE(K_EULER_DIVDP);
  euler_divdp_kernel<<<nblocks_flat(S.nelemd), TPB, 0, S.stream>>>(S.geo, S.derived_vn0, S.divdp, S.divdp_proj,
                                                                         S.nelemd);
  KERNEL_LAUNCHED(K_EULER_DIVDP);
}

static int tracer_chunk() {
  static int qc = 0;
  if (!qc) {
    const char* e = std::getenv("HXX_QCHUNK");
    qc = e ? std::max(1, std::atoi(e)) : 40;
  }
  return qc;
}

void euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt, int tavg_n0_qdp) {
  const int nq = S.p.qsize;
  if (!S.nelemd || !nq) return;
  const int mode = rhs_multiplier == 0.0 ? 0 : rhs_multiplier == 1.0 ? 1 : 2;
  if (mode == 2) S.rhs_viss = 3.0;  // compute_biharmonic_pre :196-214
  EulerArgs a{S.geo, S.tensorvisc, S.qdp, S.qtens_biharmonic, S.qlim, S.derived_dp, S.divdp_proj, S.divdp,
              S.derived_vn0, S.dpdiss_ave, S.dpdiss_biharmonic, dss_var(dss_opt), S.nelemd, nq, tracer_chunk(),
              n0_qdp, np1_qdp, dt, rhs_multiplier * dt, S.p.nu_p, S.p.nu_q, S.rhs_viss, mode, tavg_n0_qdp, S.p.limiter_option,
              S.p.consthv ? 1 : 0};
  const dim3 grid(nblocks_flat(S.nelemd), (nq + a.qchunk - 1) / a.qchunk);
  PROBE(K_EULER_QMINMAX);
  {
    constexpr size_t smem_mm = 2 * (size_t)NPSQ * TPB * sizeof(double);
    constexpr size_t smem_bih = 5 * (size_t)NPSQ * TPB * sizeof(double);
    static bool attr_mm = false;
    if (!attr_mm) {
      CUDA_OK(cudaFuncSetAttribute(euler_qminmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bih));
      attr_mm = true;
    }
    if (mode == 2) euler_qminmax_kernel<true><<<grid, TPB, smem_bih, S.stream>>>(a);
    else euler_qminmax_kernel<false><<<grid, TPB, smem_mm, S.stream>>>(a);
  }
  KERNEL_LAUNCHED(K_EULER_QMINMAX);
  if (mode == 0) {
    minmax_exchange();  // neighbor_minmax :504-507
  } else if (mode == 2) {
    // minmax_and_biharmonic :496-502 (the reference overlaps the min/max exchange with the
    // biharmonic; qlim is not touched in between, so the order is free)
    dss_exchange(fields_qtens(), true);
    minmax_exchange();
  }
  a.qlim = S.qlim;  // minmax_exchange swaps the double buffer
  static int hv_split = -1;
  if (hv_split < 0) {
    const char* e = std::getenv("HXX_HV_SPLIT");
    hv_split = e ? std::atoi(e) : 1;
  }
  const int hv = S.rhs_viss == 0.0 ? 0 : hv_split ? 2 : 1;
  const bool tavg = tavg_n0_qdp >= 0;
  auto slots = [&]() {
    return hv == 1 ? (tavg ? advect_slots<1, true>() : advect_slots<1, false>())
         : hv == 2 ? (tavg ? advect_slots<2, true>() : advect_slots<2, false>())
                   : (tavg ? advect_slots<0, true>() : advect_slots<0, false>());
  };
  const size_t smem = (size_t)slots() * TPB * sizeof(double);
  static bool attr = false;
  if (!attr) {
#define HXX_ADV_ATTR(H, T)                                                                                  \
  CUDA_OK(cudaFuncSetAttribute(euler_advect_kernel<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               advect_slots<H, T>() * TPB * (int)sizeof(double)))
    HXX_ADV_ATTR(0, false); HXX_ADV_ATTR(0, true); HXX_ADV_ATTR(1, false); HXX_ADV_ATTR(1, true);
    HXX_ADV_ATTR(2, false); HXX_ADV_ATTR(2, true);
#undef HXX_ADV_ATTR
    attr = true;
  }
  if (hv == 2) {
    PROBE(K_EULER_QMINMAX);
    euler_hvpost_kernel<<<grid, TPB, 2 * (size_t)NPSQ * TPB * sizeof(double), S.stream>>>(a);
    KERNEL_LAUNCHED(K_EULER_QMINMAX);
  }
  // divdp_proj is both the DSS variable of stage 1 and an input of compute_dp: scale it inside
  // the advection kernel only when its value no longer matters there (rhs_multiplier == 0)
  double* fdss = a.f_dss;
  const bool separate = (fdss == S.divdp_proj && a.rhsmdt != 0.0);
  if (separate) a.f_dss = nullptr;
  PROBE(K_EULER_ADVECT);
  if (hv == 1 && tavg) euler_advect_kernel<1, true><<<grid, TPB, smem, S.stream>>>(a);
  else if (hv == 1) euler_advect_kernel<1, false><<<grid, TPB, smem, S.stream>>>(a);
  else if (hv == 2 && tavg) euler_advect_kernel<2, true><<<grid, TPB, smem, S.stream>>>(a);
  else if (hv == 2) euler_advect_kernel<2, false><<<grid, TPB, smem, S.stream>>>(a);
  else if (tavg) euler_advect_kernel<0, true><<<grid, TPB, smem, S.stream>>>(a);
  else euler_advect_kernel<0, false><<<grid, TPB, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(K_EULER_ADVECT);
  if (separate) {
    PROBE(K_EULER_FDSS);
    euler_fdss_kernel<<<S.nelemd, 288, 0, S.stream>>>(fdss, S.geo);
    KERNEL_LAUNCHED(K_EULER_FDSS);
  }
  dss_exchange(fields_euler(np1_qdp, dss_opt, tavg_n0_qdp), true);  // exchange_qdp_dss_var :509-512
}

void euler_qdp_time_avg(int n0_qdp, int np1_qdp) {
  if (!S.nelemd || !S.p.qsize) return;
  PROBE(K_EULER_TAVG);
  euler_time_avg_kernel<<<dim3(S.nelemd, S.p.qsize), 288, 0, S.stream>>>(S.qdp, n0_qdp, np1_qdp);
  KERNEL_LAUNCHED(K_EULER_TAVG);
}

}  // namespace hxx
