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

// Advection kernel, per-thread storage. Shared-memory slots (slot s of thread t lives at [s][t], so a
// warp's access to one slot is 256 contiguous bytes): vstar (2x16), dpdissk (16) and its reciprocal
// (16), plus, on the hyperviscosity stage, one staged plane of the prepared term (16, cp.async).
// Registers: the limiter weights c (16), the tracer plane being advected (16) and the NEXT tracer's
// plane (16), whose loads are issued before the limiter starts and land while it runs.
// two blocks per SM: at three the register cap (168) forces spills that cost more than the extra warps give
#ifndef HXX_ADV_MINB
#define HXX_ADV_MINB 2
#endif
#ifndef HXX_ADV_MINB_HV
#define HXX_ADV_MINB_HV 2
#endif
template <bool HV>
__host__ __device__ constexpr int advect_slots() { return 64 + (HV ? NPSQ : 0); }

// compute_biharmonic_post :216-231 with rhsviss_adjustment :293-310, in place:
// qtens_biharmonic <- (-rhs_viss dt nu_q) dp0 laplace(qtens_biharmonic) / spheremp
#ifndef HXX_HVPOST_MINB
#define HXX_HVPOST_MINB 3
#endif
__global__ void __launch_bounds__(TPB, HXX_HVPOST_MINB) euler_hvpost_kernel(const EulerArgs a) {
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

// advect_and_limit :317-332 = compute_2d_advection_step (:585-626) + run_tracer_phase (:571-582).
// HV: the hyperviscosity term prepared in place by euler_hvpost_kernel is added (:216-231).
template <bool HV, bool TAVG>
__global__ void __launch_bounds__(TPB, HV ? HXX_ADV_MINB_HV : HXX_ADV_MINB) euler_advect_kernel(const EulerArgs a) {
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
  double* const s_rdpk = s_all + 48 * TPB + tid;
  double* const s_b = s_all + 64 * TPB + tid;  // HV: the prepared term of the tracer in flight
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  const double* const qin = a.qdp + off_q(ie, a.n0_qdp, 0) + k;
  const double* const qtb = a.qtens_biharmonic + (size_t)ie * QSIZE_D * NLF + k;
  const double* const qlim_in = a.qlim + (size_t)ie * QSIZE_D * 2 * NLEV + k;
  const double* const qavg = TAVG ? a.qdp + off_q(ie, a.tavg_n0, 0) + k : nullptr;
  auto stage_b = [&](int q) {
    if (HV) {
      if (q < q1) {
        const double* sb = qtb + (size_t)q * NLF;
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) cp_async8(s_b + p * TPB, sb + p * NLEV);
      }
      cp_async_commit();
    }
  };
  // next tracer's plane, qlim rows and time-average partners (qdp_time_avg :379-403, interior points)
  double xn[NPSQ], qmin_n = 0.0, qmax_n = 0.0, qa_n[4] = {0.0, 0.0, 0.0, 0.0};
  auto load_next = [&](int q) {
    if (q < q1) {
      plane_load(qin + (size_t)q * NLF, xn);
      qmin_n = qlim_in[(size_t)q * 2 * NLEV];
      qmax_n = qlim_in[(size_t)q * 2 * NLEV + NLEV];
      if (TAVG) {
        const double* pa = qavg + (size_t)q * NLF;
        qa_n[0] = pa[5 * NLEV]; qa_n[1] = pa[6 * NLEV]; qa_n[2] = pa[9 * NLEV]; qa_n[3] = pa[10 * NLEV];
      }
    }
  };
  stage_b(q0);

  const bool add_ps_diss = a.nu_p > 0 && HV;
  const double diss_fac = add_ps_diss ? -a.rhs_viss * a.dt * a.nu_q : 0.0;
  double c[NPSQ];
  {
    const double* dd = a.derived_dp + off_f(ie) + k;
    const double* dj = a.divdp_proj + off_f(ie) + k;
    const double* dv = a.divdp + off_f(ie) + k;
    const double* n0 = a.derived_vn0 + ((size_t)ie * 2 + 0) * NLF + k;
    const double* n1 = a.derived_vn0 + ((size_t)ie * 2 + 1) * NLF + k;
    const double* db = a.dpdiss_biharmonic + off_f(ie) + k;
    // two rounds of loads, so that no more than four planes are in registers at once
    double dp[NPSQ];
    {
      double r0[NPSQ], r1[NPSQ], r3[NPSQ], r4[NPSQ];
      plane_load(dd, r0);
      plane_load(dj, r1);
      plane_load(n0, r3);
      plane_load(n1, r4);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        dp[p] = r0[p] - a.rhsmdt * r1[p];
        const double rdp = 1.0 / dp[p];
        s_vs0[p * TPB] = div_rcp(r3[p], dp[p], rdp);
        s_vs1[p * TPB] = div_rcp(r4[p], dp[p], rdp);
      }
    }
    phase_fence();
    {
      double r2[NPSQ], r5[NPSQ];
      plane_load(dv, r2);
      if (add_ps_diss) plane_load(db, r5);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        const double sm_ = geo_ld(g, p, G_SPHEREMP);
        double d = dp[p] - a.dt * r2[p];
        if (add_ps_diss) d += div_rcp(diss_fac * r5[p], sm_, geo_ld(g, p, G_INV_SPHEREMP));
        s_dpk[p * TPB] = d;
        s_rdpk[p * TPB] = 1.0 / d;
        c[p] = sm_ * d;
      }
    }
  }
  phase_fence();
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
  const double alpha = -a.dt;
  double* qlp = a.qlim + ((size_t)ie * QSIZE_D + q0) * 2 * NLEV + k;
  double* out = a.qdp + off_q(ie, a.np1_qdp, q0) + k;
  load_next(q0);
  for (int q = q0; q < q1; ++q, qlp += 2 * NLEV, out += NLF) {
    const double qmin0 = qmin_n, qmax0 = qmax_n;
    const double qa[4] = {qa_n[0], qa_n[1], qa_n[2], qa_n[3]};
    // The qdp plane becomes the advected value in place, one point at a time: with the limiter's
    // weights that makes four live planes (c, x, gv0, gv1) at the widest spot.
    double x[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) x[p] = xn[p];
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
      if (HV) cp_async_wait<0>();  // this thread's copy of the prepared term has landed
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        double dx, dy;
        deriv_point(gv0, gv1, p / NP, p % NP, dx, dy);
        x[p] = x[p] + alpha * ((dx + dy) * geo_ld(g, p, G_RMETDET_R));
        if (HV) x[p] += s_b[p * TPB];
      }
    }
    // the next tracer's loads go out now and land while the limiter runs
    phase_fence();
    stage_b(q + 1);
    load_next(q + 1);
    phase_fence();
    // limiter shell :693-761; a level whose weights do not sum to a positive number is left alone
    if (!skip) {
      double qmin = qmin0, qmax = qmax0;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = div_rcp(x[p], s_dpk[p * TPB], s_rdpk[p * TPB]);
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
  if (HV) cp_async_wait<0>();
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
  const bool hv = S.rhs_viss != 0.0;
  const bool tavg = tavg_n0_qdp >= 0;
  const size_t smem = (size_t)(hv ? advect_slots<true>() : advect_slots<false>()) * TPB * sizeof(double);
  static bool attr = false;
  if (!attr) {
#define HXX_ADV_ATTR(H, T)                                                                                  \
  CUDA_OK(cudaFuncSetAttribute(euler_advect_kernel<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               advect_slots<H>() * TPB * (int)sizeof(double)))
    HXX_ADV_ATTR(false, false); HXX_ADV_ATTR(false, true); HXX_ADV_ATTR(true, false); HXX_ADV_ATTR(true, true);
#undef HXX_ADV_ATTR
    attr = true;
  }
  if (hv) {  // compute_biharmonic_post: the second Laplacian, in place
    static bool attr_hp = false;
    if (!attr_hp) {
      CUDA_OK(cudaFuncSetAttribute(euler_hvpost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   2 * NPSQ * TPB * (int)sizeof(double)));
      attr_hp = true;
    }
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
  if (hv && tavg) euler_advect_kernel<true, true><<<grid, TPB, smem, S.stream>>>(a);
  else if (hv) euler_advect_kernel<true, false><<<grid, TPB, smem, S.stream>>>(a);
  else if (tavg) euler_advect_kernel<false, true><<<grid, TPB, smem, S.stream>>>(a);
  else euler_advect_kernel<false, false><<<grid, TPB, smem, S.stream>>>(a);
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
