// Tracer advection — replaces EulerStepFunctor{,Impl}.hpp of the reference: one stage of the
// 3-stage SSP-RK2 scheme per euler_step() call (EulerStepFunctorImpl.hpp:514-561).
//
// One thread per (element, level) holds the level's 4x4 plane in registers, so the divergence, the weak
// Laplacians, the per-level min/max AND the quasi-monotone limiter (whose reductions run over the 16
// points of one level) are all thread-local — no team reductions. A block is 32 (element, level)
// lanes x 4 warps: the per-level constants of the reference's per-element set-up kernels (compute_dp
// :406-434, compute_2d_advection_step :585-626) are built once per block in shared memory instead of
// round-tripping vstar/dpdissk/dp_star through HBM, and the warps split the tracers.
// Algorithmic HBM traffic per (element, tracer, stage): min/max pass 1 tile read; advection
// 1 read + 1 write; on the stage with hyperviscosity +1 write (first Laplacian), +1 read +1 write
// (second Laplacian, in place), +1 read (the advection kernel adds the prepared term).
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

// compute_dp + compute_qmin_qmax (:406-485) and, on the hyperviscosity stage (BIH),
// compute_biharmonic_pre (:196-214, dpdiss_adjustment :251-267): Q -> laplace(Q * dpdiss_ave / dp0).
// Block shape of the advection kernel below: 32 (element, level) lanes x BIH_NW warps that share dp, its
// reciprocal and dpdiss_ave in shared memory and split the tracers; each warp stages its tracers'
// planes two ahead with cp.async. The Laplacian is finished point by point (laplace_points).
#ifndef HXX_BIH_NW
#define HXX_BIH_NW 4
#endif
#ifndef HXX_BIH_MINB
#define HXX_BIH_MINB 4
#endif
constexpr int BIH_NW = HXX_BIH_NW;
constexpr int BIH_T = 32 * BIH_NW;
constexpr int bih_smem_doubles = 3 * NPSQ * 32 + BIH_NW * 2 * NPSQ * 32;
template <bool BIH>
__global__ void __launch_bounds__(BIH_T, HXX_BIH_MINB) euler_qminmax_kernel(const EulerArgs a) {
  extern __shared__ double s_all[];
  __shared__ double s_geo[geo_span(32) * NPSQ * GEO_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e_first = (int)(((long long)blockIdx.x * 32) / NLEV);
  if (BIH) stage_geo<geo_span(32), BIH_T>(s_geo, a.geo, e_first, a.nelem);
  const long long gl = (long long)blockIdx.x * 32 + lane, glmax = (long long)a.nelem * NLEV - 1;
  const bool valid = gl <= glmax;
  const int ie = (int)((valid ? gl : glmax) / NLEV), k = (int)((valid ? gl : glmax) % NLEV);
  double* const s_dp = s_all + lane;
  double* const s_rdp = s_all + 16 * 32 + lane;
  double* const s_dave = s_all + 32 * 32 + lane;
  double* const s_q = s_all + 48 * 32 + w * 2 * NPSQ * 32 + lane;  // this warp's [2][16][32] staging
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const int q1 = a.qsize;
  const double* const qin = a.qdp + off_q(ie, a.n0_qdp, 0) + k;
  auto prefetch = [&](int q, int buf) {
    if (q < q1) {
      const double* src = qin + (size_t)q * NLF;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) cp_async8(s_q + (buf * NPSQ + p) * 32, src + p * NLEV);
    }
    cp_async_commit();
  };
  prefetch(w, 0);
  prefetch(w + BIH_NW, 1);
  const bool scale = BIH && a.nu_p > 0;
  {
    constexpr int PPW = NPSQ / BIH_NW;
    const size_t o = off_f(ie) + k;
    HXX_UNROLL
    for (int i = 0; i < PPW; ++i) {
      const int p = w * PPW + i;
      const double d = a.derived_dp[o + p * NLEV] - a.rhsmdt * a.divdp_proj[o + p * NLEV];
      s_dp[p * 32] = d;
      s_rdp[p * 32] = 1.0 / d;
      if (scale) s_dave[p * 32] = a.dpdiss_ave[o + p * NLEV];
    }
  }
  __syncthreads();
  const double dp0k = dc.dp0[k], rdp0k = 1.0 / dp0k;
  int it = 0;
  for (int q = w; q < q1; q += BIH_NW, ++it) {
    const int buf = it & 1;
    double* const ql = a.qlim + ((size_t)ie * QSIZE_D + q) * 2 * NLEV + k;
    double* const qtb = a.qtens_biharmonic + ((size_t)ie * QSIZE_D + q) * NLF + k;
    double mn = 0.0, mx = 0.0;
    if (a.rhs_mode == 1) { mn = ql[0]; mx = ql[NLEV]; }
    cp_async_wait<1>();
    double Q[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) Q[p] = s_q[(buf * NPSQ + p) * 32];
    prefetch(q + 2 * BIH_NW, buf);
    div_rcp_plane(Q, [&](int p) { return s_dp[p * 32]; }, [&](int p) { return s_rdp[p * 32]; });
    if (a.rhs_mode != 1) { mn = Q[0]; mx = Q[0]; }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) { mn = fmin(mn, Q[p]); mx = fmax(mx, Q[p]); }
    if (valid) {
      ql[0] = mn;
      ql[NLEV] = mx;
    }
    if (scale) {
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) Q[p] = Q[p] * s_dave[p * 32];
      div_rcp_plane(Q, [&](int) { return dp0k; }, [&](int) { return rdp0k; });
    }
    if (BIH)
      laplace_points<false>(g, nullptr, Q, [&](int p, double lap) {
        if (is_interior_pt(p)) lap *= geo_ld(g, p, G_RSPHEREMP);  // rspheremp of the DSS that follows
        if (valid) qtb[p * NLEV] = lap;
      });
  }
  cp_async_wait<0>();
}

// Advection kernel. A block is 32 consecutive (element, level) lanes x ADV_NW warps: the warps share the
// lanes' per-level constants of compute_2d_advection_step — vstar (2x16), dpdissk (16), its reciprocal
// (16) and the limiter weights c = spheremp dpdissk (16), built once per block in shared memory
// ([slot][lane], a warp's access to one slot is 256 contiguous bytes) — and split the tracers among
// them (warp w takes q = w, w + ADV_NW, ...). Per-thread state is then the tracer plane being advected,
// the NEXT tracer's plane (its loads are issued before the limiter starts and land while it runs) and
// two work planes, which is what lets three blocks (12 warps) live on an SM without spills.
#ifndef HXX_ADV_NW
#define HXX_ADV_NW 4
#endif
#ifndef HXX_ADV_MINB
#define HXX_ADV_MINB 3
#endif
#ifndef HXX_ADV_MINB_HV
#define HXX_ADV_MINB_HV 3
#endif
// 1: the advection kernel of the hyperviscosity stage applies the second Laplacian itself (one pass over
// qtens_biharmonic less); 0: euler_hvpost_kernel prepares the term in place first. Measured at ne30/q40:
// fused 50.1 ms per subcycle, split 49.0 ms (the fused kernel's FP64 work no longer hides its latency)
#ifndef HXX_HV_FUSED
#define HXX_HV_FUSED 0
#endif
constexpr bool HV_FUSED = HXX_HV_FUSED != 0;
constexpr int ADV_NW = HXX_ADV_NW;
constexpr int ADV_T = 32 * ADV_NW;
static_assert(NPSQ % ADV_NW == 0, "the warps split the 16 points of the set-up evenly");
// per-warp staging slots of the tracer in flight: its qdp plane (16), the two qlim rows, the four
// time-average partners and, on the hyperviscosity stage, the prepared term (16)
template <bool HV>
__host__ __device__ constexpr int advect_stage_slots() { return NPSQ + 2 + 4 + (HV ? NPSQ : 0); }
// doubles of dynamic shared memory: 5 constant planes + the weight sum (+ dp and 1/dp when the kernel also
// does the stage's min/max pass, MM), then the warps' staging slots
template <bool HV, bool MM = false>
__host__ __device__ constexpr int advect_smem_doubles() {
  return (5 * NPSQ + 1 + (MM ? 2 * NPSQ : 0)) * 32 + ADV_NW * advect_stage_slots<HV>() * 32;
}

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
    double s[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) s[p] = s_q[(buf * NPSQ + p) * TPB];
    prefetch(q + 2, buf);
    double t[NPSQ];
    auto emit = [&](int p, double lap) { t[p] = bfac * dp0k * lap; };
    if (a.consthv) laplace_points<false>(g, tv, s, emit); else laplace_points<true>(g, tv, s, emit);
    div_rcp_plane(t, [&](int p) { return geo_ld(g, p, G_SPHEREMP); }, [&](int p) { return geo_ld(g, p, G_INV_SPHEREMP); });
    plane_store(qtb + (size_t)q * NLF, t);
  }
  cp_async_wait<0>();
}

// advect_and_limit :317-332 = compute_2d_advection_step (:585-626) + run_tracer_phase (:571-582).
// HV: the hyperviscosity term prepared in place by euler_hvpost_kernel is added (:216-231).
// MM: the stage with rhs_multiplier == 1 has no neighbour exchange between its min/max pass (compute_qmin_qmax
// :436-485, qmin = min(qmin, Q), Q = Qdp / dp) and the limiter, and both read the same Qdp plane: the kernel
// forms Q from the plane it has just staged and widens the limits itself, so the tracers are read once.
template <bool HV, bool TAVG, bool MM = false>
__global__ void __launch_bounds__(ADV_T, HV ? HXX_ADV_MINB_HV : HXX_ADV_MINB) euler_advect_kernel(const EulerArgs a) {
  extern __shared__ double s_all[];
  __shared__ double s_geo[geo_span(32) * NPSQ * GEO_N];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e_first = (int)(((long long)blockIdx.x * 32) / NLEV);
  stage_geo<geo_span(32), ADV_T>(s_geo, a.geo, e_first, a.nelem);
  // lanes past the last (element, level) pair work on the last valid one and store nothing
  const long long gl = (long long)blockIdx.x * 32 + lane, glmax = (long long)a.nelem * NLEV - 1;
  const bool valid = gl <= glmax;
  const int ie = (int)((valid ? gl : glmax) / NLEV), k = (int)((valid ? gl : glmax) % NLEV);
  double* const s_vs0 = s_all + lane;
  double* const s_vs1 = s_all + 16 * 32 + lane;
  double* const s_dpk = s_all + 32 * 32 + lane;
  double* const s_rdpk = s_all + 48 * 32 + lane;
  double* const s_c = s_all + 64 * 32 + lane;
  double* const s_sumc = s_all + 80 * 32 + lane;
  double* const s_dp = s_all + 81 * 32 + lane;             // MM only: dp of compute_dp (:406-434) and 1 / dp
  double* const s_rdp = s_all + (81 + NPSQ) * 32 + lane;
  double* const s_q = s_all + (81 + (MM ? 2 * NPSQ : 0)) * 32 + w * advect_stage_slots<HV>() * 32 + lane;  // this warp's staging slots
  double* const s_l = s_q + NPSQ * 32;  // qlim rows
  double* const s_a = s_l + 2 * 32;     // time-average partners (qdp_time_avg :379-403, interior points)
  double* const s_b = s_a + 4 * 32;     // HV: the prepared term
  const GeoShared g{s_geo + (ie - e_first) * NPSQ * GEO_N};
  const double* __restrict__ tvis = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const double* const qin = a.qdp + off_q(ie, a.n0_qdp, 0) + k;
  const double* const qtb = a.qtens_biharmonic + (size_t)ie * QSIZE_D * NLF + k;
  const double* const qlim_in = a.qlim + (size_t)ie * QSIZE_D * 2 * NLEV + k;
  const double* const qavg = TAVG ? a.qdp + off_q(ie, a.tavg_n0, 0) + k : nullptr;
  const int q1 = a.qsize;
  // cp.async (LDGSTS) copies of tracer q into the staging slots: they land while the limiter of the
  // tracer before it runs, without holding registers
  auto prefetch = [&](int q) {
    if (q < q1) {
      const double* src = qin + (size_t)q * NLF;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) cp_async8(s_q + p * 32, src + p * NLEV);
      if (HV) {
        const double* sb = qtb + (size_t)q * NLF;
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) cp_async8(s_b + p * 32, sb + p * NLEV);
      }
      cp_async8(s_l, qlim_in + (size_t)q * 2 * NLEV);
      cp_async8(s_l + 32, qlim_in + (size_t)q * 2 * NLEV + NLEV);
      if (TAVG) {
        const double* pa = qavg + (size_t)q * NLF;
        cp_async8(s_a, pa + 5 * NLEV);
        cp_async8(s_a + 32, pa + 6 * NLEV);
        cp_async8(s_a + 2 * 32, pa + 9 * NLEV);
        cp_async8(s_a + 3 * 32, pa + 10 * NLEV);
      }
    }
    cp_async_commit();
  };
  prefetch(w);

  // set-up, shared by the block's warps: warp w builds points [w PPW, (w + 1) PPW)
  constexpr int PPW = NPSQ / ADV_NW;
  const bool add_ps_diss = a.nu_p > 0 && HV;
  const double diss_fac = add_ps_diss ? -a.rhs_viss * a.dt * a.nu_q : 0.0;
  {
    const size_t o = off_f(ie) + k;
    double r0[PPW], r1[PPW], r2[PPW], r3[PPW], r4[PPW], r5[PPW], rf[PPW];
    HXX_UNROLL
    for (int i = 0; i < PPW; ++i) {
      const int p = w * PPW + i;
      r0[i] = a.derived_dp[o + p * NLEV];
      r1[i] = a.divdp_proj[o + p * NLEV];
      r2[i] = a.divdp[o + p * NLEV];
      r3[i] = a.derived_vn0[((size_t)ie * 2 + 0) * NLF + k + p * NLEV];
      r4[i] = a.derived_vn0[((size_t)ie * 2 + 1) * NLF + k + p * NLEV];
      r5[i] = add_ps_diss ? a.dpdiss_biharmonic[o + p * NLEV] : 0.0;
      rf[i] = a.f_dss ? a.f_dss[o + p * NLEV] : 0.0;
    }
    HXX_UNROLL
    for (int i = 0; i < PPW; ++i) {
      const int p = w * PPW + i;
      const double sm_ = geo_ld(g, p, G_SPHEREMP);
      const double dp = r0[i] - a.rhsmdt * r1[i];
      const double rdp = 1.0 / dp;
      s_vs0[p * 32] = div_rcp(r3[i], dp, rdp);
      s_vs1[p * 32] = div_rcp(r4[i], dp, rdp);
      if (MM) { s_dp[p * 32] = dp; s_rdp[p * 32] = rdp; }
      double d = dp - a.dt * r2[i];
      if (add_ps_diss) d += div_rcp(diss_fac * r5[i], sm_, geo_ld(g, p, G_INV_SPHEREMP));
      s_dpk[p * 32] = d;
      s_rdpk[p * 32] = 1.0 / d;
      s_c[p * 32] = sm_ * d;
      if (a.f_dss && valid) {  // f_dss *= spheremp (and the interior part of the DSS rspheremp)
        double r = rf[i] * sm_;
        if (is_interior_pt(p)) r *= geo_ld(g, p, G_RSPHEREMP);
        a.f_dss[o + p * NLEV] = r;
      }
    }
  }
  __syncthreads();
  if (w == 0) {  // the limiter's weight sum does not depend on the tracer (serial order k = 0..15, as the reference)
    double sumc = s_c[0];
    HXX_UNROLL
    for (int p = 1; p < NPSQ; ++p) sumc += s_c[p * 32];
    s_sumc[0] = sumc;
  }
  __syncthreads();
  const double sumc = s_sumc[0];
  const bool skip = sumc <= 0;
  const double alpha = -a.dt;
  for (int q = w; q < q1; q += ADV_NW) {
    double* const qlp = a.qlim + ((size_t)ie * QSIZE_D + q) * 2 * NLEV + k;
    double* const out = a.qdp + off_q(ie, a.np1_qdp, q) + k;
    cp_async_wait<0>();  // this thread's copies of tracer q have landed
    const double qmin0 = s_l[0], qmax0 = s_l[32];
    double qa[4] = {0.0, 0.0, 0.0, 0.0};
    if (TAVG) { qa[0] = s_a[0]; qa[1] = s_a[32]; qa[2] = s_a[2 * 32]; qa[3] = s_a[3 * 32]; }
    // The qdp plane becomes the advected value in place, one point at a time: three live planes
    // (x, gv0, gv1) at the widest spot.
    double x[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) x[p] = s_q[p * 32];
    double qmin = qmin0, qmax = qmax0;
    if (MM) {  // compute_qmin_qmax :436-485 on the plane just staged; same quotient as euler_qminmax_kernel
      unsigned worst = 0;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) worst = max(worst, (((unsigned)__double2hiint(x[p]) >> 20) & 0x7ffu) - 423u);
      if (worst <= 1200u) {
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) {
          const double dd = s_dp[p * 32], rr = s_rdp[p * 32];
          const double q = x[p] * rr;
          const double Q = fma(fma(-dd, q, x[p]), rr, q);
          qmin = fmin(qmin, Q);
          qmax = fmax(qmax, Q);
        }
      } else {
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) {
          const double Q = div_rcp(x[p], s_dp[p * 32], s_rdp[p * 32]);
          qmin = fmin(qmin, Q);
          qmax = fmax(qmax, Q);
        }
      }
    }
    {
      // divergence_sphere_update, SphereOperators.hpp:398-444
      double gv0[NPSQ], gv1[NPSQ];
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        const double u = s_vs0[p * 32] * x[p];
        const double v = s_vs1[p * 32] * x[p];
        const double md = geo_ld(g, p, G_METDET);
        gv0[p] = (geo_ld(g, p, G_DINV00) * u + geo_ld(g, p, G_DINV10) * v) * md;
        gv1[p] = (geo_ld(g, p, G_DINV01) * u + geo_ld(g, p, G_DINV11) * v) * md;
      }
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        double dx, dy;
        deriv_point(gv0, gv1, p / NP, p % NP, dx, dy);
        x[p] = x[p] + alpha * ((dx + dy) * geo_ld(g, p, G_RMETDET_R));
        if (HV && !HV_FUSED) x[p] += s_b[p * 32];
      }
    }
    if (HV && HV_FUSED) {
      // compute_biharmonic_post :216-231 with rhsviss_adjustment :293-310 on the fly: the staged plane
      // is the assembled first Laplacian; its second Laplacian joins x point by point
      // x waits in the (consumed) staging slots of its tracer while the Laplacian has the registers
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) s_q[p * 32] = x[p];
      phase_fence();
      double sb[NPSQ];
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) sb[p] = s_b[p * 32];
      const double bfac = -a.rhs_viss * a.dt * a.nu_q, dp0k = dc.dp0[k];
      auto emit = [&](int p, double lap) {
        s_q[p * 32] += div_rcp(bfac * dp0k * lap, geo_ld(g, p, G_SPHEREMP), geo_ld(g, p, G_INV_SPHEREMP));
      };
      if (a.consthv) laplace_points<false>(g, tvis, sb, emit); else laplace_points<true>(g, tvis, sb, emit);
      phase_fence();
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = s_q[p * 32];
    }
    // the staged values are in registers: the slots take the next tracer, which lands while the limiter runs
    prefetch(q + ADV_NW);
    // limiter shell :693-761; a level whose weights do not sum to a positive number is left alone
    if (!skip) {
      div_rcp_plane(x, [&](int p) { return s_dpk[p * 32]; }, [&](int p) { return s_rdpk[p * 32]; });
      limiter_level_w(a.limiter_option, SlotPlane{s_c, 32}, sumc, x, qmin, qmax);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = x[p] * s_dpk[p * 32];
    }
    if (valid) {  // only the limits that moved (qmin0, qmax0 are what memory holds); MM: always, two registers less
      if (MM || qmin != qmin0) qlp[0] = qmin;
      if (MM || qmax != qmax0) qlp[NLEV] = qmax;
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {  // apply_spheremp :672-687
      double r = geo_ld(g, p, G_SPHEREMP) * x[p];
      if (is_interior_pt(p)) {
        r *= geo_ld(g, p, G_RSPHEREMP);
        if (TAVG) r = (qa[p == 5 ? 0 : p == 6 ? 1 : p == 9 ? 2 : 3] + 2.0 * r) / 3.0;
      }
      if (valid) out[p * NLEV] = r;
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
  // rhs_multiplier == 1 (no exchange between the min/max pass and the limiter): the advection kernel does both
  const bool fuse_mm = mode == 1 && S.rhs_viss == 0.0 && tavg_n0_qdp < 0;
  if (!fuse_mm) {
  PROBE(K_EULER_QMINMAX);
  {
    if (HXX_ONCE_PER_SESSION()) {
      CUDA_OK(cudaFuncSetAttribute(euler_qminmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   bih_smem_doubles * (int)sizeof(double)));
      CUDA_OK(cudaFuncSetAttribute(euler_qminmax_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   bih_smem_doubles * (int)sizeof(double)));
    }
    const int nb32 = (int)(((long long)S.nelemd * NLEV + 31) / 32);
    if (mode == 2) euler_qminmax_kernel<true><<<nb32, BIH_T, bih_smem_doubles * sizeof(double), S.stream>>>(a);
    else euler_qminmax_kernel<false><<<nb32, BIH_T, bih_smem_doubles * sizeof(double), S.stream>>>(a);
  }
  KERNEL_LAUNCHED(K_EULER_QMINMAX);
  }
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
  const size_t smem = (size_t)(hv ? advect_smem_doubles<true>() : fuse_mm ? advect_smem_doubles<false, true>()
                                                                           : advect_smem_doubles<false>()) * sizeof(double);
  const int adv_blocks = (int)(((long long)S.nelemd * NLEV + 31) / 32);
  if (HXX_ONCE_PER_SESSION()) {
#define HXX_ADV_ATTR(H, T)                                                                                  \
  CUDA_OK(cudaFuncSetAttribute(euler_advect_kernel<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               advect_smem_doubles<H>() * (int)sizeof(double)))
    HXX_ADV_ATTR(false, false); HXX_ADV_ATTR(false, true); HXX_ADV_ATTR(true, false); HXX_ADV_ATTR(true, true);
#undef HXX_ADV_ATTR
    CUDA_OK(cudaFuncSetAttribute(euler_advect_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 advect_smem_doubles<false, true>() * (int)sizeof(double)));
  }
  if (hv && !HV_FUSED) {  // compute_biharmonic_post: the second Laplacian, in place
    if (HXX_ONCE_PER_SESSION()) {
      CUDA_OK(cudaFuncSetAttribute(euler_hvpost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   2 * NPSQ * TPB * (int)sizeof(double)));
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
  // probe classes: euler_advect = the plain stage, euler_advect_mm = the stage that also does its min/max pass,
  // euler_advect_hv = the hyperviscosity stage (they move different bytes per launch; bench.py rooflines each)
  const int kid = hv ? K_EULER_ADVECT_HV : fuse_mm ? K_EULER_ADVECT_MM : K_EULER_ADVECT;
  PROBE(kid);
  if (hv && tavg) euler_advect_kernel<true, true><<<adv_blocks, ADV_T, smem, S.stream>>>(a);
  else if (hv) euler_advect_kernel<true, false><<<adv_blocks, ADV_T, smem, S.stream>>>(a);
  else if (tavg) euler_advect_kernel<false, true><<<adv_blocks, ADV_T, smem, S.stream>>>(a);
  else if (fuse_mm) euler_advect_kernel<false, false, true><<<adv_blocks, ADV_T, smem, S.stream>>>(a);
  else euler_advect_kernel<false, false><<<adv_blocks, ADV_T, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(kid);
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
