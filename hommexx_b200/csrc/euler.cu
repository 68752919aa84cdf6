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
// compute_biharmonic_pre (:196-214, dpdiss_adjustment :251-267): Q -> laplace(Q * dpdiss_ave / dp0)
__global__ void __launch_bounds__(TPB, 2) euler_qminmax_kernel(const EulerArgs a) {
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  const double* __restrict__ g = a.geo + (size_t)ie * NPSQ * GEO_N;
  double dps[NPSQ], dave[NPSQ];
  {
    const double* dd = a.derived_dp + off_f(ie) + k;
    const double* dj = a.divdp_proj + off_f(ie) + k;
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) dps[p] = dd[p * NLEV] - a.rhsmdt * dj[p * NLEV];
  }
  const bool bih = a.rhs_mode == 2;
  if (bih && a.nu_p > 0) plane_load(a.dpdiss_ave + off_f(ie) + k, dave);
  const double dp0k = dc.dp0[k];
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  for (int q = q0; q < q1; ++q) {
    double Q[NPSQ];
    plane_load(a.qdp + off_q(ie, a.n0_qdp, q) + k, Q);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) Q[p] = Q[p] / dps[p];
    double* ql = a.qlim + ((size_t)ie * QSIZE_D + q) * 2 * NLEV + k;
    double mn, mx;
    if (a.rhs_mode != 1) { mn = Q[0]; mx = Q[0]; }
    else { mn = ql[0]; mx = ql[NLEV]; }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) { mn = fmin(mn, Q[p]); mx = fmax(mx, Q[p]); }
    ql[0] = mn;
    ql[NLEV] = mx;
    if (bih) {
      double lap[NPSQ];
      if (a.nu_p > 0) {
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) Q[p] = Q[p] * dave[p] / dp0k;
      }
      laplace_simple(g, Q, lap);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p)
        if (is_interior_pt(p)) lap[p] *= geo_ld(g, p, G_RSPHEREMP);  // rspheremp of the DSS that follows
      plane_store(a.qtens_biharmonic + ((size_t)ie * QSIZE_D + q) * NLF + k, lap);
    }
  }
}

// advect_and_limit :317-332 = compute_2d_advection_step (:585-626) + run_tracer_phase (:571-582),
// with compute_biharmonic_post (:216-231, rhsviss_adjustment :293-310) applied on the fly.
__global__ void __launch_bounds__(TPB, 2) euler_advect_kernel(const EulerArgs a) {
  extern __shared__ double s_vs[];  // vstar: [2][16][blockDim] thread-private slots
  int ie, k;
  if (!map_thread(a.nelem, ie, k)) return;
  const int nt = TPB, tid = threadIdx.x;
  const double* __restrict__ g = a.geo + (size_t)ie * NPSQ * GEO_N;
  const double* __restrict__ tv = a.consthv ? nullptr : a.tensorvisc + (size_t)ie * 4 * NPSQ;
  const bool add_hv = a.rhs_viss != 0.0;
  const bool add_ps_diss = a.nu_p > 0 && add_hv;
  const double diss_fac = add_ps_diss ? -a.rhs_viss * a.dt * a.nu_q : 0.0;
  double dpk[NPSQ], c[NPSQ];
  {
    const double* dd = a.derived_dp + off_f(ie) + k;
    const double* dj = a.divdp_proj + off_f(ie) + k;
    const double* dv = a.divdp + off_f(ie) + k;
    const double* n0 = a.derived_vn0 + ((size_t)ie * 2 + 0) * NLF + k;
    const double* n1 = a.derived_vn0 + ((size_t)ie * 2 + 1) * NLF + k;
    const double* db = a.dpdiss_biharmonic + off_f(ie) + k;
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const double sm_ = geo_ld(g, p, G_SPHEREMP);
      const double dp = dd[p * NLEV] - a.rhsmdt * dj[p * NLEV];
      s_vs[(0 * NPSQ + p) * nt + tid] = n0[p * NLEV] / dp;
      s_vs[(1 * NPSQ + p) * nt + tid] = n1[p * NLEV] / dp;
      double d = dp - a.dt * dv[p * NLEV];
      if (add_ps_diss) d += diss_fac * db[p * NLEV] / sm_;
      dpk[p] = d;
      c[p] = sm_ * d;
    }
  }
  if (blockIdx.y == 0 && a.f_dss) {  // f_dss *= spheremp (and the interior part of the DSS rspheremp)
    double* f = a.f_dss + off_f(ie) + k;
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      double r = f[p * NLEV] * geo_ld(g, p, G_SPHEREMP);
      if (is_interior_pt(p)) r *= geo_ld(g, p, G_RSPHEREMP);
      f[p * NLEV] = r;
    }
  }
  const double dp0k = dc.dp0[k];
  const double bfac = -a.rhs_viss * a.dt * a.nu_q;
  const double alpha = -a.dt;
  const int q0 = blockIdx.y * a.qchunk, q1 = min(a.qsize, q0 + a.qchunk);
  for (int q = q0; q < q1; ++q) {
    double x[NPSQ];
    {
      double qd[NPSQ], gv0[NPSQ], gv1[NPSQ];
      plane_load(a.qdp + off_q(ie, a.n0_qdp, q) + k, qd);
      // divergence_sphere_update, SphereOperators.hpp:398-444
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        const double u = s_vs[(0 * NPSQ + p) * nt + tid] * qd[p];
        const double v = s_vs[(1 * NPSQ + p) * nt + tid] * qd[p];
        const double md = geo_ld(g, p, G_METDET);
        gv0[p] = (geo_ld(g, p, G_DINV00) * u + geo_ld(g, p, G_DINV10) * v) * md;
        gv1[p] = (geo_ld(g, p, G_DINV01) * u + geo_ld(g, p, G_DINV11) * v) * md;
      }
      double dx[NPSQ], dy[NPSQ];
      deriv_pair(gv0, gv1, dx, dy);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = qd[p] + alpha * ((dx[p] + dy[p]) * geo_ld(g, p, G_RMETDET_R));
    }
    if (add_hv) {
      double s[NPSQ], lap[NPSQ];
      plane_load(a.qtens_biharmonic + ((size_t)ie * QSIZE_D + q) * NLF + k, s);
      if (a.consthv) laplace_simple(g, s, lap); else laplace_tensor(g, tv, s, lap);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] += bfac * dp0k * lap[p] / geo_ld(g, p, G_SPHEREMP);
    }
    // limiter shell :693-761
    double* ql = a.qlim + ((size_t)ie * QSIZE_D + q) * 2 * NLEV + k;
    double qmin = ql[0], qmax = ql[NLEV];
    const double qmin0 = qmin, qmax0 = qmax;
    double xs[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) xs[p] = x[p] / dpk[p];
    if (limiter_level(a.limiter_option, c, xs, qmin, qmax)) {
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) x[p] = xs[p] * dpk[p];
      if (qmin != qmin0) ql[0] = qmin;
      if (qmax != qmax0) ql[NLEV] = qmax;
    }
    double* out = a.qdp + off_q(ie, a.np1_qdp, q) + k;  // apply_spheremp :672-687
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      double r = geo_ld(g, p, G_SPHEREMP) * x[p];
      if (is_interior_pt(p)) r *= geo_ld(g, p, G_RSPHEREMP);
      out[p * NLEV] = r;
    }
  }
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
    qc = e ? std::max(1, std::atoi(e)) : 10;
  }
  return qc;
}

void euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt) {
  const int nq = S.p.qsize;
  if (!S.nelemd || !nq) return;
  const int mode = rhs_multiplier == 0.0 ? 0 : rhs_multiplier == 1.0 ? 1 : 2;
  if (mode == 2) S.rhs_viss = 3.0;  // compute_biharmonic_pre :196-214
  EulerArgs a{S.geo, S.tensorvisc, S.qdp, S.qtens_biharmonic, S.qlim, S.derived_dp, S.divdp_proj, S.divdp,
              S.derived_vn0, S.dpdiss_ave, S.dpdiss_biharmonic, dss_var(dss_opt), S.nelemd, nq, tracer_chunk(),
              n0_qdp, np1_qdp, dt, rhs_multiplier * dt, S.p.nu_p, S.p.nu_q, S.rhs_viss, mode, S.p.limiter_option,
              S.p.consthv ? 1 : 0};
  const dim3 grid(nblocks_flat(S.nelemd), (nq + a.qchunk - 1) / a.qchunk);
  PROBE(K_EULER_QMINMAX);
  euler_qminmax_kernel<<<grid, TPB, 0, S.stream>>>(a);
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
  constexpr size_t smem = 2 * (size_t)NPSQ * TPB * sizeof(double);
  static bool attr = false;
  if (!attr) {
    CUDA_OK(cudaFuncSetAttribute(euler_advect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  // divdp_proj is both the DSS variable of stage 1 and an input of compute_dp: scale it inside
  // the advection kernel only when its value no longer matters there (rhs_multiplier == 0)
  double* fdss = a.f_dss;
  const bool separate = (fdss == S.divdp_proj && a.rhsmdt != 0.0);
  if (separate) a.f_dss = nullptr;
  PROBE(K_EULER_ADVECT);
  euler_advect_kernel<<<grid, TPB, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(K_EULER_ADVECT);
  if (separate) {
    PROBE(K_EULER_FDSS);
    euler_fdss_kernel<<<S.nelemd, 288, 0, S.stream>>>(fdss, S.geo);
    KERNEL_LAUNCHED(K_EULER_FDSS);
  }
  dss_exchange(fields_euler(np1_qdp, dss_opt), true);  // exchange_qdp_dss_var :509-512
}

void euler_qdp_time_avg(int n0_qdp, int np1_qdp) {
  if (!S.nelemd || !S.p.qsize) return;
  PROBE(K_EULER_TAVG);
  euler_time_avg_kernel<<<dim3(S.nelemd, S.p.qsize), 288, 0, S.stream>>>(S.qdp, n0_qdp, np1_qdp);
  KERNEL_LAUNCHED(K_EULER_TAVG);
}

}  // namespace hxx
