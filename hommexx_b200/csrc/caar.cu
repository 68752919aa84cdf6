// CAAR — compute_and_apply_rhs, one Runge-Kutta stage of the dynamics. Replaces
// CaarFunctor.{hpp,cpp} / CaarFunctorImpl.hpp of the reference (rsplit > 0, and rsplit = 0 as the VADV
// instantiation) plus the small
// element-wise kernels of prim_driver.cpp / prim_step.cpp / prim_advance_exp.cpp.
//
// One thread per (element, level); the 4x4 plane of the level lives in registers, every
// horizontal operator is thread-local. The three vertical integrals (pressure, hydrostatic
// geopotential, omega) run in the reference's sequential order through shared memory
// (non-CUDA branches CaarFunctorImpl.hpp:621-650, :689-729, :854-889), one thread per column.
// Planes that outlive a phase (pressure, phi, omega, Tv, the energy gradient) wait in per-thread
// shared-memory slots and the operators finish one point at a time, so nothing spills to local memory.
// Algorithmic HBM traffic per element and stage: read v,T,dp3d at n0 and nm1 (8 tiles, 4 when
// nm1 == n0), write 4 tiles, plus read-modify-write of derived_vn0 (2) and omega_p (1) when
// eta_ave_w != 0.
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

#include "hxx_sphere.cuh"

namespace hxx {

struct CaarArgs {
  const double* geo;
  double *v, *t, *dp3d, *vn0, *omega_p, *phi, *eta_dot_dpdn;
  const double* qdp;
  int nelem, nm1, n0, np1, n0_qdp;
  double dt, eta_ave_w;
  int fold_rsp, store_phi;
};

#ifndef HXX_CAAR_E
#define HXX_CAAR_E 2
#endif
#ifndef HXX_CAAR_MINB
#define HXX_CAAR_MINB 2
#endif
// Elements per block: small blocks, several resident per SM, so that while one block walks its
// columns serially (64 threads busy) the others keep the FP64 pipe and the memory system fed.
constexpr int CAAR_E = HXX_CAAR_E;

// VADV: rsplit == 0, Eulerian vertical advection (compute_phase_3 :136-148): the interface flux eta_dot_dpdn and
// the vertical-advection tendencies of T and v take four more shared-memory planes (one block per SM).
template <int E, bool VADV>
__global__ void __launch_bounds__(E* NLEV, VADV ? 1 : HXX_CAAR_MINB) caar_kernel(const CaarArgs a) {
  constexpr int LS = NLEV + 1;  // odd column stride: conflict-free column walks
  extern __shared__ double sm[];
  double* s_p = sm;                  // dp -> pressure
  double* s_x = sm + E * NPSQ * LS;  // div_vdp -> running sum; then a_k -> phi
  double* s_y = sm + 2 * E * NPSQ * LS;
  double* s_z = sm + 3 * E * NPSQ * LS;
  double* s_w = sm + 4 * E * NPSQ * LS;
  double* s_e = sm + 5 * E * NPSQ * LS;   // VADV: eta_dot_dpdn at the interface above level k
  double* s_tv = sm + 6 * E * NPSQ * LS;  // VADV: t_vadv, v_vadv
  double* s_v0 = sm + 7 * E * NPSQ * LS;
  double* s_v1 = sm + 8 * E * NPSQ * LS;
  __shared__ double s_tot[VADV ? E * NPSQ : 1];  // VADV: the column's sdot_sum
  __shared__ double s_geo[E * NPSQ * GEO_N];
  const int tid = threadIdx.x, e = tid / NLEV, k = tid % NLEV;
  // the block's geometry records: every operator below re-reads them, and shared-memory reads do
  // not queue behind the block's outstanding HBM requests
  stage_geo<E, E * NLEV>(s_geo, a.geo, blockIdx.x * E, a.nelem);
  int ie = blockIdx.x * E + e;
  const bool valid = ie < a.nelem;
  if (!valid) ie = a.nelem - 1;
  const GeoShared g{s_geo + (valid ? e : 0) * NPSQ * GEO_N};
  const double* v0p = a.v + off_v(ie, a.n0, 0) + k;
  const double* v1p = a.v + off_v(ie, a.n0, 1) + k;
  const int col0 = e * NPSQ;

  double dp[NPSQ], div[NPSQ];
  {
    double v0[NPSQ], v1[NPSQ];
    plane_load(a.dp3d + off_s(ie, a.n0) + k, dp);
    plane_load(v0p, v0);
    plane_load(v1p, v1);
    // compute_div_vdp, CaarFunctorImpl.hpp:370-396
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) { v0[p] *= dp[p]; v1[p] *= dp[p]; }
    if (a.eta_ave_w != 0.0 && valid) {
      // every read-modify-write and every nm1 -> np1 update below loads its whole plane before the
      // first store: a store between two loads of possibly aliasing arrays would serialise them
      double* n0p = a.vn0 + ((size_t)ie * 2 + 0) * NLF + k;
      double* n1p = a.vn0 + ((size_t)ie * 2 + 1) * NLF + k;
      double a0[NPSQ], a1[NPSQ];
      plane_load(n0p, a0);
      plane_load(n1p, a1);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        a0[p] += a.eta_ave_w * v0[p];
        a1[p] += a.eta_ave_w * v1[p];
      }
      plane_store(n0p, a0);
      plane_store(n1p, a1);
    }
    divergence_sphere(g, v0, v1, div);
  }
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    s_p[(col0 + p) * LS + k] = dp[p];
    s_x[(col0 + p) * LS + k] = div[p];
  }
  __syncthreads();
  if (tid < E * NPSQ) {
    // compute_pressure (:621-650) and the omega running sum (:854-889), two independent chains;
    // loads are batched eight levels at a time so only the add chains are serial
    double* cp_ = s_p + tid * LS;
    double* cx = s_x + tid * LS;
    double dp_prev = 0.0, p_prev = dc.hyai0 * dc.ps0, integ = 0.0;
    for (int k0 = 0; k0 < NLEV; k0 += 8) {
      double d[8], dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d[i] = (k0 + i < NLEV) ? cp_[k0 + i] : 0.0;
        dv[i] = (k0 + i < NLEV) ? cx[k0 + i] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k0 + i < NLEV) {
          const double pk = p_prev + 0.5 * (dp_prev + d[i]);
          cp_[k0 + i] = pk;
          p_prev = pk;
          dp_prev = d[i];
          cx[k0 + i] = integ;
          integ = integ + dv[i];
        }
      }
    }
    if (VADV) s_tot[tid] = integ;  // assign_zero_to_sdot_sum + the sums of compute_eta_dot_dpdn_vertadv_euler :255-283
  }
  __syncthreads();

  // From here on the planes that outlive a phase are parked in shared memory, in the column-major
  // slots the scans use (thread (e, k) owns [col0 + p][k]): s_p pressure, s_x the hydrostatic
  // integrand -> phi, s_y the omega integral -> omega, s_z virtual temperature. Registers then hold
  // five planes at most, which is what keeps the kernel out of local memory.
  {
    double tv[NPSQ];
    plane_load(a.t + off_s(ie, a.n0) + k, tv);
    if (a.n0_qdp >= 0) {  // :348-363 (dry: Tv = T, :333-344)
      double q[NPSQ];
      plane_load(a.qdp + off_q(ie, a.n0_qdp, 0) + k, q);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        double Qt = q[p] / dp[p];
        Qt *= (Rwater_vapor / Rgas - 1.0);
        Qt += 1.0;
        tv[p] = tv[p] * Qt;
      }
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      if (VADV)  // compute_eta_dot_dpdn_vertadv_euler :285-299: hybi(k) sdot_sum - (sum of div_vdp above level k)
        s_e[slot] = k == 0 ? 0.0 : dc.hybi[k] * s_tot[col0 + p] - s_x[slot];
      s_y[slot] = s_x[slot] + 0.5 * div[p];  // integration + 0.5*div_vdp of preq_omega_ps
      s_x[slot] = Rgas * tv[p] * (dp[p] * 0.5 / s_p[slot]);  // preq_hydrostatic :689-729
      s_z[slot] = tv[p];
    }
  }
  // compute_dp3d_np1 :468-493 (eta_dot_dpdn == 0 for rsplit > 0); stored now, dp/div die here
  if (!VADV && valid) {
    double r[NPSQ];
    plane_load(a.dp3d + off_s(ie, a.nm1) + k, r);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      r[p] = geo_ld(g, p, G_SPHEREMP) * (r[p] - div[p] * a.dt);
      if (a.fold_rsp && is_interior_pt(p)) r[p] *= geo_ld(g, p, G_RSPHEREMP);
    }
    plane_store(a.dp3d + off_s(ie, a.np1) + k, r);
  }
  __syncthreads();
  if (VADV) {
    // Every read of a neighbouring level happens here, before the barrier below and hence before any
    // thread stores T or v at np1 (which may be the n0 level itself).
    double eta[NPSQ], eta1[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      eta[p] = s_e[slot];
      eta1[p] = k + 1 < NLEV ? s_e[slot + 1] : 0.0;
    }
    if (valid) {
      {  // accumulate_eta_dot_dpdn :150-163
        double* ed = a.eta_dot_dpdn + off_f(ie) + k;
        double r[NPSQ];
        plane_load(ed, r);
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) r[p] += a.eta_ave_w * eta[p];
        plane_store(ed, r);
      }
      {  // compute_dp3d_np1 :468-493
        double r[NPSQ];
        plane_load(a.dp3d + off_s(ie, a.nm1) + k, r);
        HXX_UNROLL
        for (int p = 0; p < NPSQ; ++p) {
          double tmp = eta1[p];
          tmp += div[p];
          tmp -= eta[p];
          r[p] = geo_ld(g, p, G_SPHEREMP) * (r[p] - tmp * a.dt);
          if (a.fold_rsp && is_interior_pt(p)) r[p] *= geo_ld(g, p, G_RSPHEREMP);
        }
        plane_store(a.dp3d + off_s(ie, a.np1) + k, r);
      }
    }
    // preq_vertadv :495-597 for this level: first / interior / last level forms
    const double* tp = a.t + off_s(ie, a.n0) + k;
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      const double* f[3] = {tp + p * NLEV, v0p + p * NLEV, v1p + p * NLEV};
      double* o[3] = {s_tv + slot, s_v0 + slot, s_v1 + slot};
      double facp = 0.0, facm = 0.0;
      if (k == 0) facp = (0.5 * 1 / dp[p]) * eta1[p];
      else if (k == NLEV - 1) facm = (0.5 * (1 / dp[p])) * eta[p];
      else {
        facp = 0.5 * (1 / dp[p]) * eta1[p];
        facm = 0.5 * (1 / dp[p]) * eta[p];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double x0 = f[c][0];
        if (k == 0) *o[c] = facp * (f[c][1] - x0);
        else if (k == NLEV - 1) *o[c] = facm * (x0 - f[c][-1]);
        else *o[c] = facp * (f[c][1] - x0) + facm * (x0 - f[c][-1]);
      }
    }
  }
  if (tid < E * NPSQ) {
    double* cx = s_x + tid * LS;
    const int iec = min(blockIdx.x * E + tid / NPSQ, a.nelem - 1);
    const double phis = s_geo[((iec - blockIdx.x * E) * NPSQ + (tid % NPSQ)) * GEO_N + G_PHIS];
    double integ = 0.0;
    for (int k0 = NLEV - 1; k0 >= 0; k0 -= 8) {
      double ak[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ak[i] = (k0 - i >= 0) ? cx[k0 - i] : 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k0 - i >= 0) {
          cx[k0 - i] = phis + 2.0 * integ + ak[i];
          integ = integ + ak[i];
        }
      }
    }
  }
  __syncthreads();

  // Each operator below keeps its input plane in registers and finishes one point at a time, so a
  // result goes to its slot (or to HBM) as soon as it exists.
  double v0[NPSQ], v1[NPSQ];
  plane_load(v0p, v0);
  plane_load(v1p, v1);
  {
    double pr[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) pr[p] = s_p[(col0 + p) * LS + k];
    phase_fence();
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      double g0, g1;
      gradient_point(g, pr, p, g0, g1);  // grad p
      const double vgrad_p = v0[p] * g0 + v1[p] * g1;
      s_y[slot] = (vgrad_p - s_y[slot]) / pr[p];  // omega
      const double r = Rgas * (s_z[slot] / pr[p]);  // compute_energy_grad :98-132
      s_p[slot] = r * g0;  // the pressure is in registers: its slot and the next plane take the gradient
      s_w[slot] = r * g1;
    }
  }
  phase_fence();
  if (a.eta_ave_w != 0.0 && valid) {  // compute_omega_p :412-423
    double* om = a.omega_p + off_f(ie) + k;
    double r[NPSQ];
    plane_load(om, r);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) r[p] += a.eta_ave_w * s_y[(col0 + p) * LS + k];
    plane_store(om, r);
  }
  phase_fence();
  // compute_velocity_np1 :184-232
  {
    double ephi[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const double phi = s_x[(col0 + p) * LS + k];
      if (a.store_phi && valid) a.phi[off_f(ie) + p * NLEV + k] = phi;
      ephi[p] = 0.5 * (v0[p] * v0[p] + v1[p] * v1[p]) + phi;
    }
    phase_fence();
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {  // gradient_sphere_update
      const int slot = (col0 + p) * LS + k;
      double g0, g1;
      gradient_point(g, ephi, p, g0, g1);
      s_p[slot] += g0;
      s_w[slot] += g1;
    }
  }
  phase_fence();
  {
    double c0[NPSQ], c1[NPSQ];  // vorticity_sphere :494-533, one point at a time
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      c0[p] = geo_ld(g, p, G_D00) * v0[p] + geo_ld(g, p, G_D01) * v1[p];
      c1[p] = geo_ld(g, p, G_D10) * v0[p] + geo_ld(g, p, G_D11) * v1[p];
    }
    const double* vm0 = a.v + off_v(ie, a.nm1, 0) + k;
    const double* vm1 = a.v + off_v(ie, a.nm1, 1) + k;
    double* s_o0 = s_x;  // phi is dead: its slot takes the second velocity component
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      double dvdx, dudy;
      deriv_point(c1, c0, p / NP, p % NP, dvdx, dudy);
      const double vort = (dvdx - dudy) * geo_ld(g, p, G_RMETDET_R);
      const double vt = vort + geo_ld(g, p, G_FCOR);
      double e0 = -s_p[slot] + ((VADV ? -s_v0[slot] : 0.0) + v1[p] * vt);
      double e1 = -s_w[slot] + ((VADV ? -s_v1[slot] : 0.0) - v0[p] * vt);
      e0 = e0 * a.dt + vm0[p * NLEV];
      e1 = e1 * a.dt + vm1[p * NLEV];
      const double sm_ = geo_ld(g, p, G_SPHEREMP);
      e0 = sm_ * e0;
      e1 = sm_ * e1;
      if (a.fold_rsp && is_interior_pt(p)) {
        const double rs = geo_ld(g, p, G_RSPHEREMP);
        e0 *= rs;
        e1 *= rs;
      }
      s_p[slot] = e0;  // to HBM after the loop: a store to v(np1) between the loads of v(nm1) would
      s_o0[slot] = e1; // serialise them (the two may alias as far as the compiler knows)
    }
    phase_fence();
    if (valid) {
      double* vp0 = a.v + off_v(ie, a.np1, 0) + k;
      double* vp1 = a.v + off_v(ie, a.np1, 1) + k;
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        vp0[p * NLEV] = s_p[(col0 + p) * LS + k];
        vp1[p * NLEV] = s_o0[(col0 + p) * LS + k];
      }
    }
  }
  phase_fence();
  {  // compute_temperature_np1 :430-463
    double tn0[NPSQ], r[NPSQ];
    plane_load(a.t + off_s(ie, a.n0) + k, tn0);
    plane_load(a.t + off_s(ie, a.nm1) + k, r);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const int slot = (col0 + p) * LS + k;
      double tg0, tg1;
      gradient_point(g, tn0, p, tg0, tg1);
      const double vgrad_t = v0[p] * tg0 + v1[p] * tg1;
      const double ttens = (VADV ? -s_tv[slot] : 0.0) - vgrad_t + kappa * s_z[slot] * s_y[slot];
      r[p] = ttens * a.dt + r[p];
      r[p] *= geo_ld(g, p, G_SPHEREMP);
      if (a.fold_rsp && is_interior_pt(p)) r[p] *= geo_ld(g, p, G_RSPHEREMP);
    }
    if (valid) plane_store(a.t + off_s(ie, a.np1) + k, r);
  }
}

void caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, bool with_dss) {
  if (!S.nelemd) return;
  CaarArgs a{S.geo, S.v, S.t, S.dp3d, S.derived_vn0, S.omega_p, S.phi, S.eta_dot_dpdn, S.qdp, S.nelemd, nm1, n0, np1, n0_qdp,
             dt, eta_ave_w, with_dss ? 1 : 0, S.store_phi ? 1 : 0};
  constexpr size_t plane = (size_t)CAAR_E * NPSQ * (NLEV + 1) * sizeof(double);
  constexpr size_t smem = 5 * plane, smem_vadv = 9 * plane;
  if (HXX_ONCE_PER_SESSION()) {
    CUDA_OK(cudaFuncSetAttribute(caar_kernel<CAAR_E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(caar_kernel<CAAR_E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_vadv));
  }
  const int nb = (S.nelemd + CAAR_E - 1) / CAAR_E;
  HXX_TIMER("caar compute");
  PROBE(K_CAAR);
  if (S.p.rsplit == 0) caar_kernel<CAAR_E, true><<<nb, CAAR_E * NLEV, smem_vadv, S.stream>>>(a);
  else caar_kernel<CAAR_E, false><<<nb, CAAR_E * NLEV, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(K_CAAR);
  if (with_dss) {
    HXX_TIMER("caar_bexchV");
    dss_exchange(fields_caar(np1), true);  // CaarFunctor.cpp:113
  }
}

// ---- element-wise kernels -----------------------------------------------------------------
// prim_advance_exp.cpp:143-154 : u(nm1) = (5 u(nm1) - u(n0)) / 4 for v (2 tiles), T, dp3d
__global__ void rk_combine_kernel(double* __restrict__ v, double* __restrict__ t, double* __restrict__ dp, int nm1,
                                  int n0) {
  const int ie = blockIdx.x;
  double* vm = v + off_v(ie, nm1, 0);
  const double* vn = v + off_v(ie, n0, 0);
  double* tm = t + off_s(ie, nm1);
  const double* tn = t + off_s(ie, n0);
  double* dm = dp + off_s(ie, nm1);
  const double* dn = dp + off_s(ie, n0);
  for (int i = threadIdx.x; i < 2 * NLF; i += blockDim.x) vm[i] = (5.0 * vm[i] - vn[i]) / 4.0;
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) {
    tm[i] = (5.0 * tm[i] - tn[i]) / 4.0;
    dm[i] = (5.0 * dm[i] - dn[i]) / 4.0;
  }
}
void rk_combine(int nm1, int n0) {
  if (!S.nelemd) return;
  PROBE(K_RK_COMBINE);
  rk_combine_kernel<<<S.nelemd, 288, 0, S.stream>>>(S.v, S.t, S.dp3d, nm1, n0);
  KERNEL_LAUNCHED(K_RK_COMBINE);
}

// prim_driver.cpp:98-111
__global__ void dp3d_from_ps_kernel(double* __restrict__ dp3d, const double* __restrict__ ps_v, int n0) {
  const int ie = blockIdx.x;
  double* dp = dp3d + off_s(ie, n0);
  const double* ps = ps_v + ((size_t)ie * NTL + n0) * NPSQ;
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) {
    const int p = i / NLEV, k = i % NLEV;
    dp[i] = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps[p];
  }
}
void dp3d_from_ps(int n0) {
  if (!S.nelemd) return;
  PROBE(K_DP3D_FROM_PS);
  dp3d_from_ps_kernel<<<S.nelemd, 288, 0, S.stream>>>(S.dp3d, S.ps_v, n0);
  KERNEL_LAUNCHED(K_DP3D_FROM_PS);
}

// prim_step.cpp:51-66
__global__ void derived_dp_kernel(double* __restrict__ derived_dp, const double* __restrict__ dp3d, int n0) {
  const int ie = blockIdx.x;
  const double* s = dp3d + off_s(ie, n0);
  double* d = derived_dp + off_f(ie);
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) d[i] = s[i];
}
void prim_step_init(int n0) {
  if (!S.nelemd) return;
  const size_t f3 = (size_t)S.nelemd * NLF * sizeof(double);
  CUDA_OK(cudaMemsetAsync(S.eta_dot_dpdn, 0, f3, S.stream));
  CUDA_OK(cudaMemsetAsync(S.derived_vn0, 0, 2 * f3, S.stream));
  CUDA_OK(cudaMemsetAsync(S.omega_p, 0, f3, S.stream));
  if (S.p.nu_p > 0) {
    CUDA_OK(cudaMemsetAsync(S.dpdiss_ave, 0, f3, S.stream));
    CUDA_OK(cudaMemsetAsync(S.dpdiss_biharmonic, 0, f3, S.stream));
  }
  PROBE(K_STEP_INIT);
  derived_dp_kernel<<<S.nelemd, 288, 0, S.stream>>>(S.derived_dp, S.dp3d, n0);
  KERNEL_LAUNCHED(K_STEP_INIT);
}

// prim_driver.cpp:171-206
__global__ void update_q_kernel(double* __restrict__ Q, const double* __restrict__ qdp, const double* __restrict__ ps_v,
                                int np1_qdp, int np1) {
  const int ie = blockIdx.x, q = blockIdx.y;
  const double* qd = qdp + off_q(ie, np1_qdp, q);
  double* out = Q + ((size_t)ie * QSIZE_D + q) * NLF;
  const double* ps = ps_v + ((size_t)ie * NTL + np1) * NPSQ;
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) {
    const int p = i / NLEV, k = i % NLEV;
    const double dp = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps[p];
    out[i] = qd[i] / dp;
  }
}
void update_q(int np1_qdp, int np1) {
  if (!S.nelemd || !S.p.qsize) return;
  PROBE(K_UPDATE_Q);
  update_q_kernel<<<dim3(S.nelemd, S.p.qsize), 288, 0, S.stream>>>(S.Q, S.qdp, S.ps_v, np1_qdp, np1);
  KERNEL_LAUNCHED(K_UPDATE_Q);
}

}  // namespace hxx
