// CAM forcing and diagnostics — replaces CamForcing.cpp and Diagnostics.cpp of the reference.
//
// Forcing (CamForcing.cpp:20-174) runs once per prim_run_subcycle_c call whenever ftype is 0 or 2,
// which includes every standalone benchmark namelist (se_ftype = 0, forcing arrays all zero). The
// reference's three tracer passes (qdp += dt FQ with the negativity clamp, then Q = qdp / dp) are
// one pass here; the moist surface-pressure sum keeps the reference's ascending-level order with
// one thread per column.
// Diagnostics (Diagnostics.cpp:37-185) are host loops over device mirrors in the reference; here
// the column sums run on the device, one thread per column in the same ascending order, and only
// the sums travel to the Fortran-owned accumulators.
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

namespace hxx {

// state_forcing :20-49
__global__ void state_forcing_kernel(double* __restrict__ v, double* __restrict__ t, const double* __restrict__ fm,
                                     const double* __restrict__ ft, int np1, double dt) {
  const int ie = blockIdx.x;
  double* tt = t + off_s(ie, np1);
  const double* f = ft + off_f(ie);
  for (int i = threadIdx.x; i < NLF; i += blockDim.x) tt[i] += dt * f[i];
  double* vv = v + off_v(ie, np1, 0);
  const double* g = fm + (size_t)ie * 2 * NLF;
  for (int i = threadIdx.x; i < 2 * NLF; i += blockDim.x) vv[i] += dt * g[i];
}

__device__ __forceinline__ double clamped_increment(double qs, double v1) {  // :92-98, :118-124
  if (qs + v1 < 0.0 && v1 < 0.0) v1 = qs < 0.0 ? 0.0 : -qs;
  return v1;
}

// tracer_forcing :65-108 (moist): ps_v(np1) += sum_k clamped dt FQ(tracer 0); one thread per column
__global__ void tracer_forcing_ps_kernel(double* __restrict__ ps_v, const double* __restrict__ fq,
                                         const double* __restrict__ qdp, int nelem, int np1, int np1_qdp, double dt) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nelem * NPSQ) return;
  const int ie = g / NPSQ, p = g % NPSQ;
  const double* f = fq ? fq + (size_t)ie * QSIZE_D * NLF + p * NLEV : nullptr;
  const double* q = qdp + off_q(ie, np1_qdp, 0) + p * NLEV;
  double acc = 0.0;
  for (int k = 0; k < NLEV; ++k) acc += clamped_increment(q[k], dt * (f ? f[k] : 0.0));
  ps_v[((size_t)ie * NTL + np1) * NPSQ + p] += acc;
}

// tracer_forcing :110-146: qdp += clamped dt FQ, then Q = qdp / dp(ps_v). One thread per (element,
// point, level) walks the tracers four at a time (16 independent 8-byte requests in flight); the layer
// thickness and its reciprocal are computed once per thread, the quotient is div_rcp's.
__global__ void __launch_bounds__(256) tracer_forcing_kernel(double* __restrict__ qdp, double* __restrict__ Q,
                                                             const double* __restrict__ fq,
                                                             const double* __restrict__ ps_v, int nelem, int qsize,
                                                             int np1, int np1_qdp, double dt) {
  const long long gidx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gidx >= (long long)nelem * NLF) return;
  const int ie = (int)(gidx / NLF), i = (int)(gidx % NLF);
  const int p = i / NLEV, k = i % NLEV;
  const double dp = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps_v[((size_t)ie * NTL + np1) * NPSQ + p];
  const double rdp = 1.0 / dp;
  double* qd = qdp + off_q(ie, np1_qdp, 0) + i;
  double* out = Q + (size_t)ie * QSIZE_D * NLF + i;
  const double* f = fq ? fq + (size_t)ie * QSIZE_D * NLF + i : nullptr;  // null: no tracer forcing was ever pushed
  int q = 0;
  for (; q + 4 <= qsize; q += 4) {
    double qs[4], fv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      qs[j] = qd[(size_t)(q + j) * NLF];
      fv[j] = f ? f[(size_t)(q + j) * NLF] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double r = qs[j] + clamped_increment(qs[j], dt * fv[j]);
      if (f) qd[(size_t)(q + j) * NLF] = r;  // no forcing ever pushed: qdp + 0 is qdp, nothing to write back
      out[(size_t)(q + j) * NLF] = div_rcp(r, dp, rdp);
    }
  }
  for (; q < qsize; ++q) {
    const double qs = qd[(size_t)q * NLF];
    const double r = qs + clamped_increment(qs, dt * (f ? f[(size_t)q * NLF] : 0.0));
    if (f) qd[(size_t)q * NLF] = r;
    out[(size_t)q * NLF] = div_rcp(r, dp, rdp);
  }
}

void apply_cam_forcing(double dt, bool tracers) {
  if (!S.nelemd) return;
  const size_t f3 = (size_t)S.nelemd * NLF;
  auto zeros = [&](double*& p, size_t n) {
    if (p) return;
    CUDA_OK(cudaMalloc(&p, n * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(double), S.stream));
  };
  zeros(S.fm, f3 * 2);
  zeros(S.ft, f3);
  PROBE(K_FORCING);
  state_forcing_kernel<<<S.nelemd, 288, 0, S.stream>>>(S.v, S.t, S.fm, S.ft, S.n0, dt);
  KERNEL_LAUNCHED(K_FORCING);
  if (!tracers) return;
  // CamForcing.cpp:158-160 allocates a zero FQ on first use; here a null FQ means zero tracer forcing, so a
  // standalone run (which never pushes forcing) neither holds nor reads 40 tiles of zeros per element
  if (S.p.moist) {
    PROBE(K_FORCING);
    tracer_forcing_ps_kernel<<<(S.nelemd * NPSQ + 127) / 128, 128, 0, S.stream>>>(S.ps_v, S.fq, S.qdp, S.nelemd, S.n0,
                                                                                   S.n0_qdp, dt);
    KERNEL_LAUNCHED(K_FORCING);
  }
  if (S.p.qsize > 0) {
    PROBE(K_FORCING);
    const long long nthr = (long long)S.nelemd * NLF;
    tracer_forcing_kernel<<<(unsigned)((nthr + 255) / 256), 256, 0, S.stream>>>(S.qdp, S.Q, S.fq, S.ps_v, S.nelemd,
                                                                             S.p.qsize, S.n0, S.n0_qdp, dt);
    KERNEL_LAUNCHED(K_FORCING);
  }
}

// ---- Held-Suarez forcing on the device -------------------------------------------------------------
// hs_T_forcing (held_suarez_mod.F90:175-279, scalar branch) and hs_v_forcing (:123-173) at time level n0, one thread
// per (element, point, level); FM, FT are overwritten (the Fortran driver zeroes them before hs_forcing adds).
__global__ void __launch_bounds__(256) held_suarez_kernel(const double* __restrict__ v, const double* __restrict__ t,
                                                          const double* __restrict__ ps_v, const double* __restrict__ lat,
                                                          const double* __restrict__ hy, double* __restrict__ fm,
                                                          double* __restrict__ ft, int nelem, int n0) {
  const long long gidx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (gidx >= (long long)nelem * NLF) return;
  const int ie = (int)(gidx / NLF), i = (int)(gidx % NLF);
  const int p = i / NLEV, k = i % NLEV;
  constexpr double sigma_b = 0.70, secpday = 86400.0;
  constexpr double k_a = 1.0 / (40.0 * secpday), k_f = 1.0 / (1.0 * secpday), k_s = 1.0 / (4.0 * secpday);
  constexpr double dT_y = 60.0, dtheta_z = 10.0;
  const double hyam = hy[k], hybm = hy[NLEV + k];
  const double ps = ps_v[((size_t)ie * NTL + n0) * NPSQ + p];
  const double snlat = sin(lat[(size_t)ie * NPSQ + p]);
  const double snlatsq = snlat * snlat, cslatsq = 1.0 - snlatsq;
  const double pm = hyam * dc.ps0 + hybm * ps;
  const double logprat = log(pm) - log(dc.ps0);
  const double pratk = exp(kappa * logprat);
  const double etam = hyam + hybm;
  const double ramp = fmax(0.0, (etam - sigma_b) / (1.0 - sigma_b));
  const double k_t = k_a + (k_s - k_a) * cslatsq * cslatsq * ramp;
  const double Teq = fmax(200.0, (315.0 - dT_y * snlatsq - dtheta_z * logprat * cslatsq) * pratk);
  ft[off_f(ie) + i] = -k_t * (t[off_s(ie, n0) + i] - Teq);
  const double k_v = k_f * ramp;
  fm[((size_t)ie * 2 + 0) * NLF + i] = -k_v * v[off_v(ie, n0, 0) + i];
  fm[((size_t)ie * 2 + 1) * NLF + i] = -k_v * v[off_v(ie, n0, 1) + i];
}

void held_suarez_forcing(const double* lat, const double* hyam, const double* hybm) {
  if (!S.nelemd) return;
  const size_t f3 = (size_t)S.nelemd * NLF;
  if (!S.hs_lat) {
    if (!lat || !hyam || !hybm) runtime_abort("hxx_held_suarez_forcing: the first call of a session needs lat, hyam, hybm", 13);
    CUDA_OK(cudaMalloc(&S.hs_lat, (size_t)S.nelemd * NPSQ * sizeof(double)));
    CUDA_OK(cudaMalloc(&S.hs_hyam, 2 * NLEV * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(S.hs_lat, lat, (size_t)S.nelemd * NPSQ * sizeof(double), cudaMemcpyHostToDevice, S.stream));
    CUDA_OK(cudaMemcpyAsync(S.hs_hyam, hyam, NLEV * sizeof(double), cudaMemcpyHostToDevice, S.stream));
    CUDA_OK(cudaMemcpyAsync(S.hs_hyam + NLEV, hybm, NLEV * sizeof(double), cudaMemcpyHostToDevice, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
  }
  if (!S.fm) CUDA_OK(cudaMalloc(&S.fm, f3 * 2 * sizeof(double)));
  if (!S.ft) CUDA_OK(cudaMalloc(&S.ft, f3 * sizeof(double)));
  PROBE(K_FORCING);
  held_suarez_kernel<<<(unsigned)((f3 + 255) / 256), 256, 0, S.stream>>>(S.v, S.t, S.ps_v, S.hs_lat, S.hs_hyam, S.fm, S.ft,
                                                                      S.nelemd, S.n0);
  KERNEL_LAUNCHED(K_FORCING);
}

// ---- diagnostics ------------------------------------------------------------------------------
// prim_diag_scalars :37-90: out[0] = sum_k qdp Q, out[1] = sum_k qdp, per (element, tracer, point)
__global__ void diag_scalars_kernel(const double* __restrict__ qdp, const double* __restrict__ Q, double* __restrict__ out,
                                    int nelem, int qsize, int t2_qdp) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nelem * qsize * NPSQ) return;
  const int p = g % NPSQ, q = (g / NPSQ) % qsize, ie = g / (NPSQ * qsize);
  const double* a = qdp + off_q(ie, t2_qdp, q) + p * NLEV;
  const double* b = Q + ((size_t)ie * QSIZE_D + q) * NLF + p * NLEV;
  double s_qq = 0, s_q = 0;
  for (int k = 0; k < NLEV; ++k) {
    s_qq += a[k] * b[k];
    s_q += a[k];
  }
  out[2 * (size_t)g] = s_qq;
  out[2 * (size_t)g + 1] = s_q;
}

// prim_energy_halftimes :92-185: out = {IEner, IEner_wet, KEner, PEner} per (element, point)
__global__ void diag_energy_kernel(const double* __restrict__ v, const double* __restrict__ t,
                                   const double* __restrict__ ps_v, const double* __restrict__ qdp,
                                   const double* __restrict__ geo, double* __restrict__ out, int nelem, int t1, int t1_qdp,
                                   int use_cpstar) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nelem * NPSQ) return;
  const int ie = g / NPSQ, p = g % NPSQ;
  const double* u = v + off_v(ie, t1, 0) + p * NLEV;
  const double* w = v + off_v(ie, t1, 1) + p * NLEV;
  const double* T = t + off_s(ie, t1) + p * NLEV;
  const double* q0 = qdp + off_q(ie, t1_qdp, 0) + p * NLEV;
  const double ps = ps_v[((size_t)ie * NTL + t1) * NPSQ + p];
  const double phis = geo[((size_t)ie * NPSQ + p) * GEO_N + G_PHIS];
  constexpr double Cpwater_vapor = 1870.0;  // PhysicalConstants.hpp:21
  double IEner = 0.0, IEner_wet = 0.0, KEner = 0.0, PEner = 0.0;
  for (int k = 0; k < NLEV; ++k) {
    const double dpt1 = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps;
    double cp_star1 = cp;
    if (use_cpstar) {
      const double qval = q0[k] / dpt1;
      cp_star1 = cp * (1.0 + (Cpwater_vapor / cp - 1.0) * qval);
    }
    IEner += cp_star1 * T[k] * dpt1;
    IEner_wet += (cp_star1 - cp) * T[k] * dpt1;
    KEner += (u[k] * u[k] + w[k] * w[k]) * 0.5 * dpt1;
    PEner += phis * dpt1;
  }
  double* o = out + 4 * (size_t)g;
  o[0] = IEner; o[1] = IEner_wet; o[2] = KEner; o[3] = PEner;
}

static std::vector<double> g_host;  // staging of the sums on their way to the F90 accumulators

void prim_diag_scalars(bool before_advance, int ivar) {
  const int t2_qdp = before_advance ? S.n0_qdp : S.np1_qdp;
  if (S.p.time_step_type <= 0 || !S.nelemd || !S.diag[0]) return;
  const int n = S.nelemd, nq = S.p.qsize;
  push_Q_to_host(S.diag[0]);  // sync_to_host(tracers.Q, h_Q) :60-62
  if (!nq) return;
  const size_t cnt = (size_t)n * nq * NPSQ;
  double* d = (double*)diag_scratch(cnt * 2 * sizeof(double));
  PROBE(K_DIAG);
  diag_scalars_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, S.stream>>>(S.qdp, S.Q, d, n, nq, t2_qdp);
  KERNEL_LAUNCHED(K_DIAG);
  g_host.resize(cnt * 2);
  CUDA_OK(cudaMemcpyAsync(g_host.data(), d, cnt * 2 * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  double *Qvar = S.diag[1], *Qmass = S.diag[2], *Q1mass = S.diag[3];  // [ie][4][QSIZE_D][16], [ie][QSIZE_D][16]
  for (int ie = 0; ie < n; ++ie)
    for (int q = 0; q < nq; ++q)
      for (int p = 0; p < NPSQ; ++p) {
        const double* s = &g_host[2 * (((size_t)ie * nq + q) * NPSQ + p)];
        Qvar[(((size_t)ie * 4 + ivar) * QSIZE_D + q) * NPSQ + p] = s[0];
        Qmass[(((size_t)ie * 4 + ivar) * QSIZE_D + q) * NPSQ + p] = s[1];
        Q1mass[((size_t)ie * QSIZE_D + q) * NPSQ + p] = s[1];
      }
}

void prim_energy_halftimes(bool before_advance, int ivar) {
  const int t1 = before_advance ? S.n0 : S.np1, t1_qdp = before_advance ? S.n0_qdp : S.np1_qdp;
  if (!S.nelemd || !S.diag[4]) return;
  const int n = S.nelemd;
  const size_t cnt = (size_t)n * NPSQ;
  double* d = (double*)diag_scratch(cnt * 4 * sizeof(double));
  PROBE(K_DIAG);
  diag_energy_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, S.stream>>>(S.v, S.t, S.ps_v, S.qdp, S.geo, d, n, t1, t1_qdp,
                                                                         S.p.use_cpstar ? 1 : 0);
  KERNEL_LAUNCHED(K_DIAG);
  g_host.resize(cnt * 4);
  CUDA_OK(cudaMemcpyAsync(g_host.data(), d, cnt * 4 * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  double *IE = S.diag[4], *IEw = S.diag[5], *KE = S.diag[6], *PE = S.diag[7];  // [ie][4][16]; IEner_wet [ie][16]
  for (int ie = 0; ie < n; ++ie)
    for (int p = 0; p < NPSQ; ++p) {
      const double* s = &g_host[4 * ((size_t)ie * NPSQ + p)];
      IE[((size_t)ie * 4 + ivar) * NPSQ + p] = s[0];
      IEw[(size_t)ie * NPSQ + p] = s[1];
      KE[((size_t)ie * 4 + ivar) * NPSQ + p] = s[2];
      PE[((size_t)ie * 4 + ivar) * NPSQ + p] = s[3];
    }
}

}  // namespace hxx
