// Element-local spectral-element operators on one level held in registers: s[16], p = igp*4+jgp.
// Restates SphereOperators.hpp of the reference (file:line per function) with the reference's
// operation order, so the results are bit-identical to the CPU oracle when compiled with
// --fmad=false. The including translation unit must have HXX_DEFINE_CONSTANTS() above.
#pragma once
#include "hxx.cuh"

namespace hxx {

#define HXX_UNROLL _Pragma("unroll")

// The element's geometry record is read either from global memory through the read-only path
// (GeoGlobal) or from the block's copy of it in shared memory (GeoShared, see stage_geo). The
// operators are templates on the accessor so each kernel picks what suits its register budget.
struct GeoGlobal {
  const double* __restrict__ p;
  __device__ __forceinline__ double ld(int i) const { return __ldg(p + i); }
};
struct GeoShared {
  const double* p;
  __device__ __forceinline__ double ld(int i) const { return p[i]; }
};
template <class G>
__device__ __forceinline__ double geo_ld(const G& g, int p, int c) { return g.ld(p * GEO_N + c); }
// plain-pointer form (global memory)
__device__ __forceinline__ double geo_ld(const double* __restrict__ g, int p, int c) { return __ldg(g + p * GEO_N + c); }

// Cooperative copy of the geometry records of the (at most NE) elements a block of the flat
// (element, level) mapping touches. Every thread of the block must call it (it has a barrier).
// Shared-memory reads do not queue behind the block's outstanding HBM requests in the L1
// pipeline the way L1-hit global loads do, which matters because the operators re-read the
// geometry for every tracer / field.
template <int NE, int NT>
__device__ __forceinline__ void stage_geo(double* s_geo, const double* __restrict__ geo, int e_first, int nelem) {
  for (int i = threadIdx.x; i < NE * NPSQ * GEO_N; i += NT) {
    const int e = e_first + i / (NPSQ * GEO_N);
    if (e < nelem) s_geo[i] = __ldg(geo + (size_t)e_first * NPSQ * GEO_N + i);
  }
  __syncthreads();
}
// elements spanned by NT consecutive (element, level) threads
__host__ __device__ constexpr int geo_span(int nt) { return (nt + NLEV - 2) / NLEV + 1; }

// load / store one level of a field tile ([16][NLEV], the thread's level already added to ptr)
__device__ __forceinline__ void plane_load(const double* __restrict__ f, double (&s)[NPSQ]) {
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) s[p] = f[p * NLEV];
}
__device__ __forceinline__ void plane_store(double* __restrict__ f, const double (&s)[NPSQ]) {
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) f[p * NLEV] = s[p];
}

// dx[p] = sum_m Dvv(j,m) s(i,m) ; dy[p] = sum_m Dvv(i,m) s(m,j)   (m ascending)
__device__ __forceinline__ void deriv_pair(const double (&sx)[NPSQ], const double (&sy)[NPSQ], double (&dx)[NPSQ],
                                           double (&dy)[NPSQ]) {
  HXX_UNROLL
  for (int i = 0; i < NP; ++i) {
    HXX_UNROLL
    for (int j = 0; j < NP; ++j) {
      double a = dc.dvv[j][0] * sx[i * NP + 0];
      double b = dc.dvv[i][0] * sy[0 * NP + j];
      HXX_UNROLL
      for (int m = 1; m < NP; ++m) {
        a += dc.dvv[j][m] * sx[i * NP + m];
        b += dc.dvv[i][m] * sy[m * NP + j];
      }
      dx[i * NP + j] = a;
      dy[i * NP + j] = b;
    }
  }
}

// one point of deriv_pair (same sums, same order)
__device__ __forceinline__ void deriv_point(const double (&sx)[NPSQ], const double (&sy)[NPSQ], int i, int j, double& a,
                                            double& b) {
  a = dc.dvv[j][0] * sx[i * NP + 0];
  b = dc.dvv[i][0] * sy[0 * NP + j];
  HXX_UNROLL
  for (int m = 1; m < NP; ++m) {
    a += dc.dvv[j][m] * sx[i * NP + m];
    b += dc.dvv[i][m] * sy[m * NP + j];
  }
}
// one point of gradient_sphere (:293-319)
template <class G>
__device__ __forceinline__ void gradient_point(const G& g, const double (&s)[NPSQ], int p, double& g0, double& g1) {
  double dx, dy;
  deriv_point(s, s, p / NP, p % NP, dx, dy);
  const double v0 = dx * rrearth, v1 = dy * rrearth;
  g0 = geo_ld(g, p, G_DINV00) * v0 + geo_ld(g, p, G_DINV01) * v1;
  g1 = geo_ld(g, p, G_DINV10) * v0 + geo_ld(g, p, G_DINV11) * v1;
}
// compiler-level fence: keeps loads and stores (and with them the live ranges they start or end) on
// their side of a phase boundary
__device__ __forceinline__ void phase_fence() { asm volatile("" ::: "memory"); }

// SphereOperators.hpp:293-319
template <class G>
__device__ __forceinline__ void gradient_sphere(const G& g, const double (&s)[NPSQ],
                                                double (&g0)[NPSQ], double (&g1)[NPSQ]) {
  double dx[NPSQ], dy[NPSQ];
  deriv_pair(s, s, dx, dy);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double v0 = dx[p] * rrearth, v1 = dy[p] * rrearth;
    g0[p] = geo_ld(g, p, G_DINV00) * v0 + geo_ld(g, p, G_DINV01) * v1;
    g1[p] = geo_ld(g, p, G_DINV10) * v0 + geo_ld(g, p, G_DINV11) * v1;
  }
}

// :323-348
template <class G>
__device__ __forceinline__ void gradient_sphere_update(const G& g, const double (&s)[NPSQ],
                                                       double (&g0)[NPSQ], double (&g1)[NPSQ]) {
  double dx[NPSQ], dy[NPSQ];
  deriv_pair(s, s, dx, dy);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double v0 = dx[p] * rrearth, v1 = dy[p] * rrearth;
    g0[p] += geo_ld(g, p, G_DINV00) * v0 + geo_ld(g, p, G_DINV01) * v1;
    g1[p] += geo_ld(g, p, G_DINV10) * v0 + geo_ld(g, p, G_DINV11) * v1;
  }
}

// :352-392
template <class G>
__device__ __forceinline__ void divergence_sphere(const G& g, const double (&v0)[NPSQ],
                                                  const double (&v1)[NPSQ], double (&div)[NPSQ]) {
  double gv0[NPSQ], gv1[NPSQ];
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double md = geo_ld(g, p, G_METDET);
    gv0[p] = (geo_ld(g, p, G_DINV00) * v0[p] + geo_ld(g, p, G_DINV10) * v1[p]) * md;
    gv1[p] = (geo_ld(g, p, G_DINV01) * v0[p] + geo_ld(g, p, G_DINV11) * v1[p]) * md;
  }
  double dx[NPSQ], dy[NPSQ];
  deriv_pair(gv0, gv1, dx, dy);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) div[p] = (dx[p] + dy[p]) * geo_ld(g, p, G_RMETDET_R);
}

// :494-533
template <class G>
__device__ __forceinline__ void vorticity_sphere(const G& g, const double (&u)[NPSQ],
                                                 const double (&v)[NPSQ], double (&vort)[NPSQ]) {
  double c0[NPSQ], c1[NPSQ];
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    c0[p] = geo_ld(g, p, G_D00) * u[p] + geo_ld(g, p, G_D01) * v[p];
    c1[p] = geo_ld(g, p, G_D10) * u[p] + geo_ld(g, p, G_D11) * v[p];
  }
  double dvdx[NPSQ], dudy[NPSQ];
  deriv_pair(c1, c0, dvdx, dudy);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) vort[p] = (dvdx[p] - dudy[p]) * geo_ld(g, p, G_RMETDET_R);
}

// :538-583 (inputs are the sphere-basis vector; transformed internally)
template <class G>
__device__ __forceinline__ void divergence_sphere_wk(const G& g, const double (&v0)[NPSQ],
                                                     const double (&v1)[NPSQ], double (&div)[NPSQ]) {
  double s0[NPSQ], s1[NPSQ];  // spheremp * (Dinv^T v)
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double w0 = geo_ld(g, p, G_DINV00) * v0[p] + geo_ld(g, p, G_DINV10) * v1[p];
    const double w1 = geo_ld(g, p, G_DINV01) * v0[p] + geo_ld(g, p, G_DINV11) * v1[p];
    const double sm = geo_ld(g, p, G_SPHEREMP);
    s0[p] = sm * w0;
    s1[p] = sm * w1;
  }
  HXX_UNROLL
  for (int n = 0; n < NP; ++n) {
    HXX_UNROLL
    for (int m = 0; m < NP; ++m) {
      double dd = -((s0[n * NP + 0] * dc.dvv[0][m] + s1[0 * NP + m] * dc.dvv[0][n]) * rrearth);
      HXX_UNROLL
      for (int j = 1; j < NP; ++j)
        dd -= (s0[n * NP + j] * dc.dvv[j][m] + s1[j * NP + m] * dc.dvv[j][n]) * rrearth;
      div[n * NP + m] = dd;
    }
  }
}

// :588-597
template <class G>
__device__ __forceinline__ void laplace_simple(const G& g, const double (&s)[NPSQ],
                                               double (&lap)[NPSQ]) {
  double g0[NPSQ], g1[NPSQ];
  gradient_sphere(g, s, g0, g1);
  divergence_sphere_wk(g, g0, g1, lap);
}

// :604-635 — tv = this element's tensorVisc [2][2][16]
template <class G>
__device__ __forceinline__ void laplace_tensor(const G& g, const double* __restrict__ tv,
                                               const double (&s)[NPSQ], double (&lap)[NPSQ]) {
  double g0[NPSQ], g1[NPSQ], t0[NPSQ], t1[NPSQ];
  gradient_sphere(g, s, g0, g1);
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    t0[p] = __ldg(tv + 0 * NPSQ + p) * g0[p] + __ldg(tv + 2 * NPSQ + p) * g1[p];
    t1[p] = __ldg(tv + 1 * NPSQ + p) * g0[p] + __ldg(tv + 3 * NPSQ + p) * g1[p];
  }
  divergence_sphere_wk(g, t0, t1, lap);
}

// :714-748 — mi = this element's metinv [2][2][16]
template <class G>
__device__ __forceinline__ void grad_sphere_wk_testcov(const G& g, const double* __restrict__ mi,
                                                       const double (&s)[NPSQ], double (&g0)[NPSQ],
                                                       double (&g1)[NPSQ]) {
  HXX_UNROLL
  for (int n = 0; n < NP; ++n) {
    HXX_UNROLL
    for (int m = 0; m < NP; ++m) {
      const int p = n * NP + m;
      const double md = geo_ld(g, p, G_METDET);
      const double mi00 = __ldg(mi + 0 * NPSQ + p), mi01 = __ldg(mi + 1 * NPSQ + p), mi10 = __ldg(mi + 2 * NPSQ + p),
                   mi11 = __ldg(mi + 3 * NPSQ + p);
      double b0 = 0.0, b1 = 0.0;
      HXX_UNROLL
      for (int j = 0; j < NP; ++j) {
        const double mpnj = geo_ld(g, n * NP + j, G_MP), mpjm = geo_ld(g, j * NP + m, G_MP);
        const double snj = s[n * NP + j], sjm = s[j * NP + m];
        const double djm = dc.dvv[j][m], djn = dc.dvv[j][n];
        const double x0 = mpnj * mi00 * md * snj * djm + mpjm * mi01 * md * sjm * djn;
        const double x1 = mpnj * mi10 * md * snj * djm + mpjm * mi11 * md * sjm * djn;
        if (j == 0) { b0 = -x0; b1 = -x1; }
        else { b0 -= x0; b1 -= x1; }
      }
      g0[p] = (geo_ld(g, p, G_D00) * b0 + geo_ld(g, p, G_D10) * b1) * rrearth;
      g1[p] = (geo_ld(g, p, G_D01) * b0 + geo_ld(g, p, G_D11) * b1) * rrearth;
    }
  }
}

// :683-710
template <class G>
__device__ __forceinline__ void curl_sphere_wk_testcov_update(const G& g, double alpha, double beta,
                                                              const double (&s)[NPSQ], double (&c0)[NPSQ],
                                                              double (&c1)[NPSQ]) {
  double ms[NPSQ];
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) ms[p] = geo_ld(g, p, G_MP) * s[p];
  HXX_UNROLL
  for (int n = 0; n < NP; ++n) {
    HXX_UNROLL
    for (int m = 0; m < NP; ++m) {
      const int p = n * NP + m;
      double sb0 = -(ms[0 * NP + m] * dc.dvv[0][n]);
      double sb1 = ms[n * NP + 0] * dc.dvv[0][m];
      HXX_UNROLL
      for (int j = 1; j < NP; ++j) {
        sb0 -= ms[j * NP + m] * dc.dvv[j][n];
        sb1 += ms[n * NP + j] * dc.dvv[j][m];
      }
      c0[p] = beta * c0[p] + alpha * (geo_ld(g, p, G_D00) * sb0 + geo_ld(g, p, G_D10) * sb1) * rrearth;
      c1[p] = beta * c1[p] + alpha * (geo_ld(g, p, G_D01) * sb0 + geo_ld(g, p, G_D11) * sb1) * rrearth;
    }
  }
}

// :818-862
template <class G>
__device__ __forceinline__ void vlaplace_sphere_wk_contra(const G& g, const double* __restrict__ mi,
                                                          double nu_ratio, const double (&v0)[NPSQ],
                                                          const double (&v1)[NPSQ], double (&l0)[NPSQ],
                                                          double (&l1)[NPSQ]) {
  double sc[NPSQ], gc0[NPSQ], gc1[NPSQ];
  divergence_sphere(g, v0, v1, sc);
  if (nu_ratio > 0 && nu_ratio != 1.0) {
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) sc[p] *= nu_ratio;
  }
  grad_sphere_wk_testcov(g, mi, sc, gc0, gc1);
  vorticity_sphere(g, v0, v1, sc);
  curl_sphere_wk_testcov_update(g, -1.0, 1.0, sc, gc0, gc1);
  const double re2 = rrearth * rrearth;
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double f = 2.0 * geo_ld(g, p, G_SPHEREMP);
    l0[p] = f * v0[p] * re2 + gc0[p];
    l1[p] = f * v1[p] * re2 + gc1[p];
  }
}

// The same operator with the vector field re-read from memory (level already added to the
// pointers) for each of its three uses instead of being held in registers across the whole
// operator: 32 fewer live doubles, and the re-reads hit L1.
template <class G>
__device__ __forceinline__ void vlaplace_sphere_wk_contra_mem(const G& g, const double* __restrict__ mi,
                                                              double nu_ratio, const double* v0p, const double* v1p,
                                                              double (&l0)[NPSQ], double (&l1)[NPSQ]) {
  double sc[NPSQ];
  {
    double v0[NPSQ], v1[NPSQ];
    plane_load(v0p, v0);
    plane_load(v1p, v1);
    divergence_sphere(g, v0, v1, sc);
  }
  if (nu_ratio > 0 && nu_ratio != 1.0) {
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) sc[p] *= nu_ratio;
  }
  grad_sphere_wk_testcov(g, mi, sc, l0, l1);
  {
    double v0[NPSQ], v1[NPSQ];
    plane_load(v0p, v0);
    plane_load(v1p, v1);
    vorticity_sphere(g, v0, v1, sc);
  }
  curl_sphere_wk_testcov_update(g, -1.0, 1.0, sc, l0, l1);
  const double re2 = rrearth * rrearth;
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double f = 2.0 * geo_ld(g, p, G_SPHEREMP);
    l0[p] = f * v0p[p * NLEV] * re2 + l0[p];
    l1[p] = f * v1p[p * NLEV] * re2 + l1[p];
  }
}

// :752-814 — vs = this element's vec_sph2cart [2][3][16]
template <class G>
__device__ __forceinline__ void vlaplace_sphere_wk_cartesian(const G& g, const double* __restrict__ tv,
                                                             const double* __restrict__ vs, const double (&v0)[NPSQ],
                                                             const double (&v1)[NPSQ], double (&l0)[NPSQ],
                                                             double (&l1)[NPSQ]) {
  double acc0[NPSQ], acc1[NPSQ];
  for (int c = 0; c < 3; ++c) {
    double comp[NPSQ], lap[NPSQ];
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p)
      comp[p] = __ldg(vs + (0 * 3 + c) * NPSQ + p) * v0[p] + __ldg(vs + (1 * 3 + c) * NPSQ + p) * v1[p];
    laplace_tensor(g, tv, comp, lap);
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      const double a = __ldg(vs + (0 * 3 + c) * NPSQ + p) * lap[p], b = __ldg(vs + (1 * 3 + c) * NPSQ + p) * lap[p];
      if (c == 0) { acc0[p] = a; acc1[p] = b; }
      else { acc0[p] += a; acc1[p] += b; }
    }
  }
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    const double sm = geo_ld(g, p, G_SPHEREMP);
    l0[p] = acc0[p] + 2.0 * sm * v0[p] * rrearth * rrearth;
    l1[p] = acc1[p] + 2.0 * sm * v1[p] * rrearth * rrearth;
  }
}

// ---- the same operators finished one point at a time --------------------------------------------
// emit(p, value...) is called in ascending p as soon as point p is complete, so the caller can send
// the result to memory at once and the operator holds two work planes instead of four or five. The
// sums and their order are those of the plane-at-a-time versions above (bit-identical results).

// one point of divergence_sphere_wk's contraction (:538-583); s0, s1 = spheremp * (Dinv^T v)
__device__ __forceinline__ double div_wk_point(const double (&s0)[NPSQ], const double (&s1)[NPSQ], int n, int m) {
  double dd = -((s0[n * NP + 0] * dc.dvv[0][m] + s1[0 * NP + m] * dc.dvv[0][n]) * rrearth);
  HXX_UNROLL
  for (int j = 1; j < NP; ++j) dd -= (s0[n * NP + j] * dc.dvv[j][m] + s1[j * NP + m] * dc.dvv[j][n]) * rrearth;
  return dd;
}

// laplace_simple (:588-597) / laplace_tensor (:604-635, tv = the element's tensorVisc [2][2][16])
template <bool TENSOR, class G, class F>
__device__ __forceinline__ void laplace_points(const G& g, const double* __restrict__ tv, const double (&s)[NPSQ],
                                               F&& emit) {
  double s0[NPSQ], s1[NPSQ];
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) {
    double g0, g1;
    gradient_point(g, s, p, g0, g1);
    if (TENSOR) {
      const double t0 = __ldg(tv + 0 * NPSQ + p) * g0 + __ldg(tv + 2 * NPSQ + p) * g1;
      const double t1 = __ldg(tv + 1 * NPSQ + p) * g0 + __ldg(tv + 3 * NPSQ + p) * g1;
      g0 = t0;
      g1 = t1;
    }
    const double w0 = geo_ld(g, p, G_DINV00) * g0 + geo_ld(g, p, G_DINV10) * g1;
    const double w1 = geo_ld(g, p, G_DINV01) * g0 + geo_ld(g, p, G_DINV11) * g1;
    const double sm = geo_ld(g, p, G_SPHEREMP);
    s0[p] = sm * w0;
    s1[p] = sm * w1;
  }
  HXX_UNROLL
  for (int p = 0; p < NPSQ; ++p) emit(p, div_wk_point(s0, s1, p / NP, p % NP));
}

// vlaplace_sphere_wk_contra (:818-862). The vector field is read from memory (level already added to
// v0p, v1p) once per use; the weak gradient of the divergence waits in the caller's per-thread
// shared-memory slots park0/park1 (stride `ps` doubles between points) while the curl part is built.
// mi = the element's metinv [2][2][16]. emit(p, l0, l1).
template <class G, class MI, class F>
__device__ __forceinline__ void vlaplace_contra_points(const G& g, const MI& mi, double nu_ratio, const double* v0p,
                                                       const double* v1p, double* park0, double* park1, int ps,
                                                       F&& emit) {
  double sc[NPSQ];
  {  // divergence_sphere :352-392
    double gv0[NPSQ], gv1[NPSQ];
    {
      double v0[NPSQ], v1[NPSQ];
      plane_load(v0p, v0);
      plane_load(v1p, v1);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        const double md = geo_ld(g, p, G_METDET);
        gv0[p] = (geo_ld(g, p, G_DINV00) * v0[p] + geo_ld(g, p, G_DINV10) * v1[p]) * md;
        gv1[p] = (geo_ld(g, p, G_DINV01) * v0[p] + geo_ld(g, p, G_DINV11) * v1[p]) * md;
      }
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      double dx, dy;
      deriv_point(gv0, gv1, p / NP, p % NP, dx, dy);
      sc[p] = (dx + dy) * geo_ld(g, p, G_RMETDET_R);
    }
  }
  if (nu_ratio > 0 && nu_ratio != 1.0) {
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) sc[p] *= nu_ratio;
  }
  HXX_UNROLL
  for (int n = 0; n < NP; ++n) {  // grad_sphere_wk_testcov :714-748
    HXX_UNROLL
    for (int m = 0; m < NP; ++m) {
      const int p = n * NP + m;
      const double md = geo_ld(g, p, G_METDET);
      const double mi00 = mi.ld(0 * NPSQ + p), mi01 = mi.ld(1 * NPSQ + p), mi10 = mi.ld(2 * NPSQ + p),
                   mi11 = mi.ld(3 * NPSQ + p);
      double b0 = 0.0, b1 = 0.0;
      HXX_UNROLL
      for (int j = 0; j < NP; ++j) {
        const double mpnj = geo_ld(g, n * NP + j, G_MP), mpjm = geo_ld(g, j * NP + m, G_MP);
        const double snj = sc[n * NP + j], sjm = sc[j * NP + m];
        const double djm = dc.dvv[j][m], djn = dc.dvv[j][n];
        const double x0 = mpnj * mi00 * md * snj * djm + mpjm * mi01 * md * sjm * djn;
        const double x1 = mpnj * mi10 * md * snj * djm + mpjm * mi11 * md * sjm * djn;
        if (j == 0) { b0 = -x0; b1 = -x1; }
        else { b0 -= x0; b1 -= x1; }
      }
      park0[p * ps] = (geo_ld(g, p, G_D00) * b0 + geo_ld(g, p, G_D10) * b1) * rrearth;
      park1[p * ps] = (geo_ld(g, p, G_D01) * b0 + geo_ld(g, p, G_D11) * b1) * rrearth;
    }
  }
  phase_fence();
  {  // vorticity_sphere :494-533, then the mass-weighted copy curl_sphere_wk_testcov works on
    double c0[NPSQ], c1[NPSQ];
    {
      double v0[NPSQ], v1[NPSQ];
      plane_load(v0p, v0);
      plane_load(v1p, v1);
      HXX_UNROLL
      for (int p = 0; p < NPSQ; ++p) {
        c0[p] = geo_ld(g, p, G_D00) * v0[p] + geo_ld(g, p, G_D01) * v1[p];
        c1[p] = geo_ld(g, p, G_D10) * v0[p] + geo_ld(g, p, G_D11) * v1[p];
      }
    }
    HXX_UNROLL
    for (int p = 0; p < NPSQ; ++p) {
      double dvdx, dudy;
      deriv_point(c1, c0, p / NP, p % NP, dvdx, dudy);
      sc[p] = geo_ld(g, p, G_MP) * ((dvdx - dudy) * geo_ld(g, p, G_RMETDET_R));
    }
  }
  phase_fence();
  {  // curl_sphere_wk_testcov_update(alpha = -1, beta = 1) :683-710 and the final metric term
    double v0[NPSQ], v1[NPSQ];
    plane_load(v0p, v0);
    plane_load(v1p, v1);
    const double re2 = rrearth * rrearth;
    HXX_UNROLL
    for (int n = 0; n < NP; ++n) {
      HXX_UNROLL
      for (int m = 0; m < NP; ++m) {
        const int p = n * NP + m;
        double sb0 = -(sc[0 * NP + m] * dc.dvv[0][n]);
        double sb1 = sc[n * NP + 0] * dc.dvv[0][m];
        HXX_UNROLL
        for (int j = 1; j < NP; ++j) {
          sb0 -= sc[j * NP + m] * dc.dvv[j][n];
          sb1 += sc[n * NP + j] * dc.dvv[j][m];
        }
        const double l0 = 1.0 * park0[p * ps] + -1.0 * (geo_ld(g, p, G_D00) * sb0 + geo_ld(g, p, G_D10) * sb1) * rrearth;
        const double l1 = 1.0 * park1[p * ps] + -1.0 * (geo_ld(g, p, G_D01) * sb0 + geo_ld(g, p, G_D11) * sb1) * rrearth;
        const double f = 2.0 * geo_ld(g, p, G_SPHEREMP);
        emit(p, f * v0[p] * re2 + l0, f * v1[p] * re2 + l1);
      }
    }
  }
}

__device__ __forceinline__ bool is_interior_pt(int p) { return p == 5 || p == 6 || p == 9 || p == 10; }

}  // namespace hxx
