// Phase-level test hooks (section C of include/hommexx_b200.h) that run the SAME device
// functions the production kernels use on caller-provided data: the sphere operators
// (known-answer vectors of test/unit_tests/inputs/*.in) and the limiters (property tests of
// src/preqx/unit_tests/preqx_ut.cpp:1335-1531).
#include <cstring>

#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

#include "hxx_limiter.cuh"
#include "hxx_sphere.cuh"

namespace hxx {

enum { OP_GRAD, OP_DIV, OP_VORT, OP_LAPLACE, OP_DIV_WK, OP_VLAPLACE };

__global__ void sphere_op_kernel(int op, const double* __restrict__ geo, const double* __restrict__ mi,
                                 const double* __restrict__ in, double* __restrict__ out, double nu_ratio) {
  const int k = threadIdx.x;
  if (k >= NLEV) return;
  double a[NPSQ], b[NPSQ], r0[NPSQ], r1[NPSQ];
  plane_load(in + k, a);
  if (op != OP_GRAD && op != OP_LAPLACE) plane_load(in + NLF + k, b);
  switch (op) {
    case OP_GRAD: gradient_sphere(geo, a, r0, r1); break;
    case OP_DIV: divergence_sphere(geo, a, b, r0); break;
    case OP_VORT: vorticity_sphere(geo, a, b, r0); break;
    case OP_LAPLACE: laplace_simple(geo, a, r0); break;
    case OP_DIV_WK: divergence_sphere_wk(geo, a, b, r0); break;
    default: vlaplace_sphere_wk_contra(geo, mi, nu_ratio, a, b, r0, r1); break;
  }
  plane_store(out + k, r0);
  if (op == OP_GRAD || op == OP_VLAPLACE) plane_store(out + NLF + k, r1);
}

__global__ void limiter_kernel(int option, int nsets, const double* __restrict__ sphw, const double* __restrict__ dpmass,
                               double* __restrict__ ptens, double* __restrict__ qlim) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsets * NLEV) return;
  const int s = t / NLEV, k = t % NLEV;
  double c[NPSQ], x[NPSQ], dpm[NPSQ];
#pragma unroll
  for (int p = 0; p < NPSQ; ++p) {
    dpm[p] = dpmass[(size_t)s * NLF + p * NLEV + k];
    c[p] = sphw[s * NPSQ + p] * dpm[p];
    x[p] = ptens[(size_t)s * NLF + p * NLEV + k] / dpm[p];
  }
  double qmin = qlim[((size_t)s * 2) * NLEV + k], qmax = qlim[((size_t)s * 2 + 1) * NLEV + k];
  if (limiter_level(option, c, x, qmin, qmax)) {
#pragma unroll
    for (int p = 0; p < NPSQ; ++p) ptens[(size_t)s * NLF + p * NLEV + k] = x[p] * dpm[p];
    qlim[((size_t)s * 2) * NLEV + k] = qmin;
    qlim[((size_t)s * 2 + 1) * NLEV + k] = qmax;
  }
}

}  // namespace hxx

using namespace hxx;

extern "C" void hxx_sphere_op(const char* op, int ie, const double* in, double* out, double nu_ratio) {
  if (!S.active || !S.geo) runtime_abort("hxx_sphere_op: no initialised session", 13);
  int code, n_in, n_out;
  if (!std::strcmp(op, "gradient_sphere")) { code = OP_GRAD; n_in = 1; n_out = 2; }
  else if (!std::strcmp(op, "divergence_sphere")) { code = OP_DIV; n_in = 2; n_out = 1; }
  else if (!std::strcmp(op, "vorticity_sphere")) { code = OP_VORT; n_in = 2; n_out = 1; }
  else if (!std::strcmp(op, "laplace_simple")) { code = OP_LAPLACE; n_in = 1; n_out = 1; }
  else if (!std::strcmp(op, "divergence_sphere_wk")) { code = OP_DIV_WK; n_in = 2; n_out = 1; }
  else if (!std::strcmp(op, "vlaplace_sphere_wk_contra")) { code = OP_VLAPLACE; n_in = 2; n_out = 2; }
  else runtime_abort("hxx_sphere_op: unknown operator", 11);
  double *d_in, *d_out;
  CUDA_OK(cudaMalloc(&d_in, 2 * NLF * 8));
  CUDA_OK(cudaMalloc(&d_out, 2 * NLF * 8));
  CUDA_OK(cudaMemcpyAsync(d_in, in, (size_t)n_in * NLF * 8, cudaMemcpyHostToDevice, S.stream));
  PROBE(K_HOOK);
  sphere_op_kernel<<<1, ((NLEV + 31) / 32) * 32, 0, S.stream>>>(code, S.geo + (size_t)ie * NPSQ * GEO_N,
                                                               S.metinv + (size_t)ie * 4 * NPSQ, d_in, d_out, nu_ratio);
  KERNEL_LAUNCHED(K_HOOK);
  CUDA_OK(cudaMemcpyAsync(out, d_out, (size_t)n_out * NLF * 8, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  cudaFree(d_in);
  cudaFree(d_out);
}

extern "C" void hxx_limiter(int limiter_option, int nsets, const double* sphweights, const double* dpmass,
                            double* ptens, double* qlim) {
  if (!S.active) runtime_abort("hxx_limiter: no session", 13);
  const size_t nf = (size_t)nsets * NLF * 8, nw = (size_t)nsets * NPSQ * 8, nq = (size_t)nsets * 2 * NLEV * 8;
  double *d_w, *d_dp, *d_pt, *d_ql;
  CUDA_OK(cudaMalloc(&d_w, nw)); CUDA_OK(cudaMalloc(&d_dp, nf)); CUDA_OK(cudaMalloc(&d_pt, nf)); CUDA_OK(cudaMalloc(&d_ql, nq));
  CUDA_OK(cudaMemcpyAsync(d_w, sphweights, nw, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaMemcpyAsync(d_dp, dpmass, nf, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaMemcpyAsync(d_pt, ptens, nf, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaMemcpyAsync(d_ql, qlim, nq, cudaMemcpyHostToDevice, S.stream));
  const int nt = nsets * NLEV;
  PROBE(K_HOOK);
  limiter_kernel<<<(nt + 127) / 128, 128, 0, S.stream>>>(limiter_option, nsets, d_w, d_dp, d_pt, d_ql);
  KERNEL_LAUNCHED(K_HOOK);
  CUDA_OK(cudaMemcpyAsync(ptens, d_pt, nf, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaMemcpyAsync(qlim, d_ql, nq, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  cudaFree(d_w); cudaFree(d_dp); cudaFree(d_pt); cudaFree(d_ql);
}
