// FP64 pipe micro-benchmark: issue rate of DFMA / DMUL / DADD per SM and dependent-issue latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      if (OP == 1) x[i] = __dmul_rn(x[i], a);
      if (OP == 2) x[i] = __dadd_rn(x[i], b);
    }
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int ILP>
void run(const char* name, int warps_per_sm) {
  int nsm = 148, iters = 20000;
  double* out; cudaMalloc(&out, 8 * nsm * 32 * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid(nsm), block(32 * warps_per_sm);
  k<OP, ILP><<<grid, block>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<OP, ILP><<<grid, block>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)nsm * warps_per_sm * iters * ILP;  // warp instructions
  double clk = 1.965e9 * ms * 1e-3;
  printf("%-5s ILP=%2d warps/SM=%2d : %.3f warp-inst/clk/SM  (%.1f lanes/clk/SM)  cycles per dependent step per warp = %.2f\n", name, ILP, warps_per_sm,
         inst / clk / nsm, inst * 32 / clk / nsm, clk / iters);
  cudaFree(out);
}
int main() {
  run<0, 1>("DFMA", 4); run<0, 1>("DFMA", 8); run<0, 1>("DFMA", 16); run<0, 8>("DFMA", 4); run<0, 8>("DFMA", 12); run<0, 8>("DFMA", 32);
  run<1, 8>("DMUL", 12); run<1, 8>("DMUL", 32); run<2, 8>("DADD", 12); run<2, 8>("DADD", 32);
  run<2, 1>("DADD", 4); run<1, 1>("DMUL", 4); run<0, 2>("DFMA", 4); run<0, 4>("DFMA", 4);
  return 0;
}
