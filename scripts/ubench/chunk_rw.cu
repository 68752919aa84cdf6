// HBM behaviour behind the DSS pass: read-modify-write of CHUNK-byte pieces of a large buffer, (a) densely in address
// order, (b) 12 of every 16 columns in address order (the boundary columns of a [16][72] field tile), (c) pieces
// visited in a random order. One thread owns 16 bytes; a warp covers 512 contiguous bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chunk_rw chunk_rw.cu && ./chunk_rw
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void rmw(double2* __restrict__ buf, const int* __restrict__ order, long long nchunk, int v_per_chunk) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = g / v_per_chunk;
  if (c >= nchunk) return;
  const long long dst = (long long)order[c] * v_per_chunk + g % v_per_chunk;
  double2 x = buf[dst];
  x.x += 1.0; x.y += 1.0;
  buf[dst] = x;
}

int main(int argc, char** argv) {
  // default 4 GiB (HBM behaviour); a size below the 126 MB L2 (e.g. 48) shows what the same pass costs on L2-resident data
  const size_t bytes = argc > 1 ? size_t(atol(argv[1])) << 20 : size_t(4) << 30;
  printf("buffer %zu MiB\n", bytes >> 20);
  double2* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  std::mt19937 rng(1);
  for (int chunk : {576, 1152, 2304, 4608, 9216}) {
    const long long nchunk_all = bytes / chunk;
    const int vpc = chunk / 16;
    for (int mode = 0; mode < 3; ++mode) {
      std::vector<int> order;
      if (mode == 1 && chunk != 576) continue;
      for (long long c = 0; c < nchunk_all; ++c) {
        if (mode == 1) { const int p = c % 16; if (p == 5 || p == 6 || p == 9 || p == 10) continue; }
        order.push_back((int)c);
      }
      if (mode == 2) std::shuffle(order.begin(), order.end(), rng);
      int* d_order;
      cudaMalloc(&d_order, order.size() * sizeof(int));
      cudaMemcpy(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice);
      const long long nthreads = (long long)order.size() * vpc;
      const int nb = (int)((nthreads + 127) / 128);
      float best = 1e30f;
      for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(a);
        rmw<<<nb, 128>>>(buf, d_order, (long long)order.size(), vpc);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = std::min(best, ms);
      }
      const double gb = 2.0 * order.size() * chunk / 1e9;
      printf("chunk %5d B  %-28s %7.3f ms  %7.1f GB/s (R+W)\n", chunk,
             mode == 0 ? "dense, address order" : mode == 1 ? "12 of 16 columns, addr order" : "random order", best, gb / (best * 1e-3));
      cudaFree(d_order);
    }
  }
  return 0;
}
