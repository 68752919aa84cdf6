// Vertical remap (rsplit > 0) — replaces RemapFunctor.hpp, PpmRemap.hpp and
// VerticalRemapManager of the reference: Lagrangian levels -> reference levels with PPM
// (mirrored boundaries, optional fixed-parabola variant) for u*dp, v*dp, T*dp and every Qdp.
//
// One warp per column, lanes = levels (coalesced 576-byte column loads/stores). The column's
// grid data (partitions, kid, z2, ppmdx) are built once in shared memory and reused by all
// 3+qsize fields. The three prefix sums (source/target interfaces, tracer mass) run in the
// reference's sequential order on one lane, so results are bit-identical to the reference's
// serial path (PpmRemap.hpp:226-247,524-566).
// Algorithmic HBM traffic per column: read+write of every remapped field (2 tiles per field per
// element) plus one read of dp3d.
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

namespace hxx {

constexpr int RW = 6;  // warps (columns) per block
constexpr int PAD = 2;
constexpr int L2_ = NLEV + 2;

struct ColSmem {
  double dpo[NLEV + 4], pio[NLEV + 2], pin[NLEV + 1], z2[NLEV], tgt[NLEV];
  double ppmdx[10][NLEV + 2];
  double ao[NLEV + 4], mass_o[NLEV + 2], dma[NLEV + 2], ai[NLEV + 1], coef[3][NLEV], massn[NLEV], var[NLEV];
  int kid[NLEV];
};

__device__ __forceinline__ double integrate_parabola(double sq, double lin, double cst, double x1, double x2) {
  return (cst * (x2 - x1) + lin * (x2 * x2 - x1 * x1) / 2.0) + sq * (x2 * x2 * x2 - x1 * x1 * x1) / 3.0;  // :668-673
}

// compute_partitions :506-597 + compute_integral_bounds :600-666 + compute_grids :366-413.
// On entry c.dpo[PAD..] holds the source thickness and c.tgt the target thickness.
// Returns false when the target grid is not monotone (kid would leave the column): the reference
// has undefined behaviour there; here the index is clamped and the caller raises the abort flag.
__device__ bool ppm_column_grids(ColSmem& c, int lane) {
  bool ok = true;
  if (lane == 0) {
    double acc = 0.0;
    for (int k = 0; k < NLEV; ++k) { c.pio[k] = acc; acc += c.dpo[k + PAD]; }
    c.pio[NLEV] = c.pio[NLEV - 1] + c.dpo[NLEV - 1 + PAD];
    acc = 0.0;
    for (int k = 0; k < NLEV; ++k) { c.pin[k] = acc; acc += c.tgt[k]; }
    c.pio[NLEV + 1] = c.pio[NLEV] + 1.0;
    c.pin[NLEV] = c.pio[NLEV];
    for (int k = 0; k < 2; ++k) {
      c.dpo[PAD - 1 - k] = c.dpo[k + PAD];
      c.dpo[NLEV + PAD + k] = c.dpo[NLEV + PAD - 1 - k];
    }
  }
  __syncwarp();
  for (int k = lane; k < NLEV; k += 32) {
    int kk = k + 1;
    while (kk <= NLEV + 1 && c.pio[kk - 1] <= c.pin[k + 1]) kk++;
    kk--;
    if (kk == NLEV + 1) kk = NLEV;
    if (kk < 1) { kk = 1; ok = false; }
    c.kid[k] = kk - 1;
    c.z2[k] = (c.pin[k + 1] - (c.pio[kk - 1] + c.pio[kk]) * 0.5) / c.dpo[kk + 1 + PAD - 2];
  }
  const double* dx = c.dpo;
  for (int j = lane; j < NLEV + 2; j += 32) {
    c.ppmdx[0][j] = dx[j + 1] / (dx[j] + dx[j + 1] + dx[j + 2]);
    c.ppmdx[1][j] = (2.0 * dx[j] + dx[j + 1]) / (dx[j + 1] + dx[j + 2]);
    c.ppmdx[2][j] = (dx[j + 1] + 2.0 * dx[j + 2]) / (dx[j] + dx[j + 1]);
  }
  for (int j = lane; j < NLEV + 1; j += 32) {
    c.ppmdx[3][j] = dx[j + 1] / (dx[j + 1] + dx[j + 2]);
    c.ppmdx[4][j] = 1.0 / (dx[j] + dx[j + 1] + dx[j + 2] + dx[j + 3]);
    c.ppmdx[5][j] = (2.0 * dx[j + 1] * dx[j + 2]) / (dx[j + 1] + dx[j + 2]);
    c.ppmdx[6][j] = (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    c.ppmdx[7][j] = (dx[j + 3] + dx[j + 2]) / (2.0 * dx[j + 2] + dx[j + 1]);
    c.ppmdx[8][j] = dx[j + 1] * (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    c.ppmdx[9][j] = dx[j + 2] * (dx[j + 2] + dx[j + 3]) / (dx[j + 1] + 2.0 * dx[j + 2]);
  }
  __syncwarp();
  return !__any_sync(0xffffffffu, !ok);
}

// compute_remap_phase :203-266 for one field of the column; c.var holds the field (mass units)
// on entry and the remapped field on exit.
__device__ void ppm_column_remap(ColSmem& c, int alg, int lane) {
  for (int k = lane; k < NLEV; k += 32) c.ao[k + PAD] = c.var[k] / c.dpo[k + PAD];
  __syncwarp();
  if (lane == 0) {
    for (int k0 = 0; k0 < 2; ++k0) {  // fill_cell_means_gs, mirrored :87-101
      c.ao[PAD - 1 - k0] = c.ao[k0 + PAD];
      c.ao[NLEV + PAD + k0] = c.ao[NLEV + PAD - 1 - k0];
    }
    double acc = 0.0;
    c.mass_o[0] = 0.0;
    for (int k = 0; k < NLEV; ++k) { c.mass_o[k + 1] = acc; acc += c.var[k]; }
    c.mass_o[NLEV + 1] = c.mass_o[NLEV] + c.var[NLEV - 1];
  }
  __syncwarp();
  // compute_ppm :416-503
  for (int j = lane; j < NLEV + 2; j += 32) {
    const double a0 = c.ao[j + PAD], a1 = c.ao[j + PAD - 1], a2 = c.ao[j + PAD - 2];
    double r = 0.0;
    if ((a0 - a1) * (a1 - a2) > 0.0) {
      const double da = c.ppmdx[0][j] * (c.ppmdx[1][j] * (a0 - a1) + c.ppmdx[2][j] * (a1 - a2));
      r = fmin(fmin(fabs(da), 2.0 * fabs(a1 - a2)), 2.0 * fabs(a0 - a1)) * copysign(1.0, da);
    }
    c.dma[j] = r;
  }
  __syncwarp();
  for (int j = lane; j < NLEV + 1; j += 32) {
    const double a0 = c.ao[j + PAD], a1 = c.ao[j + PAD - 1];
    c.ai[j] = a1 + c.ppmdx[3][j] * (a0 - a1) +
              c.ppmdx[4][j] * (c.ppmdx[5][j] * (c.ppmdx[6][j] - c.ppmdx[7][j]) * (a0 - a1) -
                               c.ppmdx[8][j] * c.dma[j + 1] + c.ppmdx[9][j] * c.dma[j]);
  }
  __syncwarp();
  for (int jp = lane; jp < NLEV; jp += 32) {
    const int j = jp + 1;
    const double am = c.ao[j + PAD - 1];
    double al = c.ai[j - 1], ar = c.ai[j];
    if ((ar - am) * (am - al) <= 0.) { al = am; ar = am; }
    if ((ar - al) * (am - (al + ar) / 2.0) > (ar - al) * (ar - al) / 6.0) al = 3.0 * am - 2.0 * ar;
    if ((ar - al) * (am - (al + ar) / 2.0) < -(ar - al) * (ar - al) / 6.0) ar = 3.0 * am - 2.0 * al;
    double c0 = 1.5 * am - (al + ar) / 4.0, c1 = ar - al, c2 = 3.0 * (-2.0 * am + (al + ar));
    if (alg == 2 && (jp < 2 || jp >= NLEV - 2)) {  // PpmFixedParabola::apply_ppm_boundary :110-133
      c0 = am; c1 = 0.0; c2 = 0.0;
    }
    c.coef[0][jp] = c0; c.coef[1][jp] = c1; c.coef[2][jp] = c2;
  }
  __syncwarp();
  // compute_remap :283-324
  for (int k = lane; k < NLEV; k += 32) {
    const int kk = c.kid[k];
    const double integral = integrate_parabola(c.coef[2][kk], c.coef[1][kk], c.coef[0][kk], -0.5, c.z2[k]);
    c.massn[k] = c.mass_o[kk + 1] + integral * c.dpo[kk + PAD];
  }
  __syncwarp();
  for (int k = lane; k < NLEV; k += 32) c.var[k] = k > 0 ? c.massn[k] - c.massn[k - 1] : c.massn[0];
  __syncwarp();
}

struct RemapArgs {
  double *v, *t, *dp3d, *ps_v, *qdp;
  int nelem, np1, np1_qdp, qsize, alg;
  int* invalid;
};

__global__ void __launch_bounds__(RW * 32) remap_kernel(const RemapArgs a) {
  extern __shared__ unsigned char smraw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long col = (long long)blockIdx.x * RW + w;
  if (col >= (long long)a.nelem * NPSQ) return;
  ColSmem& c = reinterpret_cast<ColSmem*>(smraw)[w];
  const int ie = (int)(col / NPSQ), p = (int)(col % NPSQ);
  const double* src = a.dp3d + off_s(ie, a.np1) + p * NLEV;
  bool bad = false;
  for (int k = lane; k < NLEV; k += 32) {
    const double s = src[k];
    c.dpo[k + PAD] = s;
    bad |= (isnan(s) || s < 0.0);  // check_source_thickness :439-464
  }
  if (__any_sync(0xffffffffu, bad)) {
    // RemapFunctor.hpp:190-198: the run aborts after the launch; the column is left untouched
    if (lane == 0) atomicOr(a.invalid, 1);
    return;
  }
  __syncwarp();
  // compute_ps_v :367-385 (serial sum, k ascending)
  double ps = 0.0;
  if (lane == 0) {
    for (int k = 0; k < NLEV; ++k) ps += c.dpo[k + PAD];
    ps += dc.hyai0 * dc.ps0;
    a.ps_v[((size_t)ie * NTL + a.np1) * NPSQ + p] = ps;
  }
  ps = __shfl_sync(0xffffffffu, ps, 0);
  // compute_target_thickness :417-437
  for (int k = lane; k < NLEV; k += 32) c.tgt[k] = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps;
  __syncwarp();
  if (!ppm_column_grids(c, lane) && lane == 0) atomicOr(a.invalid, 1);
  const int nf = 3 + a.qsize;
  for (int f = 0; f < nf; ++f) {
    double* fld = f == 0 ? a.v + off_v(ie, a.np1, 0) : f == 1 ? a.v + off_v(ie, a.np1, 1)
                : f == 2 ? a.t + off_s(ie, a.np1) : a.qdp + off_q(ie, a.np1_qdp, f - 3);
    fld += p * NLEV;
    const bool state = f < 3;
    for (int k = lane; k < NLEV; k += 32) {
      double x = fld[k];
      if (state) x *= c.dpo[k + PAD];  // ComputeExtrinsicsTag :255-268
      c.var[k] = x;
    }
    __syncwarp();
    ppm_column_remap(c, a.alg, lane);
    for (int k = lane; k < NLEV; k += 32) {
      double x = c.var[k];
      if (state) x /= c.tgt[k];  // ComputeIntrinsicsTag :294-307
      fld[k] = x;
    }
    __syncwarp();
  }
}

void vertical_remap(int np1, int np1_qdp) {
  if (!S.nelemd) return;
  RemapArgs a{S.v, S.t, S.dp3d, S.ps_v, S.qdp, S.nelemd, np1, np1_qdp, S.p.qsize, S.p.remap_alg, S.invalid_flag};
  constexpr size_t smem = RW * sizeof(ColSmem);
  static bool attr = false;
  if (!attr) {
    CUDA_OK(cudaFuncSetAttribute(remap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const long long ncol = (long long)S.nelemd * NPSQ;
  PROBE(K_REMAP);
  remap_kernel<<<(unsigned)((ncol + RW - 1) / RW), RW * 32, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(K_REMAP);
}

void check_remap_flag() {
  if (!S.invalid_flag) return;
  CUDA_OK(cudaMemcpyAsync(S.h_invalid, S.invalid_flag, sizeof(int), cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  if (*S.h_invalid) runtime_abort("Negative (or nan) layer thickness detected, aborting!", 101);
}

// ---- test hook: remap_Q_ppm semantics on caller-provided columns ---------------------------
__global__ void __launch_bounds__(RW * 32)
    remap_columns_kernel(int alg, int ncols, int nfields, const double* __restrict__ src_dp,
                         const double* __restrict__ tgt_dp, double* __restrict__ fields) {
  extern __shared__ unsigned char smraw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * RW + w;
  if (col >= ncols) return;
  ColSmem& c = reinterpret_cast<ColSmem*>(smraw)[w];
  for (int k = lane; k < NLEV; k += 32) {
    c.dpo[k + PAD] = src_dp[(size_t)col * NLEV + k];
    c.tgt[k] = tgt_dp[(size_t)col * NLEV + k];
  }
  __syncwarp();
  ppm_column_grids(c, lane);
  for (int f = 0; f < nfields; ++f) {
    double* fld = fields + ((size_t)f * ncols + col) * NLEV;
    for (int k = lane; k < NLEV; k += 32) c.var[k] = fld[k];
    __syncwarp();
    ppm_column_remap(c, alg, lane);
    for (int k = lane; k < NLEV; k += 32) fld[k] = c.var[k];
    __syncwarp();
  }
}

}  // namespace hxx

extern "C" void hxx_remap_columns(int alg, int ncols, int nfields, const double* src_dp, const double* tgt_dp,
                                  double* fields) {
  using namespace hxx;
  if (!S.active) runtime_abort("hxx_remap_columns: no session", 13);
  const size_t nc = (size_t)ncols * NLEV * 8, nfb = nc * nfields;
  double *d_src, *d_tgt, *d_f;
  CUDA_OK(cudaMalloc(&d_src, nc)); CUDA_OK(cudaMalloc(&d_tgt, nc)); CUDA_OK(cudaMalloc(&d_f, nfb));
  CUDA_OK(cudaMemcpyAsync(d_src, src_dp, nc, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaMemcpyAsync(d_tgt, tgt_dp, nc, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaMemcpyAsync(d_f, fields, nfb, cudaMemcpyHostToDevice, S.stream));
  constexpr size_t smem = RW * sizeof(ColSmem);
  CUDA_OK(cudaFuncSetAttribute(remap_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PROBE(K_HOOK);
  remap_columns_kernel<<<(ncols + RW - 1) / RW, RW * 32, smem, S.stream>>>(alg, ncols, nfields, d_src, d_tgt, d_f);
  KERNEL_LAUNCHED(K_HOOK);
  CUDA_OK(cudaMemcpyAsync(fields, d_f, nfb, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  cudaFree(d_src); cudaFree(d_tgt); cudaFree(d_f);
}
