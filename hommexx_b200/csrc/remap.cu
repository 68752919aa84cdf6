// Vertical remap (rsplit > 0) — replaces RemapFunctor.hpp, PpmRemap.hpp and
// VerticalRemapManager of the reference: Lagrangian levels -> reference levels with PPM
// (mirrored boundaries, optional fixed-parabola variant) for u*dp, v*dp, T*dp and every Qdp,
// with update_q (prim_driver.cpp:171-206, Q = Qdp / dp) fused into the tracer store.
//
// Mapping. A block owns RC = 4 columns (one GLL row of an element) and all 3 + qsize fields.
// Phase 1: one warp per column builds the column's grid data in shared memory (source/target
// partitions, kid, the integration bounds, the ten ppmdx coefficient rows) — shared by every
// field. Phase 2: ONE THREAD PER (column, field) sweeps the column top-down once, holding the
// PPM stencil (cell means, limited slopes, interface values) in a register window; the
// remapped masses are produced by a merge of the source and target grids (kid is monotone), so
// no per-field array ever exists in shared memory. Lanes of a warp share the column, so the
// coefficient reads are shared-memory broadcasts and the merge is branch-uniform. Field data
// moves through per-warp staging rows: cp.async brings CH levels of 32 fields at a time
// (32-byte segments, two chunks ahead of the arithmetic) and results leave the same way, so
// global accesses stay sector-coalesced although each thread walks its own column.
// The prefix sums (source/target interfaces, mass above a cell) run in the reference's
// sequential order, and divisions by a value that is reused across fields use its correctly
// rounded reciprocal plus one FMA correction (bit-identical to IEEE division, see div_rcp), so
// results stay bit-identical to the reference's serial path (PpmRemap.hpp:226-247,524-566).
// Algorithmic HBM traffic per column: read+write of every remapped field (2 tiles per field per
// element), one read of dp3d and one write of Q per tracer.
#include "hxx.cuh"

HXX_DEFINE_CONSTANTS()

namespace hxx {

// unroll of the sweep: the register window rotates with periods 3 (cell means) and 2 (masses), so 6 removes
// the register moves of the rotation
#ifndef HXX_REMAP_UNROLL
#define HXX_REMAP_UNROLL 2
#endif
constexpr int REMAP_UNROLL = HXX_REMAP_UNROLL;
constexpr int PAD = 2;
constexpr int RC = 4;        // columns per block
constexpr int CH = 4;        // levels per staged chunk (32 B per column-field)
constexpr int RS = CH;       // staging row stride in doubles; the level slot inside a row is rotated by the row's
                             // index / 4 (stage_slot), so 16 lanes reading their own rows hit 16 different banks
constexpr int NBUF = 2;      // input chunks in flight per warp
constexpr int NCHUNK = (NLEV + CH - 1) / CH;
// per-warp staging: NBUF input chunks and the output chunk ([32 rows][RS] each), and the two per-row pointer
// tables (field, Q); the input rows double as phase-1 scratch. Shared memory is what bounds the occupancy of the
// production kernel: 4 columns of ColData + 6 warps of staging = 56.9 KB per block at (72, 40), four blocks per SM.
constexpr int STAGE_PER_WARP = (NBUF + 1) * 32 * RS + 64;
static_assert(STAGE_PER_WARP >= 2 * NLEV + 3, "phase-1 scratch does not fit the staging rows");
static_assert(CH == 4 && NLEV <= 255, "stage_slot and the byte-sized kid assume 4-level chunks and < 256 levels");
// slot of (row, level) inside a [32][RS] staging buffer
__device__ __forceinline__ int stage_slot(int row, int level) { return row * RS + ((level + (row >> 2)) & (CH - 1)); }

struct ColData {
  double p0[NLEV + 2], p1[NLEV + 2], p2[NLEV + 2];                               // ppmdx 0..2, j = 0..NLEV+1
  double p3[NLEV + 1], p4[NLEV + 1], p567[NLEV + 1], p8[NLEV + 1], p9[NLEV + 1];  // ppmdx 3..9, j = 0..NLEV
  double dpo[NLEV + 4], rdpo[NLEV];
  double tgt[NLEV], rtgt[NLEV];
  double d1[NLEV], d2[NLEV], d3[NLEV];  // x2-x1, x2^2-x1^2, x2^3-x1^3 of integrate_parabola, x1 = -1/2, x2 = z2
  unsigned char kid[NLEV];
  int ok;
  int pad_;
};
static_assert(sizeof(ColData) % 8 == 0, "ColData is laid out as an array of doubles");

// compute_partitions :506-597 + compute_integral_bounds :600-666 + compute_grids :366-413, by one
// warp. On entry c.dpo[PAD..] holds the source thickness and c.tgt the target thickness.
// Returns false when the target grid is not monotone (kid would leave the column): the reference
// has undefined behaviour there; here the index is clamped and the caller raises the abort flag.
// `displaced` is set when a target level lies more than MAX_LAG levels below the source cell that holds its
// lower interface: the in-place sweep of the production kernel stores a chunk only after the chunk of the same
// levels has been staged, which such a displacement (a Lagrangian surface crossing five reference layers in one
// remap interval) would break — the run aborts with its own message instead of overwriting values not yet read.
constexpr int MAX_LAG = CH + 1;
// `scratch` holds the two interface arrays pio[NLEV+2], pin[NLEV+1] (only needed here).
__device__ bool ppm_column_grids(ColData& c, double* scratch, int lane, bool* displaced = nullptr) {
  bool ok = true, lag = false;
  double* const pio = scratch;
  double* const pin = scratch + NLEV + 2;
  if (lane < 2) {
    // the two interface prefix sums, sequential (k ascending) as the reference, one lane each;
    // loads are batched eight at a time so only the add chain is serial
    const double* src = lane == 0 ? c.dpo + PAD : c.tgt;
    double* dst = lane == 0 ? pio : pin;
    double acc = 0.0;
    for (int k0 = 0; k0 < NLEV; k0 += 8) {
      double v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (k0 + i < NLEV) ? src[k0 + i] : 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (k0 + i < NLEV) { dst[k0 + i] = acc; acc += v[i]; }
    }
    if (lane == 0) {
      pio[NLEV] = pio[NLEV - 1] + c.dpo[NLEV - 1 + PAD];
      pio[NLEV + 1] = pio[NLEV] + 1.0;
      for (int k = 0; k < 2; ++k) {
        c.dpo[PAD - 1 - k] = c.dpo[k + PAD];
        c.dpo[NLEV + PAD + k] = c.dpo[NLEV + PAD - 1 - k];
      }
    }
  }
  __syncwarp();
  if (lane == 0) pin[NLEV] = pio[NLEV];
  __syncwarp();
  for (int k = lane; k < NLEV; k += 32) {
    int kk = k + 1;
    while (kk <= NLEV + 1 && pio[kk - 1] <= pin[k + 1]) kk++;
    kk--;
    if (kk == NLEV + 1) kk = NLEV;
    if (kk < 1) { kk = 1; ok = false; }
    lag |= (k - (kk - 1) > MAX_LAG);
    c.kid[k] = (unsigned char)(kk - 1);
    const double z2 = (pin[k + 1] - (pio[kk - 1] + pio[kk]) * 0.5) / c.dpo[kk + 1 + PAD - 2];
    // integrate_parabola :668-673 with x1 = -0.5: x1*x1 = 0.25 and x1*x1*x1 = -0.125 exactly
    c.d1[k] = z2 - (-0.5);
    c.d2[k] = z2 * z2 - 0.25;
    c.d3[k] = z2 * z2 * z2 - (-0.125);
    c.rdpo[k] = 1.0 / c.dpo[k + PAD];
    c.rtgt[k] = 1.0 / c.tgt[k];
  }
  const double* dx = c.dpo;
  for (int j = lane; j < NLEV + 2; j += 32) {
    c.p0[j] = dx[j + 1] / (dx[j] + dx[j + 1] + dx[j + 2]);
    c.p1[j] = (2.0 * dx[j] + dx[j + 1]) / (dx[j + 1] + dx[j + 2]);
    c.p2[j] = (dx[j + 1] + 2.0 * dx[j + 2]) / (dx[j] + dx[j + 1]);
  }
  for (int j = lane; j < NLEV + 1; j += 32) {
    c.p3[j] = dx[j + 1] / (dx[j + 1] + dx[j + 2]);
    c.p4[j] = 1.0 / (dx[j] + dx[j + 1] + dx[j + 2] + dx[j + 3]);
    const double p5 = (2.0 * dx[j + 1] * dx[j + 2]) / (dx[j + 1] + dx[j + 2]);
    const double p6 = (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    const double p7 = (dx[j + 3] + dx[j + 2]) / (2.0 * dx[j + 2] + dx[j + 1]);
    c.p567[j] = p5 * (p6 - p7);  // the leading factor of the (a0 - a1) term in compute_ppm
    c.p8[j] = dx[j + 1] * (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    c.p9[j] = dx[j + 2] * (dx[j + 2] + dx[j + 3]) / (dx[j + 1] + 2.0 * dx[j + 2]);
  }
  __syncwarp();
  lag = __any_sync(0xffffffffu, lag);
  if (displaced) *displaced = lag;
  return !__any_sync(0xffffffffu, !ok);
}

// limited slope of compute_ppm :416-447 (a0, a1, a2 = cell means j, j-1, j-2)
__device__ __forceinline__ double ppm_dma(double a0, double a1, double a2, double p0, double p1, double p2) {
  double r = 0.0;
  if ((a0 - a1) * (a1 - a2) > 0.0) {
    const double da = p0 * (p1 * (a0 - a1) + p2 * (a1 - a2));
    r = fmin(fmin(fabs(da), 2.0 * fabs(a1 - a2)), 2.0 * fabs(a0 - a1)) * copysign(1.0, da);
  }
  return r;
}
// interface value :449-466
__device__ __forceinline__ double ppm_ai(double a0, double a1, double dma_j1, double dma_j, double p3, double p4,
                                         double p567, double p8, double p9) {
  return a1 + p3 * (a0 - a1) + p4 * (p567 * (a0 - a1) - p8 * dma_j1 + p9 * dma_j);
}

// One (column, field) sweep of compute_remap_phase :203-266, top-down, PPM stencil in registers.
//   tick(ln)      called by every lane (active or not) before level ln (2 <= ln < NLEV) is read
//   load(k)       raw field value at level k; multiplied by the source thickness when `state`
//   emit(k, x)    remapped mass x of target level k (increasing k); x / target thickness (the value a state leaves
//                 with, and Q of a tracer) is div_rcp(x, c.tgt[k], c.rtgt[k]), left to the caller
template <class Tick, class Load, class Emit>
__device__ __forceinline__ void ppm_sweep(const ColData& c, int alg, bool active, bool state, Tick tick, Load load,
                                          Emit emit) {
  double Am1 = 0, A0 = 0, A1 = 0, DM0 = 0, AI0 = 0, V0 = 0, V1 = 0, Mc = 0.0, massn_prev = 0.0;
  int kt = 0, kid_next = -1;
  const double r3 = 1.0 / 3.0, r6 = 1.0 / 6.0;
  // the step for source cell cc once A(cc+2) is known (V2 = its mass, unused past the column end)
  auto body = [&](int cc, double A2, double V2) {
    const double DM1 = ppm_dma(A2, A1, A0, c.p0[cc + 2], c.p1[cc + 2], c.p2[cc + 2]);
    const double AI1 = ppm_ai(A1, A0, DM1, DM0, c.p3[cc + 1], c.p4[cc + 1], c.p567[cc + 1], c.p8[cc + 1], c.p9[cc + 1]);
    // parabola of cell cc :468-503
    const double am = A0;
    double al = AI0, ar = AI1;
    if ((ar - am) * (am - al) <= 0.) { al = am; ar = am; }
    {
      // the two overshoot tests of :488-497 share their operands unless the first one fires
      double lhs = (ar - al) * (am - (al + ar) / 2.0);
      double lim = div_rcp((ar - al) * (ar - al), 6.0, r6);
      if (lhs > lim) {
        al = 3.0 * am - 2.0 * ar;
        lhs = (ar - al) * (am - (al + ar) / 2.0);
        lim = div_rcp((ar - al) * (ar - al), 6.0, r6);
      }
      if (lhs < -lim) ar = 3.0 * am - 2.0 * al;
    }
    double c0 = 1.5 * am - (al + ar) / 4.0, c1 = ar - al, c2 = 3.0 * (-2.0 * am + (al + ar));
    if (alg == 2 && (cc < 2 || cc >= NLEV - 2)) {  // PpmFixedParabola::apply_ppm_boundary :110-133
      c0 = am; c1 = 0.0; c2 = 0.0;
    }
    // compute_remap :283-324 for every target level whose lower interface lies in cell cc
    const double dpo_c = c.dpo[cc + PAD];
    while (kid_next == cc) {
      const double integral = (c0 * c.d1[kt] + c1 * c.d2[kt] / 2.0) + div_rcp(c2 * c.d3[kt], 3.0, r3);
      const double massn = Mc + integral * dpo_c;
      const double out = kt > 0 ? massn - massn_prev : massn;
      massn_prev = massn;
      emit(kt, out);
      ++kt;
      kid_next = kt < NLEV ? c.kid[kt] : -1;
    }
    Mc += V0;  // mass above the next cell (serial sum, k ascending :226-247)
    Am1 = A0; A0 = A1; A1 = A2; DM0 = DM1; AI0 = AI1; V0 = V1; V1 = V2;
  };
  tick(0);
  if (active) {
    V0 = load(0) * (state ? c.dpo[0 + PAD] : 1.0);  // ComputeExtrinsicsTag :255-268
    V1 = load(1) * (state ? c.dpo[1 + PAD] : 1.0);
    A0 = div_rcp(V0, c.dpo[0 + PAD], c.rdpo[0]);
    A1 = div_rcp(V1, c.dpo[1 + PAD], c.rdpo[1]);
    Am1 = A0;  // mirrored ghosts :87-101: A(-1) = A(0), A(-2) = A(1)
    const double dm_m1 = ppm_dma(A0, Am1, A1, c.p0[0], c.p1[0], c.p2[0]);
    DM0 = ppm_dma(A1, A0, Am1, c.p0[1], c.p1[1], c.p2[1]);
    AI0 = ppm_ai(A0, Am1, DM0, dm_m1, c.p3[0], c.p4[0], c.p567[0], c.p8[0], c.p9[0]);
    kid_next = c.kid[0];
  }
#pragma unroll REMAP_UNROLL
  for (int cc = 0; cc < NLEV - 2; ++cc) {
    const int ln = cc + 2;  // level entering the window
    tick(ln);
    if (active) {
      const double dl = c.dpo[ln + PAD];
      const double V2 = load(ln) * (state ? dl : 1.0);
      body(cc, div_rcp(V2, dl, c.rdpo[ln]), V2);
    }
  }
  if (active) {
    body(NLEV - 2, A1, 0.0);   // A(NLEV) = A(NLEV-1)
    body(NLEV - 1, Am1, 0.0);  // A(NLEV+1) = A(NLEV-2)
  }
}

struct RemapArgs {
  double *v, *t, *dp3d, *ps_v, *qdp, *Q;
  int nelem, np1, np1_qdp, qsize, alg;
  int* invalid;
};

// thread -> (column, field) of a block: per column `full` warps of 32 fields, then the last
// NF % 32 fields of 32/G columns packed G lanes each into shared warps
struct RemapMap {
  int nf, full, rem, G, S, nwarps;
};
__host__ __device__ inline RemapMap remap_map(int qsize) {
  RemapMap m;
  m.nf = 3 + qsize;
  m.full = m.nf / 32;
  m.rem = m.nf % 32;
  m.G = 0; m.S = 0;
  if (m.rem) {
    int g = 32 / RC;  // at most RC columns share a warp
    while (g < m.rem) g *= 2;
    m.G = g;
    m.S = 32 / g;
  }
  m.nwarps = RC * m.full + (m.rem ? RC / m.S : 0);
  return m;
}

constexpr int REMAP_MAX_WARPS = RC * ((3 + QSIZE_D) / 32) + ((3 + QSIZE_D) % 32 ? RC : 0);
__global__ void __launch_bounds__(REMAP_MAX_WARPS * 32) remap_kernel(const RemapArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  ColData* cols = reinterpret_cast<ColData*>(smraw);
  double* stage_all = reinterpret_cast<double*>(smraw + RC * sizeof(ColData));
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RemapMap m = remap_map(a.qsize);
  const int ie = blockIdx.x / (NPSQ / RC), p_first = (blockIdx.x % (NPSQ / RC)) * RC;

  // ---- phase 1: column grids, one warp per column ------------------------------------------
  for (int cl = w; cl < RC; cl += m.nwarps) {
    ColData& c = cols[cl];
    const int p = p_first + cl;
    const double* src = a.dp3d + off_s(ie, a.np1) + p * NLEV;
    bool bad = false;
    for (int k = lane; k < NLEV; k += 32) {
      const double s = src[k];
      c.dpo[k + PAD] = s;
      bad |= (isnan(s) || s < 0.0);  // check_source_thickness :439-464
    }
    bad = __any_sync(0xffffffffu, bad);
    if (bad) {
      // RemapFunctor.hpp:190-198: the run aborts after the launch; the column is left untouched
      if (lane == 0) { atomicOr(a.invalid, 1); c.ok = 0; }
      continue;
    }
    __syncwarp();
    // compute_ps_v :367-385 (serial sum, k ascending)
    double ps = 0.0;
    if (lane == 0) {
      for (int k0 = 0; k0 < NLEV; k0 += 8) {
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (k0 + i < NLEV) ? c.dpo[k0 + i + PAD] : 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (k0 + i < NLEV) ps += v[i];
      }
      ps += dc.hyai0 * dc.ps0;
      a.ps_v[((size_t)ie * NTL + a.np1) * NPSQ + p] = ps;
      c.ok = 1;
    }
    ps = __shfl_sync(0xffffffffu, ps, 0);
    // compute_target_thickness :417-437
    for (int k = lane; k < NLEV; k += 32) c.tgt[k] = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps;
    __syncwarp();
    bool displaced = false;
    if (!ppm_column_grids(c, stage_all + (size_t)w * STAGE_PER_WARP, lane, &displaced) && lane == 0) atomicOr(a.invalid, 1);
    if (displaced && lane == 0) { atomicOr(a.invalid, 2); c.ok = 0; }
  }
  __syncthreads();

  // ---- phase 2: one thread per (column, field) -----------------------------------------------
  int cl, f, gsize;
  if (w < RC * m.full) { cl = w / m.full; f = (w % m.full) * 32 + lane; gsize = 32; }
  else {
    const int rw = w - RC * m.full;
    cl = rw * m.S + lane / m.G;
    f = m.full * 32 + lane % m.G;
    gsize = m.G;
    if (lane % m.G >= m.rem) f = m.nf;  // padding lane
  }
  const ColData& c = cols[cl];
  const bool active = f < m.nf && c.ok;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  if (!amask) return;
  const int gfirst = lane - lane % gsize;  // first lane of this lane's column group
  // the active lanes of the group (they share the column, hence the merge progress) and this
  // lane's rank among them; padding lanes never reach the flush
  const unsigned gmask = (gsize == 32 ? 0xffffffffu : (((1u << gsize) - 1u) << gfirst)) & amask;
  const int gact = __popc(gmask), grank = __popc(gmask & ((1u << lane) - 1u));
  const int p = p_first + cl;
  // field bases of this lane's column (row r of the warp belongs to lane r)
  double* const vbase = a.v + off_v(ie, a.np1, 0) + p * NLEV;
  double* const tbase = a.t + off_s(ie, a.np1) + p * NLEV;
  double* const qbase = a.qdp + off_q(ie, a.np1_qdp, 0) + p * NLEV;
  double* const Qbase = a.Q + ((size_t)ie * QSIZE_D * NPSQ + p) * NLEV;
  auto field_ptr = [&](int ff) -> double* {
    return ff < 2 ? vbase + (size_t)ff * NLF : ff == 2 ? tbase : qbase + (size_t)(ff - 3) * NLF;
  };
  const bool state = f < 3;
  double* const stage = stage_all + (size_t)w * STAGE_PER_WARP;
  double* const obuf = stage + NBUF * 32 * RS;
  double** const ptab = reinterpret_cast<double**>(obuf + 32 * RS);  // row -> field column
  double** const qtab = ptab + 32;                                   // row -> Q column (null for states)
  __syncwarp();
  ptab[lane] = field_ptr(f < m.nf ? f : 0);
  qtab[lane] = (f >= 3 && f < m.nf) ? Qbase + (size_t)(f - 3) * NLF : nullptr;
  __syncwarp();
  // cooperative chunk load: element e = i*32 + lane of the [32 rows][CH levels] chunk
  auto prefetch = [&](int chunk) {
    if (chunk < NCHUNK) {
      double* dstb = stage + (chunk % NBUF) * 32 * RS;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int e = i * 32 + lane, row = e / CH, lev = e % CH, level = chunk * CH + lev;
        if (((amask >> row) & 1u) && level < NLEV) cp_async8(dstb + stage_slot(row, lev), ptab[row] + level);
      }
    }
    cp_async_commit();
  };
  prefetch(0);
  // staged raw value of this lane's field at `level` (chunk must have landed)
  auto staged = [&](int level) { return stage[((level / CH) % NBUF) * 32 * RS + stage_slot(lane, level)]; };
  ppm_sweep(
      c, a.alg, active, state,
      [&](int ln) {  // uniform per step: before level ln is read
        if (ln % CH == 0) {
          cp_async_wait<0>();  // chunk ln/CH has landed
          __syncwarp();
          prefetch(ln / CH + 1);
        }
      },
      staged,
      [&](int k, double out) {
        obuf[stage_slot(lane, k)] = out;
        if ((k + 1) % CH == 0 || k + 1 == NLEV) {
          // flush the finished chunk of this column group: its gact lanes store gact rows x CH levels. The quotient
          // by the target thickness is taken here, by whichever lane stores the value: states leave as x / tgt
          // (ComputeIntrinsicsTag :294-307), tracers as Qdp = x and Q = x / tgt (update_q)
          const int chunk = k / CH;
          __syncwarp(gmask);
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int e = i * gact + grank, rowl = e / CH, lev = e % CH, level = chunk * CH + lev;
            const int row = gfirst + rowl;
            if (level < NLEV) {
              const double x = obuf[stage_slot(row, lev)];
              const double t = div_rcp(x, c.tgt[level], c.rtgt[level]);
              double* qp = qtab[row];
              if (qp) { ptab[row][level] = x; qp[level] = t; }
              else ptab[row][level] = t;
            }
          }
          __syncwarp(gmask);
        }
      });
  cp_async_wait<0>();
}

// rsplit == 0 (RemapFunctor.hpp:42-100): the dynamics stay on reference levels, so only the tracers are
// remapped, from source thickness = reference thickness + dt (eta_dot_dpdn(k+1) - eta_dot_dpdn(k)). Off the
// benchmark path: one block per column, warp 0 builds the grids, one thread per tracer sweeps the column
// with the production ppm_sweep (plain loads / stores), update_q fused as in the main kernel.
__global__ void __launch_bounds__(64) remap_eulerian_kernel(const RemapArgs a, const double* __restrict__ eta_dot_dpdn,
                                                            double dt) {
  __shared__ ColData c;
  __shared__ double scratch[2 * NLEV + 3];
  const int ie = blockIdx.x / NPSQ, p = blockIdx.x % NPSQ, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const double* dp = a.dp3d + off_s(ie, a.np1) + p * NLEV;
    const double* eta = eta_dot_dpdn + off_f(ie) + p * NLEV;
    double ps = 0.0;
    if (lane == 0) {  // compute_ps_v :367-385 (serial sum, k ascending)
      for (int k = 0; k < NLEV; ++k) ps += dp[k];
      ps += dc.hyai0 * dc.ps0;
      a.ps_v[((size_t)ie * NTL + a.np1) * NPSQ + p] = ps;
    }
    ps = __shfl_sync(0xffffffffu, ps, 0);
    bool bad = false;
    for (int k = lane; k < NLEV; k += 32) {
      const double tgt = dc.dai[k] * dc.ps0 + dc.dbi[k] * ps;  // compute_target_thickness :417-437
      const double eta_next = k + 1 < NLEV ? eta[k + 1] : 0.0;
      const double delta_dpdn = eta_next - eta[k];
      const double src = tgt + dt * delta_dpdn;                // compute_source_thickness :60-92
      c.tgt[k] = tgt;
      c.dpo[k + PAD] = src;
      bad |= (isnan(src) || src < 0.0);                        // check_source_thickness :439-464
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    if (bad) {
      if (lane == 0) { atomicOr(a.invalid, 1); c.ok = 0; }
    } else {
      if (lane == 0) c.ok = 1;
      if (!ppm_column_grids(c, scratch, lane) && lane == 0) atomicOr(a.invalid, 1);
    }
  }
  __syncthreads();
  if (!c.ok) return;
  for (int q = threadIdx.x; q < a.qsize; q += blockDim.x) {
    double* fld = a.qdp + off_q(ie, a.np1_qdp, q) + p * NLEV;
    double* Qf = a.Q + (((size_t)ie * QSIZE_D + q) * NPSQ + p) * NLEV;
    ppm_sweep(c, a.alg, true, false, [](int) {}, [&](int k) { return fld[k]; },
              [&](int k, double x) { fld[k] = x; Qf[k] = div_rcp(x, c.tgt[k], c.rtgt[k]); });
  }
}

void vertical_remap(int np1, int np1_qdp, double dt) {
  if (!S.nelemd) return;
  if (S.p.rsplit == 0) {
    RemapArgs a{S.v, S.t, S.dp3d, S.ps_v, S.qdp, S.Q, S.nelemd, np1, np1_qdp, S.p.qsize, S.p.remap_alg, S.invalid_flag};
    PROBE(K_REMAP);
    remap_eulerian_kernel<<<(unsigned)(S.nelemd * NPSQ), 64, 0, S.stream>>>(a, S.eta_dot_dpdn, dt);
    KERNEL_LAUNCHED(K_REMAP);
    return;
  }
  RemapArgs a{S.v, S.t, S.dp3d, S.ps_v, S.qdp, S.Q, S.nelemd, np1, np1_qdp, S.p.qsize, S.p.remap_alg, S.invalid_flag};
  const RemapMap m = remap_map(S.p.qsize);
  const size_t smem = RC * sizeof(ColData) + (size_t)m.nwarps * STAGE_PER_WARP * sizeof(double);
  static size_t attr = 0;
  if (HXX_ONCE_PER_SESSION()) attr = 0;
  if (smem > attr) {
    CUDA_OK(cudaFuncSetAttribute(remap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  PROBE(K_REMAP);
  remap_kernel<<<(unsigned)(S.nelemd * (NPSQ / RC)), m.nwarps * 32, smem, S.stream>>>(a);
  KERNEL_LAUNCHED(K_REMAP);
}

void check_remap_flag() {
  if (!S.invalid_flag) return;
  CUDA_OK(cudaMemcpyAsync(S.h_invalid, S.invalid_flag, sizeof(int), cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  if (*S.h_invalid & 1) runtime_abort("Negative (or nan) layer thickness detected, aborting!", 101);
  if (*S.h_invalid & 2)
    runtime_abort("vertical remap: a Lagrangian level moved more than five reference layers in one remap interval "
                  "(unsupported by the in-place sweep), aborting!", 101);
}

// ---- test hook: remap_Q_ppm semantics on caller-provided columns ---------------------------
// One block per column: warp 0 builds the grids, then one thread per field sweeps the column
// with the same ppm_sweep as the production kernel (plain loads/stores instead of the staging).
__global__ void __launch_bounds__(64)
    remap_columns_kernel(int alg, int ncols, int nfields, const double* __restrict__ src_dp,
                         const double* __restrict__ tgt_dp, double* __restrict__ fields) {
  __shared__ ColData c;
  __shared__ double scratch[2 * NLEV + 3];
  const int col = blockIdx.x, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    for (int k = lane; k < NLEV; k += 32) {
      c.dpo[k + PAD] = src_dp[(size_t)col * NLEV + k];
      c.tgt[k] = tgt_dp[(size_t)col * NLEV + k];
    }
    __syncwarp();
    ppm_column_grids(c, scratch, lane);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < nfields; f += blockDim.x) {
    double* fld = fields + ((size_t)f * ncols + col) * NLEV;
    ppm_sweep(c, alg, true, false, [](int) {}, [&](int k) { return fld[k]; },
              [&](int k, double x) { fld[k] = x; });
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
  PROBE(K_HOOK);
  remap_columns_kernel<<<ncols, 64, 0, S.stream>>>(alg, ncols, nfields, d_src, d_tgt, d_f);
  KERNEL_LAUNCHED(K_HOOK);
  CUDA_OK(cudaMemcpyAsync(fields, d_f, nfb, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
  cudaFree(d_src); cudaFree(d_tgt); cudaFree(d_f);
}
