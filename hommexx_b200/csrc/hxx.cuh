// hommexx_b200 — common declarations of the B200-native preqx dycore (sm_100a, FP64).
//
// Data layout in HBM (level index innermost, as the reference's device views,
// src/share/cxx/Elements.hpp:44-48 / Types.hpp:76):
//   v    [ie][3][2][16][NLEV]   t, dp3d [ie][3][16][NLEV]   ps_v [ie][3][16]
//   qdp  [ie][2][QSIZE_D][16][NLEV]   Q, qtens_biharmonic [ie][QSIZE_D][16][NLEV]
//   qlim [ie][QSIZE_D][2][NLEV]       every derived/scratch field [ie][(c)][16][NLEV]
// One "field tile" = 16*NLEV doubles (9216 B at NLEV=72), contiguous.
//
// Thread mapping of every element kernel: one thread per (element, level); the thread keeps the
// whole 4x4 GLL plane of its level in registers, so the horizontal spectral-element operators
// (SphereOperators.hpp) are thread-local contractions with the derivative matrix held in
// constant memory, and consecutive lanes touch consecutive levels (coalesced 8-byte accesses).
// Vertical scans go through shared memory in the reference's sequential order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hommexx_b200.h"

#ifndef HXX_NLEV
#define HXX_NLEV 72
#endif
#ifndef HXX_QSIZE_D
#define HXX_QSIZE_D 40
#endif

namespace hxx {

constexpr int NP = 4;
constexpr int NPSQ = 16;
constexpr int NLEV = HXX_NLEV;
constexpr int QSIZE_D = HXX_QSIZE_D;
constexpr int NTL = 3;   // NUM_TIME_LEVELS
constexpr int QNTL = 2;  // Q_NUM_TIME_LEVELS
constexpr int NLF = NPSQ * NLEV;  // doubles per field tile

// PhysicalConstants.hpp:17-22
constexpr double Rwater_vapor = 461.5;
constexpr double Rgas = 287.04;
constexpr double cp = 1005.0;
constexpr double kappa = Rgas / cp;
constexpr double rrearth = 1.0 / 6.376e6;

enum { DSS_ETA = 0, DSS_OMEGA = 1, DSS_DIV_VDP_AVE = 2 };  // HommexxEnums.hpp DSSOption

// Per-point geometry record (16 doubles = 128 B), [ie][16][GEO_N]. All threads of an element
// read the same record: warp-uniform loads served by one L1 transaction.
enum {
  G_DINV00 = 0, G_DINV01, G_DINV10, G_DINV11,  // Dinv(a,b) at the point
  G_D00, G_D01, G_D10, G_D11,                  // D(a,b)
  G_METDET, G_RMETDET_R,                       // metdet, (1/metdet)*rrearth
  G_SPHEREMP, G_RSPHEREMP, G_FCOR, G_PHIS, G_MP,
  G_INV_SPHEREMP,                              // 1 / spheremp (NOT rspheremp, the assembled inverse mass)
  GEO_N
};

// Elements per thread block of the element kernels (threads = EPB*NLEV).
constexpr int EPB = (NLEV * 4 <= 288) ? 4 : 2;

// kernel classes for launch accounting and the per-kernel CUDA-event probes
enum KernelId {
  K_CAAR = 0, K_DSS, K_HALO_PACK, K_RK_COMBINE, K_DP3D_FROM_PS, K_STEP_INIT, K_HV_FIRST, K_HV_SECOND, K_HV_UPDATE,
  K_EULER_DIVDP, K_EULER_QMINMAX, K_MINMAX, K_EULER_ADVECT, K_EULER_FDSS, K_EULER_TAVG, K_REMAP, K_UPDATE_Q,
  K_TRANSPOSE, K_HOOK, K_FORCING, K_DIAG, K_EULER_ADVECT_MM, K_EULER_ADVECT_HV, K_COUNT
};
extern const char* const kernel_names[K_COUNT];

struct DevConst {
  double dvv[NP][NP];
  double hyai[NLEV + 1], hybi[NLEV + 1];
  double dai[NLEV], dbi[NLEV], dp0[NLEV];
  double ps0, hyai0;
};

// ---- boundary-exchange plan (device side) -------------------------------------------------
// Node-centric DSS: one record per unique GLL boundary node. src[] are the node's member
// values: >= 0 -> local (elem*16 + pt); < 0 -> halo value index ~src in the receive buffer;
// NONE -> unused. For each LOCAL member i, ord[i][] lists, in the reference's unpack order
// (BoundaryExchange.cpp:512-524: edges S,N,W,E for k=0..3, then corners), which other members
// are added to it.
constexpr int DSS_NONE = INT32_MIN;
struct DssNode {
  int src[4];
  uint8_t ord[4][3];
  uint8_t nmem, pad[3];
};
static_assert(sizeof(DssNode) == 32, "DssNode layout");

constexpr int MAX_DSS_FIELDS = QSIZE_D + 1;
struct FieldList {
  int nf;
  // qdp_time_avg fused into the exchange: after the DSS, the first navg fields become
  // (partner + 2 * field) / 3 with partner = the same point avg_delta doubles away
  int navg = 0;
  long long avg_delta = 0;
  double* base[MAX_DSS_FIELDS];       // element 0 of each field
  long long estride[MAX_DSS_FIELDS];  // doubles between consecutive elements
};

struct Params {  // SimulationParams.hpp
  int remap_alg, limiter_option, rsplit, qsplit, time_step_type, qsize, state_frequency, ftype;
  double nu, nu_p, nu_q, nu_s, nu_div, nu_top, hypervis_scaling, nu_ratio1, nu_ratio2;
  int hypervis_order, hypervis_subcycle;
  bool moist, disable_diagnostics, use_cpstar, consthv, params_set;
};

struct ConnInfo {  // Connectivity.hpp:21-50
  int l_lid, l_gid, l_pos, r_lid, r_gid, r_pos;
  int kind;     // 0 edge, 1 corner, 2 missing
  int sharing;  // 0 local, 1 shared (remote rank), 2 missing
  int direction;
  int remote_pid;
};

struct Session {
  bool active = false;
  int rank = 0, nranks = 1, device = 0;
  bool comm_set = false;
  void* nccl = nullptr;  // ncclComm_t
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t comm_stream = nullptr;  // halo pack + NCCL transfers (multi-GPU)
  cudaStream_t copy_stream = nullptr;  // PCIe copies of the Fortran-layout arrays (overlapped with their transposes)
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_xpose[2] = {nullptr, nullptr};
  cudaEvent_t ev_produced = nullptr, ev_halo = nullptr;
  int64_t launches = 0;
  int64_t launches_by[K_COUNT] = {};
  unsigned long long profile_mask = 0;
  Params p{};
  DevConst hc{};  // host copy of the constants
  bool have_dvv = false, have_hv = false;
  int nelemd = 0;
  int nm1 = 0, n0 = 1, np1 = 2, nstep = 0, nstep0 = 0, n0_qdp = 0, np1_qdp = 1;
  double rhs_viss = 0.0;
  bool store_phi = false;
  std::vector<ConnInfo> conn;  // [nelemd][8]

  // geometry
  double *geo = nullptr, *metinv = nullptr, *tensorvisc = nullptr, *vec_sph2cart = nullptr;
  // state
  double *v = nullptr, *t = nullptr, *dp3d = nullptr, *ps_v = nullptr;
  double *phi = nullptr, *omega_p = nullptr, *eta_dot_dpdn = nullptr, *derived_vn0 = nullptr,
         *derived_dp = nullptr, *divdp = nullptr, *divdp_proj = nullptr, *dpdiss_ave = nullptr,
         *dpdiss_biharmonic = nullptr;
  double *vtens = nullptr, *ttens = nullptr, *dptens = nullptr;
  double *vstar = nullptr, *dpdissk = nullptr, *dp_star = nullptr;  // test-visible scratch
  double *qdp = nullptr, *qtens_biharmonic = nullptr, *qlim = nullptr, *qlim_x = nullptr, *Q = nullptr;
  double *hs_lat = nullptr, *hs_hyam = nullptr;  // Held-Suarez inputs: [ie][16] latitudes, [2][NLEV] hyam | hybm
  double *fm = nullptr, *ft = nullptr, *fq = nullptr;  // CAM forcing [ie][2][16][NLEV], [ie][16][NLEV], [ie][QSIZE_D][16][NLEV]
  // exchange plan
  DssNode* nodes = nullptr;  // generic remainder (cube vertices, nodes with off-rank sharers)
  int nnodes = 0;
  void *dss_pairs = nullptr, *dss_quads = nullptr;  // lean lists (dss.cu: DssPair / DssQuad)
  int npairs = 0, nquads = 0;
  int* nbr8 = nullptr;  // [nelemd][8] neighbour lid (>=0), ~halo_conn (<0) or DSS_NONE
  int* elem_order = nullptr;  // local elements, those without an off-rank neighbour first
  int n_interior = 0;
  // halo (multi-GPU)
  int n_halo_pts = 0;           // receive points (edge = 4, corner = 1 per remote connection)
  int n_send_pts = 0;
  int* send_src = nullptr;      // [n_send_pts] local elem*16+pt
  double *sendbuf = nullptr, *recvbuf = nullptr;
  std::vector<int> peer, peer_send_off, peer_send_cnt, peer_recv_off, peer_recv_cnt;  // in points
  int n_halo_conn = 0, n_send_conn = 0;  // min/max exchange: one slot per remote connection
  int* send_conn_elem = nullptr;
  std::vector<int> peer_csend_off, peer_csend_cnt, peer_crecv_off, peer_crecv_cnt;
  // P2P halo (dss.cu): the pack kernels store straight into the neighbour ranks' receive buffers over NVLink
  // (CUDA IPC mappings) and raise an epoch flag there; `halo_alloc` = [recv 0 | recv 1 | flags] of this rank
  bool p2p = false;
  void* halo_alloc = nullptr;
  size_t halo_buf_doubles = 0;          // doubles per receive buffer
  double* halo_recv[2] = {nullptr, nullptr};
  int* halo_flags = nullptr;            // [nranks], written by the peers
  std::vector<void*> peer_alloc;        // per peer slot: the peer's halo_alloc mapped here
  int *send_pt_dst = nullptr, *send_pt_peer = nullptr;      // [n_send_pts]: point index in the peer's buffer, peer slot
  int *send_conn_dst = nullptr, *send_conn_peer = nullptr;  // [n_send_conn]: the same per connection (min/max)
  unsigned halo_epoch = 0;              // exchanges so far; identical on every rank
  int* invalid_flag = nullptr;   // device flag: negative/NaN thickness in remap
  int* h_invalid = nullptr;      // pinned
  double* diag[8] = {};
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* diag_scratch = nullptr;  // device sums of the diagnostics (grown on demand, kept for the session)
  size_t diag_scratch_bytes = 0;
  // Incremented by initialize_hommexx_session: per-function attributes (dynamic shared-memory limits) are
  // per device/context, so every translation unit re-applies them when it sees a new session.
  int session_id = 0;
};

extern Session S;

[[noreturn]] void runtime_abort(const char* msg, int code);
void cuda_check(cudaError_t e, const char* what, const char* file, int line);
#define CUDA_OK(x) ::hxx::cuda_check((x), #x, __FILE__, __LINE__)
// PROBE(id) before a launch and KERNEL_LAUNCHED(id) after it: launch accounting, error check and
// (when the class is selected with hommexx_b200_profile) a CUDA-event pair on the launch stream.
void probe_begin(int id);
void probe_end(int id);
#define PROBE(id) do { if (::hxx::S.profile_mask >> (id) & 1ull) ::hxx::probe_begin(id); } while (0)
#define KERNEL_LAUNCHED(id)                                              \
  do {                                                                   \
    ++::hxx::S.launches;                                                 \
    ++::hxx::S.launches_by[id];                                          \
    CUDA_OK(cudaGetLastError());                                         \
    if (::hxx::S.profile_mask >> (id) & 1ull) ::hxx::probe_end(id);      \
  } while (0)

// True the first time a call site is reached in each session (function attributes are per device / context).
#define HXX_ONCE_PER_SESSION()                           \
  ([] {                                                  \
    static int seen = -1;                                \
    if (seen == ::hxx::S.session_id) return false;       \
    seen = ::hxx::S.session_id;                          \
    return true;                                         \
  }())

// NVTX ranges named after the reference's GPTL timers (profiling.hpp:13-51 and the start_timer calls of
// prim_driver.cpp, prim_step.cpp, prim_advance_exp.cpp, prim_advec_tracers_remap.cpp, CaarFunctor.cpp,
// HyperviscosityFunctorImpl.cpp): the reference cannot time on CUDA ("Can't use GPTL timers on CUDA"); here the
// same names show up on the host timeline of nsys / ncu --nvtx. Header-only; a no-op without a profiler attached.
struct NvtxRange {
  explicit NvtxRange(const char* name);
  ~NvtxRange();
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define HXX_TIMER_CAT2(a, b) a##b
#define HXX_TIMER_CAT(a, b) HXX_TIMER_CAT2(a, b)
#define HXX_TIMER(name) ::hxx::NvtxRange HXX_TIMER_CAT(hxx_timer_, __LINE__) { name }

// per-translation-unit constant bank (no relocatable device code: each TU owns a copy)
void register_const_uploader(void (*fn)(const DevConst&));
void upload_constants();
#define HXX_DEFINE_CONSTANTS()                                                        \
  static __constant__ ::hxx::DevConst dc;                                                 \
  static void hxx_upload_tu(const ::hxx::DevConst& h) {                               \
    CUDA_OK(cudaMemcpyToSymbolAsync(dc, &h, sizeof(h), 0, cudaMemcpyHostToDevice, ::hxx::S.stream)); \
  }                                                                                   \
  static const int hxx_reg_tu = (::hxx::register_const_uploader(&hxx_upload_tu), 0);

// field addressing helpers (host and device)
__host__ __device__ inline size_t off_v(int ie, int tl, int c) { return (((size_t)ie * NTL + tl) * 2 + c) * NLF; }
__host__ __device__ inline size_t off_s(int ie, int tl) { return ((size_t)ie * NTL + tl) * NLF; }
__host__ __device__ inline size_t off_q(int ie, int tq, int q) { return (((size_t)ie * QNTL + tq) * QSIZE_D + q) * NLF; }
__host__ __device__ inline size_t off_f(int ie) { return (size_t)ie * NLF; }

inline int nblocks_elem(int nelem) { return (nelem + EPB - 1) / EPB; }
// Threads per block of the sync-free element kernels (flat (element, level) mapping).
constexpr int TPB = 128;
inline int nblocks_flat(int nelem) { return (int)(((long long)nelem * NLEV + TPB - 1) / TPB); }

// ---- device helpers shared by the kernels ---------------------------------------------------
#ifdef __CUDACC__
// cp.async staging (LDGSTS): a thread streams 8-byte words into its own shared-memory slots
// ahead of the arithmetic, so HBM latency hides behind FP64 work without holding the in-flight
// data in registers.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// x / d given r = RN(1 / d): q = RN(x r) is within 1.5 ulp of x / d (r carries half an ulp of relative error,
// the product rounds once more), the FMA residual e = x - d q is exact, and res = RN(q + e r) adds the correction
// e / d with a relative error of 2^-53 — an absolute error below 2^-52 ulp of the quotient. res is therefore the
// correctly rounded quotient unless x / d lies within that distance of a midpoint between two doubles; Markstein's
// theorem proves the tie-free case for |q - x/d| < 1 ulp, the remaining sliver is covered empirically:
// oracle/div_rcp_check.c compares with IEEE division on operands built for it (mantissas next to powers of two,
// long runs of ones, quotients within a few ulps of representable values AND of ties formed in binary128): 0
// mismatches in 10^8 pairs, 2 x 10^7 re-run by every CPU test run (tests/test_div_rcp.py). Exactness needs the
// residual to stay normal, so x outside a wide exponent window takes the IEEE division instead (out of line, so
// that it is never speculated); zeros — common in tracer fields — return x r, the correctly signed zero.
// PRECONDITION: d is a normal number of moderate magnitude (2^-300 < |d| < 2^300): layer thicknesses, spheremp,
// small constants — every divisor of this path; a thickness that is not is caught by the remap's abort flag.
static __device__ __noinline__ double div_ieee(double x, double d) { return x / d; }
__device__ __forceinline__ double div_rcp(double x, double d, double r) {
  const unsigned ex = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
  const double q = x * r;
  const double e = fma(-d, q, x);
  double res = fma(e, r, q);
  if (ex - 423u > 1200u) {  // zero, denormal, tiny, huge, infinite or NaN
    res = q;
    if (x != 0.0) res = div_ieee(x, d);
  }
  return res;
}
// A whole plane at once: x[p] <- x[p] / d(p) with r(p) = 1 / d(p). The window test is reduced over the 16
// values first, so the common case is one branch and 48 FP64 instructions; a plane with a zero or an
// out-of-window value goes through div_rcp point by point. Same results as div_rcp.
template <int N, class D, class R>
__device__ __forceinline__ void div_rcp_plane(double (&x)[N], D d, R r) {
  unsigned worst = 0;
#pragma unroll
  for (int p = 0; p < N; ++p) worst = max(worst, (((unsigned)__double2hiint(x[p]) >> 20) & 0x7ffu) - 423u);
  if (worst <= 1200u) {
#pragma unroll
    for (int p = 0; p < N; ++p) {
      const double dd = d(p), rr = r(p);
      const double q = x[p] * rr;
      const double e = fma(-dd, q, x[p]);
      x[p] = fma(e, rr, q);
    }
  } else {
#pragma unroll
    for (int p = 0; p < N; ++p) x[p] = div_rcp(x[p], d(p), r(p));
  }
}
#endif

// ---- phases (each defined in its own .cu) -------------------------------------------------
// dss.cu
void build_exchange_plan();
void free_exchange_plan();
void dss_exchange(const FieldList& fl, bool rspheremp);  // boundary nodes only (interior folded by producer unless told)
void scale_interior_rspheremp(const FieldList& fl);      // the 4 interior points * rspheremp
void check_comm_errors();
void minmax_exchange();                                  // qlim -> qlim (neighbourhood min/max)
FieldList fields_caar(int tl);
FieldList fields_hv();
FieldList fields_euler(int tq, int dss_opt, int tavg_n0_qdp = -1);
FieldList fields_qtens();
double* dss_var(int dss_opt);
// caar.cu
void caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, bool with_dss);
void rk_combine(int nm1, int n0);
void dp3d_from_ps(int n0);
void prim_step_init(int n0);
void update_q(int np1_qdp, int np1);
// hv.cu
void hypervis_run(int np1, double dt, double eta_ave_w);
// euler.cu
void euler_precompute_divdp();
// tavg_n0_qdp >= 0 fuses qdp_time_avg(tavg_n0_qdp, np1_qdp) into this stage (its advection
// kernel does the interior points, its DSS the boundary nodes)
void euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt, int tavg_n0_qdp = -1);
void euler_qdp_time_avg(int n0_qdp, int np1_qdp);
// forcing_diag.cu
void apply_cam_forcing(double dt, bool tracers);  // CamForcing.cpp:149-174 (tracers = false: _dynamics)
void prim_diag_scalars(bool before_advance, int ivar);
void prim_energy_halftimes(bool before_advance, int ivar);
void held_suarez_forcing(const double* lat, const double* hyam, const double* hybm);  // forcing_diag.cu
void push_Q_to_host(double* host_q);  // hxx_session.cu: device Q -> F90 layout
void* diag_scratch(size_t bytes);     // hxx_session.cu: persistent device buffer of the diagnostic sums
// remap.cu
void vertical_remap(int np1, int np1_qdp, double dt);
void check_remap_flag();

}  // namespace hxx
