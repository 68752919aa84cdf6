// Session, state containers and the reference's Fortran-facing entry points (section A/B of
// include/hommexx_b200.h): the drop-in replacement for src/share/cxx/cxx_f90_interface.cpp,
// prim_driver.cpp, prim_step.cpp, prim_advance_exp.cpp, prim_advec_tracers_remap.cpp,
// mpi/mpi_cxx_f90_interface.cpp and Hommexx_Session.cpp of the reference. Host orchestration
// only; every arithmetic phase is a CUDA kernel in caar.cu / hv.cu / euler.cu / remap.cu / dss.cu.
#include <cmath>
#include <cstring>
#include <string>

#include "hxx.cuh"
#include <nvtx3/nvToolsExt.h>
#ifdef HXX_WITH_NCCL
#include <nccl.h>
#endif

namespace hxx {

Session S;

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

static std::vector<void (*)(const DevConst&)>& uploaders() {
  static std::vector<void (*)(const DevConst&)> u;
  return u;
}
void register_const_uploader(void (*fn)(const DevConst&)) { uploaders().push_back(fn); }
void upload_constants() {
  for (auto fn : uploaders()) fn(S.hc);
  CUDA_OK(cudaStreamSynchronize(S.stream));
}

const char* const kernel_names[K_COUNT] = {
    "caar", "dss", "halo_pack", "rk_combine", "dp3d_from_ps", "prim_step_init", "hv_first_laplace",
    "hv_second_laplace_pre_exchange", "hv_update_states", "euler_divdp", "euler_qminmax", "minmax",
    "euler_advect", "euler_fdss", "euler_time_avg", "remap", "update_q", "transpose", "hook", "cam_forcing",
    "diagnostics", "euler_advect_mm", "euler_advect_hv"};

// ---- per-kernel CUDA-event probes (on the launch stream) -----------------------------------
namespace {
struct ProbePair { cudaEvent_t a, b; int id; };
std::vector<ProbePair> g_probes;       // recorded pairs since the last reset
std::vector<ProbePair> g_probe_pool;   // recycled events
cudaEvent_t g_open[K_COUNT];
cudaEvent_t g_marks[16];
bool g_marks_made = false;
}  // namespace

void probe_begin(int id) {
  ProbePair pp;
  if (!g_probe_pool.empty()) { pp = g_probe_pool.back(); g_probe_pool.pop_back(); }
  else { CUDA_OK(cudaEventCreate(&pp.a)); CUDA_OK(cudaEventCreate(&pp.b)); }
  pp.id = id;
  CUDA_OK(cudaEventRecord(pp.a, S.stream));
  g_open[id] = pp.a;
  g_probes.push_back(pp);
}
void probe_end(int id) {
  for (auto it = g_probes.rbegin(); it != g_probes.rend(); ++it)
    if (it->id == id && it->a == g_open[id]) { CUDA_OK(cudaEventRecord(it->b, S.stream)); return; }
}

void runtime_abort(const char* msg, int code) {
  // ErrorDefs.cpp:23-27 (MPI_Abort -> exit: one process per GPU, torchrun tears the job down)
  std::fprintf(stderr, "%s\nExiting...\n", msg);
  std::fflush(stderr);
  std::_Exit(code);
}

void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return;
  char msg[1024];
  std::snprintf(msg, sizeof msg, "hommexx_b200: CUDA error '%s' at %s:%d in %s (no CPU fallback exists)",
                cudaGetErrorString(e), file, line, what);
  runtime_abort(msg, 1);
}

static double* dalloc(size_t n) {
  double* p = nullptr;
  CUDA_OK(cudaMalloc(&p, (n ? n : 1) * sizeof(double)));
  CUDA_OK(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(double), S.stream));
  return p;
}
template <typename T>
static void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

static void* scratch(size_t bytes) {
  if (bytes > S.scratch_bytes) {
    if (S.scratch) CUDA_OK(cudaFree(S.scratch));
    CUDA_OK(cudaMalloc(&S.scratch, bytes));
    S.scratch_bytes = bytes;
  }
  return S.scratch;
}

// ---- layout transposition (SyncUtils.hpp): src [b][R][C] -> dst [b][C][R] ------------------
__global__ void transpose_batched(const double* __restrict__ src, double* __restrict__ dst, int R, int C) {
  extern __shared__ double tile[];  // [R][C+1]
  const size_t b = blockIdx.x;
  const double* s = src + b * (size_t)R * C;
  double* d = dst + b * (size_t)R * C;
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) tile[(i / C) * (C + 1) + (i % C)] = s[i];
  __syncthreads();
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) d[i] = tile[(i % R) * (C + 1) + (i / R)];
}

static void launch_transpose(const double* src, double* dst, size_t nbatch, int R, int C) {
  if (!nbatch) return;
  const size_t smem = (size_t)R * (C + 1) * sizeof(double);
  PROBE(K_TRANSPOSE);
  transpose_batched<<<(unsigned)nbatch, 256, smem, S.stream>>>(src, dst, R, C);
  KERNEL_LAUNCHED(K_TRANSPOSE);
}

// Host <-> device moves of the Fortran-layout arrays go through two device staging buffers so that the PCIe copy
// of one chunk overlaps the layout transposition of the other (copy stream <-> compute stream, one event pair per
// buffer); with page-locked host arrays both directions then run at the bus rate instead of copy + kernel in turn.
constexpr size_t STAGE_BYTES = size_t(128) << 20;  // per staging buffer

// host F90 [b][NLEV][C] -> device [b][C][NLEV]
static void pull_field(const double* host, double* dev, size_t nbatch, int C) {
  if (!nbatch) return;
  const size_t item = (size_t)NLEV * C;
  const size_t per = std::max<size_t>(1, STAGE_BYTES / (item * 8));
  const size_t half = std::min(nbatch, per) * item;
  double* st = (double*)scratch(2 * half * 8);
  CUDA_OK(cudaEventRecord(S.ev_xpose[0], S.stream));  // the staging buffers may still be read by earlier work
  CUDA_OK(cudaStreamWaitEvent(S.copy_stream, S.ev_xpose[0], 0));
  int i = 0;
  for (size_t b0 = 0; b0 < nbatch; b0 += per, ++i) {
    const size_t nb = std::min(per, nbatch - b0);
    const int b = i & 1;
    if (i >= 2) CUDA_OK(cudaStreamWaitEvent(S.copy_stream, S.ev_xpose[b], 0));  // chunk i - 2 has left this buffer
    CUDA_OK(cudaMemcpyAsync(st + b * half, host + b0 * item, nb * item * 8, cudaMemcpyHostToDevice, S.copy_stream));
    CUDA_OK(cudaEventRecord(S.ev_copy[b], S.copy_stream));
    CUDA_OK(cudaStreamWaitEvent(S.stream, S.ev_copy[b], 0));
    launch_transpose(st + b * half, dev + b0 * item, nb, NLEV, C);
    CUDA_OK(cudaEventRecord(S.ev_xpose[b], S.stream));
  }
  CUDA_OK(cudaStreamSynchronize(S.stream));
}
// device [b][C][NLEV] -> host F90 [b][NLEV][C]
static void push_field(const double* dev, double* host, size_t nbatch, int C) {
  if (!nbatch) return;
  const size_t item = (size_t)NLEV * C;
  const size_t per = std::max<size_t>(1, STAGE_BYTES / (item * 8));
  const size_t half = std::min(nbatch, per) * item;
  double* st = (double*)scratch(2 * half * 8);
  int i = 0;
  for (size_t b0 = 0; b0 < nbatch; b0 += per, ++i) {
    const size_t nb = std::min(per, nbatch - b0);
    const int b = i & 1;
    if (i >= 2) CUDA_OK(cudaStreamWaitEvent(S.stream, S.ev_copy[b], 0));  // chunk i - 2 has reached the host
    launch_transpose(dev + b0 * item, st + b * half, nb, C, NLEV);
    CUDA_OK(cudaEventRecord(S.ev_xpose[b], S.stream));
    CUDA_OK(cudaStreamWaitEvent(S.copy_stream, S.ev_xpose[b], 0));
    CUDA_OK(cudaMemcpyAsync(host + b0 * item, st + b * half, nb * item * 8, cudaMemcpyDeviceToHost, S.copy_stream));
    CUDA_OK(cudaEventRecord(S.ev_copy[b], S.copy_stream));
  }
  CUDA_OK(cudaStreamSynchronize(S.copy_stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
}

void* diag_scratch(size_t bytes) {
  if (bytes > S.diag_scratch_bytes) {
    if (S.diag_scratch) CUDA_OK(cudaFree(S.diag_scratch));
    CUDA_OK(cudaMalloc(&S.diag_scratch, bytes));
    S.diag_scratch_bytes = bytes;
  }
  return S.diag_scratch;
}

void push_Q_to_host(double* host_q) { push_field(S.Q, host_q, (size_t)S.nelemd * QSIZE_D, NPSQ); }

static void free_all() {
  dfree(S.geo); dfree(S.metinv); dfree(S.tensorvisc); dfree(S.vec_sph2cart);
  dfree(S.v); dfree(S.t); dfree(S.dp3d); dfree(S.ps_v);
  dfree(S.phi); dfree(S.omega_p); dfree(S.eta_dot_dpdn); dfree(S.derived_vn0); dfree(S.derived_dp);
  dfree(S.divdp); dfree(S.divdp_proj); dfree(S.dpdiss_ave); dfree(S.dpdiss_biharmonic);
  dfree(S.vtens); dfree(S.ttens); dfree(S.dptens); dfree(S.vstar); dfree(S.dpdissk); dfree(S.dp_star);
  dfree(S.qdp); dfree(S.qtens_biharmonic); dfree(S.qlim); dfree(S.qlim_x); dfree(S.Q);
  dfree(S.fm); dfree(S.ft); dfree(S.fq); dfree(S.hs_lat); dfree(S.hs_hyam);
  free_exchange_plan();
  dfree(S.invalid_flag);
  if (S.h_invalid) { cudaFreeHost(S.h_invalid); S.h_invalid = nullptr; }
  if (S.scratch) { cudaFree(S.scratch); S.scratch = nullptr; S.scratch_bytes = 0; }
  if (S.diag_scratch) { cudaFree(S.diag_scratch); S.diag_scratch = nullptr; S.diag_scratch_bytes = 0; }
  for (auto& d : S.diag) d = nullptr;  // the F90 accumulators belong to the finished run
  S.conn.clear();
}

// ---- time levels (TimeLevel.hpp) ----------------------------------------------------------
static void update_tracers_levels() {  // :58-67
  const int i_temp = S.nstep / S.p.qsplit;
  if (i_temp % 2 == 0) { S.n0_qdp = 0; S.np1_qdp = 1; }
  else { S.n0_qdp = 1; S.np1_qdp = 0; }
}
static void update_dynamics_levels() {  // LEAPFROG :37-56
  const int tmp = S.np1;
  S.np1 = S.nm1; S.nm1 = S.n0; S.n0 = tmp;
  ++S.nstep;
}

// prim_advance_exp.cpp:113-161
static void u3_5stage_timestep(int nm1, int n0, int np1, int n0_qdp, double dt, double eta_ave_w) {
  HXX_TIMER("tl-ae U3-5stage_timestep");
  caar_run(n0, n0, nm1, dt / 5.0, eta_ave_w / 4.0, n0_qdp, true);
  caar_run(n0, nm1, np1, dt / 5.0, 0.0, n0_qdp, true);
  caar_run(n0, np1, np1, dt / 3.0, 0.0, n0_qdp, true);
  caar_run(n0, np1, np1, 2.0 * dt / 3.0, 0.0, n0_qdp, true);
  rk_combine(nm1, n0);
  caar_run(nm1, np1, np1, 3.0 * dt / 4.0, 3.0 * eta_ave_w / 4.0, n0_qdp, true);
}

// prim_advec_tracers_remap.cpp:32-90
static void prim_advec_tracers_remap_RK2(double dt) {
  HXX_TIMER("tl-at prim_advec_tracers_remap_RK2");
  update_tracers_levels();
  S.rhs_viss = 0.0;  // EulerStepFunctor::reset
  {
    HXX_TIMER("tl-at precompute_divdp");
    euler_precompute_divdp();
  }
  {
    HXX_TIMER("tl-at esf-0");
    euler_step(S.np1_qdp, S.n0_qdp, dt / 2.0, 0.0, DSS_DIV_VDP_AVE);
  }
  {
    HXX_TIMER("tl-at esf-1");
    euler_step(S.np1_qdp, S.np1_qdp, dt / 2.0, 1.0, DSS_ETA);
  }
  {
    // the last stage also applies qdp_time_avg(n0_qdp, np1_qdp) (:88, "tl-at qdp_time_avg"), fused into its stores
    HXX_TIMER("tl-at esf-2");
    euler_step(S.np1_qdp, S.np1_qdp, dt / 2.0, 2.0, DSS_OMEGA, S.n0_qdp);
  }
}

// prim_step.cpp:20-103
static void prim_step(double dt) {
  HXX_TIMER("tl-s prim_step");
  {
    HXX_TIMER("tl-s deep_copy+derived_dp");
    prim_step_init(S.n0);
  }
  for (int nq = 0; nq < S.p.qsplit; ++nq) {
    HXX_TIMER("tl-ae prim_advance_exp");
    if (nq > 0) update_dynamics_levels();
    // prim_advance_exp.cpp:25-111
    S.n0_qdp = -1;
    if (S.p.moist) update_tracers_levels();
    const double eta_ave_w = 1.0 / S.p.qsplit;
    u3_5stage_timestep(S.nm1, S.n0, S.np1, S.n0_qdp, dt, eta_ave_w);
    HXX_TIMER("tl-ae advance_hypervis_dp");
    hypervis_run(S.np1, dt, eta_ave_w);
  }
  if (S.p.qsize > 0) {
    HXX_TIMER("tl-s prim_advec_tracers_remap");
    prim_advec_tracers_remap_RK2(dt * S.p.qsplit);
  }
}

struct NamedField { const char* nm; double* p; size_t n; };
static bool field_by_name(const char* name, NamedField& out) {
  const size_t ne = (size_t)S.nelemd, f3 = ne * NLF;
  const NamedField tab[] = {
      {"v", S.v, f3 * NTL * 2}, {"t", S.t, f3 * NTL}, {"dp3d", S.dp3d, f3 * NTL}, {"ps_v", S.ps_v, ne * NTL * NPSQ},
      {"phi", S.phi, f3}, {"omega_p", S.omega_p, f3}, {"eta_dot_dpdn", S.eta_dot_dpdn, f3},
      {"derived_vn0", S.derived_vn0, f3 * 2}, {"derived_dp", S.derived_dp, f3}, {"divdp", S.divdp, f3},
      {"divdp_proj", S.divdp_proj, f3}, {"dpdiss_ave", S.dpdiss_ave, f3},
      {"dpdiss_biharmonic", S.dpdiss_biharmonic, f3}, {"qdp", S.qdp, f3 * QNTL * QSIZE_D},
      {"qtens_biharmonic", S.qtens_biharmonic, f3 * QSIZE_D}, {"qlim", S.qlim, ne * QSIZE_D * 2 * NLEV},
      {"Q", S.Q, f3 * QSIZE_D}, {"vtens", S.vtens, f3 * 2}, {"ttens", S.ttens, f3}, {"dptens", S.dptens, f3},
      {"vstar", S.vstar, f3 * 2}, {"dpdissk", S.dpdissk, f3}, {"dp_star", S.dp_star, f3},
      {"fm", S.fm, f3 * 2}, {"ft", S.ft, f3}, {"fq", S.fq, f3 * QSIZE_D}};
  for (const auto& t : tab)
    if (!std::strcmp(t.nm, name) && t.p) { out = t; return true; }
  return false;
}

static void need_session(const char* who) {
  if (!S.active) {
    char msg[256];
    std::snprintf(msg, sizeof msg, "%s: initialize_hommexx_session was not called", who);
    runtime_abort(msg, 13);
  }
}

static void option_error(const char* loc, const char* opt, double value) {
  char msg[512];
  std::snprintf(msg, sizeof msg, "Error in %s: unsupported value '%g' for input parameter '%s'.", loc, value, opt);
  runtime_abort(msg, 11);
}

}  // namespace hxx

using namespace hxx;

extern "C" {

// ---- section B ----------------------------------------------------------------------------
int hommexx_b200_nlev(void) { return NLEV; }
int hommexx_b200_qsize_d(void) { return QSIZE_D; }
const char* hommexx_b200_backend(void) { return "cuda-sm100a"; }
int64_t hommexx_b200_launch_count(void) { return S.launches; }
void hommexx_b200_sync(void) {
  if (S.stream) CUDA_OK(cudaStreamSynchronize(S.stream));
}

int hommexx_b200_nccl_unique_id(void* out128) {
#ifdef HXX_WITH_NCCL
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return 1;
  std::memcpy(out128, &id, sizeof id);
  return 0;
#else
  (void)out128;
  return 2;
#endif
}

void hommexx_b200_set_comm(int rank, int size, int device, const void* nccl_unique_id) {
  if (S.active) runtime_abort("hommexx_b200_set_comm must precede initialize_hommexx_session", 13);
  S.rank = rank; S.nranks = size; S.device = device; S.comm_set = true;
  if (size > 1) {
#ifdef HXX_WITH_NCCL
    if (!nccl_unique_id) runtime_abort("hommexx_b200_set_comm: size > 1 needs an ncclUniqueId", 13);
    CUDA_OK(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof id);
    ncclComm_t c;
    if (ncclCommInitRank(&c, size, id, rank) != ncclSuccess) runtime_abort("ncclCommInitRank failed", 1);
    S.nccl = c;
#else
    runtime_abort("hommexx_b200: built without NCCL; multi-GPU runs are unavailable", 12);
#endif
  }
}

void hommexx_b200_profile(unsigned long long mask) {
  for (auto& p : g_probes) g_probe_pool.push_back(p);
  g_probes.clear();
  S.profile_mask = mask;
  for (auto& c : S.launches_by) c = 0;
}

int hommexx_b200_kernel_id(const char* name) {
  for (int i = 0; i < K_COUNT; ++i)
    if (!std::strcmp(kernel_names[i], name)) return i;
  return -1;
}
const char* hommexx_b200_kernel_name(int id) { return id >= 0 && id < K_COUNT ? kernel_names[id] : nullptr; }

// total device milliseconds and launch count of one kernel class since hommexx_b200_profile()
double hommexx_b200_profile_read(int id, int64_t* launches) {
  if (S.stream) CUDA_OK(cudaStreamSynchronize(S.stream));
  double ms = 0.0;
  for (auto& p : g_probes)
    if (p.id == id) {
      float t = 0.f;
      CUDA_OK(cudaEventElapsedTime(&t, p.a, p.b));
      ms += t;
    }
  if (launches) *launches = (id >= 0 && id < K_COUNT) ? S.launches_by[id] : 0;
  return ms;
}

void hommexx_b200_event_record(int slot) {
  if (!g_marks_made) { for (auto& e : g_marks) CUDA_OK(cudaEventCreate(&e)); g_marks_made = true; }
  CUDA_OK(cudaEventRecord(g_marks[slot & 15], S.stream));
}
double hommexx_b200_event_elapsed_ms(int a, int b) {
  float t = 0.f;
  CUDA_OK(cudaEventSynchronize(g_marks[b & 15]));
  CUDA_OK(cudaEventElapsedTime(&t, g_marks[a & 15], g_marks[b & 15]));
  return t;
}

// ---- section A ----------------------------------------------------------------------------
void reset_cxx_comm(const int* f_comm) { (void)f_comm; }

void initialize_hommexx_session(void) {
  if (S.active) return;
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) runtime_abort("hommexx_b200: no CUDA device visible (there is no CPU fallback)", 1);
  if (!S.comm_set) S.device = 0;
  CUDA_OK(cudaSetDevice(S.device % ndev));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, S.device % ndev));
  if (prop.major < 10) {
    char msg[256];
    std::snprintf(msg, sizeof msg, "hommexx_b200: built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    runtime_abort(msg, 1);
  }
  CUDA_OK(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
  {
    // the halo stream outranks the compute stream: its (small) pack kernels get the next free SM slots
    int lo = 0, hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_OK(cudaStreamCreateWithPriority(&S.comm_stream, cudaStreamNonBlocking, hi));
  }
  CUDA_OK(cudaStreamCreateWithFlags(&S.copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    CUDA_OK(cudaEventCreateWithFlags(&S.ev_copy[b], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&S.ev_xpose[b], cudaEventDisableTiming));
  }
  CUDA_OK(cudaEventCreateWithFlags(&S.ev_produced, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&S.ev_halo, cudaEventDisableTiming));
  S.active = true;
  ++S.session_id;
  S.launches = 0;
  if (S.rank == 0 && std::getenv("HXX_BANNER"))
    std::printf("HOMMEXX-B200 session: %s, %d SMs, nlev=%d qsize_d=%d, ranks=%d\n", prop.name,
                prop.multiProcessorCount, NLEV, QSIZE_D, S.nranks);
}

void finalize_hommexx_session(void) {
  if (!S.active) return;
  CUDA_OK(cudaStreamSynchronize(S.stream));
  CUDA_OK(cudaStreamSynchronize(S.comm_stream));
  free_all();
#ifdef HXX_WITH_NCCL
  if (S.nccl) { ncclCommDestroy((ncclComm_t)S.nccl); S.nccl = nullptr; }
#endif
  CUDA_OK(cudaStreamDestroy(S.stream));
  CUDA_OK(cudaStreamDestroy(S.comm_stream));
  CUDA_OK(cudaStreamDestroy(S.copy_stream));
  for (int b = 0; b < 2; ++b) {
    CUDA_OK(cudaEventDestroy(S.ev_copy[b]));
    CUDA_OK(cudaEventDestroy(S.ev_xpose[b]));
  }
  S.copy_stream = nullptr;
  CUDA_OK(cudaEventDestroy(S.ev_produced));
  CUDA_OK(cudaEventDestroy(S.ev_halo));
  S.stream = S.comm_stream = nullptr;
  S.ev_produced = S.ev_halo = nullptr;
  S.active = false;
  S.comm_set = false;
  S.rank = 0; S.nranks = 1;
  S.p.params_set = false;
  S.have_dvv = S.have_hv = false;
  S.nelemd = 0;
}

void init_connectivity(const int* num_local_elems) {
  need_session("init_connectivity");
  // A session that already holds a mesh is being re-initialised (a second driver bound to this library): the
  // first mesh's arrays — including the lazily allocated forcing arrays, sized for it — go first.
  if (S.geo || S.v || S.nbr8) {
    CUDA_OK(cudaStreamSynchronize(S.stream));
    CUDA_OK(cudaStreamSynchronize(S.comm_stream));
    free_all();
  }
  // Connectivity.cpp:40-75
  S.nelemd = *num_local_elems;
  S.conn.assign((size_t)S.nelemd * 8, ConnInfo{});
  for (int ie = 0; ie < S.nelemd; ++ie)
    for (int c = 0; c < 8; ++c) {
      ConnInfo& i = S.conn[(size_t)ie * 8 + c];
      i.l_lid = ie; i.l_pos = c; i.l_gid = -1;
      i.r_lid = i.r_gid = i.r_pos = -1;
      i.kind = 2; i.sharing = 2; i.direction = 2; i.remote_pid = -1;
    }
}

void add_connection(const int* l1, const int* g1, const int* p1, const int* r1, const int* l2, const int* g2,
                    const int* p2, const int* r2) {
  if (*l1 <= 0 || *g1 <= 0 || *p1 <= 0 || *r1 <= 0 || *l2 <= 0 || *g2 <= 0 || *p2 <= 0 || *r2 <= 0)
    runtime_abort("ERROR! We were assuming F90 indices started at 1, but it appears there is an exception.", 13);
  // ConnectivityHelpers.hpp:133-142
  static const int CONNECTION_DIRECTION[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}};
  // mpi_cxx_f90_interface.cpp:47-49
  const int fep = *p1 <= 4 ? ((*p1 - 1) + 2) % 4 : *p1 - 1;
  const int sep = *p2 <= 4 ? ((*p2 - 1) + 2) % 4 : *p2 - 1;
  if (*r1 - 1 != S.rank) return;  // Connectivity.cpp:89
  if (*l1 > S.nelemd) runtime_abort("add_connection: local element id out of range", 13);
  ConnInfo& i = S.conn[(size_t)(*l1 - 1) * 8 + fep];
  i.l_lid = *l1 - 1; i.l_gid = *g1 - 1; i.l_pos = fep;
  i.r_lid = *l2 - 1; i.r_gid = *g2 - 1; i.r_pos = sep;
  i.kind = fep < 4 ? 0 : 1;
  i.direction = fep < 4 ? CONNECTION_DIRECTION[fep][sep] : 0;
  i.remote_pid = *r2 - 1;
  i.sharing = (i.remote_pid == S.rank) ? 0 : 1;
  if (i.sharing == 1 && S.nranks == 1)
    runtime_abort("add_connection: remote connection but the session has one rank (call hommexx_b200_set_comm)", 13);
}

void finalize_connectivity(void) {}

void init_derivative_c(const double* const* dvv) {
  need_session("init_derivative_c");
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) S.hc.dvv[i][j] = (*dvv)[i * NP + j];  // Derivative.cpp:20-32
  S.have_dvv = true;
  upload_constants();
}

void init_simulation_params_c(const int* remap_alg, const int* limiter_option, const int* rsplit, const int* qsplit,
                              const int* time_step_type, const int* energy_fixer, const int* qsize,
                              const int* state_frequency, const double* nu, const double* nu_p, const double* nu_q,
                              const double* nu_s, const double* nu_div, const double* nu_top,
                              const int* hypervis_order, const int* hypervis_subcycle,
                              const double* hypervis_scaling, const int* ftype, const bool* prescribed_wind,
                              const bool* moisture, const bool* disable_diagnostics, const bool* use_cpstar,
                              const bool* use_semi_lagrangian_transport) {
  const char* loc = "init_simulation_params_c";
  (void)energy_fixer;
  // cxx_f90_interface.cpp:43-52
  if (*remap_alg != 1 && *remap_alg != 2) option_error(loc, "vert_remap_q_alg", *remap_alg);
  if (*prescribed_wind) option_error(loc, "prescribed_wind", 1);
  if (*hypervis_order != 2) option_error(loc, "hypervis_order", *hypervis_order);
  if (*use_semi_lagrangian_transport) option_error(loc, "use_semi_lagrangian_transport", 1);
  if (*time_step_type != 5) option_error(loc, "time_step_type", *time_step_type);
  if (*limiter_option != 8 && *limiter_option != 9) option_error(loc, "limiter_option", *limiter_option);
  if (*ftype != -1 && *ftype != 0 && *ftype != 2) option_error(loc, "ftype", *ftype);
  if (!(*nu_p > 0.0)) option_error(loc, "nu_p", *nu_p);
  if (!(*nu > 0.0)) option_error(loc, "nu", *nu);
  if (!(*nu_div > 0.0)) option_error(loc, "nu_div", *nu_div);
  if (*qsize > QSIZE_D) runtime_abort("init_simulation_params_c: qsize exceeds the QSIZE_D of this build", 13);
  Params& p = S.p;
  p.remap_alg = *remap_alg; p.limiter_option = *limiter_option; p.rsplit = *rsplit; p.qsplit = *qsplit;
  p.time_step_type = *time_step_type; p.qsize = *qsize; p.state_frequency = *state_frequency;
  p.nu = *nu; p.nu_p = *nu_p; p.nu_q = *nu_q; p.nu_s = *nu_s; p.nu_div = *nu_div; p.nu_top = *nu_top;
  p.hypervis_order = *hypervis_order; p.hypervis_subcycle = *hypervis_subcycle;
  p.hypervis_scaling = *hypervis_scaling; p.ftype = *ftype;
  p.moist = *moisture; p.disable_diagnostics = *disable_diagnostics; p.use_cpstar = *use_cpstar;
  // :88-100
  if (p.nu != p.nu_div) {
    const double ratio = p.nu_div / p.nu;
    if (p.hypervis_scaling != 0.0) { p.nu_ratio1 = ratio * ratio; p.nu_ratio2 = 1.0; }
    else { p.nu_ratio1 = ratio; p.nu_ratio2 = ratio; }
  } else { p.nu_ratio1 = 1.0; p.nu_ratio2 = 1.0; }
  p.consthv = (p.hypervis_scaling == 0.0);
  p.params_set = true;
}

void init_hvcoord_c(const double* ps0, const double* const* am, const double* const* ai, const double* const* bm,
                    const double* const* bi) {
  need_session("init_hvcoord_c");
  (void)am; (void)bm;
  // HybridVCoord.cpp:15-53,112-156
  DevConst& c = S.hc;
  c.ps0 = *ps0;
  std::memcpy(c.hyai, *ai, (NLEV + 1) * sizeof(double));
  std::memcpy(c.hybi, *bi, (NLEV + 1) * sizeof(double));
  c.hyai0 = c.hyai[0];
  for (int k = 0; k < NLEV; ++k) {
    c.dai[k] = c.hyai[k + 1] - c.hyai[k];
    c.dbi[k] = c.hybi[k + 1] - c.hybi[k];
    c.dp0[k] = c.dai[k] * c.ps0 + c.dbi[k] * c.ps0;
  }
  S.have_hv = true;
  upload_constants();
}

void init_elements_2d_c(const int* num_elems, const double* const* D, const double* const* Dinv,
                        const double* const* fcor, const double* const* mp, const double* const* spheremp,
                        const double* const* rspheremp, const double* const* metdet, const double* const* metinv,
                        const double* const* phis, const double* const* tensorvisc,
                        const double* const* vec_sph2cart, const bool* consthv) {
  need_session("init_elements_2d_c");
  const int n = *num_elems;
  if (n != S.nelemd) runtime_abort("init_elements_2d_c: element count differs from init_connectivity", 13);
  // Elements.cpp:81-190: F90 arrays read linearly as [ie][a][b][igp][jgp]
  std::vector<double> geo((size_t)n * NPSQ * GEO_N, 0.0);
  for (int ie = 0; ie < n; ++ie)
    for (int p = 0; p < NPSQ; ++p) {
      double* g = &geo[((size_t)ie * NPSQ + p) * GEO_N];
      for (int ab = 0; ab < 4; ++ab) {
        g[G_DINV00 + ab] = (*Dinv)[((size_t)ie * 4 + ab) * NPSQ + p];
        g[G_D00 + ab] = (*D)[((size_t)ie * 4 + ab) * NPSQ + p];
      }
      const size_t s = (size_t)ie * NPSQ + p;
      g[G_METDET] = (*metdet)[s];
      g[G_RMETDET_R] = 1.0 / (*metdet)[s] * rrearth;  // SphereOperators.hpp:388, evaluated once
      g[G_SPHEREMP] = (*spheremp)[s];
      g[G_RSPHEREMP] = (*rspheremp)[s];
      g[G_FCOR] = (*fcor)[s];
      g[G_PHIS] = (*phis)[s];
      g[G_MP] = (*mp)[s];
      g[G_INV_SPHEREMP] = 1.0 / (*spheremp)[s];
    }
  S.geo = dalloc(geo.size());
  CUDA_OK(cudaMemcpyAsync(S.geo, geo.data(), geo.size() * 8, cudaMemcpyHostToDevice, S.stream));
  S.metinv = dalloc((size_t)n * 4 * NPSQ);
  CUDA_OK(cudaMemcpyAsync(S.metinv, *metinv, (size_t)n * 4 * NPSQ * 8, cudaMemcpyHostToDevice, S.stream));
  if (!*consthv) {
    S.tensorvisc = dalloc((size_t)n * 4 * NPSQ);
    S.vec_sph2cart = dalloc((size_t)n * 6 * NPSQ);
    CUDA_OK(cudaMemcpyAsync(S.tensorvisc, *tensorvisc, (size_t)n * 4 * NPSQ * 8, cudaMemcpyHostToDevice, S.stream));
    CUDA_OK(cudaMemcpyAsync(S.vec_sph2cart, *vec_sph2cart, (size_t)n * 6 * NPSQ * 8, cudaMemcpyHostToDevice, S.stream));
  }
  CUDA_OK(cudaStreamSynchronize(S.stream));
  const size_t f3 = (size_t)n * NLF;
  S.v = dalloc(f3 * NTL * 2); S.t = dalloc(f3 * NTL); S.dp3d = dalloc(f3 * NTL); S.ps_v = dalloc((size_t)n * NTL * NPSQ);
  S.phi = dalloc(f3); S.omega_p = dalloc(f3); S.eta_dot_dpdn = dalloc(f3); S.derived_vn0 = dalloc(f3 * 2);
  S.derived_dp = dalloc(f3); S.divdp = dalloc(f3); S.divdp_proj = dalloc(f3); S.dpdiss_ave = dalloc(f3);
  S.dpdiss_biharmonic = dalloc(f3);
  S.vtens = dalloc(f3 * 2); S.ttens = dalloc(f3); S.dptens = dalloc(f3);
  S.vstar = dalloc(f3 * 2); S.dpdissk = dalloc(f3); S.dp_star = dalloc(f3);
  S.qdp = dalloc(f3 * QNTL * QSIZE_D); S.qtens_biharmonic = dalloc(f3 * QSIZE_D); S.Q = dalloc(f3 * QSIZE_D);
  S.qlim = dalloc((size_t)n * QSIZE_D * 2 * NLEV); S.qlim_x = dalloc((size_t)n * QSIZE_D * 2 * NLEV);
  CUDA_OK(cudaMalloc(&S.invalid_flag, sizeof(int)));
  CUDA_OK(cudaMemsetAsync(S.invalid_flag, 0, sizeof(int), S.stream));
  CUDA_OK(cudaMallocHost(&S.h_invalid, sizeof(int)));
  *S.h_invalid = 0;
  CUDA_OK(cudaStreamSynchronize(S.stream));
}

void init_elements_states_c(const double* const* fv, const double* const* ft, const double* const* fdp,
                            const double* const* fq, const double* const* fps) {
  need_session("init_elements_states_c");
  const size_t n = S.nelemd;
  pull_field(*fv, S.v, n * NTL, 2 * NPSQ);
  pull_field(*ft, S.t, n * NTL, NPSQ);
  pull_field(*fdp, S.dp3d, n * NTL, NPSQ);
  pull_field(*fq, S.qdp, n * QNTL * QSIZE_D, NPSQ);
  CUDA_OK(cudaMemcpyAsync(S.ps_v, *fps, n * NTL * NPSQ * 8, cudaMemcpyHostToDevice, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
}

void init_diagnostics_c(double* const* a0, double* const* a1, double* const* a2, double* const* a3,
                        double* const* a4, double* const* a5, double* const* a6, double* const* a7) {
  S.diag[0] = *a0; S.diag[1] = *a1; S.diag[2] = *a2; S.diag[3] = *a3;
  S.diag[4] = *a4; S.diag[5] = *a5; S.diag[6] = *a6; S.diag[7] = *a7;
}

void init_boundary_exchanges_c(void) {
  need_session("init_boundary_exchanges_c");
  if (!S.geo) runtime_abort("init_boundary_exchanges_c: elements not initialised", 13);
  build_exchange_plan();
}

void init_time_level_c(const int* nm1, const int* n0, const int* np1, const int* nstep, const int* nstep0) {
  S.nm1 = *nm1 - 1; S.n0 = *n0 - 1; S.np1 = *np1 - 1; S.nstep = *nstep; S.nstep0 = *nstep0;
}

// prim_driver.cpp:31-156
void prim_run_subcycle_c(const double* dt, int* nstep, int* nm1, int* n0, int* np1, const int* last_time_step) {
  need_session("prim_run_subcycle_c");
  HXX_TIMER("tl-sc prim_run_subcycle_c");
  if (!S.p.params_set) runtime_abort("prim_run_subcycle_c: simulation params not set", 13);
  if (!S.nodes && S.nelemd > 0 && !S.nbr8) runtime_abort("prim_run_subcycle_c: init_boundary_exchanges_c not called", 13);
  const double dt_q = *dt * S.p.qsplit;
  double dt_remap = dt_q;
  int nstep_end = S.nstep + S.p.qsplit;
  if (S.p.rsplit > 0) {
    dt_remap = dt_q * S.p.rsplit;
    nstep_end = S.nstep + S.p.qsplit * S.p.rsplit;
  }
  // :51-64
  bool compute_diagnostics =
      nstep_end % S.p.state_frequency == 0 || nstep_end == S.nstep0 || nstep_end >= *last_time_step;
  if (S.p.disable_diagnostics) compute_diagnostics = false;
  if (compute_diagnostics) {
    update_tracers_levels();
    prim_diag_scalars(true, 3);
    prim_energy_halftimes(true, 2);
  }
  update_tracers_levels();
  // :76-82 — every standalone namelist has ftype = 0: the pass runs with zero forcing arrays
  if (S.p.ftype == 0) {
    HXX_TIMER("ApplyCAMForcing");
    apply_cam_forcing(dt_remap, true);
  } else if (S.p.ftype == 2) {
    HXX_TIMER("ApplyCAMForcing_dynamics");
    apply_cam_forcing(dt_remap, false);
  }
  if (compute_diagnostics) {
    prim_energy_halftimes(true, 0);
    prim_diag_scalars(true, 0);
  }
  {
    HXX_TIMER("tl-sc dp3d-from-ps");
    dp3d_from_ps(S.n0);  // :98-111
  }
  {
    HXX_TIMER("tl-sc prim_step-loop");
    prim_step(*dt);
    for (int r = 1; r < S.p.rsplit; ++r) {
      update_dynamics_levels();
      prim_step(*dt);
    }
  }
  update_tracers_levels();
  HXX_TIMER("tl-sc vertical_remap");
  // :131 and :138 — the remap kernel also stores Q = Qdp / dp for every tracer (update_q fused
  // into its tracer store); without tracers there is nothing to update
  vertical_remap(S.np1, S.np1_qdp, dt_remap);
  check_remap_flag();                // RemapFunctor.hpp:190-198 (one host sync per call)
  check_comm_errors();
  if (compute_diagnostics) {
    prim_diag_scalars(false, 1);
    prim_energy_halftimes(false, 1);
  }
  update_dynamics_levels();
  *nstep = S.nstep; *nm1 = S.nm1; *n0 = S.n0; *np1 = S.np1;
}

// cxx_f90_interface.cpp:126-154
void cxx_push_results_to_f90(double* const* fv, double* const* ft, double* const* fdp, double* const* fq,
                             double* const* fQ, double* const* fps, double* const* fom) {
  need_session("cxx_push_results_to_f90");
  const size_t n = S.nelemd;
  push_field(S.v, *fv, n * NTL, 2 * NPSQ);
  push_field(S.t, *ft, n * NTL, NPSQ);
  push_field(S.dp3d, *fdp, n * NTL, NPSQ);
  push_field(S.qdp, *fq, n * QNTL * QSIZE_D, NPSQ);
  push_field(S.Q, *fQ, n * QSIZE_D, NPSQ);
  push_field(S.omega_p, *fom, n, NPSQ);
  CUDA_OK(cudaMemcpyAsync(*fps, S.ps_v, n * NTL * NPSQ * 8, cudaMemcpyDeviceToHost, S.stream));
  CUDA_OK(cudaStreamSynchronize(S.stream));
}

// cxx_f90_interface.cpp:180-205: FM, FT (and FQ when ftype == 0) host -> device; Tracers::push_qdp
// (Tracers.cpp:46-51) then copies the device qdp back INTO the F90 array
void f90_push_forcing_to_cxx(double* fm, double* ft, double* fq, double* qdp) {
  need_session("f90_push_forcing_to_cxx");
  const size_t n = S.nelemd, f3 = n * NLF;
  if (!S.fm) S.fm = dalloc(f3 * 2);
  if (!S.ft) S.ft = dalloc(f3);
  pull_field(fm, S.fm, n, 2 * NPSQ);
  pull_field(ft, S.ft, n, NPSQ);
  if (S.p.ftype == 0) {
    if (!S.fq) S.fq = dalloc(f3 * QSIZE_D);
    pull_field(fq, S.fq, n * QSIZE_D, NPSQ);
  }
  push_field(S.qdp, qdp, n * QNTL * QSIZE_D, NPSQ);
}
// cxx_f90_interface.cpp:157-178
void cxx_push_forcing_to_f90(double* fm, double* ft, double* fq) {
  need_session("cxx_push_forcing_to_f90");
  const size_t n = S.nelemd, f3 = n * NLF;
  if (!S.fm) S.fm = dalloc(f3 * 2);
  if (!S.ft) S.ft = dalloc(f3);
  push_field(S.fm, fm, n, 2 * NPSQ);
  push_field(S.ft, ft, n, NPSQ);
  if (S.p.ftype == 0) {
    if (!S.fq) S.fq = dalloc(f3 * QSIZE_D);
    push_field(S.fq, fq, n * QSIZE_D, NPSQ);
  }
}

// ---- section C: phase-level entry points --------------------------------------------------
void hxx_caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, int with_dss) {
  S.store_phi = true;
  caar_run(nm1, n0, np1, dt, eta_ave_w, n0_qdp, with_dss != 0);
  S.store_phi = false;
}
void hxx_rk_combine(int nm1, int n0) { rk_combine(nm1, n0); }
void hxx_hypervis_run(int np1, double dt, double eta_ave_w) { hypervis_run(np1, dt, eta_ave_w); }
void hxx_euler_reset(void) { S.rhs_viss = 0.0; }
void hxx_euler_precompute_divdp(void) { euler_precompute_divdp(); }
void hxx_euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt) {
  euler_step(np1_qdp, n0_qdp, dt, rhs_multiplier, dss_opt);
}
void hxx_euler_qdp_time_avg(int n0_qdp, int np1_qdp) { euler_qdp_time_avg(n0_qdp, np1_qdp); }
void hxx_vertical_remap(int np1, int np1_qdp, double dt) {
  vertical_remap(np1, np1_qdp, dt);
  check_remap_flag();
}
void hxx_update_q(int np1_qdp, int np1) { update_q(np1_qdp, np1); }
void hxx_prim_step_init(int n0) { prim_step_init(n0); }
void hxx_apply_forcing(double dt) {
  update_tracers_levels();
  if (S.p.ftype == 0) apply_cam_forcing(dt, true);
  else if (S.p.ftype == 2) apply_cam_forcing(dt, false);
}
void hxx_held_suarez_forcing(const double* lat, const double* hyam, const double* hybm) {
  need_session("hxx_held_suarez_forcing");
  held_suarez_forcing(lat, hyam, hybm);
}
void hxx_diagnostics(int before_advance, int ivar_scalars, int ivar_energy) {
  update_tracers_levels();
  prim_diag_scalars(before_advance != 0, ivar_scalars);
  prim_energy_halftimes(before_advance != 0, ivar_energy);
}

void hxx_exchange(const char* field_set, int rspheremp) {
  FieldList fl;
  if (!std::strncmp(field_set, "caar:", 5)) fl = fields_caar(std::atoi(field_set + 5));
  else if (!std::strcmp(field_set, "hv")) fl = fields_hv();
  else if (!std::strncmp(field_set, "euler:", 6)) {
    int tq = 0, opt = 0;
    std::sscanf(field_set + 6, "%d:%d", &tq, &opt);
    fl = fields_euler(tq, opt);
  } else if (!std::strcmp(field_set, "qtens")) fl = fields_qtens();
  else if (!std::strcmp(field_set, "qlim")) { minmax_exchange(); return; }
  else runtime_abort("hxx_exchange: unknown field set", 11);
  dss_exchange(fl, rspheremp != 0);
  if (rspheremp) scale_interior_rspheremp(fl);
}

int64_t hxx_get_field(const char* name, double* out) {
  NamedField f;
  if (!field_by_name(name, f)) return 0;
  if (out) {
    CUDA_OK(cudaMemcpyAsync(out, f.p, f.n * 8, cudaMemcpyDeviceToHost, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
  }
  return (int64_t)f.n;
}
int64_t hxx_set_field(const char* name, const double* in) {
  NamedField f;
  if (!field_by_name(name, f)) return 0;
  if (in) {
    CUDA_OK(cudaMemcpyAsync(f.p, in, f.n * 8, cudaMemcpyHostToDevice, S.stream));
    CUDA_OK(cudaStreamSynchronize(S.stream));
  }
  return (int64_t)f.n;
}

}  // extern "C"
