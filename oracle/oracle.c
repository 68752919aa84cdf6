/* oracle.c — CPU restatement of HOMMEXX's preqx timestep.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA product in hommexx_b200/csrc. Nothing in the product
 * path may link, load or call it; only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline do. It exports the same C ABI as the product (include/hommexx_b200.h) so the same
 * driver can run either library on identical inputs.
 *
 * Every routine follows the NON-CUDA branch of the reference functor it cites, with the
 * reference's operation order (VECTOR_SIZE = 1 semantics, sums over m = 0..3 ascending,
 * accumulators starting from 0), compiled with -ffp-contract=off so no FMA contraction
 * changes the rounding. File:line citations are relative to /root/reference/src/share/cxx/.
 *
 * PINNED TO THE REFERENCE'S OWN BUILD (tests/test_oracle_vs_reference.py): oracle/_ref/libref_hommexx_*.so is the
 * reference's src/share/cxx compiled from its sources (oracle/Makefile ref_full, against the serial Kokkos stand-in of
 * oracle/ref_shim, VECTOR_SIZE 1, -ffp-contract=off); this file is bit-identical to it on v, T, dp3d, ps_v, Qdp, Q,
 * omega_p after 10-12 dynamics steps in 11 configurations (every option variant, 40 distinct tracers), in three
 * forced runs and on the diagnostics accumulators. Earlier pins, still run: the reference's known-answer vectors
 * test/unit_tests/inputs/{gradient,divergence,vorticity}_sphere_np4.in; bitwise agreement of the PPM remap with the
 * reference's plain-C++ twin src/preqx/unit_tests/remap.cpp; the limiter property tests of preqx_ut.cpp:1335-1531;
 * CAM forcing and diagnostics against a numpy restatement of the reference formulas; the rsplit = 0 path against
 * the vertically Lagrangian one. What this file adds over the reference build are the phase-level hooks of section C
 * of the ABI (the reference exposes its functors only as C++ classes), which the per-phase CUDA tests need.
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hommexx_b200.h"

#define NP 4
#define NPSQ 16
#define NTL 3  /* NUM_TIME_LEVELS */
#define QNTL 2 /* Q_NUM_TIME_LEVELS */

/* PhysicalConstants.hpp:17-22 */
static const double Rwater_vapor = 461.5;
static const double Rgas = 287.04;
static const double cp = 1005.0;
#define kappa (Rgas / cp)
#define rrearth (1.0 / 6.376e6)

enum { DSS_ETA = 0, DSS_OMEGA = 1, DSS_DIV_VDP_AVE = 2 }; /* HommexxEnums.hpp DSSOption; index = m_bes slot */

typedef struct {
  int lid, gid, pos;
} LidGidPos;
typedef struct {
  LidGidPos local, remote;
  int kind;      /* 0 edge, 1 corner, 2 missing */
  int sharing;   /* 0 local, 1 shared, 2 missing */
  int direction; /* 0 fwd, 1 bwd */
  int remote_pid;
} ConnectionInfo;

static struct Oracle {
  int nlev, qsize_d; /* set with oracle_set_dims before init */
  int nelemd;
  bool session;
  /* SimulationParams.hpp */
  int remap_alg, limiter_option, rsplit, qsplit, time_step_type, qsize, state_frequency, ftype;
  double nu, nu_p, nu_q, nu_s, nu_div, nu_top, hypervis_scaling, nu_ratio1, nu_ratio2;
  int hypervis_order, hypervis_subcycle;
  bool moist, disable_diagnostics, use_cpstar, consthv, params_set;
  /* TimeLevel.hpp */
  int nm1, n0, np1, nstep, nstep0, n0_qdp, np1_qdp;
  /* Derivative / HybridVCoord */
  double dvv[NP][NP];
  double ps0, hyai0, *hyai, *hybi, *dai, *dbi, *dp0;
  /* Elements 2d: [ie][2][2][16] and [ie][16] */
  double *d, *dinv, *metinv, *tensorvisc, *vec_sph2cart;
  double *fcor, *mp, *spheremp, *rspheremp, *metdet, *phis;
  /* Elements states: v [ie][3][2][16][nlev]; t, dp3d [ie][3][16][nlev]; ps_v [ie][3][16] */
  double *v, *t, *dp3d, *ps_v;
  double *phi, *omega_p, *eta_dot_dpdn, *derived_vn0, *derived_dp, *divdp, *divdp_proj, *dpdiss_ave,
      *dpdiss_biharmonic;
  double *vtens, *ttens, *dptens;
  double *vstar, *dpdissk, *dp_star;
  /* Tracers: qdp [ie][2][qsize_d][16][nlev]; qtens_biharmonic, Q [ie][qsize_d][16][nlev];
     qlim [ie][qsize_d][2][nlev] */
  double *qdp, *qtens_biharmonic, *qlim, *Q;
  /* CAM forcing (Elements.hpp m_fm, m_ft; Tracers.hpp fq): fm [ie][2][16][nlev], ft [ie][16][nlev],
     fq [ie][qsize_d][16][nlev] */
  double *fm, *ft, *fq;
  /* euler step data */
  double rhs_viss;
  /* Connectivity */
  ConnectionInfo* conn;
  /* retained diagnostics pointers */
  double* diag[8];
  int64_t launches;
} O = {.nlev = 72, .qsize_d = 40};

static void runtime_abort(const char* msg, int code) {
  /* ErrorDefs.cpp:23-27 (MPI_Abort replaced by exit: there is no MPI in this image) */
  fprintf(stderr, "%s\nExiting...\n", msg);
  finalize_hommexx_session();
  exit(code);
}

#define F3(ie) ((size_t)(ie) * NPSQ * O.nlev)
#define NLF ((size_t)NPSQ * O.nlev)

static double* zalloc(size_t n) {
  double* p = (double*)calloc(n ? n : 1, sizeof(double));
  if (!p) runtime_abort("oracle: out of memory", 1);
  return p;
}

/* ------------------------------------------------------------------------------------------ */
/* Section B of the ABI                                                                        */
/* ------------------------------------------------------------------------------------------ */
void oracle_set_dims(int nlev, int qsize_d) {
  O.nlev = nlev;
  O.qsize_d = qsize_d;
}
int hommexx_b200_nlev(void) { return O.nlev; }
int hommexx_b200_qsize_d(void) { return O.qsize_d; }
const char* hommexx_b200_backend(void) { return "cpu-oracle"; }
void hommexx_b200_set_comm(int rank, int size, int device, const void* id) {
  (void)rank; (void)device; (void)id;
  if (size != 1) runtime_abort("oracle: single process only", 12);
}
int64_t hommexx_b200_launch_count(void) { return 0; }
void hommexx_b200_sync(void) {}
void hommexx_b200_event_record(int slot) { (void)slot; }
double hommexx_b200_event_elapsed_ms(int a, int b) { (void)a; (void)b; return 0.0; }
void hommexx_b200_profile(unsigned long long mask) { (void)mask; }
int hommexx_b200_kernel_id(const char* name) { (void)name; return -1; }
const char* hommexx_b200_kernel_name(int id) { (void)id; return NULL; }
double hommexx_b200_profile_read(int id, int64_t* launches) { (void)id; if (launches) *launches = 0; return 0.0; }

/* ------------------------------------------------------------------------------------------ */
/* Session / init                                                                              */
/* ------------------------------------------------------------------------------------------ */
void reset_cxx_comm(const int* f_comm) { (void)f_comm; }
void initialize_hommexx_session(void) { O.session = true; }

static void free_all(void) {
  double** ptrs[] = {&O.hyai, &O.hybi, &O.dai, &O.dbi, &O.dp0, &O.d, &O.dinv, &O.metinv, &O.tensorvisc,
                     &O.vec_sph2cart, &O.fcor, &O.mp, &O.spheremp, &O.rspheremp, &O.metdet, &O.phis, &O.v,
                     &O.t, &O.dp3d, &O.ps_v, &O.phi, &O.omega_p, &O.eta_dot_dpdn, &O.derived_vn0,
                     &O.derived_dp, &O.divdp, &O.divdp_proj, &O.dpdiss_ave, &O.dpdiss_biharmonic, &O.vtens,
                     &O.ttens, &O.dptens, &O.vstar, &O.dpdissk, &O.dp_star, &O.qdp, &O.qtens_biharmonic,
                     &O.qlim, &O.Q, &O.fm, &O.ft, &O.fq};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) {
    free(*ptrs[i]);
    *ptrs[i] = NULL;
  }
  free(O.conn);
  O.conn = NULL;
}

void finalize_hommexx_session(void) {
  free_all();
  O.session = false;
  O.params_set = false;
  O.nelemd = 0;
}

void init_connectivity(const int* num_local_elems) {
  /* Connectivity.cpp:40-75: all connections MISSING until added */
  O.nelemd = *num_local_elems;
  free(O.conn);
  O.conn = (ConnectionInfo*)calloc((size_t)O.nelemd * 8, sizeof(ConnectionInfo));
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int c = 0; c < 8; ++c) {
      ConnectionInfo* info = &O.conn[ie * 8 + c];
      info->kind = 2; info->sharing = 2; info->direction = 2;
      info->local.lid = ie; info->local.pos = c; info->local.gid = -1;
      info->remote.lid = info->remote.gid = info->remote.pos = -1;
      info->remote_pid = -1;
    }
}

/* ConnectivityHelpers.hpp:133-142 */
static const int CONNECTION_DIRECTION[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, 1, 1, 0}, {1, 0, 0, 1}};

void add_connection(const int* l1, const int* g1, const int* p1, const int* r1, const int* l2, const int* g2,
                    const int* p2, const int* r2) {
  if (*l1 <= 0 || *g1 <= 0 || *p1 <= 0 || *r1 <= 0 || *l2 <= 0 || *g2 <= 0 || *p2 <= 0 || *r2 <= 0) {
    fprintf(stderr, "ERROR! We were assuming F90 indices started at 1, but it appears there is an exception.\n");
    abort();
  }
  /* mpi_cxx_f90_interface.cpp:47-49 */
  const int fep = *p1 <= 4 ? ((*p1 - 1) + 2) % 4 : *p1 - 1;
  const int sep = *p2 <= 4 ? ((*p2 - 1) + 2) % 4 : *p2 - 1;
  if (*r1 - 1 != 0) return; /* Connectivity.cpp:89 — only store if first element is ours (rank 0) */
  ConnectionInfo* info = &O.conn[(*l1 - 1) * 8 + fep];
  info->local.lid = *l1 - 1; info->local.gid = *g1 - 1; info->local.pos = fep;
  info->remote.lid = *l2 - 1; info->remote.gid = *g2 - 1; info->remote.pos = sep;
  info->kind = fep < 4 ? 0 : 1;
  info->direction = fep < 4 ? CONNECTION_DIRECTION[fep][sep] : 0;
  if (*r2 - 1 != 0) runtime_abort("oracle: remote connections are not supported (single process)", 12);
  info->sharing = 0;
}

void finalize_connectivity(void) {}

void init_derivative_c(const double* const* dvv) {
  /* Derivative.cpp:20-32 */
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) O.dvv[i][j] = (*dvv)[i * NP + j];
}

static void option_error(const char* loc, const char* opt, double value) {
  char msg[512];
  snprintf(msg, sizeof msg, "Error in %s: unsupported value '%g' for input parameter '%s'.", loc, value, opt);
  runtime_abort(msg, 12);
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
  /* cxx_f90_interface.cpp:43-52 */
  if (*remap_alg != 1 && *remap_alg != 2) option_error(loc, "vert_remap_q_alg", *remap_alg);
  if (*prescribed_wind) option_error(loc, "prescribed_wind", 1);
  if (*hypervis_order != 2) option_error(loc, "hypervis_order", *hypervis_order);
  if (*use_semi_lagrangian_transport) option_error(loc, "use_semi_lagrangian_transport", 1);
  if (*time_step_type != 5) option_error(loc, "time_step_type", *time_step_type);
  if (*limiter_option != 8 && *limiter_option != 9) option_error(loc, "limiter_option", *limiter_option);
  if (*ftype != -1 && *ftype != 0 && *ftype != 2) option_error(loc, "ftype", *ftype);
  if (!(*nu_p > 0.0)) runtime_abort("Error in init_simulation_params_c: nu_p must be > 0", 13);
  if (!(*nu > 0.0)) runtime_abort("Error in init_simulation_params_c: nu must be > 0", 13);
  if (!(*nu_div > 0.0)) runtime_abort("Error in init_simulation_params_c: nu_div must be > 0", 13);
  O.remap_alg = *remap_alg; O.limiter_option = *limiter_option; O.rsplit = *rsplit; O.qsplit = *qsplit;
  O.time_step_type = *time_step_type; O.qsize = *qsize; O.state_frequency = *state_frequency;
  O.nu = *nu; O.nu_p = *nu_p; O.nu_q = *nu_q; O.nu_s = *nu_s; O.nu_div = *nu_div; O.nu_top = *nu_top;
  O.hypervis_order = *hypervis_order; O.hypervis_subcycle = *hypervis_subcycle;
  O.hypervis_scaling = *hypervis_scaling; O.ftype = *ftype;
  O.moist = *moisture; O.disable_diagnostics = *disable_diagnostics; O.use_cpstar = *use_cpstar;
  /* :88-100 */
  if (O.nu != O.nu_div) {
    const double ratio = O.nu_div / O.nu;
    if (O.hypervis_scaling != 0.0) { O.nu_ratio1 = ratio * ratio; O.nu_ratio2 = 1.0; }
    else { O.nu_ratio1 = ratio; O.nu_ratio2 = ratio; }
  } else { O.nu_ratio1 = 1.0; O.nu_ratio2 = 1.0; }
  O.consthv = (O.hypervis_scaling == 0.0);
  O.params_set = true;
}

void init_hvcoord_c(const double* ps0, const double* const* am, const double* const* ai, const double* const* bm,
                    const double* const* bi) {
  (void)am; (void)bm;
  /* HybridVCoord.cpp:15-53,112-156 */
  const int nl = O.nlev;
  O.ps0 = *ps0;
  O.hyai = zalloc(nl + 1); O.hybi = zalloc(nl + 1); O.dai = zalloc(nl); O.dbi = zalloc(nl); O.dp0 = zalloc(nl);
  memcpy(O.hyai, *ai, (nl + 1) * sizeof(double));
  memcpy(O.hybi, *bi, (nl + 1) * sizeof(double));
  O.hyai0 = O.hyai[0];
  for (int k = 0; k < nl; ++k) {
    O.dai[k] = O.hyai[k + 1] - O.hyai[k];
    O.dbi[k] = O.hybi[k + 1] - O.hybi[k];
    O.dp0[k] = O.dai[k] * O.ps0 + O.dbi[k] * O.ps0;
  }
}

void init_elements_2d_c(const int* num_elems, const double* const* D, const double* const* Dinv,
                        const double* const* fcor, const double* const* mp, const double* const* spheremp,
                        const double* const* rspheremp, const double* const* metdet, const double* const* metinv,
                        const double* const* phis, const double* const* tensorvisc,
                        const double* const* vec_sph2cart, const bool* consthv) {
  /* Elements.cpp:20-190: the F90 arrays are read linearly into [ie][a][b][igp][jgp] */
  const int n = *num_elems, nl = O.nlev;
  if (n != O.nelemd) runtime_abort("init_elements_2d_c: element count differs from init_connectivity", 13);
  const size_t s2 = (size_t)n * NPSQ, t2 = (size_t)n * 4 * NPSQ, f3 = (size_t)n * NPSQ * nl;
  O.d = zalloc(t2); O.dinv = zalloc(t2); O.metinv = zalloc(t2);
  O.fcor = zalloc(s2); O.mp = zalloc(s2); O.spheremp = zalloc(s2); O.rspheremp = zalloc(s2);
  O.metdet = zalloc(s2); O.phis = zalloc(s2);
  memcpy(O.d, *D, t2 * 8); memcpy(O.dinv, *Dinv, t2 * 8); memcpy(O.metinv, *metinv, t2 * 8);
  memcpy(O.fcor, *fcor, s2 * 8); memcpy(O.mp, *mp, s2 * 8); memcpy(O.spheremp, *spheremp, s2 * 8);
  memcpy(O.rspheremp, *rspheremp, s2 * 8); memcpy(O.metdet, *metdet, s2 * 8); memcpy(O.phis, *phis, s2 * 8);
  if (!*consthv) {
    O.tensorvisc = zalloc(t2); O.vec_sph2cart = zalloc((size_t)n * 6 * NPSQ);
    memcpy(O.tensorvisc, *tensorvisc, t2 * 8);
    memcpy(O.vec_sph2cart, *vec_sph2cart, (size_t)n * 6 * NPSQ * 8);
  }
  O.v = zalloc(f3 * NTL * 2); O.t = zalloc(f3 * NTL); O.dp3d = zalloc(f3 * NTL); O.ps_v = zalloc(s2 * NTL);
  O.phi = zalloc(f3); O.omega_p = zalloc(f3); O.eta_dot_dpdn = zalloc(f3); O.derived_vn0 = zalloc(f3 * 2);
  O.derived_dp = zalloc(f3); O.divdp = zalloc(f3); O.divdp_proj = zalloc(f3); O.dpdiss_ave = zalloc(f3);
  O.dpdiss_biharmonic = zalloc(f3);
  O.vtens = zalloc(f3 * 2); O.ttens = zalloc(f3); O.dptens = zalloc(f3);
  O.vstar = zalloc(f3 * 2); O.dpdissk = zalloc(f3); O.dp_star = zalloc(f3);
  O.qdp = zalloc(f3 * QNTL * O.qsize_d); O.qtens_biharmonic = zalloc(f3 * O.qsize_d);
  O.Q = zalloc(f3 * O.qsize_d); O.qlim = zalloc((size_t)n * O.qsize_d * 2 * nl);
  O.fm = zalloc(f3 * 2); O.ft = zalloc(f3); O.fq = zalloc(f3 * O.qsize_d);
}

/* SyncUtils.hpp: F90 [ie][tl][lev][(2)][igp][jgp]  <->  [ie][tl][(2)][igp][jgp][lev] */
void init_elements_states_c(const double* const* fv, const double* const* ft, const double* const* fdp,
                            const double* const* fq, const double* const* fps) {
  const int n = O.nelemd, nl = O.nlev;
  for (int ie = 0; ie < n; ++ie)
    for (int tl = 0; tl < NTL; ++tl)
      for (int k = 0; k < nl; ++k)
        for (int p = 0; p < NPSQ; ++p) {
          const size_t f = (((size_t)ie * NTL + tl) * nl + k);
          O.t[((size_t)ie * NTL + tl) * NLF + p * nl + k] = (*ft)[f * NPSQ + p];
          O.dp3d[((size_t)ie * NTL + tl) * NLF + p * nl + k] = (*fdp)[f * NPSQ + p];
          for (int c = 0; c < 2; ++c)
            O.v[(((size_t)ie * NTL + tl) * 2 + c) * NLF + p * nl + k] = (*fv)[(f * 2 + c) * NPSQ + p];
        }
  for (int ie = 0; ie < n; ++ie)
    for (int tq = 0; tq < QNTL; ++tq)
      for (int q = 0; q < O.qsize_d; ++q)
        for (int k = 0; k < nl; ++k)
          for (int p = 0; p < NPSQ; ++p)
            O.qdp[(((size_t)ie * QNTL + tq) * O.qsize_d + q) * NLF + p * nl + k] =
                (*fq)[((((size_t)ie * QNTL + tq) * O.qsize_d + q) * nl + k) * NPSQ + p];
  memcpy(O.ps_v, *fps, (size_t)n * NTL * NPSQ * 8);
}

void init_diagnostics_c(double* const* a0, double* const* a1, double* const* a2, double* const* a3,
                        double* const* a4, double* const* a5, double* const* a6, double* const* a7) {
  O.diag[0] = *a0; O.diag[1] = *a1; O.diag[2] = *a2; O.diag[3] = *a3;
  O.diag[4] = *a4; O.diag[5] = *a5; O.diag[6] = *a6; O.diag[7] = *a7;
}

void init_boundary_exchanges_c(void) {}

void init_time_level_c(const int* nm1, const int* n0, const int* np1, const int* nstep, const int* nstep0) {
  O.nm1 = *nm1 - 1; O.n0 = *n0 - 1; O.np1 = *np1 - 1; O.nstep = *nstep; O.nstep0 = *nstep0;
}

void cxx_push_results_to_f90(double* const* fv, double* const* ft, double* const* fdp, double* const* fq,
                             double* const* fQ, double* const* fps, double* const* fom) {
  const int n = O.nelemd, nl = O.nlev;
  for (int ie = 0; ie < n; ++ie)
    for (int tl = 0; tl < NTL; ++tl)
      for (int k = 0; k < nl; ++k)
        for (int p = 0; p < NPSQ; ++p) {
          const size_t f = (((size_t)ie * NTL + tl) * nl + k);
          (*ft)[f * NPSQ + p] = O.t[((size_t)ie * NTL + tl) * NLF + p * nl + k];
          (*fdp)[f * NPSQ + p] = O.dp3d[((size_t)ie * NTL + tl) * NLF + p * nl + k];
          for (int c = 0; c < 2; ++c)
            (*fv)[(f * 2 + c) * NPSQ + p] = O.v[(((size_t)ie * NTL + tl) * 2 + c) * NLF + p * nl + k];
        }
  for (int ie = 0; ie < n; ++ie)
    for (int q = 0; q < O.qsize_d; ++q)
      for (int k = 0; k < nl; ++k)
        for (int p = 0; p < NPSQ; ++p) {
          for (int tq = 0; tq < QNTL; ++tq)
            (*fq)[((((size_t)ie * QNTL + tq) * O.qsize_d + q) * nl + k) * NPSQ + p] =
                O.qdp[(((size_t)ie * QNTL + tq) * O.qsize_d + q) * NLF + p * nl + k];
          (*fQ)[(((size_t)ie * O.qsize_d + q) * nl + k) * NPSQ + p] =
              O.Q[((size_t)ie * O.qsize_d + q) * NLF + p * nl + k];
        }
  memcpy(*fps, O.ps_v, (size_t)n * NTL * NPSQ * 8);
  for (int ie = 0; ie < n; ++ie)
    for (int k = 0; k < nl; ++k)
      for (int p = 0; p < NPSQ; ++p) (*fom)[((size_t)ie * nl + k) * NPSQ + p] = O.omega_p[F3(ie) + p * nl + k];
}

/* cxx_f90_interface.cpp:180-205: FM [ie][nlev][2][16], FT [ie][nlev][16], FQ [ie][qsize_d][nlev][16] (F90) to the
   device layout; FQ only for ftype == 0 (FORCING_DEBUG); Tracers::push_qdp (Tracers.cpp:46-51) copies the
   device qdp back INTO the F90 array */
void f90_push_forcing_to_cxx(double* fm, double* ft, double* fq, double* qdp) {
  const int n = O.nelemd, nl = O.nlev;
  for (int ie = 0; ie < n; ++ie)
    for (int k = 0; k < nl; ++k)
      for (int p = 0; p < NPSQ; ++p) {
        O.ft[F3(ie) + p * nl + k] = ft[((size_t)ie * nl + k) * NPSQ + p];
        for (int c = 0; c < 2; ++c)
          O.fm[((size_t)ie * 2 + c) * NLF + p * nl + k] = fm[(((size_t)ie * nl + k) * 2 + c) * NPSQ + p];
      }
  if (O.ftype == 0)
    for (int ie = 0; ie < n; ++ie)
      for (int q = 0; q < O.qsize_d; ++q)
        for (int k = 0; k < nl; ++k)
          for (int p = 0; p < NPSQ; ++p)
            O.fq[((size_t)ie * O.qsize_d + q) * NLF + p * nl + k] = fq[(((size_t)ie * O.qsize_d + q) * nl + k) * NPSQ + p];
  for (int ie = 0; ie < n; ++ie)
    for (int tq = 0; tq < QNTL; ++tq)
      for (int q = 0; q < O.qsize_d; ++q)
        for (int k = 0; k < nl; ++k)
          for (int p = 0; p < NPSQ; ++p)
            qdp[((((size_t)ie * QNTL + tq) * O.qsize_d + q) * nl + k) * NPSQ + p] =
                O.qdp[(((size_t)ie * QNTL + tq) * O.qsize_d + q) * NLF + p * nl + k];
}
/* cxx_f90_interface.cpp:157-178 */
void cxx_push_forcing_to_f90(double* fm, double* ft, double* fq) {
  const int n = O.nelemd, nl = O.nlev;
  for (int ie = 0; ie < n; ++ie)
    for (int k = 0; k < nl; ++k)
      for (int p = 0; p < NPSQ; ++p) {
        ft[((size_t)ie * nl + k) * NPSQ + p] = O.ft[F3(ie) + p * nl + k];
        for (int c = 0; c < 2; ++c)
          fm[(((size_t)ie * nl + k) * 2 + c) * NPSQ + p] = O.fm[((size_t)ie * 2 + c) * NLF + p * nl + k];
      }
  if (O.ftype == 0)
    for (int ie = 0; ie < n; ++ie)
      for (int q = 0; q < O.qsize_d; ++q)
        for (int k = 0; k < nl; ++k)
          for (int p = 0; p < NPSQ; ++p)
            fq[(((size_t)ie * O.qsize_d + q) * nl + k) * NPSQ + p] = O.fq[((size_t)ie * O.qsize_d + q) * NLF + p * nl + k];
}

/* ------------------------------------------------------------------------------------------ */
/* Sphere operators (SphereOperators.hpp), element-local, fields are [16][nlev]                */
/* ------------------------------------------------------------------------------------------ */
#define DV(a, b) O.dvv[a][b]
#define T2(arr, ie, a, b, p) (arr)[(((size_t)(ie) * 2 + (a)) * 2 + (b)) * NPSQ + (p)]
#define S2(arr, ie, p) (arr)[(size_t)(ie) * NPSQ + (p)]
#define IX(p, k) ((size_t)(p) * nlev + (k))

/* SphereOperators.hpp:293-319 */
static void gradient_sphere(int ie, const double* s, double* g0, double* g1, int nl) {
  const int nlev = O.nlev;
  for (int p = 0; p < NPSQ; ++p) {
    const int igp = p / NP, jgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double v0 = 0, v1 = 0;
      for (int kgp = 0; kgp < NP; ++kgp) {
        v0 += DV(jgp, kgp) * s[IX(igp * NP + kgp, k)];
        v1 += DV(igp, kgp) * s[IX(kgp * NP + jgp, k)];
      }
      v0 *= rrearth;
      v1 *= rrearth;
      g0[IX(p, k)] = T2(O.dinv, ie, 0, 0, p) * v0 + T2(O.dinv, ie, 0, 1, p) * v1;
      g1[IX(p, k)] = T2(O.dinv, ie, 1, 0, p) * v0 + T2(O.dinv, ie, 1, 1, p) * v1;
    }
  }
}

/* :323-348 */
static void gradient_sphere_update(int ie, const double* s, double* g0, double* g1, int nl) {
  const int nlev = O.nlev;
  for (int p = 0; p < NPSQ; ++p) {
    const int igp = p / NP, jgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double dsdx = 0, dsdy = 0;
      for (int kgp = 0; kgp < NP; ++kgp) {
        dsdx += DV(jgp, kgp) * s[IX(igp * NP + kgp, k)];
        dsdy += DV(igp, kgp) * s[IX(kgp * NP + jgp, k)];
      }
      dsdx *= rrearth;
      dsdy *= rrearth;
      g0[IX(p, k)] += T2(O.dinv, ie, 0, 0, p) * dsdx + T2(O.dinv, ie, 0, 1, p) * dsdy;
      g1[IX(p, k)] += T2(O.dinv, ie, 1, 0, p) * dsdx + T2(O.dinv, ie, 1, 1, p) * dsdy;
    }
  }
}

/* :352-392 */
static void divergence_sphere(int ie, const double* v0, const double* v1, double* div, int nl) {
  const int nlev = O.nlev;
  double* gv = (double*)malloc(2 * NLF * sizeof(double));
  double *gv0 = gv, *gv1 = gv + NLF;
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double a = v0[IX(p, k)], b = v1[IX(p, k)];
      gv0[IX(p, k)] = (T2(O.dinv, ie, 0, 0, p) * a + T2(O.dinv, ie, 1, 0, p) * b) * S2(O.metdet, ie, p);
      gv1[IX(p, k)] = (T2(O.dinv, ie, 0, 1, p) * a + T2(O.dinv, ie, 1, 1, p) * b) * S2(O.metdet, ie, p);
    }
  for (int p = 0; p < NPSQ; ++p) {
    const int igp = p / NP, jgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double dudx = 0, dvdy = 0;
      for (int kgp = 0; kgp < NP; ++kgp) {
        dudx += DV(jgp, kgp) * gv0[IX(igp * NP + kgp, k)];
        dvdy += DV(igp, kgp) * gv1[IX(kgp * NP + jgp, k)];
      }
      div[IX(p, k)] = (dudx + dvdy) * (1.0 / S2(O.metdet, ie, p) * rrearth);
    }
  }
  free(gv);
}

/* :398-444 */
static void divergence_sphere_update(int ie, double alpha, bool add_hyperviscosity, const double* vstar0,
                                     const double* vstar1, const double* qdp, double* qtens) {
  const int nlev = O.nlev;
  double* gv = (double*)malloc(2 * NLF * sizeof(double));
  double *gv0 = gv, *gv1 = gv + NLF;
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nlev; ++k) {
      const double q = qdp[IX(p, k)];
      const double a = vstar0[IX(p, k)] * q, b = vstar1[IX(p, k)] * q;
      gv0[IX(p, k)] = (T2(O.dinv, ie, 0, 0, p) * a + T2(O.dinv, ie, 1, 0, p) * b) * S2(O.metdet, ie, p);
      gv1[IX(p, k)] = (T2(O.dinv, ie, 0, 1, p) * a + T2(O.dinv, ie, 1, 1, p) * b) * S2(O.metdet, ie, p);
    }
  for (int p = 0; p < NPSQ; ++p) {
    const int igp = p / NP, jgp = p % NP;
    for (int k = 0; k < nlev; ++k) {
      double dudx = 0, dvdy = 0;
      for (int kgp = 0; kgp < NP; ++kgp) {
        dudx += DV(jgp, kgp) * gv0[IX(igp * NP + kgp, k)];
        dvdy += DV(igp, kgp) * gv1[IX(kgp * NP + jgp, k)];
      }
      const double qt0 = add_hyperviscosity ? qtens[IX(p, k)] : 0;
      qtens[IX(p, k)] =
          (qdp[IX(p, k)] + alpha * ((dudx + dvdy) * (1.0 / S2(O.metdet, ie, p) * rrearth)) + qt0);
    }
  }
  free(gv);
}

/* :494-533 */
static void vorticity_sphere(int ie, const double* u, const double* v, double* vort, int nl) {
  const int nlev = O.nlev;
  double* sb = (double*)malloc(2 * NLF * sizeof(double));
  double *c0 = sb, *c1 = sb + NLF;
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double a = u[IX(p, k)], b = v[IX(p, k)];
      c0[IX(p, k)] = T2(O.d, ie, 0, 0, p) * a + T2(O.d, ie, 0, 1, p) * b;
      c1[IX(p, k)] = T2(O.d, ie, 1, 0, p) * a + T2(O.d, ie, 1, 1, p) * b;
    }
  for (int p = 0; p < NPSQ; ++p) {
    const int igp = p / NP, jgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double dudy = 0, dvdx = 0;
      for (int kgp = 0; kgp < NP; ++kgp) {
        dvdx += DV(jgp, kgp) * c1[IX(igp * NP + kgp, k)];
        dudy += DV(igp, kgp) * c0[IX(kgp * NP + jgp, k)];
      }
      vort[IX(p, k)] = (dvdx - dudy) * (1.0 / S2(O.metdet, ie, p) * rrearth);
    }
  }
  free(sb);
}

/* :538-583 — v is overwritten (the reference aliases its temporary on the input) */
static void divergence_sphere_wk(int ie, double* v0, double* v1, double* div, int nl) {
  const int nlev = O.nlev;
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double a = v0[IX(p, k)], b = v1[IX(p, k)];
      v0[IX(p, k)] = T2(O.dinv, ie, 0, 0, p) * a + T2(O.dinv, ie, 1, 0, p) * b;
      v1[IX(p, k)] = T2(O.dinv, ie, 0, 1, p) * a + T2(O.dinv, ie, 1, 1, p) * b;
    }
  for (int p = 0; p < NPSQ; ++p) {
    const int ngp = p / NP, mgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double dd = 0;
      for (int jgp = 0; jgp < NP; ++jgp) {
        dd -= (S2(O.spheremp, ie, ngp * NP + jgp) * v0[IX(ngp * NP + jgp, k)] * DV(jgp, mgp) +
               S2(O.spheremp, ie, jgp * NP + mgp) * v1[IX(jgp * NP + mgp, k)] * DV(jgp, ngp)) *
              rrearth;
      }
      div[IX(p, k)] = dd;
    }
  }
}

/* :588-597 — field and laplace may alias */
static void laplace_simple(int ie, const double* field, double* laplace, int nl) {
  double* g = (double*)malloc(2 * NLF * sizeof(double));
  gradient_sphere(ie, field, g, g + NLF, nl);
  divergence_sphere_wk(ie, g, g + NLF, laplace, nl);
  free(g);
}

/* :604-635 */
static void laplace_tensor(int ie, const double* field, double* laplace, int nl) {
  const int nlev = O.nlev;
  double* g = (double*)malloc(4 * NLF * sizeof(double));
  double *g0 = g, *g1 = g + NLF, *s0 = g + 2 * NLF, *s1 = g + 3 * NLF;
  gradient_sphere(ie, field, g0, g1, nl);
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double a = g0[IX(p, k)], b = g1[IX(p, k)];
      s0[IX(p, k)] = T2(O.tensorvisc, ie, 0, 0, p) * a + T2(O.tensorvisc, ie, 1, 0, p) * b;
      s1[IX(p, k)] = T2(O.tensorvisc, ie, 0, 1, p) * a + T2(O.tensorvisc, ie, 1, 1, p) * b;
    }
  divergence_sphere_wk(ie, s0, s1, laplace, nl);
  free(g);
}

/* :683-710 */
static void curl_sphere_wk_testcov_update(int ie, double alpha, double beta, const double* s, double* c0,
                                          double* c1, int nl) {
  const int nlev = O.nlev;
  for (int p = 0; p < NPSQ; ++p) {
    const int ngp = p / NP, mgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double sb0 = 0, sb1 = 0;
      for (int jgp = 0; jgp < NP; ++jgp) {
        sb0 -= S2(O.mp, ie, jgp * NP + mgp) * s[IX(jgp * NP + mgp, k)] * DV(jgp, ngp);
        sb1 += S2(O.mp, ie, ngp * NP + jgp) * s[IX(ngp * NP + jgp, k)] * DV(jgp, mgp);
      }
      c0[IX(p, k)] = beta * c0[IX(p, k)] +
                     alpha * (T2(O.d, ie, 0, 0, p) * sb0 + T2(O.d, ie, 1, 0, p) * sb1) * rrearth;
      c1[IX(p, k)] = beta * c1[IX(p, k)] +
                     alpha * (T2(O.d, ie, 0, 1, p) * sb0 + T2(O.d, ie, 1, 1, p) * sb1) * rrearth;
    }
  }
}

/* :714-748 */
static void grad_sphere_wk_testcov(int ie, const double* s, double* g0, double* g1, int nl) {
  const int nlev = O.nlev;
  for (int p = 0; p < NPSQ; ++p) {
    const int ngp = p / NP, mgp = p % NP;
    for (int k = 0; k < nl; ++k) {
      double b0 = 0, b1 = 0;
      for (int jgp = 0; jgp < NP; ++jgp) {
        const double mpnj = S2(O.mp, ie, ngp * NP + jgp), mpjm = S2(O.mp, ie, jgp * NP + mgp);
        const double md = S2(O.metdet, ie, p);
        const double snj = s[IX(ngp * NP + jgp, k)], sjm = s[IX(jgp * NP + mgp, k)];
        const double djm = DV(jgp, mgp), djn = DV(jgp, ngp);
        b0 -= (mpnj * T2(O.metinv, ie, 0, 0, p) * md * snj * djm + mpjm * T2(O.metinv, ie, 0, 1, p) * md * sjm * djn);
        b1 -= (mpnj * T2(O.metinv, ie, 1, 0, p) * md * snj * djm + mpjm * T2(O.metinv, ie, 1, 1, p) * md * sjm * djn);
      }
      g0[IX(p, k)] = (T2(O.d, ie, 0, 0, p) * b0 + T2(O.d, ie, 1, 0, p) * b1) * rrearth;
      g1[IX(p, k)] = (T2(O.d, ie, 0, 1, p) * b0 + T2(O.d, ie, 1, 1, p) * b1) * rrearth;
    }
  }
}

/* :818-862 — vector and laplace may alias */
static void vlaplace_sphere_wk_contra(int ie, double nu_ratio, const double* v0, const double* v1, double* l0,
                                      double* l1, int nl) {
  const int nlev = O.nlev;
  double* buf = (double*)malloc(3 * NLF * sizeof(double));
  double *sc = buf, *gc0 = buf + NLF, *gc1 = buf + 2 * NLF;
  divergence_sphere(ie, v0, v1, sc, nl);
  if (nu_ratio > 0 && nu_ratio != 1.0)
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nl; ++k) sc[IX(p, k)] *= nu_ratio;
  grad_sphere_wk_testcov(ie, sc, gc0, gc1, nl);
  vorticity_sphere(ie, v0, v1, sc, nl);
  curl_sphere_wk_testcov_update(ie, -1.0, 1.0, sc, gc0, gc1, nl);
  const double re2 = rrearth * rrearth;
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double f = 2.0 * S2(O.spheremp, ie, p);
      const double a = v0[IX(p, k)], b = v1[IX(p, k)];
      l0[IX(p, k)] = f * a * re2 + gc0[IX(p, k)];
      l1[IX(p, k)] = f * b * re2 + gc1[IX(p, k)];
    }
  free(buf);
}

/* :752-814 */
static void vlaplace_sphere_wk_cartesian(int ie, const double* v0, const double* v1, double* l0, double* l1,
                                         int nl) {
  const int nlev = O.nlev;
  double* lap = (double*)malloc(3 * NLF * sizeof(double));
  const double* vs = O.vec_sph2cart + (size_t)ie * 6 * NPSQ; /* [2][3][16] */
  for (int c = 0; c < 3; ++c)
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nl; ++k)
        lap[c * NLF + IX(p, k)] = vs[(0 * 3 + c) * NPSQ + p] * v0[IX(p, k)] + vs[(1 * 3 + c) * NPSQ + p] * v1[IX(p, k)];
  for (int c = 0; c < 3; ++c) laplace_tensor(ie, lap + c * NLF, lap + c * NLF, nl);
  for (int p = 0; p < NPSQ; ++p)
    for (int k = 0; k < nl; ++k) {
      const double a = v0[IX(p, k)], b = v1[IX(p, k)];
      l0[IX(p, k)] = vs[(0 * 3 + 0) * NPSQ + p] * lap[0 * NLF + IX(p, k)] + vs[(0 * 3 + 1) * NPSQ + p] * lap[1 * NLF + IX(p, k)] +
                     vs[(0 * 3 + 2) * NPSQ + p] * lap[2 * NLF + IX(p, k)] +
                     2.0 * S2(O.spheremp, ie, p) * a * (rrearth) * (rrearth);
      l1[IX(p, k)] = vs[(1 * 3 + 0) * NPSQ + p] * lap[0 * NLF + IX(p, k)] + vs[(1 * 3 + 1) * NPSQ + p] * lap[1 * NLF + IX(p, k)] +
                     vs[(1 * 3 + 2) * NPSQ + p] * lap[2 * NLF + IX(p, k)] +
                     2.0 * S2(O.spheremp, ie, p) * b * (rrearth) * (rrearth);
    }
  free(lap);
}

/* ------------------------------------------------------------------------------------------ */
/* Boundary exchange (mpi/BoundaryExchange.cpp): pack -> buffers -> ordered unpack             */
/* ------------------------------------------------------------------------------------------ */
static const int EDGE_PTS_FWD[4][4] = {{0, 1, 2, 3}, {12, 13, 14, 15}, {0, 4, 8, 12}, {3, 7, 11, 15}};
static const int CORNER_PTS[4] = {0, 3, 12, 15};

typedef struct {
  double* base;    /* first element's field */
  size_t estride;  /* doubles between consecutive elements */
} FieldRef;

/* exchange() for nf 3-D fields; BoundaryExchange.cpp:285-420 (pack), :505-537 (unpack, CPU branch
 * :539-596 with the fixed order S,N,W,E per k then corners), optional rspheremp scaling. */
static void exchange(const FieldRef* fields, int nf, bool rspheremp) {
  const int n = O.nelemd, nlev = O.nlev;
  /* recv buffer: [ie][field][conn][4 pts][nlev], kept between calls as the reference's BuffersManager keeps
     its buffers (mpi/BuffersManager.cpp:95-135); slots of MISSING connections are never read (the reference
     points them at a zero-filled "blackhole" instead, BoundaryExchange.cpp:1003-1020: adding 0 is a no-op) */
  const size_t per_f = (size_t)8 * NP * nlev;
  static double* recv = NULL;
  static size_t recv_cap = 0;
  if ((size_t)n * nf * per_f > recv_cap) {
    free(recv);
    recv_cap = (size_t)n * nf * per_f;
    recv = (double*)malloc(recv_cap * sizeof(double));
    if (!recv) runtime_abort("oracle: out of memory", 1);
  }
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie)
    for (int f = 0; f < nf; ++f)
      for (int c = 0; c < 8; ++c) {
        const ConnectionInfo* info = &O.conn[ie * 8 + c];
        if (info->kind == 2) continue;
        const double* src = fields[f].base + (size_t)ie * fields[f].estride;
        /* LOCAL connection: write straight into the remote element's slot (:327) */
        double* dst = recv + ((size_t)info->remote.lid * nf + f) * per_f + (size_t)info->remote.pos * NP * nlev;
        const int npts = info->kind == 0 ? NP : 1;
        for (int k = 0; k < npts; ++k) {
          int pt;
          if (info->kind == 0) pt = EDGE_PTS_FWD[c][info->direction ? 3 - k : k];
          else pt = CORNER_PTS[c - 4];
          memcpy(dst + (size_t)k * nlev, src + (size_t)pt * nlev, nlev * sizeof(double));
        }
      }
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie)
    for (int f = 0; f < nf; ++f) {
      double* fld = fields[f].base + (size_t)ie * fields[f].estride;
      const double* rb = recv + ((size_t)ie * nf + f) * per_f;
      for (int k = 0; k < NP; ++k)
        for (int e = 0; e < 4; ++e) {
          if (O.conn[ie * 8 + e].kind == 2) continue;
          double* fp = fld + (size_t)EDGE_PTS_FWD[e][k] * nlev;
          const double* rp = rb + ((size_t)e * NP + k) * nlev;
          for (int l = 0; l < nlev; ++l) fp[l] += rp[l];
        }
      for (int c = 0; c < 4; ++c) {
        if (O.conn[ie * 8 + 4 + c].kind == 2) continue;
        double* fp = fld + (size_t)CORNER_PTS[c] * nlev;
        const double* rp = rb + ((size_t)(4 + c) * NP) * nlev;
        for (int l = 0; l < nlev; ++l) fp[l] += rp[l];
      }
      if (rspheremp)
        for (int p = 0; p < NPSQ; ++p) {
          const double r = S2(O.rspheremp, ie, p);
          for (int l = 0; l < nlev; ++l) fld[IX(p, l)] *= r;
        }
    }
}

/* exchange_min_max on qlim (BoundaryExchange.cpp:620-849) */
static void exchange_min_max(void) {
  const int n = O.nelemd, nlev = O.nlev, nq = O.qsize;
  const size_t sz = (size_t)n * O.qsize_d * 2 * nlev;
  double* snap = (double*)malloc(sz * sizeof(double));
  memcpy(snap, O.qlim, sz * sizeof(double));
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie)
    for (int q = 0; q < nq; ++q) {
      double* mine = O.qlim + ((size_t)ie * O.qsize_d + q) * 2 * nlev;
      for (int c = 0; c < 8; ++c) {
        const ConnectionInfo* info = &O.conn[ie * 8 + c];
        if (info->kind == 2) continue;
        const double* theirs = snap + ((size_t)info->remote.lid * O.qsize_d + q) * 2 * nlev;
        for (int l = 0; l < nlev; ++l) {
          mine[l] = fmin(mine[l], theirs[l]);
          mine[nlev + l] = fmax(mine[nlev + l], theirs[nlev + l]);
        }
      }
    }
  free(snap);
}

/* ------------------------------------------------------------------------------------------ */
/* CAAR (CaarFunctorImpl.hpp)                                                                  */
/* ------------------------------------------------------------------------------------------ */
#define VFLD(ie, tl, c) (O.v + (((size_t)(ie) * NTL + (tl)) * 2 + (c)) * NLF)
#define TFLD(ie, tl) (O.t + ((size_t)(ie) * NTL + (tl)) * NLF)
#define DPFLD(ie, tl) (O.dp3d + ((size_t)(ie) * NTL + (tl)) * NLF)
#define QDPFLD(ie, tq, q) (O.qdp + (((size_t)(ie) * QNTL + (tq)) * O.qsize_d + (q)) * NLF)

static void caar_element(int ie, int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp) {
  const int nlev = O.nlev;
  double* buf = (double*)malloc(18 * NLF * sizeof(double));
  double *t_virt = buf, *vdp0 = buf + NLF, *vdp1 = buf + 2 * NLF, *div_vdp = buf + 3 * NLF, *pressure = buf + 4 * NLF,
         *pgrad0 = buf + 5 * NLF, *pgrad1 = buf + 6 * NLF, *omega_p = buf + 7 * NLF, *tgrad0 = buf + 8 * NLF,
         *tgrad1 = buf + 9 * NLF, *egrad0 = buf + 10 * NLF, *egrad1 = buf + 11 * NLF, *ephi = buf + 12 * NLF,
         *vort = buf + 13 * NLF;
  const double *v0 = VFLD(ie, n0, 0), *v1 = VFLD(ie, n0, 1), *t0 = TFLD(ie, n0), *dp0 = DPFLD(ie, n0);
  double* phi = O.phi + F3(ie);

  /* compute_temperature_div_vdp :402-409 */
  if (n0_qdp < 0) {
    for (size_t i = 0; i < NLF; ++i) t_virt[i] = t0[i]; /* :333-344 */
  } else {
    const double* q = QDPFLD(ie, n0_qdp, 0); /* :348-363 */
    for (size_t i = 0; i < NLF; ++i) {
      double Qt = q[i] / dp0[i];
      Qt *= (Rwater_vapor / Rgas - 1.0);
      Qt += 1.0;
      t_virt[i] = t0[i] * Qt;
    }
  }
  /* compute_div_vdp :370-396 */
  double *vn00 = O.derived_vn0 + ((size_t)ie * 2 + 0) * NLF, *vn01 = O.derived_vn0 + ((size_t)ie * 2 + 1) * NLF;
  for (size_t i = 0; i < NLF; ++i) {
    vdp0[i] = v0[i] * dp0[i];
    vdp1[i] = v1[i] * dp0[i];
    vn00[i] += eta_ave_w * vdp0[i];
    vn01[i] += eta_ave_w * vdp1[i];
  }
  divergence_sphere(ie, vdp0, vdp1, div_vdp, nlev);

  /* compute_pressure, non-CUDA :621-650 */
  for (int p = 0; p < NPSQ; ++p) {
    double dp_prev = 0;
    double p_prev = O.hyai0 * O.ps0;
    for (int k = 0; k < nlev; ++k) {
      const double pk = p_prev + 0.5 * (dp_prev + dp0[IX(p, k)]);
      pressure[IX(p, k)] = pk;
      p_prev = pk;
      dp_prev = dp0[IX(p, k)];
    }
  }
  /* preq_hydrostatic, non-CUDA :689-729 */
  for (int p = 0; p < NPSQ; ++p) {
    double integration = 0;
    const double phis = S2(O.phis, ie, p);
    for (int k = nlev - 1; k >= 0; --k) {
      const double a = Rgas * t_virt[IX(p, k)] * (dp0[IX(p, k)] * 0.5 / pressure[IX(p, k)]);
      phi[IX(p, k)] = phis + 2.0 * integration + a;
      integration = integration + a;
    }
  }
  /* preq_omega_ps, non-CUDA :854-889 */
  gradient_sphere(ie, pressure, pgrad0, pgrad1, nlev);
  for (int p = 0; p < NPSQ; ++p) {
    double integration = 0;
    for (int k = 0; k < nlev; ++k) {
      const double vgrad_p = v0[IX(p, k)] * pgrad0[IX(p, k)] + v1[IX(p, k)] * pgrad1[IX(p, k)];
      const double dv = div_vdp[IX(p, k)];
      omega_p[IX(p, k)] = (vgrad_p - (integration + 0.5 * dv)) / pressure[IX(p, k)];
      integration = integration + dv;
    }
  }

  /* compute_phase_3 :136-148; rsplit == 0: Eulerian vertical advection */
  double *eta_buf = buf + 14 * NLF, *t_vadv = buf + 15 * NLF, *v_vadv0 = buf + 16 * NLF, *v_vadv1 = buf + 17 * NLF;
  const bool vadv = (O.rsplit == 0);
  if (vadv) {
    for (int p = 0; p < NPSQ; ++p) {
      /* assign_zero_to_sdot_sum + compute_eta_dot_dpdn_vertadv_euler :255-300 */
      double sdot = 0;
      for (int k = 0; k < nlev - 1; ++k) {
        sdot += div_vdp[IX(p, k)];
        eta_buf[IX(p, k + 1)] = sdot;
      }
      sdot += div_vdp[IX(p, nlev - 1)];
      for (int k = 1; k < nlev; ++k) eta_buf[IX(p, k)] = O.hybi[k] * sdot - eta_buf[IX(p, k)];
      eta_buf[IX(p, 0)] = 0.0;
      /* preq_vertadv :495-597 */
      {
        double facp = (0.5 * 1 / dp0[IX(p, 0)]) * eta_buf[IX(p, 1)], facm;
        t_vadv[IX(p, 0)] = facp * (t0[IX(p, 1)] - t0[IX(p, 0)]);
        v_vadv0[IX(p, 0)] = facp * (v0[IX(p, 1)] - v0[IX(p, 0)]);
        v_vadv1[IX(p, 0)] = facp * (v1[IX(p, 1)] - v1[IX(p, 0)]);
        for (int k = 1; k < nlev - 1; ++k) {
          facp = 0.5 * (1 / dp0[IX(p, k)]) * eta_buf[IX(p, k + 1)];
          facm = 0.5 * (1 / dp0[IX(p, k)]) * eta_buf[IX(p, k)];
          t_vadv[IX(p, k)] = facp * (t0[IX(p, k + 1)] - t0[IX(p, k)]) + facm * (t0[IX(p, k)] - t0[IX(p, k - 1)]);
          v_vadv0[IX(p, k)] = facp * (v0[IX(p, k + 1)] - v0[IX(p, k)]) + facm * (v0[IX(p, k)] - v0[IX(p, k - 1)]);
          v_vadv1[IX(p, k)] = facp * (v1[IX(p, k + 1)] - v1[IX(p, k)]) + facm * (v1[IX(p, k)] - v1[IX(p, k - 1)]);
        }
        const int k = nlev - 1;
        facm = (0.5 * (1 / dp0[IX(p, k)])) * eta_buf[IX(p, k)];
        t_vadv[IX(p, k)] = facm * (t0[IX(p, k)] - t0[IX(p, k - 1)]);
        v_vadv0[IX(p, k)] = facm * (v0[IX(p, k)] - v0[IX(p, k - 1)]);
        v_vadv1[IX(p, k)] = facm * (v1[IX(p, k)] - v1[IX(p, k - 1)]);
      }
    }
    /* accumulate_eta_dot_dpdn :150-163 */
    double* eta = O.eta_dot_dpdn + F3(ie);
    for (size_t i = 0; i < NLF; ++i) eta[i] += eta_ave_w * eta_buf[i];
  }
  /* compute_omega_p :412-423 */
  double* om = O.omega_p + F3(ie);
  for (size_t i = 0; i < NLF; ++i) om[i] += eta_ave_w * omega_p[i];

  /* compute_temperature_np1 :430-463 */
  gradient_sphere(ie, t0, tgrad0, tgrad1, nlev);
  {
    const double* tm1 = TFLD(ie, nm1);
    double* tp1 = TFLD(ie, np1);
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k) {
        const size_t i = IX(p, k);
        const double vgrad_t = v0[i] * tgrad0[i] + v1[i] * tgrad1[i];
        const double ttens = (vadv ? -t_vadv[i] : 0) - vgrad_t + kappa * t_virt[i] * omega_p[i];
        double temp_np1 = ttens * dt + tm1[i];
        temp_np1 *= S2(O.spheremp, ie, p);
        tp1[i] = temp_np1;
      }
  }
  /* compute_velocity_np1 :184-232 with compute_energy_grad :98-132 */
  for (size_t i = 0; i < NLF; ++i) {
    egrad0[i] = Rgas * (t_virt[i] / pressure[i]) * pgrad0[i];
    egrad1[i] = Rgas * (t_virt[i] / pressure[i]) * pgrad1[i];
    const double k_energy = 0.5 * (v0[i] * v0[i] + v1[i] * v1[i]);
    ephi[i] = k_energy + phi[i];
  }
  gradient_sphere_update(ie, ephi, egrad0, egrad1, nlev);
  vorticity_sphere(ie, v0, v1, vort, nlev);
  {
    const double *vm0 = VFLD(ie, nm1, 0), *vm1 = VFLD(ie, nm1, 1);
    double *vp0 = VFLD(ie, np1, 0), *vp1 = VFLD(ie, np1, 1);
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k) {
        const size_t i = IX(p, k);
        vort[i] += S2(O.fcor, ie, p);
        egrad0[i] *= -1;
        egrad0[i] += (vadv ? -v_vadv0[i] : 0) + v1[i] * vort[i];
        egrad1[i] *= -1;
        egrad1[i] += (vadv ? -v_vadv1[i] : 0) - v0[i] * vort[i];
        egrad0[i] *= dt;
        egrad0[i] += vm0[i];
        egrad1[i] *= dt;
        egrad1[i] += vm1[i];
        /* v0/v1 may alias vp0/vp1 (n0 == np1): both components are read above before any write
           at this point, and the stencil reads finished in vorticity_sphere. */
        const double r0 = S2(O.spheremp, ie, p) * egrad0[i];
        const double r1 = S2(O.spheremp, ie, p) * egrad1[i];
        vp0[i] = r0;
        vp1[i] = r1;
      }
  }
  /* compute_dp3d_np1 :468-493 (eta_dot_dpdn_buf == 0 for rsplit > 0) */
  {
    const double* dpm1 = DPFLD(ie, nm1);
    double* dpp1 = DPFLD(ie, np1);
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k) {
        const size_t i = IX(p, k);
        double tmp = (vadv && k + 1 < nlev) ? eta_buf[IX(p, k + 1)] : 0.0;
        tmp += div_vdp[i];
        tmp -= vadv ? eta_buf[i] : 0.0;
        tmp = dpm1[i] - tmp * dt;
        dpp1[i] = S2(O.spheremp, ie, p) * tmp;
      }
  }
  free(buf);
}

void hxx_caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, int with_dss) {
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) caar_element(ie, nm1, n0, np1, dt, eta_ave_w, n0_qdp);
  if (with_dss) {
    /* CaarFunctorImpl.hpp:68-79 registration order: v (2), t, dp3d */
    const size_t es_v = NTL * 2 * NLF, es = NTL * NLF;
    FieldRef f[4] = {{VFLD(0, np1, 0), es_v}, {VFLD(0, np1, 1), es_v}, {TFLD(0, np1), es}, {DPFLD(0, np1), es}};
    exchange(f, 4, true);
  }
}

/* prim_advance_exp.cpp:143-154 */
void hxx_rk_combine(int nm1, int n0) {
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    double *tm = TFLD(ie, nm1), *v0m = VFLD(ie, nm1, 0), *v1m = VFLD(ie, nm1, 1), *dpm = DPFLD(ie, nm1);
    const double *t0 = TFLD(ie, n0), *v00 = VFLD(ie, n0, 0), *v10 = VFLD(ie, n0, 1), *dp0 = DPFLD(ie, n0);
    for (size_t i = 0; i < NLF; ++i) {
      tm[i] = (5.0 * tm[i] - t0[i]) / 4.0;
      v0m[i] = (5.0 * v0m[i] - v00[i]) / 4.0;
      v1m[i] = (5.0 * v1m[i] - v10[i]) / 4.0;
      dpm[i] = (5.0 * dpm[i] - dp0[i]) / 4.0;
    }
  }
}

/* prim_advance_exp.cpp:113-161 */
static void u3_5stage_timestep(int nm1, int n0, int np1, int n0_qdp, double dt, double eta_ave_w) {
  hxx_caar_run(n0, n0, nm1, dt / 5.0, eta_ave_w / 4.0, n0_qdp, 1);
  hxx_caar_run(n0, nm1, np1, dt / 5.0, 0.0, n0_qdp, 1);
  hxx_caar_run(n0, np1, np1, dt / 3.0, 0.0, n0_qdp, 1);
  hxx_caar_run(n0, np1, np1, 2.0 * dt / 3.0, 0.0, n0_qdp, 1);
  hxx_rk_combine(nm1, n0);
  hxx_caar_run(nm1, np1, np1, 3.0 * dt / 4.0, 3.0 * eta_ave_w / 4.0, n0_qdp, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Hyperviscosity (HyperviscosityFunctorImpl.{hpp,cpp})                                        */
/* ------------------------------------------------------------------------------------------ */
static void hv_exchange(bool rspheremp) {
  /* HyperviscosityFunctorImpl.cpp:44-54: vtens (2), ttens, dptens */
  FieldRef f[4] = {{O.vtens, 2 * NLF}, {O.vtens + NLF, 2 * NLF}, {O.ttens, NLF}, {O.dptens, NLF}};
  exchange(f, 4, rspheremp);
}

void hxx_hypervis_run(int np1, double dt_in, double eta_ave_w) {
  const int nlev = O.nlev, n = O.nelemd;
  const double dt = dt_in / O.hypervis_subcycle; /* .cpp:59 */
  const double nu_scale_top[3] = {4.0 * O.nu_top, 2.0 * O.nu_top, 1.0 * O.nu_top}; /* .cpp:24-38 */
  for (int icycle = 0; icycle < O.hypervis_subcycle; ++icycle) {
    /* biharmonic_wk_dp3d .cpp:87-112 — TagFirstLaplaceHV .hpp:72-87 */
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie) {
      laplace_simple(ie, TFLD(ie, np1), O.ttens + F3(ie), nlev);
      laplace_simple(ie, DPFLD(ie, np1), O.dptens + F3(ie), nlev);
      vlaplace_sphere_wk_contra(ie, O.nu_ratio1, VFLD(ie, np1, 0), VFLD(ie, np1, 1),
                                O.vtens + ((size_t)ie * 2) * NLF, O.vtens + ((size_t)ie * 2 + 1) * NLF, nlev);
    }
    hv_exchange(true);
    /* TagSecondLaplaceConstHV .hpp:92-107 / TagSecondLaplaceTensorHV :111-131 */
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie) {
      double *vt0 = O.vtens + ((size_t)ie * 2) * NLF, *vt1 = vt0 + NLF;
      if (O.consthv) {
        laplace_simple(ie, O.ttens + F3(ie), O.ttens + F3(ie), nlev);
        laplace_simple(ie, O.dptens + F3(ie), O.dptens + F3(ie), nlev);
        vlaplace_sphere_wk_contra(ie, O.nu_ratio2, vt0, vt1, vt0, vt1, nlev);
      } else {
        laplace_tensor(ie, O.ttens + F3(ie), O.ttens + F3(ie), nlev);
        laplace_tensor(ie, O.dptens + F3(ie), O.dptens + F3(ie), nlev);
        vlaplace_sphere_wk_cartesian(ie, vt0, vt1, vt0, vt1, nlev);
      }
    }
    /* TagHyperPreExchange .hpp:161-257 */
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie) {
      double *vt0 = O.vtens + ((size_t)ie * 2) * NLF, *vt1 = vt0 + NLF, *tt = O.ttens + F3(ie), *dpt = O.dptens + F3(ie);
      const double* dp = DPFLD(ie, np1);
      double *dave = O.dpdiss_ave + F3(ie), *dbih = O.dpdiss_biharmonic + F3(ie);
      for (size_t i = 0; i < NLF; ++i) {
        dave[i] += eta_ave_w * dp[i] / O.hypervis_subcycle;
        dbih[i] += eta_ave_w * dpt[i] / O.hypervis_subcycle;
      }
      double* lap = (double*)calloc(4 * NLF, sizeof(double));
      double *lv0 = lap, *lv1 = lap + NLF, *lt = lap + 2 * NLF, *ldp = lap + 3 * NLF;
      if (O.nu_top > 0) {
        vlaplace_sphere_wk_contra(ie, 1.0, VFLD(ie, np1, 0), VFLD(ie, np1, 1), lv0, lv1, 3);
        laplace_simple(ie, TFLD(ie, np1), lt, 3);
        laplace_simple(ie, DPFLD(ie, np1), ldp, 3);
      }
      for (int p = 0; p < NPSQ; ++p) {
        for (int k = 0; k < nlev; ++k) {
          const size_t i = IX(p, k);
          vt0[i] *= -O.nu;
          vt1[i] *= -O.nu;
          tt[i] *= -O.nu_s;
          dpt[i] *= -O.nu_p;
        }
        if (O.nu_top > 0)
          for (int k = 0; k < 3; ++k) {
            const size_t i = IX(p, k);
            vt0[i] += nu_scale_top[k] * lv0[i];
            vt1[i] += nu_scale_top[k] * lv1[i];
            tt[i] += nu_scale_top[k] * lt[i];
            dpt[i] += nu_scale_top[k] * ldp[i];
          }
        for (int k = 0; k < nlev; ++k) {
          const size_t i = IX(p, k);
          dpt[i] *= dt;
          dpt[i] += dp[i] * S2(O.spheremp, ie, p);
        }
      }
      free(lap);
    }
    hv_exchange(false);
    /* TagUpdateStates .hpp:134-158 */
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie) {
      double *vt0 = O.vtens + ((size_t)ie * 2) * NLF, *vt1 = vt0 + NLF, *tt = O.ttens + F3(ie), *dpt = O.dptens + F3(ie);
      double *v0 = VFLD(ie, np1, 0), *v1 = VFLD(ie, np1, 1), *t = TFLD(ie, np1), *dp = DPFLD(ie, np1);
      for (int p = 0; p < NPSQ; ++p) {
        const double rs = S2(O.rspheremp, ie, p);
        for (int k = 0; k < nlev; ++k) {
          const size_t i = IX(p, k);
          vt0[i] = (dt * vt0[i] * rs);
          vt1[i] = (dt * vt1[i] * rs);
          v0[i] += vt0[i];
          v1[i] += vt1[i];
          tt[i] = (dt * tt[i] * rs);
          const double heating = vt0[i] * v0[i] + vt1[i] * v1[i];
          t[i] = t[i] + tt[i] - heating / cp;
          dp[i] = (dpt[i] * rs);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Euler step (EulerStepFunctorImpl.hpp)                                                       */
/* ------------------------------------------------------------------------------------------ */
void hxx_euler_reset(void) { O.rhs_viss = 0.0; } /* :118-135 */

/* :348-377 */
void hxx_euler_precompute_divdp(void) {
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    divergence_sphere(ie, O.derived_vn0 + ((size_t)ie * 2) * NLF, O.derived_vn0 + ((size_t)ie * 2 + 1) * NLF,
                      O.divdp + F3(ie), O.nlev);
    memcpy(O.divdp_proj + F3(ie), O.divdp + F3(ie), NLF * sizeof(double));
  }
}

/* :379-403 */
void hxx_euler_qdp_time_avg(int n0_qdp, int np1_qdp) {
  const double rkstage = 3.0;
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int q = 0; q < O.qsize; ++q) {
      const double* a = QDPFLD(ie, n0_qdp, q);
      double* b = QDPFLD(ie, np1_qdp, q);
      for (size_t i = 0; i < NLF; ++i) b[i] = (a[i] + (rkstage - 1) * b[i]) / rkstage;
    }
}

/* limiter shell + lim8 / lim9 for one level: x, c are the 16 point values (:693-884), sums in
 * k = 0..15 ascending as in the serial Dispatch<>::parallel_reduce_NP2. */
static void limiter_level(int limiter_option, const double* sphweights, const double* dpmass, size_t stride,
                          double* ptens, double* qlim_min, double* qlim_max) {
  double x[NPSQ], c[NPSQ];
  for (int k = 0; k < NPSQ; ++k) {
    const double dpm = dpmass[k * stride];
    c[k] = sphweights[k] * dpm;
    x[k] = ptens[k * stride] / dpm;
  }
  double mass = 0, sumc = 0;
  for (int k = 0; k < NPSQ; ++k) {
    mass += x[k] * c[k];
    sumc += c[k];
  }
  if (sumc <= 0) return;
  double minp = *qlim_min, maxp = *qlim_max;
  if (minp < 0) minp = *qlim_min = 0;
  if (mass < minp * sumc) minp = *qlim_min = mass / sumc;
  if (mass > maxp * sumc) maxp = *qlim_max = mass / sumc;

  if (limiter_option == 8) { /* :766-823 */
    const int maxiter = NP * NP - 1;
    const double tol_limiter = 5e-14;
    for (int iter = 0; iter < maxiter; ++iter) {
      double addmass = 0;
      for (int k = 0; k < NPSQ; ++k) {
        double delta = 0;
        if (x[k] > maxp) { delta = x[k] - maxp; x[k] = maxp; }
        else if (x[k] < minp) { delta = x[k] - minp; x[k] = minp; }
        addmass += delta * c[k];
      }
      if (fabs(addmass) <= tol_limiter * fabs(mass)) break;
      if (addmass > 0) {
        double weightssum = 0;
        for (int k = 0; k < NPSQ; ++k)
          if (x[k] < maxp) weightssum += c[k];
        const double adw = addmass / weightssum;
        for (int k = 0; k < NPSQ; ++k) x[k] += (x[k] < maxp) ? adw : 0;
      } else {
        double weightssum = 0;
        for (int k = 0; k < NPSQ; ++k)
          if (x[k] > minp) weightssum += c[k];
        const double adw = addmass / weightssum;
        for (int k = 0; k < NPSQ; ++k) x[k] += (x[k] > minp) ? adw : 0;
      }
    }
  } else { /* limiter 9 :826-884 */
    double addmass = 0;
    for (int k = 0; k < NPSQ; ++k) {
      double delta = 0;
      if (x[k] > maxp) { delta = x[k] - maxp; x[k] = maxp; }
      else if (x[k] < minp) { delta = x[k] - minp; x[k] = minp; }
      addmass += delta * c[k];
    }
    if (addmass != 0) {
      if (addmass > 0) {
        double fac = 0;
        for (int k = 0; k < NPSQ; ++k) fac += c[k] * (maxp - x[k]);
        if (fac > 0) {
          fac = addmass / fac;
          for (int k = 0; k < NPSQ; ++k) x[k] += fac * (maxp - x[k]);
        }
      } else {
        double fac = 0;
        for (int k = 0; k < NPSQ; ++k) fac += c[k] * (x[k] - minp);
        if (fac > 0) {
          fac = addmass / fac;
          for (int k = 0; k < NPSQ; ++k) x[k] += fac * (x[k] - minp);
        }
      }
    }
  }
  for (int k = 0; k < NPSQ; ++k) ptens[k * stride] = x[k] * dpmass[k * stride];
}

void hxx_limiter(int limiter_option, int nsets, const double* sphweights, const double* dpmass, double* ptens,
                 double* qlim) {
  const int nlev = O.nlev;
  for (int s = 0; s < nsets; ++s)
    for (int k = 0; k < nlev; ++k)
      limiter_level(limiter_option, sphweights + (size_t)s * NPSQ, dpmass + (size_t)s * NLF + k, nlev,
                    ptens + (size_t)s * NLF + k, qlim + ((size_t)s * 2) * nlev + k, qlim + ((size_t)s * 2 + 1) * nlev + k);
}

static double* dss_var(int dss_opt) {
  return dss_opt == DSS_ETA ? O.eta_dot_dpdn : dss_opt == DSS_OMEGA ? O.omega_p : O.divdp_proj;
}

/* euler_step :514-561 */
void hxx_euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt) {
  const int nlev = O.nlev, n = O.nelemd, nq = O.qsize;
  /* compute_dp :406-434 */
  const double rhsmdt = rhs_multiplier * dt;
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie) {
    double* buf = O.dp_star + F3(ie);
    const double *dp = O.derived_dp + F3(ie), *dproj = O.divdp_proj + F3(ie);
    for (size_t i = 0; i < NLF; ++i) buf[i] = dp[i] - rhsmdt * dproj[i];
  }
  /* compute_qmin_qmax :436-485 */
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie)
    for (int q = 0; q < nq; ++q) {
      const double *dp_t = O.dp_star + F3(ie), *qdp_t = QDPFLD(ie, n0_qdp, q);
      double* qt = O.qtens_biharmonic + ((size_t)ie * O.qsize_d + q) * NLF;
      double* ql = O.qlim + ((size_t)ie * O.qsize_d + q) * 2 * nlev;
      if (rhs_multiplier != 1.0)
        for (int k = 0; k < nlev; ++k) {
          const double v = qdp_t[IX(0, k)] / dp_t[IX(0, k)];
          qt[IX(0, k)] = v; ql[k] = v; ql[nlev + k] = v;
        }
      for (int p = 0; p < NPSQ; ++p)
        for (int k = 0; k < nlev; ++k) {
          const double v = qdp_t[IX(p, k)] / dp_t[IX(p, k)];
          qt[IX(p, k)] = v;
          ql[k] = fmin(ql[k], v);
          ql[nlev + k] = fmax(ql[nlev + k], v);
        }
    }
  if (rhs_multiplier == 0.0) {
    exchange_min_max(); /* neighbor_minmax :504-507 */
  } else if (rhs_multiplier == 2.0) {
    /* minmax_and_biharmonic :496-502. The min/max exchange is started before and finished after
       the biharmonic; qlim is not touched in between, so one exchange at the end is identical. */
    O.rhs_viss = 3.0; /* compute_biharmonic_pre :196-214 */
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie)
      for (int q = 0; q < nq; ++q) {
        double* qt = O.qtens_biharmonic + ((size_t)ie * O.qsize_d + q) * NLF;
        if (O.nu_p > 0) { /* dpdiss_adjustment :251-267 */
          const double* dave = O.dpdiss_ave + F3(ie);
          for (int p = 0; p < NPSQ; ++p)
            for (int k = 0; k < nlev; ++k) qt[IX(p, k)] = qt[IX(p, k)] * dave[IX(p, k)] / O.dp0[k];
        }
        laplace_simple(ie, qt, qt, nlev);
      }
    {
      FieldRef* f = (FieldRef*)malloc(nq * sizeof(FieldRef)); /* m_mmqb_be :155-161 */
      for (int q = 0; q < nq; ++q) { f[q].base = O.qtens_biharmonic + (size_t)q * NLF; f[q].estride = (size_t)O.qsize_d * NLF; }
      exchange(f, nq, true);
      free(f);
    }
    /* compute_biharmonic_post :216-231 + rhsviss_adjustment :293-310 */
    const double fac = -O.rhs_viss * dt * O.nu_q;
#pragma omp parallel for
    for (int ie = 0; ie < n; ++ie)
      for (int q = 0; q < nq; ++q) {
        double* qt = O.qtens_biharmonic + ((size_t)ie * O.qsize_d + q) * NLF;
        if (O.consthv) laplace_simple(ie, qt, qt, nlev);
        else laplace_tensor(ie, qt, qt, nlev);
        for (int p = 0; p < NPSQ; ++p)
          for (int k = 0; k < nlev; ++k)
            qt[IX(p, k)] = (fac * O.dp0[k] * qt[IX(p, k)] / S2(O.spheremp, ie, p));
      }
    exchange_min_max();
  }
  /* advect_and_limit :317-332 — AALSetupPhase = compute_2d_advection_step :585-626 */
  const bool add_ps_diss = O.nu_p > 0 && O.rhs_viss != 0.0;
  const double diss_fac = add_ps_diss ? -O.rhs_viss * dt * O.nu_q : 0;
  double* f_dss = dss_var(dss_opt);
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie) {
    const double *dps = O.dp_star + F3(ie), *vn00 = O.derived_vn0 + ((size_t)ie * 2) * NLF, *vn01 = vn00 + NLF;
    double *vs0 = O.vstar + ((size_t)ie * 2) * NLF, *vs1 = vs0 + NLF, *dpk = O.dpdissk + F3(ie);
    const double *ddiv = O.divdp + F3(ie), *dbih = O.dpdiss_biharmonic + F3(ie);
    double* fd = f_dss + F3(ie);
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k) {
        const size_t i = IX(p, k);
        const double dp = dps[i];
        vs0[i] = vn00[i] / dp;
        vs1[i] = vn01[i] / dp;
        dpk[i] = dp - dt * ddiv[i];
        if (add_ps_diss) dpk[i] += diss_fac * dbih[i] / S2(O.spheremp, ie, p);
        fd[i] *= S2(O.spheremp, ie, p);
      }
  }
  /* AALTracerPhase = run_tracer_phase :571-582 */
#pragma omp parallel for
  for (int ie = 0; ie < n; ++ie)
    for (int q = 0; q < nq; ++q) {
      double* qt = O.qtens_biharmonic + ((size_t)ie * O.qsize_d + q) * NLF;
      const double* vs0 = O.vstar + ((size_t)ie * 2) * NLF;
      divergence_sphere_update(ie, -dt, O.rhs_viss != 0.0, vs0, vs0 + NLF, QDPFLD(ie, n0_qdp, q), qt);
      double* ql = O.qlim + ((size_t)ie * O.qsize_d + q) * 2 * nlev;
      for (int k = 0; k < nlev; ++k)
        limiter_level(O.limiter_option, O.spheremp + (size_t)ie * NPSQ, O.dpdissk + F3(ie) + k, nlev, qt + k,
                      ql + k, ql + nlev + k);
      double* out = QDPFLD(ie, np1_qdp, q); /* apply_spheremp :672-687 */
      for (int p = 0; p < NPSQ; ++p)
        for (int k = 0; k < nlev; ++k) out[IX(p, k)] = S2(O.spheremp, ie, p) * qt[IX(p, k)];
    }
  /* exchange_qdp_dss_var :509-512 — fields: qdp(np1_qdp, 0..qsize-1), then the DSS variable */
  {
    FieldRef* f = (FieldRef*)malloc((nq + 1) * sizeof(FieldRef));
    for (int q = 0; q < nq; ++q) { f[q].base = QDPFLD(0, np1_qdp, q); f[q].estride = (size_t)QNTL * O.qsize_d * NLF; }
    f[nq].base = f_dss; f[nq].estride = NLF;
    exchange(f, nq + 1, true);
    free(f);
  }
}

/* prim_advec_tracers_remap_RK2, prim_advec_tracers_remap.cpp:32-90 */
static void update_tracers_levels(void) { /* TimeLevel.hpp:58-67 */
  const int i_temp = O.nstep / O.qsplit;
  if (i_temp % 2 == 0) { O.n0_qdp = 0; O.np1_qdp = 1; }
  else { O.n0_qdp = 1; O.np1_qdp = 0; }
}
static void update_dynamics_levels(void) { /* LEAPFROG, TimeLevel.hpp:37-56 */
  const int tmp = O.np1;
  O.np1 = O.nm1; O.nm1 = O.n0; O.n0 = tmp;
  ++O.nstep;
}

static void prim_advec_tracers_remap_RK2(double dt) {
  update_tracers_levels();
  hxx_euler_reset();
  hxx_euler_precompute_divdp();
  hxx_euler_step(O.np1_qdp, O.n0_qdp, dt / 2.0, 0.0, DSS_DIV_VDP_AVE);
  hxx_euler_step(O.np1_qdp, O.np1_qdp, dt / 2.0, 1.0, DSS_ETA);
  hxx_euler_step(O.np1_qdp, O.np1_qdp, dt / 2.0, 2.0, DSS_OMEGA);
  hxx_euler_qdp_time_avg(O.n0_qdp, O.np1_qdp);
}

/* ------------------------------------------------------------------------------------------ */
/* Vertical remap (RemapFunctor.hpp, PpmRemap.hpp), rsplit > 0                                 */
/* ------------------------------------------------------------------------------------------ */
#define PPM_PAD 2 /* only the 2 ghost cells matter; the reference pads further for alignment */

typedef struct {
  double *dpo, *pio, *pin, *ppmdx, *z2;
  int* kid;
} ColumnGrid;

/* compute_partitions :506-597 + compute_integral_bounds :600-666 + compute_grids :366-413 */
static void ppm_column_grids(int nlev, const double* src, size_t sstride, const double* tgt, size_t tstride,
                             double* dpo /*nlev+4*/, double* pio /*nlev+2*/, double* pin /*nlev+1*/,
                             double* ppmdx /*10*(nlev+2)*/, double* z2, int* kid) {
  double acc = 0;
  for (int k = 0; k < nlev; ++k) { pio[k] = acc; acc += src[k * sstride]; }
  pio[nlev] = pio[nlev - 1] + src[(nlev - 1) * sstride];
  acc = 0;
  for (int k = 0; k < nlev; ++k) { pin[k] = acc; acc += tgt[k * tstride]; }
  pin[nlev] = pin[nlev - 1] + tgt[(nlev - 1) * tstride];
  pio[nlev + 1] = pio[nlev] + 1.0;
  pin[nlev] = pio[nlev];
  for (int k = 0; k < nlev; ++k) dpo[k + PPM_PAD] = src[k * sstride];
  for (int k = 0; k < 2; ++k) {
    dpo[PPM_PAD - 1 - k] = dpo[k + PPM_PAD];
    dpo[nlev + PPM_PAD + k] = dpo[nlev + PPM_PAD - 1 - k];
  }
  for (int k = 0; k < nlev; ++k) {
    int kk = k + 1;
    while (pio[kk - 1] <= pin[k + 1]) kk++;
    kk--;
    if (kk == nlev + 1) kk = nlev;
    kid[k] = kk - 1;
    z2[k] = (pin[k + 1] - (pio[kk - 1] + pio[kk]) * 0.5) / dpo[kk + 1 + PPM_PAD - 2];
  }
  const int L = nlev + 2;
  const double* dx = dpo; /* dpo_offset = PAD - gs = 0 */
  for (int j = 0; j < nlev + 2; ++j) {
    ppmdx[0 * L + j] = dx[j + 1] / (dx[j] + dx[j + 1] + dx[j + 2]);
    ppmdx[1 * L + j] = (2.0 * dx[j] + dx[j + 1]) / (dx[j + 1] + dx[j + 2]);
    ppmdx[2 * L + j] = (dx[j + 1] + 2.0 * dx[j + 2]) / (dx[j] + dx[j + 1]);
  }
  for (int j = 0; j < nlev + 1; ++j) {
    ppmdx[3 * L + j] = dx[j + 1] / (dx[j + 1] + dx[j + 2]);
    ppmdx[4 * L + j] = 1.0 / (dx[j] + dx[j + 1] + dx[j + 2] + dx[j + 3]);
    ppmdx[5 * L + j] = (2.0 * dx[j + 1] * dx[j + 2]) / (dx[j + 1] + dx[j + 2]);
    ppmdx[6 * L + j] = (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    ppmdx[7 * L + j] = (dx[j + 3] + dx[j + 2]) / (2.0 * dx[j + 2] + dx[j + 1]);
    ppmdx[8 * L + j] = dx[j + 1] * (dx[j] + dx[j + 1]) / (2.0 * dx[j + 1] + dx[j + 2]);
    ppmdx[9 * L + j] = dx[j + 2] * (dx[j + 2] + dx[j + 3]) / (dx[j + 1] + 2.0 * dx[j + 2]);
  }
}

static double integrate_parabola(double sq, double lin, double cst, double x1, double x2) { /* :668-673 */
  return (cst * (x2 - x1) + lin * (x2 * x2 - x1 * x1) / 2.0) + sq * (x2 * x2 * x2 - x1 * x1 * x1) / 3.0;
}

/* compute_remap_phase :203-266 for one column of one field (in place) */
static void ppm_column_remap(int alg, int nlev, const double* dpo, const double* ppmdx, const double* z2,
                             const int* kid, double* var, size_t vstride, double* work) {
  const int L = nlev + 2;
  double *ao = work, *mass_o = work + (nlev + 4), *dma = mass_o + (nlev + 2), *ai = dma + (nlev + 2),
         *coef = ai + (nlev + 1);
  for (int k = 0; k < nlev; ++k) ao[k + PPM_PAD] = var[k * vstride] / dpo[k + PPM_PAD];
  /* fill_cell_means_gs: mirrored (alg 1 and 2) :87-101 */
  for (int k0 = 0; k0 < 2; ++k0) {
    ao[PPM_PAD - 1 - k0] = ao[k0 + PPM_PAD];
    ao[nlev + PPM_PAD + k0] = ao[nlev + PPM_PAD - 1 - k0];
  }
  {
    double acc = 0;
    mass_o[0] = 0;
    for (int k = 0; k < nlev; ++k) { mass_o[k + 1] = acc; acc += var[k * vstride]; }
    mass_o[nlev + 1] = mass_o[nlev] + var[(nlev - 1) * vstride];
  }
  /* compute_ppm :416-503 */
  for (int j = 0; j < nlev + 2; ++j) {
    if ((ao[j + PPM_PAD] - ao[j + PPM_PAD - 1]) * (ao[j + PPM_PAD - 1] - ao[j + PPM_PAD - 2]) > 0.0) {
      const double da = ppmdx[0 * L + j] * (ppmdx[1 * L + j] * (ao[j + PPM_PAD] - ao[j + PPM_PAD - 1]) +
                                            ppmdx[2 * L + j] * (ao[j + PPM_PAD - 1] - ao[j + PPM_PAD - 2]));
      dma[j] = fmin(fmin(fabs(da), 2.0 * fabs(ao[j + PPM_PAD - 1] - ao[j + PPM_PAD - 2])),
                    2.0 * fabs(ao[j + PPM_PAD] - ao[j + PPM_PAD - 1])) *
               copysign(1.0, da);
    } else {
      dma[j] = 0.0;
    }
  }
  for (int j = 0; j < nlev + 1; ++j) {
    ai[j] = ao[j + PPM_PAD - 1] + ppmdx[3 * L + j] * (ao[j + PPM_PAD] - ao[j + PPM_PAD - 1]) +
            ppmdx[4 * L + j] * (ppmdx[5 * L + j] * (ppmdx[6 * L + j] - ppmdx[7 * L + j]) *
                                    (ao[j + PPM_PAD] - ao[j + PPM_PAD - 1]) -
                                ppmdx[8 * L + j] * dma[j + 1] + ppmdx[9 * L + j] * dma[j]);
  }
  for (int jp = 0; jp < nlev; ++jp) {
    const int j = jp + 1;
    const double am = ao[j + PPM_PAD - 1];
    double al = ai[j - 1], ar = ai[j];
    if ((ar - am) * (am - al) <= 0.) { al = am; ar = am; }
    if ((ar - al) * (am - (al + ar) / 2.0) > (ar - al) * (ar - al) / 6.0) al = 3.0 * am - 2.0 * ar;
    if ((ar - al) * (am - (al + ar) / 2.0) < -(ar - al) * (ar - al) / 6.0) ar = 3.0 * am - 2.0 * al;
    coef[0 * nlev + j - 1] = 1.5 * am - (al + ar) / 4.0;
    coef[1 * nlev + j - 1] = ar - al;
    coef[2 * nlev + j - 1] = 3.0 * (-2.0 * am + (al + ar));
  }
  if (alg == 2) { /* PpmFixedParabola::apply_ppm_boundary :110-133 */
    coef[0 * nlev + 0] = ao[PPM_PAD]; coef[0 * nlev + 1] = ao[PPM_PAD + 1];
    coef[0 * nlev + nlev - 2] = ao[PPM_PAD + nlev - 2]; coef[0 * nlev + nlev - 1] = ao[PPM_PAD + nlev - 1];
    for (int c = 1; c < 3; ++c) {
      coef[c * nlev + 0] = 0.0; coef[c * nlev + 1] = 0.0;
      coef[c * nlev + nlev - 2] = 0.0; coef[c * nlev + nlev - 1] = 0.0;
    }
  }
  /* compute_remap, non-GPU :283-324 (the mass buffer is reused in place, k ascending) */
  for (int k = 0; k < nlev; ++k) {
    const int kk = kid[k];
    const double integral = integrate_parabola(coef[2 * nlev + kk], coef[1 * nlev + kk], coef[0 * nlev + kk], -0.5, z2[k]);
    mass_o[k] = mass_o[kk + 1] + integral * dpo[kk + PPM_PAD];
  }
  for (int k = nlev - 1; k > 0; --k) var[k * vstride] = mass_o[k] - mass_o[k - 1];
  var[0] = mass_o[0];
}

void hxx_remap_columns(int alg, int ncols, int nfields, const double* src_dp, const double* tgt_dp, double* fields) {
  const int nlev = O.nlev;
#pragma omp parallel for
  for (int c = 0; c < ncols; ++c) {
    double* g = (double*)malloc(sizeof(double) * ((nlev + 4) + (nlev + 2) + (nlev + 1) + 10 * (nlev + 2) + nlev + 8 * (nlev + 4)));
    double *dpo = g, *pio = dpo + nlev + 4, *pin = pio + nlev + 2, *ppmdx = pin + nlev + 1, *z2 = ppmdx + 10 * (nlev + 2),
           *work = z2 + nlev;
    int* kid = (int*)malloc(sizeof(int) * nlev);
    ppm_column_grids(nlev, src_dp + (size_t)c * nlev, 1, tgt_dp + (size_t)c * nlev, 1, dpo, pio, pin, ppmdx, z2, kid);
    for (int f = 0; f < nfields; ++f)
      ppm_column_remap(alg, nlev, dpo, ppmdx, z2, kid, fields + ((size_t)f * ncols + c) * nlev, 1, work);
    free(kid);
    free(g);
  }
}

/* RemapFunctor::run_remap :306-331. rsplit == 0 (:42-100): the source thickness is the reference thickness
   plus dt times the increment of the mean vertical flux, and only the tracers are remapped. */
void hxx_vertical_remap(int np1, int np1_qdp, double dt) {
  const int nlev = O.nlev, n = O.nelemd, nq = O.qsize;
  const bool lagrangian = O.rsplit > 0;
  int invalid_any = 0;
#pragma omp parallel for reduction(| : invalid_any)
  for (int ie = 0; ie < n; ++ie) {
    double* tgt = (double*)malloc(2 * NLF * sizeof(double));
    const double* src = DPFLD(ie, np1);
    const double* dp_np1 = src;
    double* ps = O.ps_v + ((size_t)ie * NTL + np1) * NPSQ;
    /* compute_ps_v :367-385 */
    for (int p = 0; p < NPSQ; ++p) {
      ps[p] = 0.0;
      for (int k = 0; k < nlev; ++k) ps[p] += dp_np1[IX(p, k)];
      ps[p] += O.hyai0 * O.ps0;
    }
    /* compute_target_thickness :417-437 */
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k)
        tgt[IX(p, k)] = (O.hyai[k + 1] - O.hyai[k]) * O.ps0 + (O.hybi[k + 1] - O.hybi[k]) * ps[p];
    if (!lagrangian) { /* compute_source_thickness :60-92 */
      double* s0 = tgt + NLF;
      const double* eta = O.eta_dot_dpdn + F3(ie);
      for (int p = 0; p < NPSQ; ++p)
        for (int k = 0; k < nlev; ++k) {
          const double eta_next = k + 1 < nlev ? eta[IX(p, k + 1)] : 0;
          const double delta_dpdn = eta_next - eta[IX(p, k)];
          s0[IX(p, k)] = tgt[IX(p, k)] + dt * delta_dpdn;
        }
      src = s0;
    }
    /* check_source_thickness :439-464 */
    int invalid = 0;
    for (size_t i = 0; i < NLF; ++i) invalid |= (isnan(src[i]) || src[i] < 0.0);
    invalid_any |= invalid;
    const int nst = lagrangian ? 3 : 0;
    if (!invalid && (nq + nst) > 0) {
      /* ComputeExtrinsicsTag :255-268 */
      double* st[3] = {VFLD(ie, np1, 0), VFLD(ie, np1, 1), TFLD(ie, np1)};
      for (int s = 0; s < nst; ++s)
        for (size_t i = 0; i < NLF; ++i) st[s][i] *= src[i];
      double* g = (double*)malloc(sizeof(double) * ((nlev + 4) + (nlev + 2) + (nlev + 1) + 10 * (nlev + 2) + nlev + 8 * (nlev + 4)));
      double *dpo = g, *pio = dpo + nlev + 4, *pin = pio + nlev + 2, *ppmdx = pin + nlev + 1, *z2 = ppmdx + 10 * (nlev + 2),
             *work = z2 + nlev;
      int* kid = (int*)malloc(sizeof(int) * nlev);
      for (int p = 0; p < NPSQ; ++p) {
        ppm_column_grids(nlev, src + IX(p, 0), 1, tgt + IX(p, 0), 1, dpo, pio, pin, ppmdx, z2, kid);
        for (int s = 0; s < nst; ++s) ppm_column_remap(O.remap_alg, nlev, dpo, ppmdx, z2, kid, st[s] + IX(p, 0), 1, work);
        for (int q = 0; q < nq; ++q)
          ppm_column_remap(O.remap_alg, nlev, dpo, ppmdx, z2, kid, QDPFLD(ie, np1_qdp, q) + IX(p, 0), 1, work);
      }
      free(kid);
      free(g);
      /* ComputeIntrinsicsTag :294-307 */
      for (int s = 0; s < nst; ++s)
        for (size_t i = 0; i < NLF; ++i) st[s][i] /= tgt[i];
    }
    free(tgt);
  }
  if (invalid_any) runtime_abort("Negative (or nan) layer thickness detected, aborting!", 101);
}

/* prim_driver.cpp:171-206 */
void hxx_update_q(int np1_qdp, int np1) {
  const int nlev = O.nlev;
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int q = 0; q < O.qsize; ++q) {
      const double* qd = QDPFLD(ie, np1_qdp, q);
      double* Q = O.Q + ((size_t)ie * O.qsize_d + q) * NLF;
      for (int p = 0; p < NPSQ; ++p)
        for (int k = 0; k < nlev; ++k) {
          const double dp = O.dai[k] * O.ps0 + O.dbi[k] * O.ps_v[((size_t)ie * NTL + np1) * NPSQ + p];
          Q[IX(p, k)] = qd[IX(p, k)] / dp;
        }
    }
}

/* prim_step.cpp:51-66 */
void hxx_prim_step_init(int n0) {
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    memset(O.eta_dot_dpdn + F3(ie), 0, NLF * 8);
    memset(O.derived_vn0 + (size_t)ie * 2 * NLF, 0, 2 * NLF * 8);
    memset(O.omega_p + F3(ie), 0, NLF * 8);
    if (O.nu_p > 0) {
      memset(O.dpdiss_ave + F3(ie), 0, NLF * 8);
      memset(O.dpdiss_biharmonic + F3(ie), 0, NLF * 8);
    }
    memcpy(O.derived_dp + F3(ie), DPFLD(ie, n0), NLF * 8);
  }
}

/* prim_step.cpp:20-103 */
static void prim_step(double dt) {
  hxx_prim_step_init(O.n0);
  for (int nq = 0; nq < O.qsplit; ++nq) {
    if (nq > 0) update_dynamics_levels();
    /* prim_advance_exp, prim_advance_exp.cpp:25-111 */
    O.n0_qdp = -1;
    if (O.moist) update_tracers_levels();
    const double eta_ave_w = 1.0 / O.qsplit;
    u3_5stage_timestep(O.nm1, O.n0, O.np1, O.n0_qdp, dt, eta_ave_w);
    hxx_hypervis_run(O.np1, dt, eta_ave_w);
  }
  if (O.qsize > 0) prim_advec_tracers_remap_RK2(dt * O.qsplit);
}

/* CamForcing.cpp:20-49 state_forcing */
static void state_forcing(int np1, double dt) {
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    double* t = TFLD(ie, np1);
    const double* ft = O.ft + F3(ie);
    for (size_t i = 0; i < NLF; ++i) t[i] += dt * ft[i];
    for (int c = 0; c < 2; ++c) {
      double* v = VFLD(ie, np1, c);
      const double* fm = O.fm + ((size_t)ie * 2 + c) * NLF;
      for (size_t i = 0; i < NLF; ++i) v[i] += dt * fm[i];
    }
  }
}

/* CamForcing.cpp:51-147 tracer_forcing (np1 = tl.n0, np1_qdp = tl.n0_qdp) */
static void tracer_forcing(double dt) {
  const int nlev = O.nlev, np1 = O.n0, np1_qdp = O.n0_qdp;
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    double* ps = O.ps_v + ((size_t)ie * NTL + np1) * NPSQ;
    if (O.moist) { /* :65-108, reads qdp before the update below */
      const double* fq0 = O.fq + (size_t)ie * O.qsize_d * NLF;
      const double* q0 = QDPFLD(ie, np1_qdp, 0);
      for (int p = 0; p < NPSQ; ++p) {
        double acc = 0.0;
        for (int k = 0; k < nlev; ++k) {
          double v1 = dt * fq0[IX(p, k)];
          const double qs = q0[IX(p, k)];
          if (qs + v1 < 0.0 && v1 < 0.0) v1 = qs < 0.0 ? 0.0 : -qs;
          acc += v1;
        }
        ps[p] += acc;
      }
    }
    for (int q = 0; q < O.qsize; ++q) { /* :110-131 */
      const double* fq = O.fq + ((size_t)ie * O.qsize_d + q) * NLF;
      double* qd = QDPFLD(ie, np1_qdp, q);
      for (size_t i = 0; i < NLF; ++i) {
        double v1 = dt * fq[i];
        if (qd[i] + v1 < 0.0 && v1 < 0.0) v1 = qd[i] < 0.0 ? 0.0 : -qd[i];
        qd[i] += v1;
      }
    }
    for (int q = 0; q < O.qsize; ++q) { /* :133-146 */
      const double* qd = QDPFLD(ie, np1_qdp, q);
      double* Q = O.Q + ((size_t)ie * O.qsize_d + q) * NLF;
      for (int p = 0; p < NPSQ; ++p)
        for (int k = 0; k < nlev; ++k) {
          const double dp = O.dai[k] * O.ps0 + O.dbi[k] * ps[p];
          Q[IX(p, k)] = qd[IX(p, k)] / dp;
        }
    }
  }
}

/* CamForcing.cpp:149-174 */
static void apply_cam_forcing(double dt) {
  state_forcing(O.n0, dt);
  tracer_forcing(dt);
}
static void apply_cam_forcing_dynamics(double dt) { state_forcing(O.n0, dt); }

/* Diagnostics.cpp:37-90. diag[] = {Q, Qvar, Qmass, Q1mass, IEner, IEner_wet, KEner, PEner} (F90-owned):
   Q [ie][qsize_d][nlev][16]; Qvar, Qmass [ie][4][qsize_d][16]; Q1mass [ie][qsize_d][16];
   IEner, KEner, PEner [ie][4][16]; IEner_wet [ie][16] */
static void prim_diag_scalars(bool before_advance, int ivar) {
  const int nlev = O.nlev, qd_ = O.qsize_d;
  update_tracers_levels();
  const int t2_qdp = before_advance ? O.n0_qdp : O.np1_qdp;
  if (O.time_step_type <= 0) return;
  double *hQ = O.diag[0], *Qvar = O.diag[1], *Qmass = O.diag[2], *Q1mass = O.diag[3];
  for (int ie = 0; ie < O.nelemd; ++ie) /* sync_to_host(tracers.Q, h_Q): all QSIZE_D tracers */
    for (int q = 0; q < qd_; ++q)
      for (int k = 0; k < nlev; ++k)
        for (int p = 0; p < NPSQ; ++p)
          hQ[(((size_t)ie * qd_ + q) * nlev + k) * NPSQ + p] = O.Q[((size_t)ie * qd_ + q) * NLF + IX(p, k)];
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int q = 0; q < O.qsize; ++q) {
      const double* qdp = QDPFLD(ie, t2_qdp, q);
      for (int p = 0; p < NPSQ; ++p) {
        double accum_qdp_q = 0, accum_qdp = 0;
        for (int k = 0; k < nlev; ++k) {
          accum_qdp_q += qdp[IX(p, k)] * hQ[(((size_t)ie * qd_ + q) * nlev + k) * NPSQ + p];
          accum_qdp += qdp[IX(p, k)];
        }
        Qvar[(((size_t)ie * 4 + ivar) * qd_ + q) * NPSQ + p] = accum_qdp_q;
        Qmass[(((size_t)ie * 4 + ivar) * qd_ + q) * NPSQ + p] = accum_qdp;
        Q1mass[((size_t)ie * qd_ + q) * NPSQ + p] = accum_qdp;
      }
    }
}

/* Diagnostics.cpp:92-185 */
static void prim_energy_halftimes(bool before_advance, int ivar) {
  const int nlev = O.nlev;
  const double cp = 1005.0, cpwv = 1870.0; /* PhysicalConstants.hpp */
  update_tracers_levels();
  const int t1 = before_advance ? O.n0 : O.np1, t1_qdp = before_advance ? O.n0_qdp : O.np1_qdp;
  double *IE = O.diag[4], *IEw = O.diag[5], *KE = O.diag[6], *PE = O.diag[7];
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int p = 0; p < NPSQ; ++p) {
      double IEner = 0.0, IEner_wet = 0.0, KEner = 0.0, PEner = 0.0;
      const double *u = VFLD(ie, t1, 0), *v = VFLD(ie, t1, 1), *T = TFLD(ie, t1);
      const double ps = O.ps_v[((size_t)ie * NTL + t1) * NPSQ + p], phis = S2(O.phis, ie, p);
      for (int k = 0; k < nlev; ++k) {
        const double dpt1 = O.dai[k] * O.ps0 + O.dbi[k] * ps;
        double cp_star1 = cp;
        if (O.use_cpstar) {
          const double qval = QDPFLD(ie, t1_qdp, 0)[IX(p, k)] / dpt1;
          cp_star1 = cp * (1.0 + (cpwv / cp - 1.0) * qval);
        }
        IEner += cp_star1 * T[IX(p, k)] * dpt1;
        IEner_wet += (cp_star1 - cp) * T[IX(p, k)] * dpt1;
        KEner += (u[IX(p, k)] * u[IX(p, k)] + v[IX(p, k)] * v[IX(p, k)]) * 0.5 * dpt1;
        PEner += phis * dpt1;
      }
      IE[((size_t)ie * 4 + ivar) * NPSQ + p] = IEner;
      IEw[(size_t)ie * NPSQ + p] = IEner_wet;
      KE[((size_t)ie * 4 + ivar) * NPSQ + p] = KEner;
      PE[((size_t)ie * 4 + ivar) * NPSQ + p] = PEner;
    }
}

/* Held-Suarez forcing, physics/heldsuarez/held_suarez_mod.F90:123-279 (scalar branches of hs_v_forcing and
 * hs_T_forcing) on time level n0; FM, FT overwritten. The checker of the product's device kernel. */
void hxx_held_suarez_forcing(const double* lat, const double* hyam, const double* hybm) {
  const int nlev = O.nlev, n0 = O.n0;
  const double sigma_b = 0.70, secpday = 86400.0;
  const double k_a = 1.0 / (40.0 * secpday), k_f = 1.0 / (1.0 * secpday), k_s = 1.0 / (4.0 * secpday);
  const double dT_y = 60.0, dtheta_z = 10.0;
  if (!O.fm) O.fm = zalloc((size_t)O.nelemd * 2 * NLF);
  if (!O.ft) O.ft = zalloc((size_t)O.nelemd * NLF);
  if (!O.fq) O.fq = zalloc((size_t)O.nelemd * O.qsize_d * NLF);
  const double logps0 = log(O.ps0);
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie)
    for (int p = 0; p < NPSQ; ++p) {
      const double ps = O.ps_v[((size_t)ie * NTL + n0) * NPSQ + p];
      const double snlat = sin(lat[(size_t)ie * NPSQ + p]);
      const double snlatsq = snlat * snlat, cslatsq = 1.0 - snlatsq;
      for (int k = 0; k < nlev; ++k) {
        const double pm = hyam[k] * O.ps0 + hybm[k] * ps;
        const double logprat = log(pm) - logps0;
        const double pratk = exp(kappa * logprat);
        const double etam = hyam[k] + hybm[k];
        const double ramp = fmax(0.0, (etam - sigma_b) / (1.0 - sigma_b));
        const double k_t = k_a + (k_s - k_a) * cslatsq * cslatsq * ramp;
        const double Teq = fmax(200.0, (315.0 - dT_y * snlatsq - dtheta_z * logprat * cslatsq) * pratk);
        O.ft[F3(ie) + IX(p, k)] = -k_t * (TFLD(ie, n0)[IX(p, k)] - Teq);
        const double k_v = k_f * ramp;
        O.fm[((size_t)ie * 2 + 0) * NLF + IX(p, k)] = -k_v * VFLD(ie, n0, 0)[IX(p, k)];
        O.fm[((size_t)ie * 2 + 1) * NLF + IX(p, k)] = -k_v * VFLD(ie, n0, 1)[IX(p, k)];
      }
    }
}

void hxx_apply_forcing(double dt) {
  update_tracers_levels();
  if (O.ftype == 0) apply_cam_forcing(dt);
  else if (O.ftype == 2) apply_cam_forcing_dynamics(dt);
}
void hxx_diagnostics(int before_advance, int ivar_scalars, int ivar_energy) {
  prim_diag_scalars(before_advance != 0, ivar_scalars);
  prim_energy_halftimes(before_advance != 0, ivar_energy);
}

/* prim_driver.cpp:31-156 */
void prim_run_subcycle_c(const double* dt, int* nstep, int* nm1, int* n0, int* np1, const int* last_time_step) {
  const int nlev = O.nlev;
  if (!O.params_set) runtime_abort("prim_run_subcycle_c: simulation params not set", 13);
  const double dt_q = *dt * O.qsplit;
  double dt_remap = dt_q;
  int nstep_end = O.nstep + O.qsplit;
  if (O.rsplit > 0) {
    dt_remap = dt_q * O.rsplit;
    nstep_end = O.nstep + O.qsplit * O.rsplit;
  }
  /* :51-64 */
  bool compute_diagnostics =
      nstep_end % O.state_frequency == 0 || nstep_end == O.nstep0 || nstep_end >= *last_time_step;
  if (O.disable_diagnostics) compute_diagnostics = false;
  if (compute_diagnostics) {
    prim_diag_scalars(true, 3);
    prim_energy_halftimes(true, 2);
  }
  update_tracers_levels();
  /* :76-82: standalone runs carry ftype = 0 with zero FM/FT/FQ, and still take this pass */
  if (O.ftype == 0) apply_cam_forcing(dt_remap);
  else if (O.ftype == 2) apply_cam_forcing_dynamics(dt_remap);
  if (compute_diagnostics) {
    prim_energy_halftimes(true, 0);
    prim_diag_scalars(true, 0);
  }
  /* dp3d from ps_v :98-111 */
#pragma omp parallel for
  for (int ie = 0; ie < O.nelemd; ++ie) {
    double* dp = DPFLD(ie, O.n0);
    for (int p = 0; p < NPSQ; ++p)
      for (int k = 0; k < nlev; ++k)
        dp[IX(p, k)] = O.dai[k] * O.ps0 + O.dbi[k] * O.ps_v[((size_t)ie * NTL + O.n0) * NPSQ + p];
  }
  prim_step(*dt);
  for (int r = 1; r < O.rsplit; ++r) {
    update_dynamics_levels();
    prim_step(*dt);
  }
  update_tracers_levels();
  hxx_vertical_remap(O.np1, O.np1_qdp, dt_remap);
  hxx_update_q(O.np1_qdp, O.np1);
  if (compute_diagnostics) {
    prim_diag_scalars(false, 1);
    prim_energy_halftimes(false, 1);
  }
  update_dynamics_levels();
  *nstep = O.nstep; *nm1 = O.nm1; *n0 = O.n0; *np1 = O.np1;
}

/* ------------------------------------------------------------------------------------------ */
/* Test hooks                                                                                  */
/* ------------------------------------------------------------------------------------------ */
void hxx_exchange(const char* field_set, int rspheremp) {
  const int nq = O.qsize;
  if (!strncmp(field_set, "caar:", 5)) {
    const int tl = atoi(field_set + 5);
    const size_t es_v = NTL * 2 * NLF, es = NTL * NLF;
    FieldRef f[4] = {{VFLD(0, tl, 0), es_v}, {VFLD(0, tl, 1), es_v}, {TFLD(0, tl), es}, {DPFLD(0, tl), es}};
    exchange(f, 4, rspheremp);
  } else if (!strcmp(field_set, "hv")) {
    hv_exchange(rspheremp);
  } else if (!strncmp(field_set, "euler:", 6)) {
    int tq = 0, opt = 0;
    sscanf(field_set + 6, "%d:%d", &tq, &opt);
    FieldRef* f = (FieldRef*)malloc((nq + 1) * sizeof(FieldRef));
    for (int q = 0; q < nq; ++q) { f[q].base = QDPFLD(0, tq, q); f[q].estride = (size_t)QNTL * O.qsize_d * NLF; }
    f[nq].base = dss_var(opt); f[nq].estride = NLF;
    exchange(f, nq + 1, rspheremp);
    free(f);
  } else if (!strcmp(field_set, "qtens")) {
    FieldRef* f = (FieldRef*)malloc(nq * sizeof(FieldRef));
    for (int q = 0; q < nq; ++q) { f[q].base = O.qtens_biharmonic + (size_t)q * NLF; f[q].estride = (size_t)O.qsize_d * NLF; }
    exchange(f, nq, rspheremp);
    free(f);
  } else if (!strcmp(field_set, "qlim")) {
    exchange_min_max();
  } else {
    runtime_abort("hxx_exchange: unknown field set", 11);
  }
}

static double* field_by_name(const char* name, size_t* n) {
  const size_t ne = (size_t)O.nelemd, f3 = ne * NLF;
  struct { const char* nm; double* p; size_t n; } tab[] = {
      {"v", O.v, f3 * NTL * 2}, {"t", O.t, f3 * NTL}, {"dp3d", O.dp3d, f3 * NTL}, {"ps_v", O.ps_v, ne * NTL * NPSQ},
      {"phi", O.phi, f3}, {"omega_p", O.omega_p, f3}, {"eta_dot_dpdn", O.eta_dot_dpdn, f3},
      {"derived_vn0", O.derived_vn0, f3 * 2}, {"derived_dp", O.derived_dp, f3}, {"divdp", O.divdp, f3},
      {"divdp_proj", O.divdp_proj, f3}, {"dpdiss_ave", O.dpdiss_ave, f3},
      {"dpdiss_biharmonic", O.dpdiss_biharmonic, f3}, {"qdp", O.qdp, f3 * QNTL * O.qsize_d},
      {"qtens_biharmonic", O.qtens_biharmonic, f3 * O.qsize_d}, {"qlim", O.qlim, ne * O.qsize_d * 2 * O.nlev},
      {"Q", O.Q, f3 * O.qsize_d}, {"vtens", O.vtens, f3 * 2}, {"ttens", O.ttens, f3}, {"dptens", O.dptens, f3},
      {"vstar", O.vstar, f3 * 2}, {"dpdissk", O.dpdissk, f3}, {"dp_star", O.dp_star, f3},
      {"fm", O.fm, f3 * 2}, {"ft", O.ft, f3}, {"fq", O.fq, f3 * O.qsize_d}};
  for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i)
    if (!strcmp(tab[i].nm, name)) { *n = tab[i].n; return tab[i].p; }
  *n = 0;
  return NULL;
}

int64_t hxx_get_field(const char* name, double* out) {
  size_t n;
  double* p = field_by_name(name, &n);
  if (p && out) memcpy(out, p, n * sizeof(double));
  return (int64_t)n;
}
int64_t hxx_set_field(const char* name, const double* in) {
  size_t n;
  double* p = field_by_name(name, &n);
  if (p && in) memcpy(p, in, n * sizeof(double));
  return (int64_t)n;
}

void hxx_sphere_op(const char* op, int ie, const double* in, double* out, double nu_ratio) {
  const int nlev = O.nlev;
  if (!strcmp(op, "gradient_sphere")) gradient_sphere(ie, in, out, out + NLF, nlev);
  else if (!strcmp(op, "divergence_sphere")) divergence_sphere(ie, in, in + NLF, out, nlev);
  else if (!strcmp(op, "vorticity_sphere")) vorticity_sphere(ie, in, in + NLF, out, nlev);
  else if (!strcmp(op, "laplace_simple")) laplace_simple(ie, in, out, nlev);
  else if (!strcmp(op, "divergence_sphere_wk")) {
    double* tmp = (double*)malloc(2 * NLF * sizeof(double));
    memcpy(tmp, in, 2 * NLF * sizeof(double));
    divergence_sphere_wk(ie, tmp, tmp + NLF, out, nlev);
    free(tmp);
  } else if (!strcmp(op, "vlaplace_sphere_wk_contra"))
    vlaplace_sphere_wk_contra(ie, nu_ratio, in, in + NLF, out, out + NLF, nlev);
  else runtime_abort("hxx_sphere_op: unknown operator", 11);
}
