/* hommexx_b200 — C ABI of the B200-native preqx dycore.
 *
 * Drop-in boundary: every symbol in section A is exactly what HOMME's Fortran driver binds
 * through iso_c_binding (interfaces in src/prim_main.F90:55-82,
 * src/share/prim_driver_mod.F90:586-697 and src/share/prim_cxx_driver_mod.F90:20-60 of the
 * reference). gfortran passes every bind(c) argument by reference, so each C++ `const T&`
 * of the reference is a `const T*` here and each `CF90Ptr&` (const double* const&) is a
 * `const double* const*`. Plain pointers and sizes only; no torch / CUDA types.
 *
 * Array shapes are the Fortran memory layouts read as C row-major (outer -> inner), with
 * np = 4, nlev = hommexx_b200_nlev(), QSIZE_D = hommexx_b200_qsize_d() (compile-time in the
 * reference too: -DPLEV, -DQSIZE_D, src/preqx/CMakeLists.txt:243-261):
 *   D, Dinv, metinv, tensorvisc  [nelemd][2][2][np][np]      vec_sph2cart [nelemd][2][3][np][np]
 *   fcor, mp, spheremp, rspheremp, metdet, phis               [nelemd][np][np]
 *   v [nelemd][3][nlev][2][np][np]    T, dp3d [nelemd][3][nlev][np][np]    ps_v [nelemd][3][np][np]
 *   Qdp [nelemd][2][QSIZE_D][nlev][np][np]   Q [nelemd][QSIZE_D][nlev][np][np]
 *   omega_p [nelemd][nlev][np][np]
 *
 * Errors follow the reference (src/share/cxx/ErrorDefs.cpp:23-27): message on stderr, session
 * finalised, process aborted with the reference's codes (11/12/13/101). There are no return
 * codes on the section-A symbols because the Fortran interfaces have none.
 */
#ifndef HOMMEXX_B200_H
#define HOMMEXX_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * A. The reference's Fortran-facing entry points (same names, same argument meaning).
 * ---------------------------------------------------------------------------------------- */

/* mpi/mpi_cxx_f90_interface.cpp:20 — first C call; f_comm is a Fortran MPI handle. This build
 * has no MPI: the handle is recorded, and the rank/size wiring comes from
 * hommexx_b200_set_comm() (section B). Single-rank runs need neither. */
void reset_cxx_comm(const int* f_comm);
/* Hommexx_Session.cpp:35 / :75 */
void initialize_hommexx_session(void);
void finalize_hommexx_session(void);
/* mpi/mpi_cxx_f90_interface.cpp:27,34,57 — all indices 1-based; pos 1..8 = W,E,S,N,SW,SE,NW,NE */
void init_connectivity(const int* num_local_elems);
void add_connection(const int* first_elem_lid, const int* first_elem_gid, const int* first_elem_pos,
                    const int* first_elem_pid, const int* second_elem_lid, const int* second_elem_gid,
                    const int* second_elem_pos, const int* second_elem_pid);
void finalize_connectivity(void);
/* cxx_f90_interface.cpp:207 — dvv[np][np] */
void init_derivative_c(const double* const* dvv);
/* cxx_f90_interface.cpp:33 — validates the option set exactly as the reference (:43-52) */
void init_simulation_params_c(const int* remap_alg, const int* limiter_option, const int* rsplit,
                              const int* qsplit, const int* time_step_type, const int* energy_fixer,
                              const int* qsize, const int* state_frequency, const double* nu,
                              const double* nu_p, const double* nu_q, const double* nu_s,
                              const double* nu_div, const double* nu_top, const int* hypervis_order,
                              const int* hypervis_subcycle, const double* hypervis_scaling,
                              const int* ftype, const bool* prescribed_wind, const bool* moisture,
                              const bool* disable_diagnostics, const bool* use_cpstar,
                              const bool* use_semi_lagrangian_transport);
/* cxx_f90_interface.cpp:224 */
void init_elements_2d_c(const int* num_elems, const double* const* D, const double* const* Dinv,
                        const double* const* fcor, const double* const* mp,
                        const double* const* spheremp, const double* const* rspheremp,
                        const double* const* metdet, const double* const* metinv,
                        const double* const* phis, const double* const* tensorvisc,
                        const double* const* vec_sph2cart, const bool* consthv);
/* cxx_f90_interface.cpp:235 */
void init_elements_states_c(const double* const* elem_state_v, const double* const* elem_state_temp,
                            const double* const* elem_state_dp3d, const double* const* elem_state_Qdp,
                            const double* const* elem_state_ps_v);
/* cxx_f90_interface.cpp:253 — pointers are retained (diagnostics write into them later) */
void init_diagnostics_c(double* const* elem_state_q, double* const* elem_accum_qvar,
                        double* const* elem_accum_qmass, double* const* elem_accum_q1mass,
                        double* const* elem_accum_iener, double* const* elem_accum_iener_wet,
                        double* const* elem_accum_kener, double* const* elem_accum_pener);
/* cxx_f90_interface.cpp:119 — only hyai/hybi (nlev+1 each) are used, as in HybridVCoord.cpp:15-53 */
void init_hvcoord_c(const double* ps0, const double* const* hybrid_am, const double* const* hybrid_ai,
                    const double* const* hybrid_bm, const double* const* hybrid_bi);
/* cxx_f90_interface.cpp:264 */
void init_boundary_exchanges_c(void);
/* cxx_f90_interface.cpp:213 — 1-based in, stored 0-based */
void init_time_level_c(const int* nm1, const int* n0, const int* np1, const int* nstep,
                       const int* nstep0);
/* prim_driver.cpp:31 — the timestep; nstep/nm1/n0/np1 are outputs (0-based time levels) */
void prim_run_subcycle_c(const double* dt, int* nstep, int* nm1, int* n0, int* np1,
                         const int* last_time_step);
/* cxx_f90_interface.cpp:126 — device -> Fortran layout, all time levels */
void cxx_push_results_to_f90(double* const* elem_state_v, double* const* elem_state_temp,
                             double* const* elem_state_dp3d, double* const* elem_state_Qdp,
                             double* const* elem_Q, double* const* elem_state_ps_v,
                             double* const* elem_derived_omega_p);
/* cxx_f90_interface.cpp:180 / :157 — CAM coupling; arguments are by VALUE in the reference */
void f90_push_forcing_to_cxx(double* elem_derived_FM, double* elem_derived_FT, double* elem_derived_FQ,
                             double* elem_state_Qdp);
void cxx_push_forcing_to_f90(double* elem_derived_FM, double* elem_derived_FT, double* elem_derived_FQ);

/* ------------------------------------------------------------------------------------------
 * B. Build information and multi-GPU wiring (replaces what MPI provided to the reference).
 * ---------------------------------------------------------------------------------------- */
int hommexx_b200_nlev(void);     /* PLEV of this build */
int hommexx_b200_qsize_d(void);  /* QSIZE_D of this build */
/* "cuda-sm100a" for the product, "cpu-oracle" for oracle/liboracle.so */
const char* hommexx_b200_backend(void);
/* One process per GPU. rank/size as in MPI_Comm_rank/size; device = CUDA device ordinal;
 * nccl_unique_id = the 128-byte ncclUniqueId created on rank 0 and broadcast by the host
 * (torch.distributed / MPI_Bcast). Must be called BEFORE initialize_hommexx_session (the session
 * binds the device and creates its streams there); a call on an active session aborts with code 13. */
void hommexx_b200_set_comm(int rank, int size, int device, const void* nccl_unique_id);
/* Fills out128 with a fresh ncclUniqueId (call on rank 0, broadcast, pass to set_comm).
 * Returns 0 on success, nonzero if NCCL is unavailable in this build. */
int hommexx_b200_nccl_unique_id(void* out128);
/* Number of kernels this library has launched since session start (bench's gpu_launches). */
int64_t hommexx_b200_launch_count(void);
/* Blocks until all device work issued so far has completed. */
void hommexx_b200_sync(void);
/* CUDA-event timing on the library's own launch stream (torch.cuda.Event only sees torch's
 * stream): record into slot 0..15, elapsed milliseconds between two recorded slots. */
void hommexx_b200_event_record(int slot);
double hommexx_b200_event_elapsed_ms(int slot_a, int slot_b);
/* Per-kernel probes: bit i of mask selects kernel class i (hommexx_b200_kernel_id("euler_advect"),
 * ...); every launch of a selected class is bracketed by a CUDA-event pair. Calling it again
 * resets the per-class launch counters and the recorded pairs. profile_read returns the summed
 * device milliseconds of one class and its launch count since the last reset. */
void hommexx_b200_profile(unsigned long long mask);
int hommexx_b200_kernel_id(const char* name);
const char* hommexx_b200_kernel_name(int id);
double hommexx_b200_profile_read(int id, int64_t* launches);

/* ------------------------------------------------------------------------------------------
 * C. Phase-level entry points: the public run methods of the reference's functors, exposed
 *    so each phase can be checked on its own (the reference's unit tests call the C++ functor
 *    methods directly, src/preqx/unit_tests/preqx_ut.cpp). Time levels are 0-based.
 * ---------------------------------------------------------------------------------------- */
/* CaarFunctor::run, CaarFunctor.cpp:96-116. with_dss=0 stops before the boundary exchange. */
void hxx_caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, int with_dss);
/* RK combine of u3_5stage_timestep, prim_advance_exp.cpp:143-154 */
void hxx_rk_combine(int nm1, int n0);
/* HyperviscosityFunctorImpl::run, HyperviscosityFunctorImpl.cpp:56-85 */
void hxx_hypervis_run(int np1, double dt, double eta_ave_w);
/* EulerStepFunctor::{reset,precompute_divdp,euler_step,qdp_time_avg} */
void hxx_euler_reset(void);
void hxx_euler_precompute_divdp(void);
void hxx_euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt);
void hxx_euler_qdp_time_avg(int n0_qdp, int np1_qdp);
/* VerticalRemapManager::run_remap, RemapFunctor.hpp:306-331 */
void hxx_vertical_remap(int np1, int np1_qdp, double dt);
/* update_q, prim_driver.cpp:171-206 */
void hxx_update_q(int np1_qdp, int np1);
/* prim_step's zeroing kernel, prim_step.cpp:51-66 */
void hxx_prim_step_init(int n0);
/* apply_cam_forcing (ftype 0) / apply_cam_forcing_dynamics (ftype 2), CamForcing.cpp:149-174, on
 * time level n0 / n0_qdp with the FM, FT, FQ last pushed by f90_push_forcing_to_cxx */
void hxx_apply_forcing(double dt);
/* Held-Suarez (1994) forcing evaluated where the state lives (hs_forcing, hs_T_forcing, hs_v_forcing of
 * physics/heldsuarez/held_suarez_mod.F90:36-279, dry part): FM, FT of the session are OVERWRITTEN with the
 * Newtonian relaxation of T(n0) towards T_eq(lat, p) and the Rayleigh friction on v(n0) below sigma_b = 0.7,
 * FQ is left alone (never pushed = zero). The GPU-resident equivalent of the Fortran physics followed by
 * f90_push_forcing_to_cxx, without the host round trip of every tracer; the next prim_run_subcycle_c applies it
 * (ftype = 0). lat = [nelemd][np][np] latitudes, read on the first call of a session (may be null afterwards);
 * hyam, hybm = the [nlev] mid-level coefficients init_hvcoord_c was given. */
void hxx_held_suarez_forcing(const double* lat, const double* hyam, const double* hybm);
/* Diagnostics::prim_diag_scalars + prim_energy_halftimes, Diagnostics.cpp:37-185, into the arrays
 * registered with init_diagnostics_c */
void hxx_diagnostics(int before_advance, int ivar_scalars, int ivar_energy);
/* BoundaryExchange::exchange on a named field set: "caar:<tl>", "hv", "euler:<np1_qdp>:<dssopt>",
 * "qtens", and exchange_min_max for "qlim". rspheremp!=0 applies the inverse mass afterwards. */
void hxx_exchange(const char* field_set, int rspheremp);

/* Copy a named device array out/in, in the level-innermost device layout
 * [nelemd][...][np][np][nlev]; returns the number of doubles (0 if the name is unknown).
 * Names: v t dp3d ps_v phi omega_p eta_dot_dpdn derived_vn0 derived_dp divdp divdp_proj
 *        dpdiss_ave dpdiss_biharmonic qdp qtens_biharmonic qlim Q vtens ttens dptens
 *        vstar dpdissk dp_star fm ft fq */
int64_t hxx_get_field(const char* name, double* out);
int64_t hxx_set_field(const char* name, const double* in);

/* Element-local operator checks (SphereOperators.hpp) on caller-provided [np][np][nlev] data of
 * element `ie`; used by the known-answer tests. op in {"gradient_sphere","divergence_sphere",
 * "vorticity_sphere","laplace_simple","vlaplace_sphere_wk_contra","divergence_sphere_wk"}.
 * in has n_in fields, out n_out fields of np*np*nlev doubles. */
void hxx_sphere_op(const char* op, int ie, const double* in, double* out, double nu_ratio);
/* limiter_optim_iter_full / limiter_clip_and_sum (EulerStepFunctorImpl.hpp:766-884) on
 * nsets independent problems laid out [set][np*np][nlev]; qlim is [set][2][nlev] (updated). */
void hxx_limiter(int limiter_option, int nsets, const double* sphweights /*[set][16]*/,
                 const double* dpmass, double* ptens, double* qlim);
/* PPM remap of nfields fields of one element-set: remap_Q_ppm semantics
 * (src/share/vertremap_mod_base.F90:524-643; PpmRemap.hpp). Layout [ncol][nlev] columns. */
void hxx_remap_columns(int alg, int ncols, int nfields, const double* src_dp, const double* tgt_dp,
                       double* fields /*[nfields][ncols][nlev]*/);

#ifdef __cplusplus
}
#endif
#endif /* HOMMEXX_B200_H */
