// Host-side driver: the stand-in for HOMME's Fortran driver (prim_main / prim_driver_mod).
//
// The reference keeps mesh generation, partitioning, initial conditions and the time loop in
// Fortran (src/prim_main.F90:113-331, src/share/prim_driver_mod.F90:185-1105) and calls the
// dycore through 17 extern "C" symbols. There is no Fortran toolchain in this image, so this
// C++ harness restates the slice of that infrastructure the hot path needs as input:
//   - equi-angular cubed-sphere mesh, metric terms, mass matrix   (src/share/cube_mod.F90:215-704,
//     src/share/mass_matrix_mod.F90:28-124, src/share/prim_driver_mod.F90:393-419)
//   - GLL quadrature and derivative matrix                         (src/share/quadrature_mod.F90,
//     src/share/derivative_mod_base.F90:440-475)
//   - element connectivity as add_connection() tuples              (src/share/prim_cxx_driver_mod.F90:43-140)
//   - space-filling-curve partition                                (src/share/spacecurve_mod.F90:1218-1273)
//   - Jablonowski-Williamson baroclinic initial state              (src/test_src/baroclinic_inst_mod.F90:56-271)
// and then drives ANY library that exports the reference's C-ABI (include/hommexx_b200.h):
// the CUDA product (libhommexx_b200_*.so) or the CPU oracle (oracle/liboracle.so).
// All arrays are held in the Fortran memory layout the ABI expects.
#pragma once
#include <cstdint>

extern "C" {

struct HommeDriver;

struct HommeParams {
  // mesh / sizes
  int ne, nlev, qsize_d, qsize;
  int npart, part_id;             // SFC partition: this rank owns part `part_id` of `npart`
  // ctl_nl namelist values forwarded through init_simulation_params_c
  int remap_alg, limiter_option, rsplit, qsplit, time_step_type, energy_fixer, state_frequency;
  double nu, nu_p, nu_q, nu_s, nu_div, nu_top;
  int hypervis_order, hypervis_subcycle;
  double hypervis_scaling;
  int ftype;
  int prescribed_wind, moisture, disable_diagnostics, use_cpstar, use_semi_lagrangian_transport;
  double tstep;
  double u_perturb;               // JW perturbation amplitude (namelist u_perturb)
};

// Build mesh + partition + geometry for this rank. hyai/hybi have nlev+1 entries, hyam/hybm nlev.
HommeDriver* hd_create(const HommeParams* p, const double* hyai, const double* hybi,
                       const double* hyam, const double* hybm);
void hd_destroy(HommeDriver* h);

// Jablonowski-Williamson baroclinic wave initial state (also fills Qdp).
void hd_init_jw(HommeDriver* h);

// dlopen a dycore library exporting the reference C-ABI. Returns 0 on success; on failure
// returns nonzero and hd_last_error() describes the missing library/symbol.
int hd_bind(HommeDriver* h, const char* libpath);
const char* hd_last_error();

// Calls the init_* entry points in the reference's order (prim_driver_mod.F90:1043-1104).
void hd_init_dycore(HommeDriver* h);
// (Re-)upload the Fortran-side state: init_elements_states_c + init_time_level_c.
void hd_upload_state(HommeDriver* h);
// One prim_run_subcycle_c call (= qsplit*max(rsplit,1) dynamics steps). Returns nstep after it.
int hd_run_subcycle(HommeDriver* h);
// cxx_push_results_to_f90 into the driver's Fortran-layout arrays.
void hd_push_results(HommeDriver* h);
// CAM-coupling calls of the Fortran prim_run_subcycle wrapper (prim_driver_mod.F90:1380,1402):
// f90_push_forcing_to_cxx(FM, FT, FQ, Qdp) and cxx_push_forcing_to_f90(FM, FT, FQ).
void hd_push_forcing(HommeDriver* h);
void hd_pull_forcing(HommeDriver* h);
// nEndStep passed to prim_run_subcycle_c as last_time_step (default: never)
void hd_set_last_step(HommeDriver* h, int nEndStep);
void hd_finalize_dycore(HommeDriver* h);

// Introspection for tests/bench (pointers stay valid until hd_destroy).
int hd_nelemd(const HommeDriver* h);
int hd_nelem_global(const HommeDriver* h);
// name in {D,Dinv,fcor,mp,spheremp,rspheremp,metdet,metinv,phis,v,T,dp3d,Qdp,Q,ps_v,omega_p,dvv,
//          lat,lon,gid,FM,FT,FQ,Qvar,Qmass,Q1mass,IEner,IEner_wet,KEner,PEner}; returns element count of the array through *n.
double* hd_array(HommeDriver* h, const char* name, int64_t* n);
// connectivity tuples as passed to add_connection: 8 ints each (all 1-based); returns count.
int hd_connections(const HommeDriver* h, const int** tuples);
void hd_time_levels(const HommeDriver* h, int* nstep, int* nm1, int* n0, int* np1);
// Global ids (0-based, SFC-ordered within the rank) of the local elements.
const int* hd_local_gids(const HommeDriver* h);
// rank that owns each global element (size nelem_global)
const int* hd_owner(const HommeDriver* h);

}  // extern "C"
