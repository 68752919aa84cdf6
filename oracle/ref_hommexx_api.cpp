// Section B of include/hommexx_b200.h for the REFERENCE library: the few build-information symbols the host driver
// (hommexx_b200/driver) asks every dycore library for, so that oracle/_ref/libref_hommexx_<PLEV>_<QSIZE_D>.so — the
// reference's own src/share/cxx sources compiled against the serial Kokkos stand-in of oracle/ref_shim — binds to the
// same driver as the product and the oracle. Everything in section A comes from the reference's own
// cxx_f90_interface.cpp / prim_driver.cpp / mpi_cxx_f90_interface.cpp. TEST INFRASTRUCTURE.
//
// Section C (phase-level hooks) is bound to the public run methods of the reference's functors, so that the oracle
// can be pinned PER FUNCTOR as well as per run: hxx_caar_run -> CaarFunctor::run, hxx_hypervis_run ->
// HyperviscosityFunctor::run, hxx_euler_* -> EulerStepFunctor::{reset,precompute_divdp,euler_step,qdp_time_avg},
// hxx_vertical_remap -> VerticalRemapManager::run_remap, hxx_update_q -> update_q (prim_driver.cpp); and
// hxx_get_field / hxx_set_field copy the reference's Views out / in. With HOMMEXX_VECTOR_SIZE = 1 a View
// Scalar*[..][NP][NP][NUM_LEV] is exactly the level-innermost [nelemd][..][np][np][nlev] layout of the ABI.
#include <cstdint>
#ifdef REF_API_PHASE_HOOKS  // the serial VECTOR_SIZE = 1 builds only
#include <cstring>
#include <string>

#include "CaarFunctor.hpp"
#include "CamForcing.hpp"
#include "Diagnostics.hpp"
#include "Context.hpp"
#include "Elements.hpp"
#include "EulerStepFunctor.hpp"
#include "EulerStepFunctorImpl.hpp"
#include "HyperviscosityFunctor.hpp"
#include "Derivative.hpp"
#include "KernelVariables.hpp"
#include "SimulationParams.hpp"
#include "SphereOperators.hpp"
#include "TimeLevel.hpp"
#include "Tracers.hpp"
#include "VerticalRemapManager.hpp"

namespace Homme {
void update_q(const int np1_qdp, const int np1);
}

namespace {
using namespace Homme;
static_assert(sizeof(Scalar) == sizeof(double), "the field hooks need the scalar (VECTOR_SIZE = 1) build");

struct Span {
  double* p = nullptr;
  size_t n = 0;
};
template <typename V>
Span span_of(const V& v) {
  return Span{reinterpret_cast<double*>(v.data()), static_cast<size_t>(v.size())};
}

Span ref_field(const char* name) {
  Elements& e = Context::singleton().get_elements();
  Tracers& t = Context::singleton().get_tracers();
  auto is = [&](const char* s) { return std::strcmp(name, s) == 0; };
  if (is("v")) return span_of(e.m_v);
  if (is("t")) return span_of(e.m_t);
  if (is("dp3d")) return span_of(e.m_dp3d);
  if (is("ps_v")) return span_of(e.m_ps_v);
  if (is("phi")) return span_of(e.m_phi);
  if (is("omega_p")) return span_of(e.m_omega_p);
  if (is("eta_dot_dpdn")) return span_of(e.m_eta_dot_dpdn);
  if (is("derived_vn0")) return span_of(e.m_derived_vn0);
  if (is("derived_dp")) return span_of(e.m_derived_dp);
  if (is("divdp")) return span_of(e.m_derived_divdp);
  if (is("divdp_proj")) return span_of(e.m_derived_divdp_proj);
  if (is("dpdiss_ave")) return span_of(e.m_derived_dpdiss_ave);
  if (is("dpdiss_biharmonic")) return span_of(e.m_derived_dpdiss_biharmonic);
  if (is("vtens")) return span_of(e.buffers.vtens);
  if (is("ttens")) return span_of(e.buffers.ttens);
  if (is("dptens")) return span_of(e.buffers.dptens);
  if (is("vstar")) return span_of(e.buffers.vstar);
  if (is("dpdissk")) return span_of(e.buffers.dpdissk);
  if (is("fm")) return span_of(e.m_fm);
  if (is("ft")) return span_of(e.m_ft);
  if (is("qdp")) return span_of(t.qdp);
  if (is("qtens_biharmonic")) return span_of(t.qtens_biharmonic);
  if (is("qlim")) return span_of(t.qlim);
  if (is("Q")) return span_of(t.Q);
  if (is("fq")) return span_of(t.fq);
  return Span{};
}
}  // namespace
#endif

extern "C" {
#ifdef REF_API_PHASE_HOOKS
int64_t hxx_get_field(const char* name, double* out) {
  const Span s = ref_field(name);
  if (s.p && out) std::memcpy(out, s.p, s.n * sizeof(double));
  return static_cast<int64_t>(s.n);
}
int64_t hxx_set_field(const char* name, const double* in) {
  const Span s = ref_field(name);
  if (s.p && in) std::memcpy(s.p, in, s.n * sizeof(double));
  return static_cast<int64_t>(s.n);
}
void hxx_caar_run(int nm1, int n0, int np1, double dt, double eta_ave_w, int n0_qdp, int with_dss) {
  CaarFunctor& f = Context::singleton().get_caar_functor();
  f.set_n0_qdp(n0_qdp);
  if (with_dss) {
    f.run(nm1, n0, np1, dt, eta_ave_w, false);
  } else {
    f.set_rk_stage_data(nm1, n0, np1, dt, eta_ave_w, false);
    f.run();
  }
}
void hxx_hypervis_run(int np1, double dt, double eta_ave_w) {
  Context::singleton().get_hyperviscosity_functor().run(np1, dt, eta_ave_w);
}
void hxx_euler_reset(void) {
  Context::singleton().get_euler_step_functor().reset(Context::singleton().get_simulation_params());
}
void hxx_euler_precompute_divdp(void) { Context::singleton().get_euler_step_functor().precompute_divdp(); }
void hxx_euler_step(int np1_qdp, int n0_qdp, double dt, double rhs_multiplier, int dss_opt) {
  const DSSOption opt = dss_opt == 0 ? DSSOption::ETA : dss_opt == 1 ? DSSOption::OMEGA : DSSOption::DIV_VDP_AVE;
  Context::singleton().get_euler_step_functor().euler_step(np1_qdp, n0_qdp, dt, rhs_multiplier, opt);
}
void hxx_euler_qdp_time_avg(int n0_qdp, int np1_qdp) {
  Context::singleton().get_euler_step_functor().qdp_time_avg(n0_qdp, np1_qdp);
}
void hxx_vertical_remap(int np1, int np1_qdp, double dt) {
  Context::singleton().get_vertical_remap_manager().run_remap(np1, np1_qdp, dt);
}
void hxx_update_q(int np1_qdp, int np1) { Homme::update_q(np1_qdp, np1); }
// apply_cam_forcing / apply_cam_forcing_dynamics (CamForcing.cpp:149-174) the way prim_run_subcycle_c calls them
// (prim_driver.cpp:66-82): tracer time levels from nstep first, then the pass the namelist's ftype selects
void hxx_apply_forcing(double dt) {
  SimulationParams& params = Context::singleton().get_simulation_params();
  Context::singleton().get_time_level().update_tracers_levels(params.qsplit);
  if (params.ftype == ForcingAlg::FORCING_DEBUG) Homme::apply_cam_forcing(dt);
  else if (params.ftype == ForcingAlg::FORCING_2) Homme::apply_cam_forcing_dynamics(dt);
}
// Diagnostics::prim_diag_scalars + prim_energy_halftimes (Diagnostics.cpp:37-185) into the registered F90 arrays
void hxx_diagnostics(int before_advance, int ivar_scalars, int ivar_energy) {
  Diagnostics& d = Context::singleton().get_diagnostics();
  d.prim_diag_scalars(before_advance != 0, ivar_scalars);
  d.prim_energy_halftimes(before_advance != 0, ivar_energy);
}
// One element-local operator of SphereOperators.hpp on caller-provided [np][np][nlev] fields of element `ie`
// (a team per element as in the functors; only the team of `ie` works). Returns silently on an unknown name.
void hxx_sphere_op(const char* op, int ie, const double* in, double* out, double nu_ratio) {
  Elements& e = Context::singleton().get_elements();
  SphereOperators sph(e, Context::singleton().get_derivative());
  const auto policy = Homme::get_default_team_policy<ExecSpace>(e.num_elems());
  sph.allocate_buffers(policy);
  ExecViewManaged<Scalar[2][NP][NP][NUM_LEV]> vin("sphere_op in"), vout("sphere_op out");
  const size_t F = NP * NP * NUM_LEV;
  const std::string name(op);
  const int n_in = (name == "gradient_sphere" || name == "laplace_simple") ? 1 : 2;
  const int n_out = (name == "gradient_sphere" || name == "vlaplace_sphere_wk_contra") ? 2 : 1;
  std::memcpy(vin.data(), in, n_in * F * sizeof(double));
  Kokkos::parallel_for(policy, KOKKOS_LAMBDA(const TeamMember& team) {
    KernelVariables kv(team);
    if (kv.ie != ie) return;
    if (name == "gradient_sphere") sph.gradient_sphere(kv, Homme::subview(vin, 0), vout);
    else if (name == "divergence_sphere") sph.divergence_sphere(kv, vin, Homme::subview(vout, 0));
    else if (name == "vorticity_sphere")
      sph.vorticity_sphere(kv, Homme::subview(vin, 0), Homme::subview(vin, 1), Homme::subview(vout, 0));
    else if (name == "laplace_simple") sph.laplace_simple(kv, Homme::subview(vin, 0), Homme::subview(vout, 0));
    else if (name == "divergence_sphere_wk") sph.divergence_sphere_wk(kv, vin, Homme::subview(vout, 0));
    else if (name == "vlaplace_sphere_wk_contra") sph.vlaplace_sphere_wk_contra(kv, nu_ratio, vin, vout);
  });
  std::memcpy(out, vout.data(), n_out * F * sizeof(double));
}
// The limiters of EulerStepFunctorImpl.hpp on nsets independent problems [set][np*np][nlev], qlim [set][2][nlev]:
// limiter_option 8 / 9 = what limiter_optim_iter_full(kv) / limiter_clip_and_sum(kv) dispatch to on a host
// execution space (SerialLimiter::run<8|9>, :640-666); 108 / 109 = the team implementations the reference's unit
// tests and its GPU build call (:766-884).
void hxx_limiter(int limiter_option, int nsets, const double* sphweights, const double* dpmass, double* ptens,
                 double* qlim) {
  const size_t F = NP * NP * NUM_LEV;
  ExecViewManaged<Real[NP][NP]> sw("sphweights");
  ExecViewManaged<Scalar[NP][NP][NUM_LEV]> dm("dpmass"), pt("ptens"), wrk("rwrk");
  ExecViewManaged<Scalar[2][NUM_LEV]> ql("qlim");
  const auto policy = Kokkos::TeamPolicy<ExecSpace>(1, 1, 1);
  for (int s = 0; s < nsets; ++s) {
    std::memcpy(sw.data(), sphweights + (size_t)s * NP * NP, NP * NP * sizeof(double));
    std::memcpy(dm.data(), dpmass + s * F, F * sizeof(double));
    std::memcpy(pt.data(), ptens + s * F, F * sizeof(double));
    std::memcpy(ql.data(), qlim + (size_t)s * 2 * NUM_LEV, 2 * NUM_LEV * sizeof(double));
    Kokkos::parallel_for(policy, KOKKOS_LAMBDA(const TeamMember& team) {
      if (limiter_option == 8) SerialLimiter<ExecSpace>::run<8>(sw, dm, ql, pt, wrk);
      else if (limiter_option == 9) SerialLimiter<ExecSpace>::run<9>(sw, dm, ql, pt, wrk);
      else if (limiter_option == 108) EulerStepFunctorImpl::limiter_optim_iter_full(team, sw, dm, ql, pt);
      else if (limiter_option == 109) EulerStepFunctorImpl::limiter_clip_and_sum(team, sw, dm, ql, pt);
    });
    std::memcpy(ptens + s * F, pt.data(), F * sizeof(double));
    std::memcpy(qlim + (size_t)s * 2 * NUM_LEV, ql.data(), 2 * NUM_LEV * sizeof(double));
  }
}
#endif

int hommexx_b200_nlev(void) { return PLEV; }
int hommexx_b200_qsize_d(void) { return QSIZE_D; }
const char* hommexx_b200_backend(void) { return "reference-serial"; }
void hommexx_b200_set_comm(int, int, int, const void*) {}
int64_t hommexx_b200_launch_count(void) { return 0; }
void hommexx_b200_sync(void) {}
}
