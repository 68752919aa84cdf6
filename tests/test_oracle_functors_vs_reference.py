"""THE PIN OF THE ORACLE, PER FUNCTOR: each public run method of the reference's functors against the oracle's
restatement of it, from the same state, every named array compared bit for bit.

The reference library (oracle/_ref/libref_hommexx_*.so, the reference's own sources, tests/reference_lib.py) exports
the phase-level hooks of include/hommexx_b200.h section C bound to the reference's objects (oracle/ref_hommexx_api.cpp):
hxx_caar_run = CaarFunctor::run (CaarFunctor.cpp:96-116), hxx_hypervis_run = HyperviscosityFunctor::run
(HyperviscosityFunctorImpl.cpp:56-85), hxx_euler_* = EulerStepFunctor::{reset, precompute_divdp, euler_step,
qdp_time_avg} (EulerStepFunctorImpl.hpp), hxx_vertical_remap = VerticalRemapManager::run_remap (RemapFunctor.hpp:306-331),
hxx_update_q = update_q (prim_driver.cpp:171-206); hxx_get_field / hxx_set_field copy its Views. The whole-run pin
(test_oracle_vs_reference.py) sees the prognostic arrays after 10 steps; this one sees every functor's own outputs —
derived_vn0, eta_dot_dpdn, omega_p, phi, dpdiss_*, divdp, divdp_proj, qtens_biharmonic, qlim — directly."""
import numpy as np
import pytest

from functor_pair import FIELDS, Pair
from hommexx_b200 import homme
from oracle import oraclelib

FIELDS_WITH_FORCING = {0: FIELDS + ["fm", "ft", "fq"], 2: FIELDS + ["fm", "ft"]}


@pytest.fixture(scope="class")
def ne4():
    """Class scope: closed before the option variants open their own sessions on the same two libraries."""
    p = Pair(homme.preset("ne4"), oraclelib.ORACLE_LIB, b"cpu-oracle")
    yield p
    p.close()


class TestNe4:
    def test_the_field_hooks_see_the_same_state_after_a_run(self, ne4):
        """One more full call on each library from the common state: every array, not only the prognostic ones."""
        ne4.reset()
        # the oracle's time levels are still those of init: bring it to the reference's by running it once, then reset
        if ne4.ho.time_levels() != ne4.hr.time_levels():
            ne4.ho.run_subcycle()
            ne4.reset()
        assert ne4.ho.time_levels() == ne4.hr.time_levels()
        ne4.hr.run_subcycle()
        ne4.ho.run_subcycle()
        ne4.same("prim_run_subcycle_c", changed=("v", "t", "dp3d", "qdp", "derived_vn0", "qlim"))

    @pytest.mark.parametrize("nm1,n0,np1,dt,w,n0_qdp,dss", [
        (1, 1, 0, 360.0, 0.25, -1, 1),   # RK stage 1 shape (accumulates eta_ave_w / 4)
        (1, 0, 2, 360.0, 0.0, -1, 1),    # stage 2
        (1, 2, 2, 600.0, 0.0, -1, 1),    # stages 3 / 4: in place
        (0, 2, 2, 1350.0, 0.75, -1, 1),  # stage 5
        (0, 1, 2, 60.0, 0.25, 1, 1),     # moist: virtual temperature from Qdp(n0_qdp)
        (1, 1, 0, 360.0, 0.25, -1, 0),   # the functor alone, no boundary exchange
    ])
    def test_caar_functor(self, ne4, nm1, n0, np1, dt, w, n0_qdp, dss):
        ne4.reset()
        ne4.call("hxx_caar_run", nm1, n0, np1, dt, w, n0_qdp, dss)
        ne4.same(f"CaarFunctor::run {(nm1, n0, np1, n0_qdp, dss)}", changed=("v", "t", "dp3d", "phi"))

    def test_hyperviscosity_functor(self, ne4):
        ne4.reset()
        ne4.call("hxx_hypervis_run", 2, 1800.0, 1.0)
        ne4.same("HyperviscosityFunctor::run", changed=("v", "t", "dp3d", "dpdiss_ave", "dpdiss_biharmonic"))

    def test_euler_step_functor(self, ne4):
        ne4.reset()
        ne4.call("hxx_euler_reset")
        for h in (ne4.hr, ne4.ho):  # the snapshot already holds this state's divdp: wipe it to see the functor write it
            h.set_field("divdp", np.zeros_like(ne4.snap["divdp"]))
            h.set_field("divdp_proj", np.zeros_like(ne4.snap["divdp"]))
        ne4.call("hxx_euler_precompute_divdp")
        ne4.same("precompute_divdp")
        assert np.array_equal(ne4.hr.get_field("divdp"), ne4.snap["divdp"])
        for (np1q, n0q, rhs, opt) in [(1, 0, 0.0, 2), (1, 1, 1.0, 0), (1, 1, 2.0, 1)]:
            ne4.call("hxx_euler_step", np1q, n0q, 900.0, rhs, opt)
            ne4.same(f"euler_step rhs_multiplier={rhs}", changed=("qdp",))
        ne4.call("hxx_euler_qdp_time_avg", 0, 1)
        ne4.same("qdp_time_avg")

    def test_remap_functor_and_update_q(self, ne4):
        ne4.reset()
        # a forward step of level 1 into level 2 first: Lagrangian layers that really left the reference grid
        ne4.call("hxx_caar_run", 1, 1, 2, 1800.0, 1.0, -1, 1)
        before = {n: ne4.hr.get_field(n) for n in ("qdp", "t", "v")}  # (dp3d(np1) is left Lagrangian, as the reference leaves it)
        ne4.call("hxx_vertical_remap", 2, 1, 5400.0)
        ne4.same("VerticalRemapManager::run_remap")
        for n, a in before.items():
            assert not np.array_equal(ne4.hr.get_field(n), a), (n, "not remapped")
        ne4.call("hxx_update_q", 1, 2)
        ne4.same("update_q")


VARIANTS = {
    # limiter 9 + PPM mirrored boundaries at the reference's prtcA sizes
    "prtcA-lim9-alg2": dict(base="prtcA", limiter_option=9, remap_alg=2),
    # Eulerian vertical coordinate: CAAR with vertical advection, tracer-only remap
    "ne4-r0-moist": dict(base="ne4", rsplit=0, moisture=1),
    # tensor hyperviscosity with nu_div != nu
    "prtcA-tensorhv": dict(base="prtcA", hypervis_scaling=3.2, nu=1e-9, nu_p=1e-9, nu_q=1e-9, nu_s=1e-9, nu_div=2.5e-9,
                           hypervis_subcycle=2),
    # the benchmarked dimensions, 40 distinct tracers
    "ne4-q40": dict(base="ne4", qsize=40, qsize_d=40),
}


@pytest.mark.parametrize("case", list(VARIANTS))
def test_every_functor_in_the_option_variants(case):
    over = dict(VARIANTS[case])
    cfg = homme.preset(over.pop("base"), **over)
    p = Pair(cfg, oraclelib.ORACLE_LIB, b"cpu-oracle")
    try:
        moist = 1 if cfg.moisture else -1
        p.reset()
        p.call("hxx_caar_run", 1, 1, 0, cfg.tstep / 5.0, 0.25, moist, 1)
        p.same(case + " caar stage 1", changed=("v", "t", "dp3d"))
        p.call("hxx_caar_run", 0, 2, 2, 0.75 * cfg.tstep, 0.75, moist, 1)
        p.same(case + " caar stage 5")
        p.call("hxx_hypervis_run", 2, cfg.tstep, 1.0)
        p.same(case + " hypervis", changed=("v", "t"))
        p.reset()
        p.call("hxx_euler_reset")
        p.call("hxx_euler_precompute_divdp")
        dtq = cfg.tstep * cfg.qsplit
        for (np1q, n0q, rhs, opt) in [(1, 0, 0.0, 2), (1, 1, 1.0, 0), (1, 1, 2.0, 1)]:
            p.call("hxx_euler_step", np1q, n0q, dtq / 2.0, rhs, opt)
            p.same(f"{case} euler_step rhs_multiplier={rhs}", changed=("qdp",))
        p.call("hxx_euler_qdp_time_avg", 0, 1)
        p.same(case + " qdp_time_avg")
        p.reset()
        p.call("hxx_caar_run", 1, 1, 2, cfg.tstep, 1.0, moist, 1)
        p.call("hxx_vertical_remap", 2, 1, dtq * max(cfg.rsplit, 1))
        p.call("hxx_update_q", 1, 2)
        p.same(case + " remap + update_q", changed=("qdp",))
    finally:
        p.close()


# ---- CamForcing.cpp and Diagnostics.cpp, one call at a time ---------------------------------------------------------
@pytest.mark.parametrize("moist,ftype", [(0, 0), (1, 0), (0, 2)])
def test_cam_forcing_pass(moist, ftype):
    """apply_cam_forcing (ftype 0: states and tracers, the negativity clamp, the moist ps_v / dp3d update) and
    apply_cam_forcing_dynamics (ftype 2) on point-wise random tendencies pushed with f90_push_forcing_to_cxx."""
    from forcing_inputs import fill_forcing
    cfg = homme.preset("ne4", moisture=moist, ftype=ftype)
    p = Pair(cfg, oraclelib.ORACLE_LIB, b"cpu-oracle", fields=FIELDS_WITH_FORCING[ftype])
    try:
        p.ho.run_subcycle()                      # both at the same time levels (Pair warmed the reference only)
        p.reset()
        for h in (p.hr, p.ho):
            fill_forcing(h)
            h.push_forcing()
        p.call("hxx_apply_forcing", cfg.tstep * cfg.rsplit)
        p.same(f"apply_cam_forcing moist={moist} ftype={ftype}", changed=("v", "t") + (("qdp",) if ftype == 0 else ()))
    finally:
        p.close()


def test_diagnostics_calls():
    """prim_diag_scalars + prim_energy_halftimes into the F90 accumulators, before and after an advance."""
    cfg = homme.preset("ne4", disable_diagnostics=0, state_frequency=9999, moisture=1, use_cpstar=1)
    p = Pair(cfg, oraclelib.ORACLE_LIB, b"cpu-oracle")
    try:
        p.ho.run_subcycle()
        p.reset()
        for (before, ivs, ive) in [(1, 3, 2), (1, 0, 0), (0, 1, 1)]:
            p.call("hxx_diagnostics", before, ivs, ive)
            a, b = p.hr.accum(), p.ho.accum()
            for k in a:
                assert np.isfinite(a[k]).all(), k
                assert np.array_equal(a[k], b[k]), ((before, ivs, ive), k, float(np.abs(a[k] - b[k]).max()))
        assert np.abs(a["KEner"]).max() > 0 and np.abs(a["IEner"]).max() > 0 and np.abs(a["Qmass"]).max() > 0
        p.same("diagnostics leave the state alone")
    finally:
        p.close()
