"""GPU parity of the timestep phases and of whole runs against the CPU oracle, through the C ABI.
north_star tolerance: normalised relative L2 <= 1e-11 on v, T, dp3d, ps, Q after 10 steps. The
kernels reproduce the oracle's operation order (and --fmad=false), so the tests demand MORE:
bit-identical fields (tol = 0) wherever stated."""
import numpy as np
import pytest

import parity
from hommexx_b200 import homme
from oracle import oraclelib

pytestmark = pytest.mark.gpu
TOL = 1e-11  # north_star


@pytest.fixture(scope="module")
def warm_pair():
    """ne4 / nlev72 / qsize4 after one subcycle call: a state with every field non-trivial."""
    cfg = homme.preset("ne4")
    hc, ho = parity.pair(cfg)
    ho.run_subcycle()
    snap = {n: ho.get_field(n) for n in parity.STATE_FIELDS}
    yield hc, ho, snap
    hc.close(); ho.close()


def _reset(warm):
    """Both libraries back to the same physical state (the oracle's after one subcycle call)."""
    hc, ho, snap = warm
    for n, a in snap.items():
        hc.set_field(n, a); ho.set_field(n, a)
    return hc, ho


def test_caar_stage_parity(warm_pair):
    for (nm1, n0, np1, dt, w, dss) in [(1, 1, 0, 360.0, 0.25, 0), (1, 0, 2, 360.0, 0.0, 1), (1, 2, 2, 600.0, 0.0, 1),
                                       (0, 2, 2, 1350.0, 0.75, 1)]:
        hc, ho = _reset(warm_pair)
        for h in (hc, ho):
            h.lib.hxx_caar_run(nm1, n0, np1, dt, w, -1, dss)
        parity.compare_fields(hc, ho, parity.STATE_FIELDS + ["phi"], tol=0.0, what=f"caar {(nm1, n0, np1, dss)}")


def test_caar_moist_parity(warm_pair):
    hc, ho = _reset(warm_pair)
    for h in (hc, ho):
        h.lib.hxx_caar_run(0, 1, 2, 60.0, 0.25, 1, 1)
    parity.compare_fields(hc, ho, tol=0.0, what="caar moist")


def test_rk_combine_and_step_init_parity(warm_pair):
    hc, ho = _reset(warm_pair)
    for h in (hc, ho):
        h.lib.hxx_rk_combine(0, 1)
        h.lib.hxx_prim_step_init(2)
    parity.compare_fields(hc, ho, tol=0.0, what="rk_combine/prim_step_init")


def test_hypervis_parity(warm_pair):
    hc, ho = _reset(warm_pair)
    for h in (hc, ho):
        h.lib.hxx_hypervis_run(2, 1800.0, 1.0)
    # vtens/ttens are scratch: the reference's TagUpdateStates leaves dt*tens*rspheremp in them
    # (HyperviscosityFunctorImpl.hpp:141-150), the CUDA kernel keeps that product in registers
    names = [f for f in parity.STATE_FIELDS if f not in ("vtens", "ttens")]
    parity.compare_fields(hc, ho, names, tol=0.0, what="hypervis")


def test_euler_stages_parity(warm_pair):
    hc, ho = _reset(warm_pair)
    for h in (hc, ho):
        h.lib.hxx_euler_reset()
        h.lib.hxx_euler_precompute_divdp()
    parity.compare_fields(hc, ho, tol=0.0, what="precompute_divdp")
    names = [f for f in parity.STATE_FIELDS]
    for (np1q, n0q, rhs, opt) in [(1, 0, 0.0, 2), (1, 1, 1.0, 0), (1, 1, 2.0, 1)]:
        for h in (hc, ho):
            h.lib.hxx_euler_step(np1q, n0q, 900.0, rhs, opt)
        parity.compare_fields(hc, ho, names, tol=0.0, what=f"euler_step rhs_mult={rhs}")
    for h in (hc, ho):
        h.lib.hxx_euler_qdp_time_avg(0, 1)
    parity.compare_fields(hc, ho, names, tol=0.0, what="qdp_time_avg")


def test_remap_and_update_q_parity(warm_pair):
    hc, ho = _reset(warm_pair)
    qdp0 = ho.get_field("qdp").reshape(ho.nelemd, 2, 4, 16, 72).copy()
    for h in (hc, ho):
        h.lib.hxx_vertical_remap(2, 1, 5400.0)
        h.lib.hxx_update_q(1, 2)
    parity.compare_fields(hc, ho, tol=0.0, what="vertical_remap/update_q")
    # tracer column mass is conserved by the remap to round-off
    qdp1 = hc.get_field("qdp").reshape(hc.nelemd, 2, 4, 16, 72)
    assert np.allclose(qdp1[:, 1].sum(-1), qdp0[:, 1].sum(-1), rtol=1e-13, atol=0)


def test_caar_eulerian_vertical_parity():
    """rsplit = 0: every RK stage shape of CAAR with the vertical-advection terms, then the tracer-only remap."""
    cfg = homme.preset("ne4", rsplit=0)
    hc, ho = parity.pair(cfg)
    ho.run_subcycle()
    snap = {n: ho.get_field(n) for n in parity.STATE_FIELDS}
    for (nm1, n0, np1, dt, w, dss) in [(1, 1, 0, 360.0, 0.25, 0), (1, 0, 2, 360.0, 0.0, 1), (1, 2, 2, 600.0, 0.0, 1),
                                       (0, 2, 2, 1350.0, 0.75, 1)]:
        for h in (hc, ho):
            for n, a in snap.items():
                h.set_field(n, a)
            h.lib.hxx_caar_run(nm1, n0, np1, dt, w, -1, dss)
        parity.compare_fields(hc, ho, parity.STATE_FIELDS + ["phi"], tol=0.0, what=f"caar rsplit=0 {(nm1, n0, np1, dss)}")
    for h in (hc, ho):
        h.lib.hxx_vertical_remap(2, 1, 1800.0)
        h.lib.hxx_update_q(1, 2)   # fused into the CUDA remap's tracer store; a separate pass in the oracle
    parity.compare_fields(hc, ho, tol=0.0, what="eulerian remap")
    hc.close(); ho.close()


CASES = {
    "ne4": dict(),                                              # BASELINE configs[0]: ne4, nlev 72, qsize 4
    "prtcA": dict(),                                            # the reference's prtcA_c sizes: nlev 26, qsize 4
    "prtcA-lim9-alg2": dict(base="prtcA", limiter_option=9, remap_alg=2),
    "prtcA-moist-nudiv-q2": dict(base="prtcA", moisture=1, nu_div=1.75e16, qsplit=2, rsplit=2),
    "ne8": dict(),
    # rsplit = 0: Eulerian vertical advection in CAAR, tracer-only remap (test-list.cmake's r0 variants)
    "ne4-r0": dict(base="ne4", rsplit=0),
    "prtcA-r0-moist-q3": dict(base="prtcA", rsplit=0, moisture=1, qsplit=3),
    # tensor hyperviscosity (prtcB-r3-tensorhv-dry.nl: hypervis_scaling 3.2, nu 1e-9, hypervis_subcycle 2), with the
    # driver's tensorVisc / vec_sphere2cart (cube_mod.F90:172-178,315-428)
    "prtcA-tensorhv": dict(base="prtcA", hypervis_scaling=3.2, nu=1e-9, nu_p=1e-9, nu_q=1e-9, nu_s=1e-9, nu_div=1e-9,
                           hypervis_subcycle=2),
    "ne4-tensorhv-nudiv": dict(base="ne4", hypervis_scaling=3.2, nu=5e-8, nu_p=5e-8, nu_q=5e-8, nu_s=5e-8, nu_div=1.25e-7),
    # no tracers at all (the remap then carries only the three state fields; forcing and update_q have nothing to do)
    "prtcA-q0": dict(base="prtcA", qsize=0),
    "ne4-r0-q0": dict(base="ne4", rsplit=0, qsize=0),
}


@pytest.mark.parametrize("case", list(CASES))
def test_ten_step_parity(case):
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    hc, ho = parity.pair(cfg)
    nstep = 0
    while nstep < 10:
        nstep = hc.run_subcycle()
        assert ho.run_subcycle() == nstep
    assert hc.time_levels() == ho.time_levels()
    hc.push_results(); ho.push_results()
    sc, so = hc.state(), ho.state()
    errs = {k: (0.0 if np.array_equal(sc[k], so[k]) else parity.rel_l2(sc[k], so[k])) for k in parity.PROGNOSTIC}
    print(case, "steps", nstep, "rel-L2 vs oracle:", errs)
    assert not any(np.isnan(so[k]).any() for k in so)
    assert max(errs.values()) <= TOL, errs
    assert hc.lib.hommexx_b200_launch_count() > 0
    hc.close(); ho.close()


def test_ne30_full_size_properties_and_parity():
    """BASELINE configs[1] (ne30, nlev 72, qsize 40): one subcycle call = 3 dynamics steps on the
    GPU. Size-independent properties — global tracer mass and dry-air mass conserved to round-off,
    the q=1 tracer stays 1, nothing NaN — plus a full parity check of a 4-tracer twin against the oracle."""
    parity.need_gpu()
    cfg = homme.preset("ne30")
    hc = homme.Homme(cfg, parity.cuda_lib(72, 40))
    hc.init_dycore()
    s = hc.state()
    sph = hc.array("spheremp").reshape(-1, 1, 1, 4, 4)
    m0 = (s["Qdp"][:, 0] * sph).sum(axis=(0, 2, 3, 4))
    hc.run_subcycle()
    hc.push_results()
    s = hc.state()
    nstep, nm1, n0, np1 = hc.time_levels()
    assert nstep == 3 and not any(np.isnan(v).any() for v in s.values())
    tq = (nstep // cfg.qsplit) % 2    # n0_qdp after the step
    m1 = (s["Qdp"][:, tq] * sph).sum(axis=(0, 2, 3, 4))
    assert np.abs(m1 - m0).max() <= 1e-12 * np.abs(m0).max(), (m0, m1)
    assert np.abs(s["Q"][:, 2] - 1.0).max() <= 1e-4            # q3 == 1 initially: stays 1 to truncation error
    ps = s["ps_v"][:, n0 - 1]
    area = hc.array("spheremp").reshape(-1, 4, 4)
    assert abs((ps * area).sum() / area.sum() - 1e5) <= 1e-8 * 1e5   # mean surface pressure
    hc.close()
    cfg4 = homme.preset("ne30", qsize=4, qsize_d=4)
    h4, ho = parity.pair(cfg4)
    h4.run_subcycle(); ho.run_subcycle()
    h4.push_results(); ho.push_results()
    a, b = h4.state(), ho.state()
    errs = {k: (0.0 if np.array_equal(a[k], b[k]) else parity.rel_l2(a[k], b[k])) for k in parity.PROGNOSTIC}
    print("ne30 q4 rel-L2 vs oracle:", errs)
    assert max(errs.values()) <= TOL, errs
    h4.close(); ho.close()


@pytest.mark.parametrize("preset,calls", [("ne8", 4), ("ne30q4", 2)])
def test_dcmip11_tracer_stress_parity(preset, calls):
    """BASELINE configs[2]: the tracer shapes of DCMIP 2012 test 1-1 (cosine bells, correlated field, slotted
    cylinders, constant) advected by the evolving JW flow. The reference's C++ path rejects prescribed winds
    (cxx_f90_interface.cpp:44), so the stress is on what it does run — EulerStep + limiter + PPM remap with
    the limiter iterating on sharp edges: CUDA vs oracle bit-identical, global bounds kept, mass conserved."""
    import dcmip_tracers
    cfg = homme.preset("ne30", qsize=4, qsize_d=4) if preset == "ne30q4" else homme.preset(preset)
    parity.need_gpu()
    hc = homme.Homme(cfg, parity.cuda_lib(cfg.nlev, cfg.qsize_d))
    ho = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    for h in (hc, ho):
        dcmip_tracers.install(h)
        h.init_dycore()
    sph = hc.array("spheremp").reshape(-1, 1, 1, 4, 4)
    m0 = (hc.state()["Qdp"][:, 0] * sph).sum(axis=(0, 2, 3, 4))
    for _ in range(calls):
        hc.run_subcycle(); ho.run_subcycle()
    hc.push_results(); ho.push_results()
    a, b = hc.state(), ho.state()
    for k in parity.PROGNOSTIC:
        assert np.array_equal(a[k], b[k]), k
    nstep = hc.time_levels()[0]
    tq = (nstep // cfg.qsplit) % 2
    m1 = (a["Qdp"][:, tq] * sph).sum(axis=(0, 2, 3, 4))
    assert np.abs(m1 - m0).max() <= 1e-12 * np.abs(m0).max()
    Q = a["Q"]
    # the limiter bounds Qdp / (tracer-consistent dp); Q = Qdp / dp(ps_v) carries the tracer/dynamics
    # consistency error on top (the q = 1 tracer shows it: ~2e-4 after a few steps), hence the 5e-4 margins
    eps = 5e-4
    assert Q[:, 0].min() >= -1e-12 and Q[:, 0].max() <= 1.0 + eps              # cosine bells stay in [0, 1]
    assert Q[:, 2].min() >= 0.1 * (1 - eps) and Q[:, 2].max() <= 1.0 + eps    # slotted cylinders stay in [0.1, 1]
    assert np.abs(Q[:, 3] - 1.0).max() <= eps                                  # the constant stays constant
    assert (np.abs(a["Qdp"][:, tq, 2] - a["Qdp"][:, 1 - tq, 2]) > 0).any()    # and something moved
    hc.close(); ho.close()
