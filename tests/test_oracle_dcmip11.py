"""DCMIP 2012 test 1-1 (3-D deformational flow, BASELINE configs[2]) as a tracer-only run on the oracle through the
functors' run methods (hommexx_b200/dcmip.py): the analytic winds and tracer shapes of the reference's
dcmip2012_test1_2_3.F90, one simulated day at ne8. The reference's C++ path has no prescribed-wind mode to
compare with (cxx_f90_interface.cpp:44), so the checks are the case's own invariants."""
import numpy as np

from hommexx_b200 import dcmip, homme
from oracle import oraclelib


def test_initial_fields_follow_the_dcmip_definition():
    lon = np.array([5 * np.pi / 6, 7 * np.pi / 6, 0.3]).reshape(3, 1, 1)
    lat = np.zeros((3, 1, 1))
    p5km = dcmip.P0 * np.exp(-5000.0 / dcmip.H)
    q = dcmip.tracers(lon, lat, np.array([p5km]).reshape(1, 1, 1))
    assert np.allclose(q[0, :2], 1.0) and q[0, 2] == 0.0                # bell centres and far field
    assert np.allclose(q[1, :2], 0.1) and np.isclose(q[1, 2, 0, 0], 0.9)
    assert (q[2, :2] == 1.0).all() and q[2, 2] == 0.1
    assert np.allclose(q[3], 1.0 - 0.3 * (q[0] + q[1] + q[2]))
    # the flow at t = 0: solid-body part 2 pi a / 12 days at the equator where sin(lon') = 0
    u, v = dcmip.winds(0.0, np.zeros((1, 1, 1)), np.zeros((1, 1, 1)), np.array([dcmip.P0 * 0.5]).reshape(1, 1, 1))
    ptop = dcmip.P0 * np.exp(-12000.0 / dcmip.H)
    ud = (23000 * np.pi / dcmip.TAU * dcmip.A) / (0.2 * ptop) * (-np.exp((0.5 * dcmip.P0 - dcmip.P0) / (0.2 * ptop))
                                                              + np.exp((ptop - 0.5 * dcmip.P0) / (0.2 * ptop)))
    assert np.isclose(u[0, 0, 0], 2 * np.pi * dcmip.A / dcmip.TAU + ud) and v[0, 0, 0] == 0.0


def test_one_day_of_deformational_flow():
    d = dcmip.Dcmip11(8, 26, oraclelib.ORACLE_LIB, tstep=600.0)
    m0, q0 = d.masses(), d.q()
    for _ in range(144):
        d.step()
    q, m = d.q(), d.masses()
    ps = d.h.get_field("ps_v").reshape(d.n, 3, 16)[:, 2]
    d.close()
    assert np.isfinite(q).all()
    assert np.abs(m / m0 - 1.0).max() <= 1e-13                          # tracer mass: round-off
    # the limiter keeps every tracer inside its initial global range (margin: tracer / thickness consistency)
    for i in range(4):
        lo, hi = q0[i].min(), q0[i].max()
        assert q[i].min() >= lo - 1e-2 * (hi - lo) and q[i].max() <= hi + 1e-2 * (hi - lo), i
    assert np.abs(ps - 1e5).max() < 20.0                                # the flow is non-divergent in the column
    # and it really deforms the fields: a day moves the bells by 30 degrees and stretches them
    assert np.abs(q[0] - q0[0]).max() > 0.3


def test_dcmip11_on_the_reference_functors_matches_the_oracle():
    """The same tracer-only driver on the REFERENCE's own build: its prim_run_subcycle_c refuses prescribed winds, but
    its functors' run methods (EulerStepFunctor::euler_step x 3, qdp_time_avg, VerticalRemapManager::run_remap,
    update_q — bound to the section-C hooks by oracle/ref_hommexx_api.cpp) advect the DCMIP 1-1 tracers through the
    analytic deformational flow all the same. 24 steps at ne8: every tracer array bit-identical to the oracle."""
    from reference_lib import reference_lib
    runs = []
    for lib in (reference_lib(26, 4), oraclelib.ORACLE_LIB):
        d = dcmip.Dcmip11(8, 26, lib, tstep=600.0)
        m0 = d.masses()
        for _ in range(24):
            d.step()
        runs.append((d.q(), d.masses() / m0, d.h.get_field("qdp"), d.h.get_field("Q"), d.h.get_field("dp3d")))
        d.close()
    (qr, mr, qdpr, Qr, dpr), (qo, mo, qdpo, Qo, dpo) = runs
    assert np.isfinite(qr).all() and np.abs(mr - 1.0).max() <= 1e-13
    assert np.abs(qr[0] - qr[0].mean()).max() > 0.1                      # a non-trivial field
    for a, b, nm in ((qr, qo, "q"), (mr, mo, "mass"), (qdpr, qdpo, "qdp"), (Qr, Qo, "Q"), (dpr, dpo, "dp3d")):
        assert np.array_equal(a, b), (nm, float(np.abs(a - b).max()))
