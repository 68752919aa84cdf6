"""Property tests of the oracle's quasi-monotone limiters, after the reference's own
(src/preqx/unit_tests/preqx_ut.cpp:1263-1532): results stay inside [qmin,qmax]*dp to 10 eps,
mass is conserved to 100 eps, and lim8 / lim9 (clip-and-sum) agree in 1-norm distance."""
import numpy as np
import pytest

from hommexx_b200 import homme
from oracle import oraclelib
from limiter_problems import EPS, check_limited, feasible_problem, run_limiter


@pytest.mark.parametrize("nlev", [72, 26])
@pytest.mark.parametrize("option", [8, 9])
def test_limiter_bounds_and_mass(nlev, option):
    lib = oraclelib.load_oracle(nlev, 4)
    for seed in range(4):
        sph, dpm, pt, ql, mass = feasible_problem(6, nlev, 1000 + seed)
        out, ql_out = run_limiter(lib, option, sph, dpm, pt, ql)
        check_limited(sph, dpm, out, ql_out, mass)
        assert np.allclose(ql_out, ql, rtol=1e-12, atol=0)  # feasible: limits move by round-off at most


def test_limiter_is_identity_inside_bounds():
    nlev = 26
    lib = oraclelib.load_oracle(nlev, 4)
    sph, dpm, _, ql, _ = feasible_problem(3, nlev, 5)
    pt = np.random.default_rng(5).uniform(0.1, 0.9, dpm.shape) * dpm
    ql[:, 0] = 0.0; ql[:, 1] = 10.0
    for option in (8, 9):
        out, _ = run_limiter(lib, option, sph, dpm, pt, ql)
        assert np.allclose(out, pt, rtol=4 * EPS, atol=0)


def test_limiter_relaxes_infeasible_bounds_and_clamps_negative_min():
    """EulerStepFunctorImpl.hpp:729-742: qmin<0 -> 0; bounds relaxed to the mean when infeasible."""
    nlev = 8
    lib = oraclelib.load_oracle(nlev, 4)
    rng = np.random.default_rng(3)
    sph = rng.uniform(1 / 16, 2 / 16, (1, 16)); dpm = rng.uniform(0.5, 1.0, (1, 16, nlev))
    pt = rng.uniform(0.4, 0.6, (1, 16, nlev)) * dpm
    ql = np.zeros((1, 2, nlev)); ql[0, 0] = -0.5; ql[0, 1] = 0.45   # mean ~0.5 > qmax
    out, ql_out = run_limiter(lib, 8, sph, dpm, pt, ql)
    assert (ql_out[0, 0] == 0.0).all()
    c = sph[:, :, None] * dpm
    mean = (pt / dpm * c).sum(1) / c.sum(1)
    assert np.allclose(ql_out[0, 1], np.maximum(mean[0], 0.45), rtol=1e-14)
    assert np.allclose((sph[:, :, None] * out).sum(1), (sph[:, :, None] * pt).sum(1), rtol=1e-13)


def test_lim8_vs_caas_one_norm():
    """preqx_ut.cpp:1483-1531: both limiters solve the same problem; lim8's 1-norm change is not larger
    than CAAS's by more than round-off."""
    nlev = 26
    lib = oraclelib.load_oracle(nlev, 4)
    sph, dpm, pt, ql, mass = feasible_problem(8, nlev, 77)
    o8, _ = run_limiter(lib, 8, sph, dpm, pt, ql)
    o9, _ = run_limiter(lib, 9, sph, dpm, pt, ql)
    w = sph[:, :, None]
    n8 = (w * np.abs(o8 - pt)).sum(1); n9 = (w * np.abs(o9 - pt)).sum(1)
    assert (n8 <= n9 * (1 + 1e3 * EPS) + 1e3 * EPS).all()
