"""The oracle's limiters against the reference's own, problem by problem: the reference build exports hxx_limiter bound
to SerialLimiter::run<8|9> — what limiter_optim_iter_full(kv) / limiter_clip_and_sum(kv) dispatch to on a host
execution space (EulerStepFunctorImpl.hpp:640-666) — and, as options 108 / 109, to the team implementations its
unit tests call (:766-884, preqx_ut.cpp:1386-1411). Feasible problems of LimiterTester (tests/limiter_problems.py),
infeasible bounds, negative minima, constant fields: ptens and qlim bit-identical to the serial path, and the
team path of the reference within round-off of it."""
import numpy as np
import pytest

from hommexx_b200 import homme
from limiter_problems import EPS, check_limited, feasible_problem, run_limiter
from oracle import oraclelib
from reference_lib import reference_lib


def problems(nlev):
    for seed in range(3):
        yield f"feasible-{seed}", feasible_problem(6, nlev, 4000 + seed)[:4]
    rng = np.random.default_rng(9)
    sph = rng.uniform(1 / 16, 2 / 16, (4, 16)); dpm = rng.uniform(0.5, 1.0, (4, 16, nlev))
    pt = rng.uniform(0.4, 0.6, (4, 16, nlev)) * dpm
    ql = np.zeros((4, 2, nlev)); ql[:, 0] = -0.5; ql[:, 1] = 0.45       # mean above qmax, negative qmin
    yield "infeasible-high", (sph, dpm, pt, ql)
    ql2 = np.zeros((4, 2, nlev)); ql2[:, 0] = 0.55; ql2[:, 1] = 0.9     # mean below qmin
    yield "infeasible-low", (sph, dpm, pt, ql2)
    pt3 = 0.3 * dpm                                                     # constant mixing ratio, bounds touching it
    ql3 = np.zeros((4, 2, nlev)); ql3[:, 0] = 0.3; ql3[:, 1] = 0.3
    yield "constant", (sph, dpm, pt3, ql3)
    pt4 = pt.copy(); pt4[:, ::3] = 0.0                                  # exact zeros, as in a cosine-bell field
    ql4 = np.zeros((4, 2, nlev)); ql4[:, 1] = 0.5
    yield "zeros", (sph, dpm, pt4, ql4)


@pytest.mark.parametrize("nlev", [72, 26])
@pytest.mark.parametrize("option", [8, 9])
def test_limiter_is_bit_identical_to_the_reference(nlev, option):
    ref = homme.load_dycore(reference_lib(nlev, 4))
    ora = oraclelib.load_oracle(nlev, 4)
    for name, (sph, dpm, pt, ql) in problems(nlev):
        pr, qr = run_limiter(ref, option, sph, dpm, pt, ql)
        po, qo = run_limiter(ora, option, sph, dpm, pt, ql)
        assert np.isfinite(pr).all(), name
        assert np.array_equal(pr, po), (name, option, float(np.abs(pr - po).max()))
        assert np.array_equal(qr, qo), (name, option, float(np.abs(qr - qo).max()))


@pytest.mark.parametrize("option", [8, 9])
def test_reference_team_limiter_agrees_with_its_serial_limiter(option):
    """The two implementations the reference holds (host serial, team): same answers to round-off; both satisfy the
    acceptance thresholds of its own unit test."""
    nlev = 72
    ref = homme.load_dycore(reference_lib(nlev, 4))
    sph, dpm, pt, ql, mass = feasible_problem(6, nlev, 4100)
    ps, qs = run_limiter(ref, option, sph, dpm, pt, ql)
    ptm, qtm = run_limiter(ref, 100 + option, sph, dpm, pt, ql)
    check_limited(sph, dpm, ps, qs, mass)
    check_limited(sph, dpm, ptm, qtm, mass)
    assert np.abs(ps - ptm).max() <= 1e3 * EPS
