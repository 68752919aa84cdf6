"""Feasible limiter problems, after LimiterTester::init_feasible of the reference's unit tests
(src/preqx/unit_tests/preqx_ut.cpp:1335-1384), with fixed seeds instead of std::random_device."""
import numpy as np

EPS = np.finfo(float).eps


def feasible_problem(nsets, nlev, seed):
    rng = np.random.default_rng(seed)
    sphweights = rng.uniform(1.0 / 16, 2.0 / 16, (nsets, 16))
    dpmass = rng.uniform(0.5, 1.0, (nsets, 16, nlev))
    ptens = rng.uniform(0.0, 1.0, (nsets, 16, nlev)) * dpmass
    qlim = np.sort(rng.uniform(0.0, 1.0, (nsets, 2, nlev)), axis=1)
    w = sphweights[:, :, None]
    m = (w * ptens).sum(1)
    lo = (w * qlim[:, 0:1, :] * dpmass).sum(1)
    hi = (w * qlim[:, 1:2, :] * dpmass).sum(1)
    dm = np.where(m < lo, lo - m, np.where(m > hi, hi - m, 0.0))
    ptens = ptens + (1 + 1e2 * EPS) * dm[:, None, :] / (16 * w)
    mass = (w * ptens).sum(1)
    assert (mass >= (1 - 10 * EPS) * lo).all() and (mass <= (1 + 10 * EPS) * hi).all()
    return sphweights, dpmass, ptens, qlim, mass


def check_limited(sphweights, dpmass, ptens, qlim, mass):
    """preqx_ut.cpp:1413-1429: bounds to 10 eps, mass to 100 eps."""
    lo = (1 - 10 * EPS) * qlim[:, 0:1, :] * dpmass
    hi = (1 + 10 * EPS) * qlim[:, 1:2, :] * dpmass
    assert (ptens >= lo).all(), float((lo - ptens).max())
    assert (ptens <= hi).all(), float((ptens - hi).max())
    m = (sphweights[:, :, None] * ptens).sum(1)
    assert (np.abs(m - mass) <= 1e2 * EPS * mass).all(), float((np.abs(m - mass) / mass).max())


def run_limiter(lib, option, sphweights, dpmass, ptens, qlim):
    nsets = sphweights.shape[0]
    s = np.ascontiguousarray(sphweights); d = np.ascontiguousarray(dpmass)
    p = np.ascontiguousarray(ptens).copy(); q = np.ascontiguousarray(qlim).copy()
    lib.hxx_limiter(option, nsets, s.ctypes.data, d.ctypes.data, p.ctypes.data, q.ctypes.data)
    return p, q
