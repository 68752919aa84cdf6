"""Tensor hyperviscosity (hypervis_scaling != 0: laplace_tensor, vlaplace_sphere_wk_cartesian,
SphereOperators.hpp:604-635,752-814) in the oracle, with the driver's tensorVisc (cube_mod.F90:315-428).
The conversion nu_tensor = nu_const (2 rearth / ((np-1) dx))^hv_scaling rearth^-4 (cube_mod.F90:372-384) makes the
tensor operator as strong as the constant-coefficient one on a uniform mesh, so the two runs must stay
close — and must not be the same arithmetic."""
import numpy as np

from hommexx_b200 import homme
from oracle import oraclelib


def _run(cfg, calls=4):
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    tv = h.array("tensorvisc").copy()
    for _ in range(calls):
        h.run_subcycle()
    h.push_results()
    st = {k: v.copy() for k, v in h.state().items()}
    n0 = h.time_levels()[2] - 1
    h.close()
    return st, n0, tv


def test_tensor_hv_tracks_constant_hv():
    base = homme.preset("ne4")
    ne, hs = base.ne, 3.2
    dx = 2 * np.pi * 6.376e6 / (3 * 4 * ne)
    nu_t = base.nu * (2 * 6.376e6 / (3 * dx)) ** hs * 6.376e6 ** -4.0
    assert 3e-8 < nu_t < 8e-8
    t, n0, tv = _run(homme.preset("ne4", hypervis_scaling=hs, nu=nu_t, nu_p=nu_t, nu_q=nu_t, nu_s=nu_t))
    c, n0c, tv0 = _run(base)
    assert n0 == n0c and np.abs(tv).max() > 0 and np.abs(tv0).max() == 0
    assert all(np.isfinite(v).all() for v in t.values())
    assert not np.array_equal(t["T"][:, n0], c["T"][:, n0])
    assert np.abs(t["T"][:, n0] - c["T"][:, n0]).max() <= 0.2             # K, of ~300
    assert np.abs(t["v"][:, n0] - c["v"][:, n0]).max() <= 1.0             # m/s, of ~35
    assert np.abs(t["Q"] - c["Q"]).max() <= 5e-2 * np.abs(c["Q"]).max()
