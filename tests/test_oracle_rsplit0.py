"""rsplit = 0 (Eulerian vertical advection, CaarFunctorImpl.hpp:136-143,236-300,495-597; tracer-only remap,
RemapFunctor.hpp:42-100) in the oracle, cross-checked against the vertically Lagrangian path (rsplit = 3):
two different discretisations of the same equations must agree to truncation error after the same
simulated time, and the Eulerian path must conserve tracer mass to round-off."""
import numpy as np

from hommexx_b200 import homme
from oracle import oraclelib


def _run(rsplit, calls):
    cfg = homme.preset("ne4", rsplit=rsplit)
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    sph = h.array("spheremp").reshape(-1, 1, 1, 4, 4).copy()
    m0 = (h.state()["Qdp"][:, 0] * sph).sum(axis=(0, 2, 3, 4))
    for _ in range(calls):
        h.run_subcycle()
    h.push_results()
    nstep, nm1, n0, np1 = h.time_levels()
    st = {k: v.copy() for k, v in h.state().items()}
    tq = (nstep // cfg.qsplit) % 2
    m1 = (st["Qdp"][:, tq] * sph).sum(axis=(0, 2, 3, 4))
    h.close()
    return st, n0 - 1, nstep, m0, m1


def test_eulerian_vertical_matches_lagrangian_to_truncation_error():
    e, ne0, nstep_e, m0, m1 = _run(0, 6)
    l, nl0, nstep_l, _, _ = _run(3, 2)
    assert nstep_e == nstep_l == 6
    assert all(np.isfinite(v).all() for v in e.values())
    assert np.abs(e["T"][:, ne0] - l["T"][:, nl0]).max() <= 1e-4 * np.abs(l["T"][:, nl0]).max()
    assert np.abs(e["v"][:, ne0] - l["v"][:, nl0]).max() <= 2e-3 * np.abs(l["v"][:, nl0]).max()
    assert np.abs(e["ps_v"][:, ne0] - l["ps_v"][:, nl0]).max() <= 1e-5 * 1e5
    # the two paths are NOT the same arithmetic
    assert not np.array_equal(e["T"][:, ne0], l["T"][:, nl0])
    # tracer mass: conserved to round-off by advection + tracer-only remap
    assert np.abs(m1 - m0).max() <= 1e-12 * np.abs(m0).max(), (m0, m1)
