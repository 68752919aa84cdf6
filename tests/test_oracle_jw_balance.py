"""End-to-end physics check of the oracle + driver: the UNPERTURBED Jablonowski-Williamson state is an exact
steady solution of the primitive equations, so the discrete model must hold it up to truncation error, and
that error must shrink when the mesh is refined (Jablonowski & Williamson 2006, section 4a). This exercises
the initial state (baroclinic_inst_mod.F90), the metric terms, CAAR, hyperviscosity and the remap together,
independently of any stored reference output."""
import time

import numpy as np

from hommexx_b200 import homme
from oracle import oraclelib


def _drift(ne):
    s = (4.0 / ne)
    nu = 7e15 * s ** 3.2
    cfg = homme.preset("prtcA", ne=ne, qsize=0, u_perturb=0.0, tstep=600.0 * s, nu=nu, nu_p=nu, nu_div=nu)
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    s0 = {k: v.copy() for k, v in h.state().items()}
    for _ in range(4 * ne // 4):      # 2 simulated hours
        h.run_subcycle()
    h.push_results()
    st, n0 = h.state(), h.time_levels()[2] - 1
    out = dict(du=np.abs(st["v"][:, n0, :, 0] - s0["v"][:, 0, :, 0]).max(), v=np.abs(st["v"][:, n0, :, 1]).max(),
               dps=np.abs(st["ps_v"][:, n0] - 1e5).max(), dT=np.abs(st["T"][:, n0] - s0["T"][:, 0]).max())
    h.close()
    return out


def test_unperturbed_jw_state_stays_balanced_and_converges():
    t0 = time.time()
    c, f = _drift(4), _drift(8)
    print("JW balance drift after 2 h: ne4", c, "ne8", f, f"({time.time() - t0:.1f}s)")
    # a 35 m/s jet and a 1000 hPa surface: the drift is truncation error
    assert c["du"] < 1.0 and c["v"] < 1.0 and c["dps"] < 150.0 and c["dT"] < 0.5
    for k in c:
        assert f[k] < c[k] / 2.5, (k, c[k], f[k])
