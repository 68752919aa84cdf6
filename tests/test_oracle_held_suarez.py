"""Held-Suarez forcing harness (hommexx_b200/held_suarez.py; reference physics/heldsuarez/held_suarez_mod.F90)
driving the oracle through f90_push_forcing_to_cxx + prim_run_subcycle_c (ftype = 0)."""
import numpy as np

from hommexx_b200 import held_suarez as hs
from oracle import oraclelib
from hommexx_b200 import homme
from oracle import oraclelib


def test_forcing_formulas():
    hyai, hybi, hyam, hybm = homme.read_vcoord(26)
    lat = np.linspace(-1.5, 1.5, 32).reshape(2, 4, 4)
    ps = np.full((2, 4, 4), 1.0e5)
    T = np.full((2, 26, 4, 4), 250.0)
    ft, Teq = hs.hs_T_forcing(hyam, hybm, ps, T, lat)
    assert Teq.min() >= 200.0 and Teq.max() <= 315.0 + 1e-9
    # equilibrium temperature: warmest at the equatorial surface, 200 K floor aloft
    assert Teq[:, -1].max() > 300.0 and np.isclose(Teq[:, 0].min(), 200.0)
    # relaxation pulls towards Teq at rate between k_a and k_s
    rate = -ft / (T - Teq)
    assert rate.min() >= hs.K_A * (1 - 1e-12) and rate.max() <= hs.K_S * (1 + 1e-12)
    v = np.ones((2, 26, 2, 4, 4))
    fm = hs.hs_v_forcing(hyam, hybm, v)
    sigma = hyam + hybm
    assert (fm[:, sigma <= hs.SIGMA_B] == 0).all()                   # free atmosphere: no friction
    assert np.isclose(fm[:, -1].min(), -hs.K_F * (sigma[-1] - 0.7) / 0.3)


def test_held_suarez_forced_run_on_the_oracle():
    cfg = homme.preset("prtcA", qsize=0, ftype=0)
    runs = {}
    for forced in (False, True):
        h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
        h.init_dycore()
        for _ in range(6):
            if forced:
                hs.forced_step(h)
            else:
                h.run_subcycle()
        h.push_results()
        n0 = h.time_levels()[2] - 1
        runs[forced] = {k: v[:, n0].copy() for k, v in h.state().items() if k in ("T", "v", "ps_v")}
        lat = h.array("lat").reshape(h.nelemd, 4, 4).copy()
        vc = h.vcoord
        h.close()
    f, u = runs[True], runs[False]
    assert all(np.isfinite(a).all() for a in f.values())
    assert not np.array_equal(f["T"], u["T"])
    # Rayleigh friction: the boundary-layer winds are weaker than in the unforced run, the free atmosphere barely moves
    sigma = vc[2] + vc[3]
    bl = sigma > 0.85
    ke = lambda a, m: (a[:, m] ** 2).sum()
    assert ke(f["v"], bl) < 0.97 * ke(u["v"], bl)
    assert abs(ke(f["v"], ~bl) / ke(u["v"], ~bl) - 1.0) < 0.05
    # Newtonian cooling: T moved towards Teq where it was far from it (6 calls = 3 h at k_T <= 1/4 day^-1)
    _, Teq = hs.hs_T_forcing(vc[2], vc[3], u["ps_v"], u["T"], lat)
    far = np.abs(u["T"] - Teq) > 20.0
    assert far.any()
    assert (np.abs(f["T"] - Teq)[far]).mean() < (np.abs(u["T"] - Teq)[far]).mean()


def test_oracle_hs_forcing_entry_point_matches_the_numpy_formulas():
    """hxx_held_suarez_forcing (the checker of the product's device kernel) against the numpy restatement above."""
    cfg = homme.preset("prtcA", qsize=0, ftype=0)
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    h.run_subcycle()
    h.push_results()
    n0 = h.time_levels()[2] - 1
    st = h.state()
    lat = h.array("lat").reshape(h.nelemd, 4, 4)
    ft, _ = hs.hs_T_forcing(h.vcoord[2], h.vcoord[3], st["ps_v"][:, n0], st["T"][:, n0], lat)
    fm = hs.hs_v_forcing(h.vcoord[2], h.vcoord[3], st["v"][:, n0])
    h.held_suarez_forcing()
    n, nlev = h.nelemd, cfg.nlev
    got_ft = h.get_field("ft").reshape(n, 16, nlev).transpose(0, 2, 1).reshape(n, nlev, 4, 4)
    got_fm = h.get_field("fm").reshape(n, 2, 16, nlev).transpose(0, 3, 1, 2).reshape(n, nlev, 2, 4, 4)
    assert np.abs(got_ft - ft).max() <= 1e-13 * np.abs(ft).max()
    assert np.abs(got_fm - fm).max() <= 1e-13 * np.abs(fm).max()
    h.close()


def test_held_suarez_forced_run_matches_the_reference_build():
    """The CAM-coupled wrapper's sequence (forcing in, step, results out; prim_driver_mod.F90:1380-1402) with the
    Held-Suarez tendencies, on the reference's own build and on the oracle: the same numpy physics feeds both, so
    after eight forced calls (with four tracers riding along) every prognostic array must agree bit for bit."""
    from reference_lib import reference_lib
    cfg = homme.preset("prtcA", ftype=0)
    out = []
    for lib in (reference_lib(cfg.nlev, cfg.qsize_d), oraclelib.ORACLE_LIB):
        h = homme.Homme(cfg, lib)
        h.init_dycore()
        for _ in range(8):
            hs.forced_step(h)
        out.append({k: v.copy() for k, v in h.state().items()})
        h.close()
    ref, ora = out
    for k in ("v", "T", "dp3d", "ps_v", "Qdp", "Q", "omega_p"):
        assert np.isfinite(ref[k]).all(), k
        assert np.array_equal(ref[k], ora[k]), (k, float(np.abs(ref[k] - ora[k]).max()))
