"""CPU tests of the oracle's CAM forcing (CamForcing.cpp:20-174) and diagnostics
(Diagnostics.cpp:37-185) against a direct numpy restatement of the reference formulas, through the
reference's ABI (f90_push_forcing_to_cxx / prim_run_subcycle_c / init_diagnostics_c pointers)."""
import numpy as np
import pytest

from forcing_inputs import fill_forcing
from hommexx_b200 import homme
from oracle import oraclelib

CP, CPWV = 1005.0, 1870.0


def _dp(h, ps):
    hyai, hybi = h.vcoord[0], h.vcoord[1]
    dai, dbi = np.diff(hyai), np.diff(hybi)
    return dai[None, :, None, None] * 1e5 + dbi[None, :, None, None] * ps[:, None]


@pytest.mark.parametrize("moist,ftype", [(0, 0), (1, 0), (0, 2)])
def test_forcing_matches_formulas(moist, ftype):
    cfg = homme.preset("ne4", nlev=26, vcoord="cam-26", moisture=moist, ftype=ftype)
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    fill_forcing(h)
    st0 = {k: v.copy() for k, v in h.state().items()}
    f = {k: v.copy() for k, v in h.forcing().items()}
    h.push_forcing()
    lib = h.lib
    nstep, nm1, n0, np1 = h.time_levels()
    dt_remap = cfg.tstep * cfg.qsplit * cfg.rsplit
    # expected, from the reference formulas
    n0c = n0 - 1
    T_exp = st0["T"][:, n0c] + dt_remap * f["FT"]
    v_exp = st0["v"][:, n0c] + dt_remap * f["FM"]
    q_exp = st0["Qdp"][:, 0].copy()
    ps_exp = st0["ps_v"][:, n0c].copy()
    if ftype == 0:
        v1 = dt_remap * f["FQ"]
        clamp = (q_exp + v1 < 0.0) & (v1 < 0.0)
        v1 = np.where(clamp, np.where(q_exp < 0.0, 0.0, -q_exp), v1)
        if moist:
            acc = np.zeros_like(ps_exp)
            for k in range(cfg.nlev):
                acc = acc + v1[:, 0, k]
            ps_exp = ps_exp + acc
        q_exp[:, :cfg.qsize] = (q_exp + v1)[:, :cfg.qsize]
    # run the pass through the library's own phase hook
    lib.hxx_apply_forcing(dt_remap)
    nl = cfg.nlev
    t_dev = h.get_field("t").reshape(h.nelemd, 3, 16, nl)[:, n0c].reshape(h.nelemd, 4, 4, nl).transpose(0, 3, 1, 2)
    v_dev = h.get_field("v").reshape(h.nelemd, 3, 2, 16, nl)[:, n0c].reshape(h.nelemd, 2, 4, 4, nl).transpose(0, 4, 1, 2, 3)
    assert np.array_equal(t_dev, T_exp)
    assert np.array_equal(v_dev, v_exp)
    q_dev = h.get_field("qdp").reshape(h.nelemd, 2, cfg.qsize_d, 4, 4, nl)[:, 0].transpose(0, 1, 4, 2, 3)
    ps_dev = h.get_field("ps_v").reshape(h.nelemd, 3, 4, 4)[:, n0c]
    assert np.array_equal(q_dev, q_exp)
    assert np.array_equal(ps_dev, ps_exp)
    if ftype == 0:
        assert (q_dev[:, :cfg.qsize] >= np.minimum(st0["Qdp"][:, 0, :cfg.qsize], 0.0)).all()  # clamp never overshoots
        Q_dev = h.get_field("Q").reshape(h.nelemd, cfg.qsize_d, 4, 4, nl).transpose(0, 1, 4, 2, 3)
        Q_exp = q_exp / _dp(h, ps_exp)[:, None]
        assert np.array_equal(Q_dev[:, :cfg.qsize], Q_exp[:, :cfg.qsize])
    h.close()


def test_forcing_roundtrip_and_qdp_push():
    cfg = homme.preset("ne4", nlev=26, vcoord="cam-26")
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    fill_forcing(h)
    f = {k: v.copy() for k, v in h.forcing().items()}
    q0 = h.state()["Qdp"].copy()
    h.state()["Qdp"][...] = -1.0          # f90_push_forcing_to_cxx overwrites the F90 Qdp with the device copy
    h.push_forcing()
    assert np.array_equal(h.state()["Qdp"], q0)
    for a in h.forcing().values():
        a[...] = 0.0
    h.pull_forcing()
    for k, a in h.forcing().items():
        assert np.array_equal(a, f[k]), k
    h.close()


@pytest.mark.parametrize("cpstar", [0, 1])
def test_diagnostics_match_formulas(cpstar):
    cfg = homme.preset("ne4", nlev=26, vcoord="cam-26", disable_diagnostics=0, use_cpstar=cpstar, state_frequency=3)
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    st0 = {k: v.copy() for k, v in h.state().items()}
    _, _, n0, _ = h.time_levels()
    h.run_subcycle()        # nstep_end = 3: divisible by state_frequency -> diagnostics on
    h.push_results()
    acc = h.accum()
    st1 = h.state()
    nstep, nm1, n0b, np1b = h.time_levels()
    phis = h.array("phis").reshape(h.nelemd, 4, 4)

    def energies(st, tl, tq):
        dp = _dp(h, st["ps_v"][:, tl])
        T, u, v = st["T"][:, tl], st["v"][:, tl, :, 0], st["v"][:, tl, :, 1]
        cps = np.full_like(dp, CP)
        if cpstar:
            cps = CP * (1.0 + (CPWV / CP - 1.0) * (st["Qdp"][:, tq, 0] / dp))
        IE = np.zeros_like(phis); KE = np.zeros_like(phis); PE = np.zeros_like(phis); IW = np.zeros_like(phis)
        for k in range(cfg.nlev):
            IE = IE + cps[:, k] * T[:, k] * dp[:, k]
            IW = IW + (cps[:, k] - CP) * T[:, k] * dp[:, k]
            KE = KE + (u[:, k] * u[:, k] + v[:, k] * v[:, k]) * 0.5 * dp[:, k]
            PE = PE + phis * dp[:, k]
        return IE, IW, KE, PE

    # before the advance (ivar 0; ivar 2 holds the same numbers since the forcing is zero)
    IE, IW, KE, PE = energies(st0, n0 - 1, 0)
    for iv in (0, 2):
        assert np.array_equal(acc["IEner"][:, iv], IE)
        assert np.array_equal(acc["KEner"][:, iv], KE)
        assert np.array_equal(acc["PEner"][:, iv], PE)
    # after the advance (ivar 1): state at the level that is n0 after the final rotation
    IE, IW, KE, PE = energies(st1, n0b - 1, 1)
    assert np.array_equal(acc["IEner"][:, 1], IE)
    assert np.array_equal(acc["KEner"][:, 1], KE)
    assert np.array_equal(acc["PEner"][:, 1], PE)
    assert np.array_equal(acc["IEner_wet"], IW)
    # tracer mass / variance after the advance: Qmass = sum_k qdp, Qvar = sum_k qdp Q
    qdp, Q = st1["Qdp"][:, 1, :cfg.qsize], st1["Q"][:, :cfg.qsize]
    qm = np.zeros_like(qdp[:, :, 0]); qv = np.zeros_like(qm)
    for k in range(cfg.nlev):
        qv = qv + qdp[:, :, k] * Q[:, :, k]
        qm = qm + qdp[:, :, k]
    assert np.array_equal(acc["Qmass"][:, 1, :cfg.qsize], qm)
    assert np.array_equal(acc["Qvar"][:, 1, :cfg.qsize], qv)
    assert np.array_equal(acc["Q1mass"][:, :cfg.qsize], qm)
    h.close()
