"""hommexx_b200.run: the namelist-driven stand-in for prim_main, on the oracle (CPU). Checks the namelist
mapping (namelist_mod.F90 names) and that the diagnostics the dycore writes give conserved global integrals."""
import io
import pathlib

import numpy as np

from hommexx_b200 import homme, run
from oracle import oraclelib

NL = pathlib.Path(run.__file__).parent / "namelists"


def test_namelist_mapping():
    nl = run.parse_namelist((NL / "homme-ne30-v1.nl").read_text())
    cfg = run.config_from_namelist(nl)
    ref = homme.preset("ne30")
    for k in ("ne", "nlev", "qsize", "qsize_d", "rsplit", "qsplit", "tstep", "nu", "nu_p", "nu_q", "nu_s", "nu_div",
              "nu_top", "hypervis_subcycle", "limiter_option", "remap_alg", "ftype", "time_step_type",
              "state_frequency", "disable_diagnostics", "hypervis_scaling"):
        assert getattr(cfg, k) == getattr(ref, k), k
    small = run.config_from_namelist(run.parse_namelist((NL / "prtcA-r3-dry.nl").read_text()))
    assert (small.ne, small.nlev, small.qsize, small.qsize_d, small.state_frequency) == (4, 26, 4, 4, 3)
    assert small.disable_diagnostics == 0 and small.nu == 7e15 and small.nu_s == -1.0
    # Fortran spellings
    g = run.parse_namelist("&ctl_nl\n a = .true., b=1.5d3 ! c\n s = 'x,y'\n/\n")["ctl_nl"]
    assert g == {"a": True, "b": 1500.0, "s": "x,y"}


def test_prtcA_run_prints_conserved_integrals():
    nl = run.parse_namelist((NL / "prtcA-r3-dry.nl").read_text())
    cfg = run.config_from_namelist(nl)
    buf = io.StringIO()
    hist = run.run(cfg, oraclelib.ORACLE_LIB, nmax=12, out=buf)
    assert [h["nstep"] for h in hist] == [3, 6, 9, 12]
    text = buf.getvalue()
    assert "TOTE" in text and "Q1  mass" in text
    tote = np.array([h["TOTE"] for h in hist])
    assert np.isfinite(tote).all() and tote.min() > 1e9               # ~2.5e9 J/m^2 for an Earth-like atmosphere
    assert np.abs(tote / tote[0] - 1.0).max() < 1e-5                   # adiabatic, a few steps: energy nearly conserved
    qm = np.array([h["Qmass"] for h in hist])
    assert np.abs(qm / qm[0] - 1.0).max() < 1e-12                      # tracer mass conserved to round-off
