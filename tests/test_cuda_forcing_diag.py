"""GPU parity of the CAM forcing pass and of the diagnostics against the CPU oracle, through the C
ABI (f90_push_forcing_to_cxx, prim_run_subcycle_c, the init_diagnostics_c arrays). Bit-identical."""
import numpy as np
import pytest

import parity
from forcing_inputs import fill_forcing, fill_smooth_forcing
from hommexx_b200 import homme

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("moist,ftype", [(0, 0), (1, 0), (0, 2)])
def test_forcing_pass_parity(moist, ftype):
    cfg = homme.preset("ne4", moisture=moist, ftype=ftype)
    hc, ho = parity.pair(cfg)
    for h in (hc, ho):
        fill_forcing(h)
        h.push_forcing()
        h.lib.hxx_apply_forcing(cfg.tstep * cfg.rsplit)
    parity.compare_fields(hc, ho, parity.STATE_FIELDS + ["fm", "ft"] + (["fq"] if ftype == 0 else []), tol=0.0,
                          what=f"forcing moist={moist} ftype={ftype}")
    # round trip of the forcing arrays and the Qdp push-back
    for h in (hc, ho):
        for a in h.forcing().values():
            a[...] = 0.0
        h.pull_forcing()
    for k in ("FM", "FT") + (("FQ",) if ftype == 0 else ()):
        assert np.array_equal(hc.forcing()[k], ho.forcing()[k]), k
        assert np.abs(hc.forcing()[k]).max() > 0
    hc.close(); ho.close()


def test_forced_run_parity():
    """Three forced subcycle calls (moist, ftype 0, forcing re-pushed every call as the CAM wrapper does)."""
    cfg = homme.preset("ne4", moisture=1, ftype=0)
    hc, ho = parity.pair(cfg)
    for h in (hc, ho):
        fill_smooth_forcing(h)
    for _ in range(3):
        for h in (hc, ho):
            h.push_forcing()
            h.run_subcycle()
    # vtens/ttens are scratch (see test_hypervis_parity)
    names = [f for f in parity.STATE_FIELDS if f not in ("vtens", "ttens")]
    parity.compare_fields(hc, ho, names, tol=0.0, what="forced run")
    for h in (hc, ho):
        h.push_results()
    for k in parity.PROGNOSTIC:
        assert np.array_equal(hc.state()[k], ho.state()[k]), k
    hc.close(); ho.close()


@pytest.mark.parametrize("cpstar", [0, 1])
def test_diagnostics_parity(cpstar):
    cfg = homme.preset("ne4", disable_diagnostics=0, use_cpstar=cpstar, state_frequency=3, moisture=cpstar)
    hc, ho = parity.pair(cfg)
    for _ in range(2):
        for h in (hc, ho):
            h.run_subcycle()
    a, b = hc.accum(), ho.accum()
    for k in b:
        assert np.array_equal(a[k], b[k]), k
        assert np.isfinite(b[k]).all()
    assert np.abs(b["KEner"]).max() > 0 and np.abs(b["Qmass"]).max() > 0
    assert np.array_equal(hc.state()["Q"], ho.state()["Q"])  # prim_diag_scalars syncs Q to the F90 array
    # last_time_step also switches the diagnostics on (prim_driver.cpp:55-57)
    for h in (hc, ho):
        for arr in h.accum().values():
            arr[...] = 0.0
        h.set_last_step(7)
        h.run_subcycle()
    a, b = hc.accum(), ho.accum()
    for k in b:
        assert np.array_equal(a[k], b[k]), k
    assert np.abs(b["IEner"]).max() > 0
    hc.close(); ho.close()


def test_held_suarez_device_forcing_parity():
    """BASELINE configs[4]'s forcing: hxx_held_suarez_forcing evaluated on the GPU against the oracle's C twin (and
    through it the numpy formulas of tests/test_oracle_held_suarez.py), then six forced prim_run_subcycle_c calls.
    log / exp / sin come from different math libraries on the two sides, so FT is held to 1e-13 instead of bit for bit;
    the forced run to the north-star tolerance."""
    import numpy as np
    cfg = homme.preset("ne4", ftype=0)
    hc, ho = parity.pair(cfg)
    for h in (hc, ho):
        h.run_subcycle()
        h.held_suarez_forcing()
    for name in ("fm", "ft"):
        a, b = hc.get_field(name), ho.get_field(name)
        assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max(), name
    for _ in range(6):
        for h in (hc, ho):
            h.held_suarez_forcing()
            h.run_subcycle()
    hc.push_results(); ho.push_results()
    sc, so = hc.state(), ho.state()
    errs = {k: parity.rel_l2(sc[k], so[k]) for k in ("v", "T", "dp3d", "ps_v", "Qdp", "Q")}
    print("held-suarez forced run, rel-L2 vs oracle:", errs)
    assert max(errs.values()) <= 1e-11, errs
    hc.close(); ho.close()
