"""Checkpoint / resume through the C ABI: cxx_push_results_to_f90 -> file -> init_elements_states_c + init_time_level_c
in a new session continues BIT-IDENTICALLY, on the oracle and on the reference's own build (the property the
reference's restart files rely on: everything the next prim_run_subcycle_c reads is in elem%state and tl)."""
import pytest

import restart_check
from hommexx_b200 import homme, run as hrun
from oracle import oraclelib
from reference_lib import reference_lib

CASES = {
    "ne4": dict(),
    "prtcA-moist-q2": dict(base="prtcA", moisture=1, qsplit=2, rsplit=2),   # odd tracer time level at the checkpoint
    "prtcA-r0": dict(base="prtcA", rsplit=0),
}


@pytest.mark.parametrize("which", ["oracle", "reference"])
@pytest.mark.parametrize("case", list(CASES))
def test_restart_run_is_bit_identical(tmp_path, case, which):
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    lib = oraclelib.ORACLE_LIB if which == "oracle" else reference_lib(cfg.nlev, cfg.qsize_d)
    restart_check.restart_is_bit_identical(cfg, lib, tmp_path / "R.npz", first=3 if cfg.rsplit == 0 else 1, more=2)


def test_restart_file_is_checked_against_the_run(tmp_path):
    cfg = homme.preset("prtcA")
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    h.run_subcycle()
    h.write_restart(tmp_path / "R.npz")
    h.close()
    other = homme.Homme(homme.preset("prtcA", qsize=2), oraclelib.ORACLE_LIB)
    with pytest.raises(ValueError, match="written for"):
        other.read_restart(tmp_path / "R.npz")
    other.close()


def test_namelist_run_with_restart_files(tmp_path):
    """run(): restartfreq writes R<nstep>.npz; a runtype = 1 run from one of them ends where the long run ended."""
    import io
    cfg = homme.preset("prtcA", state_frequency=6, disable_diagnostics=0)
    dyn = cfg.qsplit * max(cfg.rsplit, 1)
    full = hrun.run(cfg, oraclelib.ORACLE_LIB, 4 * dyn, out=io.StringIO(), restartfreq=2 * dyn, restartdir=str(tmp_path))
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == [f"R{2 * dyn:09d}.npz", f"R{4 * dyn:09d}.npz"]
    again = hrun.run(cfg, oraclelib.ORACLE_LIB, 4 * dyn, out=io.StringIO(), restart_in=str(tmp_path / files[0]))
    assert again[-1]["nstep"] == full[-1]["nstep"] == 4 * dyn
    for k in ("KEner", "IEner", "PEner", "TOTE"):
        assert again[-1][k] == full[-1][k], k
    assert again[-1]["Qmass"] == full[-1]["Qmass"]
