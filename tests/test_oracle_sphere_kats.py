"""Pins the oracle's SphereOperators restatement against the reference's own known-answer
vectors (test/unit_tests/inputs/*_sphere_np4.in -> tests/golden/sphere_kats.json, imported by
scripts/import_sphere_kats.py). The vectors come from the Fortran operators
(src/share/derivative_mod_base.F90:1020-1052, :1135-1180, :1232-1290), which the reference's C++
unit tests accept to a relative error of a few hundred eps (preqx_ut_sphere_op_ml.cpp)."""
import json
import pathlib

import numpy as np
import pytest

import abi
from hommexx_b200 import homme
from oracle import oraclelib

KATS = json.loads((pathlib.Path(__file__).parent / "golden" / "sphere_kats.json").read_text())
NLEV = 8


@pytest.fixture()
def lib():
    lib = oraclelib.load_oracle(NLEV, 4)
    yield lib
    lib.finalize_hommexx_session()


def _f90_tensor(flat):
    # F90 X(np,np,2,2) column-major == C [c][r][j][i]: exactly the memory image the ABI takes
    return np.asarray(flat, dtype=np.float64).reshape(1, 2, 2, 4, 4)


def _levels(field16):
    return np.repeat(np.asarray(field16, dtype=np.float64).reshape(-1, 16, 1), NLEV, axis=2)


def test_gradient_sphere_kat(lib):
    k = KATS["gradient"]
    dinv = _f90_tensor(k["elem_Dinv"])
    abi.isolated_elements_session(lib, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv, metdet=1.0)
    out = abi.sphere_op(lib, "gradient_sphere", 0, _levels(k["s"]), 2, NLEV)
    ref = np.asarray(k["Gradient_Sphere_result"]).reshape(2, 16)
    # bit-for-bit, including the first component, a cancellation to ~1e-8 of terms ~1e-3
    for lev in range(NLEV):
        assert np.array_equal(out[:, :, lev], ref)


def test_divergence_sphere_kat(lib):
    k = KATS["divergence"]
    dinv = _f90_tensor(k["elem_Dinv"])
    metdet = np.reshape(k["elem_metdet"], (1, 4, 4))
    abi.isolated_elements_session(lib, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv, metdet=metdet)
    v = np.asarray(k["v"]).reshape(2, 16)
    out = abi.sphere_op(lib, "divergence_sphere", 0, _levels(v), 1, NLEV)
    ref = np.asarray(k["Divergence_Sphere_result"])
    # the result (~1e-11) is the cancellation residue of terms ~1e-6: it still matches bit-for-bit
    assert np.array_equal(out[0, :, 0], ref)


def test_vorticity_sphere_kat(lib):
    k = KATS["vorticity"]
    d = _f90_tensor(k["elem_D"])
    metdet = 1.0 / np.reshape(k["elem_rmetdet"], (1, 4, 4))
    abi.isolated_elements_session(lib, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=d, Dinv=d, metdet=metdet)
    v = np.asarray(k["v"]).reshape(2, 16)
    out = abi.sphere_op(lib, "vorticity_sphere", 0, _levels(v), 1, NLEV)
    ref = np.asarray(k["Vorticity_Sphere_result"])
    # metdet is rebuilt as 1/rmetdet (the Fortran multiplies by rmetdet): allow 4 ulp
    assert np.abs(out[0, :, 0] - ref).max() <= 4 * np.finfo(float).eps * np.abs(ref).max()


def test_dvv_matches_reference_table():
    """The driver's GLL derivative matrix equals the deriv_Dvv block printed by the reference."""
    h = homme.Homme(homme.preset("ne4", ne=2), oraclelib.ORACLE_LIB, init="none")
    dvv = h.array("dvv").copy()
    h.close()
    assert np.abs(dvv - np.asarray(KATS["gradient"]["deriv_Dvv"])).max() < 4e-16
