"""The oracle's sphere operators against the reference's own SphereOperators.hpp, operator by operator: the reference
build (oracle/_ref, tests/reference_lib.py) exports hxx_sphere_op bound to SphereOperators::{gradient_sphere,
divergence_sphere, vorticity_sphere, laplace_simple, divergence_sphere_wk, vlaplace_sphere_wk_contra} run by one team
on caller-provided fields (oracle/ref_hommexx_api.cpp). Same isolated-element sessions as the reference's functor
unit tests (random metric terms, no connections): bit-identical. And the reference's golden vectors
(test/unit_tests/inputs/*_sphere_np4.in, tests/golden/sphere_kats.json) are reproduced by the reference build itself,
which checks the stand-in Kokkos the build runs on."""
import json
import pathlib

import numpy as np
import pytest

import abi
from hommexx_b200 import homme
from oracle import oraclelib
from reference_lib import reference_lib

KATS = json.loads((pathlib.Path(__file__).parent / "golden" / "sphere_kats.json").read_text())
NLEV = 26


@pytest.fixture()
def libs():
    ref = homme.load_dycore(reference_lib(NLEV, 4))
    ora = oraclelib.load_oracle(NLEV, 4)
    yield ref, ora
    ref.finalize_hommexx_session()
    ora.finalize_hommexx_session()


@pytest.mark.parametrize("op,n_in,n_out", [("gradient_sphere", 1, 2), ("divergence_sphere", 2, 1),
                                           ("vorticity_sphere", 2, 1), ("laplace_simple", 1, 1),
                                           ("divergence_sphere_wk", 2, 1), ("vlaplace_sphere_wk_contra", 2, 2)])
def test_sphere_operator_is_bit_identical_to_the_reference(libs, op, n_in, n_out):
    ref, ora = libs
    rng = np.random.default_rng(23)
    n = 4
    D = rng.uniform(0.5, 1.5, (n, 2, 2, 4, 4)); D[:, 0, 1] *= 0.1; D[:, 1, 0] *= 0.1
    det = D[:, 0, 0] * D[:, 1, 1] - D[:, 0, 1] * D[:, 1, 0]
    Dinv = np.stack([np.stack([D[:, 1, 1], -D[:, 0, 1]], 1), np.stack([-D[:, 1, 0], D[:, 0, 0]], 1)], 1) / det[:, None, None]
    kw = dict(D=D, Dinv=Dinv, metdet=np.abs(det), metinv=rng.uniform(0.5, 1.5, (n, 2, 2, 4, 4)),
              mp=rng.uniform(0.1, 1.0, (n, 4, 4)))
    dvv = np.reshape(KATS["gradient"]["deriv_Dvv"], (4, 4))
    for lib in (ref, ora):
        abi.isolated_elements_session(lib, n, NLEV, dvv, **kw)
    x = rng.standard_normal((n_in, 16, NLEV))
    for ie in range(n):
        for nu_ratio in ((1.0, 2.5) if op.startswith("vlaplace") else (1.0,)):
            a = abi.sphere_op(ref, op, ie, x, n_out, NLEV, nu_ratio)
            b = abi.sphere_op(ora, op, ie, x, n_out, NLEV, nu_ratio)
            assert np.abs(a).max() > 0
            assert np.array_equal(a, b), (op, ie, float(np.abs(a - b).max()))


def _levels(field16):
    return np.repeat(np.asarray(field16, dtype=np.float64).reshape(-1, 16, 1), NLEV, axis=2)


def test_the_reference_build_reproduces_its_own_golden_vectors(libs):
    ref, _ = libs
    k = KATS["gradient"]
    dinv = np.asarray(k["elem_Dinv"], dtype=np.float64).reshape(1, 2, 2, 4, 4)
    abi.isolated_elements_session(ref, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv, metdet=1.0)
    out = abi.sphere_op(ref, "gradient_sphere", 0, _levels(k["s"]), 2, NLEV)
    want = np.asarray(k["Gradient_Sphere_result"]).reshape(2, 16)
    for lev in range(NLEV):
        assert np.array_equal(out[:, :, lev], want)
    ref.finalize_hommexx_session()
    k = KATS["divergence"]
    dinv = np.asarray(k["elem_Dinv"], dtype=np.float64).reshape(1, 2, 2, 4, 4)
    abi.isolated_elements_session(ref, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv,
                                  metdet=np.reshape(k["elem_metdet"], (1, 4, 4)))
    out = abi.sphere_op(ref, "divergence_sphere", 0, _levels(np.asarray(k["v"]).reshape(2, 16)), 1, NLEV)
    assert np.array_equal(out[0, :, 0], np.asarray(k["Divergence_Sphere_result"]))
    assert np.array_equal(out[0, :, NLEV - 1], np.asarray(k["Divergence_Sphere_result"]))
    ref.finalize_hommexx_session()
    k = KATS["vorticity"]   # written by the Fortran with 1 / metdet precomputed: equal to a few ulp, not bitwise
    d = np.asarray(k["elem_D"], dtype=np.float64).reshape(1, 2, 2, 4, 4)
    abi.isolated_elements_session(ref, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=d, Dinv=d,
                                  metdet=1.0 / np.reshape(k["elem_rmetdet"], (1, 4, 4)))
    out = abi.sphere_op(ref, "vorticity_sphere", 0, _levels(np.asarray(k["v"]).reshape(2, 16)), 1, NLEV)
    want = np.asarray(k["Vorticity_Sphere_result"])
    assert np.abs(out[0, :, 0] - want).max() <= 4 * np.finfo(float).eps * np.abs(want).max()
