"""Pins the oracle's PPM vertical remap (PpmRemap.hpp restatement) BITWISE against the
reference's own plain-C++ remap twin, src/preqx/unit_tests/remap.cpp, compiled in place from
/root/reference into oracle/_ref/libref_remap_<nlev>.so by oracle/Makefile. The reference asserts
that twin bitwise equal to both its Fortran and its Kokkos remap (preqx_ut_remap.cpp:122,209,359)."""
import ctypes as C
import pathlib

import numpy as np
import pytest

from hommexx_b200 import homme
from oracle import oraclelib

REF_DIR = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_ref"


def _ref(nlev):
    p = REF_DIR / f"libref_remap_{nlev}.so"
    if not p.exists():
        pytest.skip(f"{p} not built (needs the reference tree at build time)")
    lib = C.CDLL(str(p))
    assert lib.ref_remap_nlev() == nlev
    lib.ref_remap_Q_ppm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def _problem(nlev, qsize, seed):
    rng = np.random.default_rng(seed)
    dp1 = rng.uniform(0.5, 2.0, (nlev, 4, 4))            # source thickness, [lev][j][i]
    w = rng.uniform(0.5, 2.0, (nlev, 4, 4))
    dp2 = w / w.sum(0) * dp1.sum(0)                       # same column mass, different partition
    q = rng.uniform(0.0, 1.0, (qsize, nlev, 4, 4)) * dp1  # Qdp
    return dp1, dp2, q


@pytest.mark.parametrize("nlev", [72, 26, 8])
@pytest.mark.parametrize("alg", [1, 2])
def test_remap_bitwise_vs_reference_twin(nlev, alg):
    ref = _ref(nlev)
    ora = oraclelib.load_oracle(nlev, 4)
    for seed in range(3):
        qsize = 5
        dp1, dp2, q = _problem(nlev, qsize, 100 * nlev + seed)
        qref = q.copy()
        ref.ref_remap_Q_ppm(qref.ctypes.data, qsize, dp1.ctypes.data, dp2.ctypes.data, alg)
        # oracle layout: columns [ncol][nlev], fields [nf][ncol][nlev]
        src = np.ascontiguousarray(dp1.reshape(nlev, 16).T)
        tgt = np.ascontiguousarray(dp2.reshape(nlev, 16).T)
        f = np.ascontiguousarray(q.reshape(qsize, nlev, 16).transpose(0, 2, 1))
        ora.hxx_remap_columns(alg, 16, qsize, src.ctypes.data, tgt.ctypes.data, f.ctypes.data)
        got = f.transpose(0, 2, 1).reshape(qsize, nlev, 4, 4)
        assert np.array_equal(got, qref), np.abs(got - qref).max()
        # conservation to round-off: column mass of every field
        assert np.allclose(got.sum(1), q.sum(1), rtol=1e-13, atol=0)


def test_remap_identity_when_grids_match():
    nlev = 26
    ora = oraclelib.load_oracle(nlev, 4)
    dp1, _, q = _problem(nlev, 3, 7)
    src = np.ascontiguousarray(dp1.reshape(nlev, 16).T)
    f = np.ascontiguousarray(q.reshape(3, nlev, 16).transpose(0, 2, 1))
    f0 = f.copy()
    ora.hxx_remap_columns(1, 16, 3, src.ctypes.data, src.ctypes.data, f.ctypes.data)
    assert np.allclose(f, f0, rtol=1e-12)
