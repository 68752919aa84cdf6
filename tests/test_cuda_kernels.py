"""GPU parity of the building blocks, through the C ABI, against (a) the reference's known-answer
vectors, (b) the reference's own PPM twin compiled into oracle/_ref, (c) the CPU oracle on the
same seeded inputs. Integer/index work and — because the kernels keep the reference's operation
order and are compiled with --fmad=false — all floating-point results are required BIT-EXACT."""
import ctypes as C
import json
import pathlib

import numpy as np
import pytest

import abi
import parity
from hommexx_b200 import homme
from oracle import oraclelib
from limiter_problems import EPS, check_limited, feasible_problem, run_limiter

pytestmark = pytest.mark.gpu
KATS = json.loads((pathlib.Path(__file__).parent / "golden" / "sphere_kats.json").read_text())
NLEV = 26  # the (26, 4) build: PLEV of the reference's prtcA executables


@pytest.fixture()
def cuda():
    parity.need_gpu()
    lib = homme.load_dycore(parity.cuda_lib(NLEV, 4))
    yield lib
    lib.finalize_hommexx_session()


@pytest.fixture()
def oracle():
    lib = oraclelib.load_oracle(NLEV, 4)
    yield lib
    lib.finalize_hommexx_session()


def _f90_tensor(flat):
    return np.asarray(flat, dtype=np.float64).reshape(1, 2, 2, 4, 4)


def _levels(field16):
    return np.repeat(np.asarray(field16, dtype=np.float64).reshape(-1, 16, 1), NLEV, axis=2)


# ---- (a) the reference's golden vectors, test/unit_tests/inputs/*_sphere_np4.in ---------------
def test_gradient_sphere_kat(cuda):
    k = KATS["gradient"]
    dinv = _f90_tensor(k["elem_Dinv"])
    abi.isolated_elements_session(cuda, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv, metdet=1.0)
    out = abi.sphere_op(cuda, "gradient_sphere", 0, _levels(k["s"]), 2, NLEV)
    ref = np.asarray(k["Gradient_Sphere_result"]).reshape(2, 16)
    for lev in range(NLEV):
        assert np.array_equal(out[:, :, lev], ref)


def test_divergence_sphere_kat(cuda):
    k = KATS["divergence"]
    dinv = _f90_tensor(k["elem_Dinv"])
    metdet = np.reshape(k["elem_metdet"], (1, 4, 4))
    abi.isolated_elements_session(cuda, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=dinv, Dinv=dinv, metdet=metdet)
    out = abi.sphere_op(cuda, "divergence_sphere", 0, _levels(np.asarray(k["v"]).reshape(2, 16)), 1, NLEV)
    assert np.array_equal(out[0, :, 0], np.asarray(k["Divergence_Sphere_result"]))
    assert np.array_equal(out[0, :, NLEV - 1], np.asarray(k["Divergence_Sphere_result"]))


def test_vorticity_sphere_kat(cuda):
    k = KATS["vorticity"]
    d = _f90_tensor(k["elem_D"])
    metdet = 1.0 / np.reshape(k["elem_rmetdet"], (1, 4, 4))
    abi.isolated_elements_session(cuda, 1, NLEV, np.reshape(k["deriv_Dvv"], (4, 4)), D=d, Dinv=d, metdet=metdet)
    out = abi.sphere_op(cuda, "vorticity_sphere", 0, _levels(np.asarray(k["v"]).reshape(2, 16)), 1, NLEV)
    ref = np.asarray(k["Vorticity_Sphere_result"])
    assert np.abs(out[0, :, 0] - ref).max() <= 4 * EPS * np.abs(ref).max()


# ---- (c) every sphere operator vs the oracle on random data, bit-exact -------------------------
@pytest.mark.parametrize("op,n_in,n_out", [("gradient_sphere", 1, 2), ("divergence_sphere", 2, 1),
                                           ("vorticity_sphere", 2, 1), ("laplace_simple", 1, 1),
                                           ("divergence_sphere_wk", 2, 1), ("vlaplace_sphere_wk_contra", 2, 2)])
def test_sphere_ops_match_oracle_bitwise(cuda, oracle, op, n_in, n_out):
    rng = np.random.default_rng(11)
    n = 3
    D = rng.uniform(0.5, 1.5, (n, 2, 2, 4, 4)); D[:, 0, 1] *= 0.1; D[:, 1, 0] *= 0.1
    det = D[:, 0, 0] * D[:, 1, 1] - D[:, 0, 1] * D[:, 1, 0]
    Dinv = np.stack([np.stack([D[:, 1, 1], -D[:, 0, 1]], 1), np.stack([-D[:, 1, 0], D[:, 0, 0]], 1)], 1) / det[:, None, None]
    kw = dict(D=D, Dinv=Dinv, metdet=np.abs(det), metinv=rng.uniform(0.5, 1.5, (n, 2, 2, 4, 4)),
              mp=rng.uniform(0.1, 1.0, (n, 4, 4)))
    dvv = np.reshape(KATS["gradient"]["deriv_Dvv"], (4, 4))
    for lib in (cuda, oracle):
        abi.isolated_elements_session(lib, n, NLEV, dvv, **kw)
    x = rng.standard_normal((n_in, 16, NLEV))
    for ie in range(n):
        for nu_ratio in ((1.0, 2.5) if op.startswith("vlaplace") else (1.0,)):
            a = abi.sphere_op(cuda, op, ie, x, n_out, NLEV, nu_ratio)
            b = abi.sphere_op(oracle, op, ie, x, n_out, NLEV, nu_ratio)
            assert np.array_equal(a, b), (op, ie, float(np.abs(a - b).max()))


# ---- limiters: the reference's property tests + bit-exact vs the oracle ------------------------
@pytest.mark.parametrize("nlev", [72, 26])
@pytest.mark.parametrize("option", [8, 9])
def test_limiter_properties_and_oracle_parity(nlev, option):
    parity.need_gpu()
    cu = homme.load_dycore(parity.cuda_lib(nlev, 4))
    cu.initialize_hommexx_session()
    ora = oraclelib.load_oracle(nlev, 4)
    rng = np.random.default_rng(9)
    for seed in range(4):
        sph, dpm, pt, ql, mass = feasible_problem(50, nlev, 2000 + seed)
        if seed == 3:  # infeasible / negative-minimum limits exercise the relaxation branch
            ql[:, 0] -= rng.uniform(0.0, 0.6, ql[:, 0].shape)
            ql[:, 1] *= rng.uniform(0.3, 1.0, ql[:, 1].shape)
        a, qa = run_limiter(cu, option, sph, dpm, pt, ql)
        b, qb = run_limiter(ora, option, sph, dpm, pt, ql)
        assert np.array_equal(a, b) and np.array_equal(qa, qb)
        if seed < 3:
            check_limited(sph, dpm, a, qa, mass)
    cu.finalize_hommexx_session()


# ---- PPM remap: bit-exact vs the reference's own twin and vs the oracle ------------------------
@pytest.mark.parametrize("nlev", [72, 26])
@pytest.mark.parametrize("alg", [1, 2])
def test_remap_columns_bitwise(nlev, alg):
    parity.need_gpu()
    cu = homme.load_dycore(parity.cuda_lib(nlev, 4))
    cu.initialize_hommexx_session()
    ora = oraclelib.load_oracle(nlev, 4)
    ref_path = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_ref" / f"libref_remap_{nlev}.so"
    ref = C.CDLL(str(ref_path)) if ref_path.exists() else None
    rng = np.random.default_rng(100 + nlev)
    nf, ncol = 5, 16 * 7
    dp1 = rng.uniform(0.5, 2.0, (ncol, nlev))
    w = rng.uniform(0.9, 1.1, (ncol, nlev)) * dp1          # target grid: moved by less than a layer
    dp2 = w / w.sum(1, keepdims=True) * dp1.sum(1, keepdims=True)
    f = rng.uniform(0.0, 1.0, (nf, ncol, nlev)) * dp1
    fa, fb = f.copy(), f.copy()
    cu.hxx_remap_columns(alg, ncol, nf, dp1.ctypes.data, dp2.ctypes.data, fa.ctypes.data)
    ora.hxx_remap_columns(alg, ncol, nf, dp1.ctypes.data, dp2.ctypes.data, fb.ctypes.data)
    assert np.array_equal(fa, fb), float(np.abs(fa - fb).max())
    assert np.allclose(fa.sum(2), f.sum(2), rtol=1e-13, atol=0)   # column mass conserved to round-off
    if ref is not None:
        ref.ref_remap_Q_ppm.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        for e in range(ncol // 16):
            sl = slice(16 * e, 16 * e + 16)
            q = np.ascontiguousarray(f[:, sl].transpose(0, 2, 1)).reshape(nf, nlev, 4, 4)
            d1 = np.ascontiguousarray(dp1[sl].T).reshape(nlev, 4, 4)
            d2 = np.ascontiguousarray(dp2[sl].T).reshape(nlev, 4, 4)
            ref.ref_remap_Q_ppm(q.ctypes.data, nf, d1.ctypes.data, d2.ctypes.data, alg)
            assert np.array_equal(q.reshape(nf, nlev, 16).transpose(0, 2, 1), fa[:, sl])
    cu.finalize_hommexx_session()


# ---- DSS / min-max exchange on the real cubed-sphere connectivity -------------------------------
@pytest.fixture(scope="module")
def mesh_pair():
    cfg = homme.preset("prtcA", ne=3)
    hc, ho = parity.pair(cfg)
    yield hc, ho
    hc.close(); ho.close()


@pytest.mark.parametrize("fset,rsp", [("caar:0", 1), ("caar:2", 0), ("hv", 1), ("hv", 0), ("euler:1:0", 1),
                                      ("euler:0:1", 1), ("euler:0:2", 1), ("qtens", 1), ("qlim", 0)])
def test_exchange_matches_oracle_bitwise(mesh_pair, fset, rsp):
    hc, ho = mesh_pair
    rng = np.random.default_rng(5)
    for name in ("v", "t", "dp3d", "vtens", "ttens", "dptens", "qdp", "qtens_biharmonic", "eta_dot_dpdn",
                 "omega_p", "divdp_proj", "qlim"):
        x = rng.standard_normal(ho.field_size(name))
        hc.set_field(name, x); ho.set_field(name, x)
    hc.lib.hxx_exchange(fset.encode(), rsp)
    ho.lib.hxx_exchange(fset.encode(), rsp)
    parity.compare_fields(hc, ho, ["v", "t", "dp3d", "vtens", "ttens", "dptens", "qdp", "qtens_biharmonic",
                                   "eta_dot_dpdn", "omega_p", "divdp_proj", "qlim"], tol=0.0, what=fset)


def test_dss_continuity_and_conservation(mesh_pair):
    """boundary_exchange_ut.cpp: after DSS*rspheremp of spheremp*f every sharer holds one value
    (to 1e-13) and the global integral is conserved."""
    hc, _ = mesh_pair
    n, nlev = hc.nelemd, hc.cfg.nlev
    rng = np.random.default_rng(0)
    sph = hc.array("spheremp").reshape(n, 16, 1)
    f = rng.standard_normal((n, 3, 16, nlev))
    t = f.copy(); t[:, 2] *= sph
    hc.set_field("t", t)
    hc.lib.hxx_exchange(b"caar:2", 1)
    out = hc.get_field("t").reshape(n, 3, 16, nlev)
    assert np.array_equal(out[:, :2], f[:, :2])
    lat, lon = hc.array("lat"), hc.array("lon")
    xyz = np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], 1)
    _, pid = np.unique(np.round(xyz * 1e9).astype(np.int64), axis=0, return_inverse=True)
    pid = pid.reshape(-1)
    o2 = out[:, 2].reshape(n * 16, nlev)
    for gpt in np.unique(pid):
        rows = o2[pid == gpt]
        assert np.abs(rows - rows[0]).max() <= 1e-13 * np.abs(rows[0]).max()
    before = (sph * f[:, 2]).sum(); after = (sph * out[:, 2]).sum()
    assert abs(before - after) <= 1e-12 * np.abs(sph * f[:, 2]).sum()
