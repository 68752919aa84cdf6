"""DSS / boundary-exchange checks of the oracle on the real cubed-sphere connectivity, after
test_execs/share_ut/boundary_exchange_ut.cpp (tol 1e-13): shared GLL points end up with one
value on every element that owns them, the exchange of a C0 field with rspheremp-weighting is
the identity, global sums are conserved, and the min/max exchange equals a brute-force
neighbourhood min/max."""
import numpy as np
import pytest

from hommexx_b200 import homme
from oracle import oraclelib


@pytest.fixture(scope="module")
def sess():
    cfg = homme.preset("ne4", ne=3, nlev=8, qsize=2, qsize_d=2, vcoord="")
    h = homme.Homme(cfg, oraclelib.ORACLE_LIB)
    h.init_dycore()
    yield h
    h.close()


def _points(h):
    lat, lon = h.array("lat"), h.array("lon")
    xyz = np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], 1)
    key = np.round(xyz * 1e9).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    return inv.reshape(-1)  # global point id per (elem, pt)


def test_dss_makes_fields_continuous_and_conserves(sess):
    h = sess
    n, nlev = h.nelemd, h.cfg.nlev
    rng = np.random.default_rng(0)
    sph = h.array("spheremp").reshape(n, 16, 1)
    f = rng.standard_normal((n, 3, 16, nlev))
    t = f.copy()
    t[:, 2] *= sph                                   # field(np1=2) = spheremp * f
    h.set_field("t", t)
    h.lib.hxx_exchange(b"caar:2", 1)
    out = h.get_field("t").reshape(n, 3, 16, nlev)
    assert np.array_equal(out[:, :2], f[:, :2])      # other time levels untouched
    pid = _points(h)
    o2 = out[:, 2].reshape(n * 16, nlev)
    for g in np.unique(pid):
        rows = o2[pid == g]
        assert np.abs(rows - rows[0]).max() <= 1e-13 * np.abs(rows[0]).max()
    # integral conservation: sum(spheremp * f) before == after
    before = (sph * f[:, 2]).sum(); after = (sph * out[:, 2]).sum()
    assert abs(before - after) <= 1e-12 * np.abs(sph * f[:, 2]).sum()


def test_dss_of_continuous_field_is_identity(sess):
    h = sess
    n, nlev = h.nelemd, h.cfg.nlev
    lat, lon = h.array("lat").reshape(n, 16, 1), h.array("lon").reshape(n, 16, 1)
    sph = h.array("spheremp").reshape(n, 16, 1)
    c0 = np.sin(lat) * np.cos(2 * lon) * np.cos(lat) + 0.1 * np.arange(nlev).reshape(1, 1, nlev)
    t = np.zeros((n, 3, 16, nlev)); t[:, 0] = c0 * sph
    h.set_field("t", t)
    h.lib.hxx_exchange(b"caar:0", 1)
    out = h.get_field("t").reshape(n, 3, 16, nlev)[:, 0]
    assert np.abs(out - c0).max() <= 1e-13 * np.abs(c0).max()


def test_minmax_exchange_is_neighbourhood_minmax(sess):
    h = sess
    n, nlev, q = h.nelemd, h.cfg.nlev, h.cfg.qsize_d
    rng = np.random.default_rng(1)
    ql = rng.standard_normal((n, q, 2, nlev)); ql[:, :, 1] += 3.0
    h.set_field("qlim", ql)
    h.lib.hxx_exchange(b"qlim", 0)
    out = h.get_field("qlim").reshape(n, q, 2, nlev)
    conn = h.connections()                            # 1-based (lid1,gid1,pos1,pid1,lid2,...)
    nbrs = [[i] for i in range(n)]
    for c in conn:
        nbrs[c[0] - 1].append(c[4] - 1)
    for i in range(n):
        assert np.array_equal(out[i, :, 0], ql[nbrs[i], :, 0].min(0))
        assert np.array_equal(out[i, :, 1], ql[nbrs[i], :, 1].max(0))
    assert sorted(len(b) - 1 for b in nbrs).count(7) == 24   # the 24 cube-vertex elements miss one corner
