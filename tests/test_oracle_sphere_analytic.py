"""Analytic checks of the C++ driver's cubed-sphere geometry (cube_mod.F90 restatement) together with the
oracle's sphere operators: identities that hold for the continuous operators must hold to spectral
truncation error on the mesh, and converge when the mesh is refined. Independent of the reference's
known-answer vectors (tests/test_oracle_sphere_kats.py), which pin the operators on ONE element."""
import numpy as np
import pytest

import abi
from hommexx_b200 import homme
from oracle import oraclelib

A = 6.376e6
NLEV = 26


def _mesh(ne):
    h = homme.Homme(homme.preset("prtcA", ne=ne, qsize=0), oraclelib.ORACLE_LIB)
    h.init_dycore()
    n = h.nelemd
    return h, h.array("lat").reshape(n, 16).copy(), h.array("lon").reshape(n, 16).copy()


def _apply(h, op, fields, n_out):
    """fields: [n_in][nelem][16] -> [n_out][nelem][16] (level 0 of NLEV identical levels)."""
    n = h.nelemd
    out = np.zeros((n_out, n, 16))
    for ie in range(n):
        fin = np.repeat(np.stack([f[ie] for f in fields])[:, :, None], NLEV, axis=2)
        out[:, ie] = abi.sphere_op(h.lib, op, ie, fin, n_out, NLEV)[:, :, 0]
    return out


def test_sphere_area_and_mass_matrix():
    h, lat, lon = _mesh(4)
    sph = h.array("spheremp").reshape(-1, 16)
    rsp = h.array("rspheremp").reshape(-1, 16)
    assert abs(sph.sum() - 4 * np.pi) <= 1e-12 * 4 * np.pi           # the alpha correction makes the area exact
    # rspheremp is the inverse of the ASSEMBLED mass: interior points 1/spheremp, shared points smaller
    interior = [5, 6, 9, 10]
    assert np.allclose(rsp[:, interior] * sph[:, interior], 1.0, rtol=1e-14)
    edge = [1, 2, 4, 7, 8, 11, 13, 14]
    assert (rsp[:, edge] * sph[:, edge] < 0.75).all() and (rsp[:, edge] * sph[:, edge] > 0.25).all()
    h.close()


def _errors(ne):
    h, lat, lon = _mesh(ne)
    u0 = 30.0
    # gradient of f = sin(lat) cos(lon) ... keep it pole-safe: f = a-independent smooth function of (x,y,z)
    x, y, z = np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)
    f = z + 0.5 * x * y
    g = _apply(h, "gradient_sphere", [f], 2)
    # analytic: grad f = P (df/dX) / a with P the tangential projection; components on (e_lon, e_lat)
    dfdx, dfdy, dfdz = 0.5 * y, 0.5 * x, np.ones_like(z)
    e_lon = (-np.sin(lon), np.cos(lon), np.zeros_like(lon))
    e_lat = (-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat))
    g_lon = (dfdx * e_lon[0] + dfdy * e_lon[1] + dfdz * e_lon[2]) / A
    g_lat = (dfdx * e_lat[0] + dfdy * e_lat[1] + dfdz * e_lat[2]) / A
    pole_safe = np.abs(lat) < 1.5
    e_grad = max(np.abs(g[0] - g_lon)[pole_safe].max(), np.abs(g[1] - g_lat)[pole_safe].max()) * A
    # solid-body rotation: divergence 0, vorticity 2 u0 sin(lat) / a
    u, v = u0 * np.cos(lat), np.zeros_like(lat)
    div = _apply(h, "divergence_sphere", [u, v], 1)[0]
    vort = _apply(h, "vorticity_sphere", [u, v], 1)[0]
    e_div = np.abs(div).max() * A / u0
    e_vort = np.abs(vort - 2 * u0 * np.sin(lat) / A).max() * A / u0
    # weak Laplacian: its global integral (= sum over elements, the test function is 1) vanishes
    lap = _apply(h, "laplace_simple", [f], 1)[0]
    e_lap = abs(lap.sum()) / np.abs(lap).sum()
    h.close()
    return e_grad, e_div, e_vort, e_lap


def test_operators_converge_to_the_analytic_fields():
    coarse, fine = _errors(4), _errors(8)
    for name, c, f in zip(("gradient", "divergence", "vorticity"), coarse, fine):
        assert f < 3e-3, (name, f)                # ne8, np4: relative truncation error (measured 5e-4 .. 2e-3)
        assert f < c / 5.0, (name, c, f)          # ~third order under mesh refinement (measured ratios 5.9 .. 8.4)
    assert coarse[3] < 1e-12 and fine[3] < 1e-12  # conservation of the weak Laplacian to round-off
