"""Held-Suarez (1994) forcing as a HOST-side harness (BASELINE configs[4], SURVEY.md 8d caveat C3).

In HOMME the Held-Suarez physics is Fortran (`physics/heldsuarez/held_suarez_mod.F90:37-279`): it fills
`elem%derived%FM / FT` and the dycore applies them as CAM forcing (`ftype = 0`, CamForcing.cpp:20-49). The
dycore library does the second half; this module restates the first half with numpy on the driver's
Fortran-layout arrays — Newtonian relaxation of T towards T_eq(lat, p) and Rayleigh friction on the winds
below sigma_b — and hands the result to the library through the reference's own entry point
(`f90_push_forcing_to_cxx`). Host glue only: no arithmetic of the hot path happens here.
"""
from __future__ import annotations

import numpy as np

SECPDAY = 86400.0
SIGMA_B = 0.70
K_A = 1.0 / (40.0 * SECPDAY)
K_F = 1.0 / (1.0 * SECPDAY)
K_S = 1.0 / (4.0 * SECPDAY)
DT_Y = 60.0
DTHETA_Z = 10.0
P0 = 1.0e5
KAPPA = 287.04 / 1005.0


def hs_T_forcing(hyam, hybm, ps, T, lat):
    """held_suarez_mod.F90:175-279. ps, lat [n,4,4]; T [n,nlev,4,4] -> FT [n,nlev,4,4]."""
    snlatsq = np.sin(lat) ** 2
    cslatsq = 1.0 - snlatsq
    p = hyam[None, :, None, None] * P0 + hybm[None, :, None, None] * ps[:, None]
    logprat = np.log(p) - np.log(P0)
    pratk = np.exp(KAPPA * logprat)
    etam = (hyam + hybm)[None, :, None, None]
    k_t = K_A + (K_S - K_A) * (cslatsq * cslatsq)[:, None] * np.maximum(0.0, (etam - SIGMA_B) / (1.0 - SIGMA_B))
    Teq = np.maximum(200.0, (315.0 - DT_Y * snlatsq[:, None] - DTHETA_Z * logprat * cslatsq[:, None]) * pratk)
    return -k_t * (T - Teq), Teq


def hs_v_forcing(hyam, hybm, v):
    """held_suarez_mod.F90:123-173. v [n,nlev,2,4,4] -> FM [n,nlev,2,4,4]."""
    etam = (hyam + hybm)[None, :, None, None, None]
    k_v = K_F * np.maximum(0.0, (etam - SIGMA_B) / (1.0 - SIGMA_B))
    return -k_v * v


def fill_forcing(h) -> None:
    """Set the driver's FM, FT from its current state at time level n0 (call push_results() first if the
    dycore has stepped) and leave FQ zero; then h.push_forcing() hands them to the dycore."""
    nstep, nm1, n0, np1 = h.time_levels()
    st, f = h.state(), h.forcing()
    n = h.nelemd
    lat = h.array("lat").reshape(n, 4, 4)
    hyam, hybm = h.vcoord[2], h.vcoord[3]
    ft, _ = hs_T_forcing(hyam, hybm, st["ps_v"][:, n0 - 1], st["T"][:, n0 - 1], lat)
    f["FT"][...] = ft
    f["FM"][...] = hs_v_forcing(hyam, hybm, st["v"][:, n0 - 1])
    f["FQ"][...] = 0.0


def forced_step(h) -> int:
    """One prim_run_subcycle_c call with Held-Suarez forcing, in the order of the CAM-coupled wrapper
    (prim_driver_mod.F90:1380-1402): forcing in, step, results out."""
    fill_forcing(h)
    h.push_forcing()
    nstep = h.run_subcycle()
    h.push_results()
    return nstep
