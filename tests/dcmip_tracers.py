"""Tracer fields of DCMIP 2012 test 1-1 (3-D deformational flow; reference: src/preqx/dcmip_tests.F90:40-188,
dcmip2012_test1_2_3.F90 test1_advection_deformation): two cosine bells, a correlated field, two slotted
cylinders and a constant — the sharp-edged shapes that keep the quasi-monotone limiter iterating.
Horizontal shapes are DCMIP's (centres (5pi/6, 0) and (7pi/6, 0), radius a/2); the vertical profile is a
band of levels instead of a height window, so the helper does not need the hydrostatic height."""
import numpy as np


def dcmip11_mixing_ratios(lat, lon, nlev):
    """lat, lon: [n, 4, 4] -> q [n, 4, nlev, 4, 4] (tracers q1..q4)."""
    n = lat.shape[0]
    lam = [5 * np.pi / 6, 7 * np.pi / 6]
    d = [np.arccos(np.clip(np.cos(lat) * np.cos(lon - l), -1.0, 1.0)) for l in lam]   # great-circle distance, phi_c = 0
    R = 0.5
    k = np.arange(nlev)
    band = ((k >= nlev // 3) & (k < 2 * nlev // 3)).astype(float).reshape(1, nlev, 1, 1)
    bells = sum(0.5 * (1 + np.cos(np.pi * np.minimum(di / R, 1.0))) for di in d)          # q1 horizontal part
    q1 = bells[:, None] * band
    q2 = 0.9 - 0.8 * q1 ** 2
    cyl = np.full(lat.shape, 0.1)
    for i, di in enumerate(d):
        inside = di < R
        slot = (np.abs(lon - lam[i]) < R / 6) & ((lat < 5 * R / 12) if i == 0 else (lat > -5 * R / 12))
        cyl = np.where(inside & ~slot, 1.0, cyl)
    q3 = np.where(band > 0, cyl[:, None], 0.1)
    q4 = np.ones((n, nlev, 4, 4))
    return np.stack([np.broadcast_to(q, (n, nlev, 4, 4)) for q in (q1, q2, q3, q4)], axis=1)


def install(h):
    """Overwrite the driver's Qdp (both time levels) with q * dp3d of the initial state. Call BEFORE init_dycore."""
    st = h.state()
    n, nlev = h.nelemd, h.cfg.nlev
    lat = h.array("lat").reshape(n, 4, 4)
    lon = h.array("lon").reshape(n, 4, 4)
    q = dcmip11_mixing_ratios(lat, lon, nlev)
    nq = min(4, h.cfg.qsize)
    dp = st["dp3d"][:, 0]
    for tl in range(2):
        st["Qdp"][:, tl, :nq] = q[:, :nq] * dp[:, None]
    return q[:, :nq]
