"""Pairwise-distinct tracer fields for the qsize = 40 parity tests.

The Jablonowski-Williamson initial state of the reference (baroclinic_inst_mod.F90:195-236) sets every
tracer but q2 and q3 to T/400, so 37 of the 40 tracers of the headline configuration are identical and
a kernel that mixed up two tracers' buffers would pass any comparison. Here tracer i is one of the four
DCMIP 2012 test 1-1 shapes (tests/dcmip_tracers.py: cosine bells, correlated field, slotted cylinders,
constant) rotated in longitude by i * 2 pi / qsize, confined to a band of levels that moves with i, and
scaled by an amplitude that depends on i - no two tracers agree at any level."""
import numpy as np

import dcmip_tracers


def mixing_ratios(lat, lon, nlev, qsize):
    """lat, lon: [n, 4, 4] -> q [n, qsize, nlev, 4, 4], every tracer different from every other."""
    n = lat.shape[0]
    out = np.empty((n, qsize, nlev, 4, 4))
    k = np.arange(nlev).reshape(1, nlev, 1, 1)
    for i in range(qsize):
        shapes = dcmip_tracers.dcmip11_mixing_ratios(lat, lon - i * 2.0 * np.pi / qsize, nlev)
        base = shapes[:, i % 4]
        # a smooth vertical modulation whose phase depends on the tracer, so even the "constant" shape differs
        vert = 1.0 + 0.25 * np.cos(2.0 * np.pi * (k + 3.0 * i) / nlev)
        out[:, i] = (0.05 + 0.95 * (i + 1.0) / qsize) * base * vert + 1e-3 * i
    return out


def install(h):
    """Overwrite the driver's Qdp (both time levels) and Q. Call BEFORE init_dycore."""
    st = h.state()
    n, nlev, nq = h.nelemd, h.cfg.nlev, h.cfg.qsize
    lat = h.array("lat").reshape(n, 4, 4)
    lon = h.array("lon").reshape(n, 4, 4)
    q = mixing_ratios(lat, lon, nlev, nq)
    dp = st["dp3d"][:, 0]
    for tl in range(2):
        st["Qdp"][:, tl, :nq] = q * dp[:, None]
    st["Q"][:, :nq] = q
    return q


def assert_distinct(q):
    """No two tracers of the installed field coincide (guards the guard)."""
    nq = q.shape[1]
    flat = q.transpose(1, 0, 2, 3, 4).reshape(nq, -1)
    for i in range(nq):
        for j in range(i + 1, nq):
            assert not np.array_equal(flat[i], flat[j]), (i, j)
