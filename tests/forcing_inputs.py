"""Seeded CAM-forcing arrays (Fortran layout) used by the forcing / diagnostics tests."""
import numpy as np


def fill_forcing(h, seed=7):
    """Point-wise random tendencies; tracer tendencies comparable to qdp / dt_remap so that the
    negativity clamp of CamForcing.cpp:92-98 / :118-124 fires. For single passes only (the fields are
    discontinuous across elements)."""
    rng = np.random.default_rng(seed)
    f = h.forcing()
    st = h.state()
    f["FM"][...] = 1e-4 * rng.standard_normal(f["FM"].shape)
    f["FT"][...] = 1e-5 * rng.standard_normal(f["FT"].shape)
    scale = np.abs(st["Qdp"][:, 0]).mean() / (h.cfg.tstep * h.cfg.rsplit * h.cfg.qsplit)
    f["FQ"][...] = 2.0 * scale * rng.standard_normal(f["FQ"].shape)
    f["FQ"][:, h.cfg.qsize:] = 0.0


def fill_smooth_forcing(h):
    """Tendencies that are smooth functions of (lon, lat, level), continuous across elements, for forced
    multi-step runs: a weak momentum/heat source, a gentle moisture source (tracer 0, moves ps_v in moist
    runs) and a strong sink of tracer 1 in one region that drives it into the clamp."""
    f = h.forcing()
    st = h.state()
    n, nl = h.nelemd, h.cfg.nlev
    lat = h.array("lat").reshape(n, 1, 4, 4)
    lon = h.array("lon").reshape(n, 1, 4, 4)
    prof = np.sin(np.pi * (np.arange(nl) + 0.5) / nl).reshape(1, nl, 1, 1)
    dt_remap = h.cfg.tstep * h.cfg.rsplit * h.cfg.qsplit
    f["FM"][:, :, 0] = 2e-5 * np.cos(lat) * np.sin(2 * lon) * prof
    f["FM"][:, :, 1] = 1e-5 * np.cos(lat) ** 2 * np.cos(lon) * prof
    f["FT"][...] = 2e-5 * np.cos(lat) * np.cos(3 * lon) * prof
    f["FQ"][...] = 0.0
    q = st["Qdp"][:, 0]
    f["FQ"][:, 0] = 1e-3 * np.abs(q[:, 0]).mean() / dt_remap * np.cos(lat) * np.sin(lon) * prof
    if h.cfg.qsize > 1:
        f["FQ"][:, 1] = -0.6 * q[:, 1] / dt_remap * np.exp(-4.0 * ((lat - 0.5) ** 2 + (lon - 2.0) ** 2))
        f["FQ"][:, 1] -= 0.7 * np.abs(q[:, 1]) / dt_remap * (np.exp(-8.0 * ((lat + 0.3) ** 2 + (lon - 4.0) ** 2)) > 0.5)
