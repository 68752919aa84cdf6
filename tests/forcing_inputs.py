"""Seeded CAM-forcing arrays (Fortran layout) used by the forcing / diagnostics tests: large negative
tracer tendencies so that the negativity clamp of CamForcing.cpp:92-98 / :118-124 fires."""
import numpy as np


def fill_forcing(h, seed=7):
    rng = np.random.default_rng(seed)
    f = h.forcing()
    st = h.state()
    f["FM"][...] = 1e-4 * rng.standard_normal(f["FM"].shape)
    f["FT"][...] = 1e-5 * rng.standard_normal(f["FT"].shape)
    # tendencies comparable to qdp / dt_remap: some push qdp below zero (clamped), some do not
    scale = np.abs(st["Qdp"][:, 0]).mean() / (h.cfg.tstep * h.cfg.rsplit * h.cfg.qsplit)
    f["FQ"][...] = 2.0 * scale * rng.standard_normal(f["FQ"].shape)
    f["FQ"][:, h.cfg.qsize:] = 0.0
