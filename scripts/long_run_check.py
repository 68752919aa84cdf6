#!/usr/bin/env python
"""Long-run sanity of the CUDA dycore at BASELINE configs[1] (ne30, nlev 72, qsize 40, 40 distinct tracers): N calls of
prim_run_subcycle_c with both builds (strict and FMA), then dry-air mass and every tracer's global mass against their
initial values, the range of the q = const tracer, and how far the two builds have drifted apart (round-off growing
in a baroclinically unstable flow: reported, not asserted). One JSON line.
    python scripts/long_run_check.py [--calls 200] [--ne 30]"""
import argparse
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import distinct_tracers  # noqa: E402
from hommexx_b200 import homme  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--calls", type=int, default=200)
ap.add_argument("--ne", type=int, default=30)
a = ap.parse_args()
cfg = homme.preset("ne30", ne=a.ne)
out, states = {"calls": a.calls, "ne": a.ne, "dynamics_steps": a.calls * cfg.rsplit * cfg.qsplit,
               "simulated_days": a.calls * cfg.rsplit * cfg.qsplit * cfg.tstep / 86400.0}, {}
for flav in ("", "fma"):
    path = homme.cuda_lib_path(cfg.nlev, cfg.qsize_d, flav)
    h = homme.Homme(cfg, path)
    distinct_tracers.install(h)
    h.init_dycore()
    sph = h.array("spheremp").reshape(-1, 1, 1, 4, 4)
    s = h.state()
    m0 = (s["Qdp"][:, 0] * sph).sum(axis=(0, 2, 3, 4))
    dry0 = (s["dp3d"][:, 0] * sph[:, 0]).sum()
    for _ in range(a.calls):
        nstep = h.run_subcycle()
    h.push_results()
    s = h.state()
    n0 = h.time_levels()[2] - 1
    tq = (nstep // cfg.qsplit) % 2
    m1 = (s["Qdp"][:, tq] * sph).sum(axis=(0, 2, 3, 4))
    dry1 = (s["dp3d"][:, n0] * sph[:, 0]).sum()
    states[flav] = {k: s[k].copy() for k in ("T", "ps_v", "Q")} | {"v": s["v"][:, n0].copy(), "T0": s["T"][:, n0].copy()}
    out[flav or "strict"] = {
        "finite": bool(all(np.isfinite(v).all() for v in s.values())),
        "tracer_mass_rel_drift_max": float(np.abs(m1 / m0 - 1.0).max()),
        "dry_mass_rel_drift": float(abs(dry1 / dry0 - 1.0)),
        "T_range": [float(s["T"][:, n0].min()), float(s["T"][:, n0].max())],
        "ps_range": [float(s["ps_v"][:, n0].min()), float(s["ps_v"][:, n0].max())],
        "max_wind": float(np.abs(s["v"][:, n0]).max()), "library": path.name}
    h.close()
rel = lambda x, y: float(np.sqrt(((x - y) ** 2).sum()) / np.sqrt((y ** 2).sum()))
out["fma_vs_strict_rel_l2"] = {k: rel(states["fma"][k], states[""][k]) for k in ("v", "T0", "Q")}
print(json.dumps(out))
