"""DCMIP 2012 test 1-1 (3-D deformational flow) as a tracer-only run: BASELINE configs[2].

The reference's C++ path rejects `prescribed_wind = 1` (cxx_f90_interface.cpp:44), so there is no
`prim_run_subcycle_c` behaviour to match for this case (SURVEY.md 8d caveat C2). What the case exercises —
EulerStepFunctor with the quasi-monotone limiter, its DSS / min-max exchanges and the PPM remap — is driven
here through the functors' public run methods (include/hommexx_b200.h section C), with the analytic winds
of `src/test_src/dcmip2012_test1_2_3.F90:87-270` evaluated by this harness each step, the way
`dcmip_tests.F90:40-100` + `set_prescribed_wind` do on the Fortran side:

    mass flux    derived_vn0 = (u, v)(t + dt/2) dp_ref           derived_dp = dp_ref
    tracers      three SSP-RK2 stages + time average            (prim_advec_tracers_remap.cpp:32-90)
    thickness    dp3d(np1) = dp_ref - dt DSS(div(derived_vn0))  (the stages' own divdp_proj: exactly consistent)
    remap        qdp from dp3d(np1) back to the reference levels, Q = qdp / dp_ref

Host orchestration only (numpy for the analytic fields); the arithmetic of the path is the library's.
Works with either library (CUDA product or CPU oracle): both export the same hooks.
"""
from __future__ import annotations

import numpy as np

from . import homme

A = 6.376e6
P0 = 1.0e5
RD, G, T0 = 287.04, 9.80616, 300.0
H = RD * T0 / G
TAU = 12.0 * 86400.0
DSS_ETA, DSS_OMEGA, DSS_DIV_VDP_AVE = 0, 1, 2


def winds(time, lon, lat, p):
    """u, v of test1_advection_deformation (:164-177); lon, lat [n,16,1], p [1,1,nlev] broadcast."""
    u0, k0, omega0 = 2 * np.pi * A / TAU, 10 * A / TAU, 23000 * np.pi / TAU
    ptop = P0 * np.exp(-12000.0 / H)
    bs = 0.2
    lonp = lon - 2 * np.pi * time / TAU
    ud = (omega0 * A) / (bs * ptop) * np.cos(lonp) * np.cos(lat) ** 2 * np.cos(2 * np.pi * time / TAU) * \
        (-np.exp((p - P0) / (bs * ptop)) + np.exp((ptop - p) / (bs * ptop)))
    u = k0 * np.sin(lonp) ** 2 * np.sin(2 * lat) * np.cos(np.pi * time / TAU) + u0 * np.cos(lat) + ud
    v = k0 * np.sin(2 * lonp) * np.cos(lat) * np.cos(np.pi * time / TAU) + 0.0 * p
    return u, v


def tracers(lon, lat, p):
    """q1..q4 of test1_advection_deformation (:204-248) at t = 0."""
    z = H * np.log(P0 / p) + 0.0 * lon
    RR, ZZ, z0 = 0.5, 1000.0, 5000.0
    r1 = np.arccos(np.clip(np.cos(lat) * np.cos(lon - 5 * np.pi / 6), -1, 1))
    r2 = np.arccos(np.clip(np.cos(lat) * np.cos(lon - 7 * np.pi / 6), -1, 1))
    d1 = np.minimum(1.0, (r1 / RR) ** 2 + ((z - z0) / ZZ) ** 2)
    d2 = np.minimum(1.0, (r2 / RR) ** 2 + ((z - z0) / ZZ) ** 2)
    q1 = 0.5 * (1 + np.cos(np.pi * d1)) + 0.5 * (1 + np.cos(np.pi * d2))
    q2 = 0.9 - 0.8 * q1 ** 2
    q3 = np.where((d1 <= RR) | (d2 <= RR), 1.0, 0.1)
    q3 = np.where((z > z0) & (np.abs(lat) < 0.125), 0.1, q3)
    q4 = 1.0 - 0.3 * (q1 + q2 + q3)
    return np.stack([q1, q2, q3, q4])


class Dcmip11:
    """Tracer-only DCMIP 1-1 on a bound dycore library. Fields live in the device layout [ie][..][16][nlev]."""

    def __init__(self, ne: int, nlev: int, libpath, tstep: float, nu_q: float = 0.0, limiter_option: int = 8):
        nu = max(nu_q, 1e-30)
        self.cfg = homme.Config(ne=ne, nlev=nlev, qsize=4, qsize_d=4, vcoord="dcmip-z12km", tstep=tstep, rsplit=1,
                                qsplit=1, nu=nu, nu_p=nu, nu_q=nu, nu_s=nu, nu_top=0.0, limiter_option=limiter_option,
                                hypervis_subcycle=1, u_perturb=0.0)
        self.h = h = homme.Homme(self.cfg, libpath)
        n = self.n = h.nelemd
        self.lat = h.array("lat").reshape(n, 16, 1).copy()
        self.lon = h.array("lon").reshape(n, 16, 1).copy()
        hyai, hybi, hyam, hybm = h.vcoord
        self.pm = (P0 * (hyam + hybm)).reshape(1, 1, nlev)
        self.dp = np.broadcast_to((P0 * np.diff(hyai + hybi)).reshape(1, 1, nlev), (n, 16, nlev)).copy()
        # prescribed state in the driver arrays (Fortran layout), then the usual init sequence
        st = h.state()
        st["ps_v"][...] = P0
        st["T"][...] = T0
        st["v"][...] = 0.0
        dpf = self.dp.reshape(n, 4, 4, nlev).transpose(0, 3, 1, 2)
        st["dp3d"][...] = dpf[:, None]
        self.q0 = tracers(self.lon, self.lat, self.pm)                      # [4][n][16][nlev]
        qdp = (self.q0 * self.dp[None]).reshape(4, n, 4, 4, nlev).transpose(1, 0, 4, 2, 3)
        st["Qdp"][...] = qdp[:, None]
        h.init_dycore()
        h.set_field("dpdiss_ave", self.dp)      # what the dynamics would have averaged (prim_advance_hypervis)
        self.time, self.n0_qdp = 0.0, 0

    def step(self):
        h, lib, dt = self.h, self.h.lib, self.cfg.tstep
        u, v = winds(self.time + 0.5 * dt, self.lon, self.lat, self.pm)
        h.set_field("derived_dp", self.dp)
        h.set_field("derived_vn0", np.stack([u * self.dp, v * self.dp], axis=1))
        n0q, np1q = self.n0_qdp, 1 - self.n0_qdp
        lib.hxx_euler_reset()
        lib.hxx_euler_precompute_divdp()
        lib.hxx_euler_step(np1q, n0q, dt / 2, 0.0, DSS_DIV_VDP_AVE)
        lib.hxx_euler_step(np1q, np1q, dt / 2, 1.0, DSS_ETA)
        lib.hxx_euler_step(np1q, np1q, dt / 2, 2.0, DSS_OMEGA)
        lib.hxx_euler_qdp_time_avg(n0q, np1q)
        # Lagrangian thickness after the step, consistent with the tracer stages' own divergence
        div = h.get_field("divdp_proj").reshape(self.n, 16, -1)
        dp3d = h.get_field("dp3d").reshape(self.n, 3, 16, -1)
        dp3d[:, 2] = self.dp - dt * div
        h.set_field("dp3d", dp3d)
        lib.hxx_vertical_remap(2, np1q, dt)
        lib.hxx_update_q(np1q, 2)
        self.n0_qdp = np1q
        self.time += dt

    def q(self):
        """Mixing ratios [4][n][16][nlev] of the current tracer level."""
        qdp = self.h.get_field("qdp").reshape(self.n, 2, 4, 16, -1)[:, self.n0_qdp]
        return (qdp / self.dp[:, None]).transpose(1, 0, 2, 3)

    def masses(self):
        sph = self.h.array("spheremp").reshape(self.n, 1, 16, 1)
        qdp = self.h.get_field("qdp").reshape(self.n, 2, 4, 16, -1)[:, self.n0_qdp]
        return (qdp * sph).sum(axis=(0, 2, 3))

    def close(self):
        self.h.close()
