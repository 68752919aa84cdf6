"""Python handle on the host-side driver (hommexx_b200/driver) and on dycore libraries.

The driver (C++, libhomme_driver.so) plays the role of HOMME's Fortran driver: it owns the
Fortran-layout arrays and calls a dycore library through the reference's C ABI
(include/hommexx_b200.h). This module only wraps it with ctypes so tests and bench.py can
orchestrate runs; no arithmetic of the hot path happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
from dataclasses import dataclass, field

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
PKG = pathlib.Path(__file__).resolve().parent
DRIVER_LIB = PKG / "driver" / "libhomme_driver.so"


def cuda_lib_path(nlev: int, qsize_d: int, flavour: str = "") -> pathlib.Path:
    """flavour "" = the strict build (--fmad=false, bit-identical to the oracle); "fma" = the same sources with
    multiply-adds contracted (parity <= 1e-11). HXX_VARIANT selects an experimental build of the same sources
    (scripts/build_variant.py) instead."""
    var = os.environ.get("HXX_VARIANT", "")
    base = PKG / "csrc" / "variants" / var if var else PKG / "csrc"
    return base / f"libhommexx_b200_nlev{nlev}_q{qsize_d}{'_' + flavour if flavour else ''}.so"


class HommeParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("ne", "nlev", "qsize_d", "qsize", "npart", "part_id",
                                      "remap_alg", "limiter_option", "rsplit", "qsplit", "time_step_type",
                                      "energy_fixer", "state_frequency")] + \
               [(n, C.c_double) for n in ("nu", "nu_p", "nu_q", "nu_s", "nu_div", "nu_top")] + \
               [("hypervis_order", C.c_int), ("hypervis_subcycle", C.c_int), ("hypervis_scaling", C.c_double),
                ("ftype", C.c_int)] + \
               [(n, C.c_int) for n in ("prescribed_wind", "moisture", "disable_diagnostics", "use_cpstar",
                                      "use_semi_lagrangian_transport")] + \
               [("tstep", C.c_double), ("u_perturb", C.c_double)]


@dataclass
class Config:
    """ctl_nl namelist values (defaults = test/reg_test/benchmarks/v1/homme-ne30-v1.nl)."""
    ne: int = 30
    nlev: int = 72
    qsize_d: int = 40
    qsize: int = 40
    npart: int = 1
    part_id: int = 0
    remap_alg: int = 1
    limiter_option: int = 8
    rsplit: int = 3
    qsplit: int = 1
    time_step_type: int = 5
    energy_fixer: int = -1
    state_frequency: int = 9999
    nu: float = 1e15
    nu_p: float = 1e15
    nu_q: float = 1e15
    nu_s: float = 1e15
    nu_div: float = -1.0
    nu_top: float = 2.5e5
    hypervis_order: int = 2
    hypervis_subcycle: int = 3
    hypervis_scaling: float = 0.0
    ftype: int = 0
    prescribed_wind: int = 0
    moisture: int = 0
    disable_diagnostics: int = 1
    use_cpstar: int = 0
    use_semi_lagrangian_transport: int = 0
    tstep: float = 300.0
    u_perturb: float = 1.0
    vcoord: str = ""  # "" -> by nlev

    def struct(self) -> HommeParams:
        p = HommeParams()
        for n, _ in HommeParams._fields_:
            setattr(p, n, getattr(self, n))
        return p


# Named configurations. ne4: benchmarks/v1/homme-ne4-v1.nl physics values.
def preset(name: str, **over) -> Config:
    base = {
        # BASELINE configs[0]: prtcA-like CPU-runnable case, nlev 72, qsize 4
        "ne4": dict(ne=4, qsize=4, qsize_d=4, tstep=1800.0, nu=4.5e17, nu_p=4.5e17, nu_q=4.5e17, nu_s=4.5e17),
        # reference prtcA_c executable: PLEV=26, QSIZE_D=4 (test/reg_test/namelists/prtcA-r3-dry.nl)
        "prtcA": dict(ne=4, nlev=26, qsize=4, qsize_d=4, tstep=600.0, nu=7e15, nu_p=7e15, nu_q=7e15, nu_s=-1.0,
                      nu_div=7e15),
        "ne8": dict(ne=8, qsize=4, qsize_d=4, tstep=900.0, nu=5.5e16, nu_p=5.5e16, nu_q=5.5e16, nu_s=5.5e16),
        # BASELINE configs[1]: the headline single-B200 case
        "ne30": dict(),
        # BASELINE configs[3]: homme-ne120-v1.nl
        "ne120": dict(ne=120, tstep=75.0, rsplit=2, nu=1e13, nu_p=1e13, nu_q=1e13, nu_s=1e13, hypervis_subcycle=4),
    }[name]
    base.update(over)
    return Config(**base)


def read_vcoord(nlev: int, name: str = ""):
    if name == "dcmip-z12km":
        # DCMIP 2012 test 1-x grid (dcmip_tests.F90:62-68): nlev layers evenly spaced in z up to 12 km in an
        # isothermal 300 K atmosphere, eta = exp(-z/H); hybrid coefficients with p(eta, ps = p0) = p0 eta
        H = 287.04 * 300.0 / 9.80616
        etai = np.exp(-np.linspace(12000.0, 0.0, nlev + 1) / H)
        hybi = (etai - etai[0]) / (1.0 - etai[0])
        hyai = etai - hybi
        hyam, hybm = 0.5 * (hyai[1:] + hyai[:-1]), 0.5 * (hybi[1:] + hybi[:-1])
        return hyai, hybi, hyam, hybm
    if not name:
        name = {72: "acme-72", 26: "cam-26"}.get(nlev, "")
    if name:
        rows = np.loadtxt(PKG / "data" / f"vcoord-{name}.txt")
        assert rows.shape[0] == 2 * nlev + 1, (rows.shape, nlev)
        hyai, hybi = rows[: nlev + 1, 0].copy(), rows[: nlev + 1, 1].copy()
        hyam, hybm = rows[nlev + 1:, 0].copy(), rows[nlev + 1:, 1].copy()
        return hyai, hybi, hyam, hybm
    # Synthetic smooth hybrid grid for odd test sizes (e.g. nlev=8): pure sigma above a small top.
    s = np.linspace(0.0, 1.0, nlev + 1) ** 1.5
    ptop = 0.002
    hyai = ptop * (1.0 - s)
    hybi = s.copy()
    hybi[0] = 0.0
    hyam, hybm = 0.5 * (hyai[1:] + hyai[:-1]), 0.5 * (hybi[1:] + hybi[:-1])
    return hyai, hybi, hyam, hybm


_drv = None


def driver_lib() -> C.CDLL:
    global _drv
    if _drv is None:
        if not DRIVER_LIB.exists():
            raise RuntimeError(f"{DRIVER_LIB} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        d = C.CDLL(str(DRIVER_LIB))
        d.hd_create.restype = C.c_void_p
        d.hd_create.argtypes = [C.POINTER(HommeParams)] + [C.c_void_p] * 4
        for f in ("hd_destroy", "hd_init_jw", "hd_init_dycore", "hd_upload_state", "hd_push_results",
                  "hd_finalize_dycore", "hd_push_forcing", "hd_pull_forcing"):
            getattr(d, f).argtypes = [C.c_void_p]
            getattr(d, f).restype = None
        d.hd_set_last_step.argtypes = [C.c_void_p, C.c_int]
        d.hd_set_last_step.restype = None
        d.hd_bind.argtypes = [C.c_void_p, C.c_char_p]
        d.hd_bind.restype = C.c_int
        d.hd_last_error.restype = C.c_char_p
        d.hd_run_subcycle.argtypes = [C.c_void_p]
        d.hd_run_subcycle.restype = C.c_int
        d.hd_nelemd.argtypes = [C.c_void_p]
        d.hd_nelem_global.argtypes = [C.c_void_p]
        d.hd_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        d.hd_array.restype = C.POINTER(C.c_double)
        d.hd_connections.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int))]
        d.hd_time_levels.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        d.hd_set_time_levels.argtypes = [C.c_void_p] + [C.c_int] * 4
        d.hd_set_time_levels.restype = None
        d.hd_local_gids.argtypes = [C.c_void_p]
        d.hd_local_gids.restype = C.POINTER(C.c_int)
        d.hd_owner.argtypes = [C.c_void_p]
        d.hd_owner.restype = C.POINTER(C.c_int)
        _drv = d
    return _drv


def load_dycore(path) -> C.CDLL:
    """dlopen a dycore library and declare the section B/C signatures. Fails loudly."""
    path = pathlib.Path(path)
    if not path.exists():
        raise RuntimeError(f"dycore library {path} not built; the product has no CPU fallback")
    lib = C.CDLL(str(path))
    lib.hommexx_b200_backend.restype = C.c_char_p
    lib.hommexx_b200_launch_count.restype = C.c_int64
    lib.hommexx_b200_set_comm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    # section C (phase-level hooks): every product / oracle library has them; the reference's own build under
    # oracle/_ref binds the functor-level ones to the reference's objects (oracle/ref_hommexx_api.cpp)
    sigs = {
        "hxx_caar_run": ([C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int], None),
        "hxx_rk_combine": ([C.c_int, C.c_int], None),
        "hxx_hypervis_run": ([C.c_int, C.c_double, C.c_double], None),
        "hxx_euler_step": ([C.c_int, C.c_int, C.c_double, C.c_double, C.c_int], None),
        "hxx_euler_qdp_time_avg": ([C.c_int, C.c_int], None),
        "hxx_vertical_remap": ([C.c_int, C.c_int, C.c_double], None),
        "hxx_update_q": ([C.c_int, C.c_int], None),
        "hxx_prim_step_init": ([C.c_int], None),
        "hxx_apply_forcing": ([C.c_double], None),
        "hxx_held_suarez_forcing": ([C.c_void_p, C.c_void_p, C.c_void_p], None),
        "hxx_diagnostics": ([C.c_int, C.c_int, C.c_int], None),
        "hxx_exchange": ([C.c_char_p, C.c_int], None),
        "hxx_get_field": ([C.c_char_p, C.c_void_p], C.c_int64),
        "hxx_set_field": ([C.c_char_p, C.c_void_p], C.c_int64),
        "hxx_sphere_op": ([C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double], None),
        "hxx_limiter": ([C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], None),
        "hxx_remap_columns": ([C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], None),
        "hxx_euler_reset": ([], None), "hxx_euler_precompute_divdp": ([], None),
        "hommexx_b200_sync": ([], None), "finalize_hommexx_session": ([], None),
    }
    for name, (argtypes, restype) in sigs.items():
        if hasattr(lib, name):
            f = getattr(lib, name)
            f.argtypes, f.restype = argtypes, restype
    return lib


class Homme:
    """One rank of a run: mesh + initial state + a bound dycore library."""

    def __init__(self, cfg: Config, libpath, init: str = "jw"):
        self.cfg = cfg
        self.d = driver_lib()
        hy = [np.ascontiguousarray(a, dtype=np.float64) for a in read_vcoord(cfg.nlev, cfg.vcoord)]
        self.vcoord = hy
        p = cfg.struct()
        self.h = C.c_void_p(self.d.hd_create(C.byref(p), *[a.ctypes.data for a in hy]))
        if init == "jw":
            self.d.hd_init_jw(self.h)
        self.libpath = str(libpath)
        self.lib = load_dycore(libpath)
        if hasattr(self.lib, "oracle_set_dims"):
            self.lib.oracle_set_dims.argtypes = [C.c_int, C.c_int]
            self.lib.oracle_set_dims(cfg.nlev, cfg.qsize_d)
        rc = self.d.hd_bind(self.h, self.libpath.encode())
        if rc != 0:
            raise RuntimeError(f"hd_bind({libpath}) failed: {self.d.hd_last_error().decode()}")
        self.nelemd = self.d.hd_nelemd(self.h)
        self.nelem = self.d.hd_nelem_global(self.h)
        self._initialised = False

    # -- lifecycle ---------------------------------------------------------------------------
    def init_dycore(self):
        self.d.hd_init_dycore(self.h)
        self._initialised = True

    def upload_state(self):
        self.d.hd_upload_state(self.h)

    def run_subcycle(self) -> int:
        return self.d.hd_run_subcycle(self.h)

    def push_results(self):
        self.d.hd_push_results(self.h)

    def push_forcing(self):
        """f90_push_forcing_to_cxx(FM, FT, FQ, Qdp): forcing arrays to the dycore, Qdp back to the driver."""
        self.d.hd_push_forcing(self.h)

    def pull_forcing(self):
        self.d.hd_pull_forcing(self.h)

    def held_suarez_forcing(self):
        """hxx_held_suarez_forcing: FM, FT from the dycore's own state at n0, evaluated where the state lives."""
        self.lib.hxx_held_suarez_forcing(self.array("lat").ctypes.data, self.vcoord[2].ctypes.data,
                                         self.vcoord[3].ctypes.data)

    def set_last_step(self, n_end_step: int):
        self.d.hd_set_last_step(self.h, n_end_step)

    def forcing(self) -> dict:
        c, n = self.cfg, self.nelemd
        return {"FM": self.array("FM").reshape(n, c.nlev, 2, 4, 4), "FT": self.array("FT").reshape(n, c.nlev, 4, 4),
                "FQ": self.array("FQ").reshape(n, c.qsize_d, c.nlev, 4, 4)}

    def accum(self) -> dict:
        """elem%accum diagnostics written by the dycore (Diagnostics.cpp:17-35 shapes)."""
        c, n = self.cfg, self.nelemd
        qd = max(1, c.qsize_d)
        out = {}
        for nm, shp in (("Qvar", (4, qd, 4, 4)), ("Qmass", (4, qd, 4, 4)), ("Q1mass", (qd, 4, 4)),
                        ("IEner", (4, 4, 4)), ("IEner_wet", (4, 4)), ("KEner", (4, 4, 4)), ("PEner", (4, 4, 4))):
            a = self.array(nm)
            k = int(np.prod(shp))
            out[nm] = a[: n * k].reshape((n,) + shp)
        return out

    def finalize(self):
        if self._initialised:
            self.d.hd_finalize_dycore(self.h)
            self._initialised = False

    def close(self):
        self.finalize()
        if self.h:
            self.d.hd_destroy(self.h)
            self.h = None

    # -- data --------------------------------------------------------------------------------
    def array(self, name: str) -> np.ndarray:
        """Fortran-layout driver array as a flat numpy view (no copy)."""
        n = C.c_int64()
        ptr = self.d.hd_array(self.h, name.encode(), C.byref(n))
        if not ptr:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def state(self) -> dict:
        """Shaped views of the prognostic arrays in the Fortran/ABI layout."""
        c, n = self.cfg, self.nelemd
        return {
            "v": self.array("v").reshape(n, 3, c.nlev, 2, 4, 4),
            "T": self.array("T").reshape(n, 3, c.nlev, 4, 4),
            "dp3d": self.array("dp3d").reshape(n, 3, c.nlev, 4, 4),
            "ps_v": self.array("ps_v").reshape(n, 3, 4, 4),
            "Qdp": self.array("Qdp").reshape(n, 2, c.qsize_d, c.nlev, 4, 4),
            "Q": self.array("Q").reshape(n, c.qsize_d, c.nlev, 4, 4),
            "omega_p": self.array("omega_p").reshape(n, c.nlev, 4, 4),
        }

    def time_levels(self):
        v = [C.c_int() for _ in range(4)]
        self.d.hd_time_levels(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)  # nstep, nm1, n0, np1 (1-based levels)

    # -- restart (WriteRestart / ReadRestart of restart_io_mod.F90:626-698: elem%state and the time levels) ---------
    RESTART_ARRAYS = ("v", "T", "dp3d", "ps_v", "Qdp", "Q")

    def write_restart(self, path):
        """Pull the state (cxx_push_results_to_f90) and write it with the time levels; one file per rank."""
        self.push_results()
        c = self.cfg
        np.savez(path, tl=np.array(self.time_levels(), dtype=np.int64), gids=self.local_gids(),
                 dims=np.array([c.ne, c.nlev, c.qsize, c.qsize_d], dtype=np.int64),
                 **{k: self.array(k) for k in self.RESTART_ARRAYS})

    def read_restart(self, path):
        """Restore the driver's arrays and time levels from write_restart; call BEFORE init_dycore (a restart run
        initialises the dycore from the restored elem%state, prim_main.F90:181-262) or follow it with upload_state()."""
        c = self.cfg
        with np.load(path) as z:
            if tuple(z["dims"]) != (c.ne, c.nlev, c.qsize, c.qsize_d):
                raise ValueError(f"{path}: written for (ne, nlev, qsize, qsize_d) = {tuple(z['dims'])}")
            if not np.array_equal(z["gids"], self.local_gids()):
                raise ValueError(f"{path}: written for a different partition (rank's element list differs)")
            for k in self.RESTART_ARRAYS:
                self.array(k)[:] = z[k]
            self.d.hd_set_time_levels(self.h, *[int(x) for x in z["tl"]])

    def local_gids(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.d.hd_local_gids(self.h), shape=(self.nelemd,)).copy()

    def connections(self) -> np.ndarray:
        p = C.POINTER(C.c_int)()
        n = self.d.hd_connections(self.h, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n, 8)).copy()

    # -- device-layout fields through the phase-level hooks -------------------------------------
    def field_size(self, name: str) -> int:
        return int(self.lib.hxx_get_field(name.encode(), None))

    def get_field(self, name: str) -> np.ndarray:
        n = self.field_size(name)
        if n == 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.lib.hxx_get_field(name.encode(), out.ctypes.data)
        return out

    def set_field(self, name: str, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64).ravel()
        assert a.size == self.field_size(name), (name, a.size, self.field_size(name))
        self.lib.hxx_set_field(name.encode(), a.ctypes.data)
