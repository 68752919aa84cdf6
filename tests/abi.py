"""Direct ctypes calls into the reference's C ABI (include/hommexx_b200.h), gfortran style:
every argument by reference. Shared by the oracle tests and the CUDA parity tests."""
import ctypes as C

import numpy as np


def _ri(x):
    return C.byref(C.c_int(int(x)))


def _rd(x):
    return C.byref(C.c_double(float(x)))


def _rb(x):
    return C.byref(C.c_bool(bool(x)))


def _rp(a):
    """const double* const& : pointer to a pointer"""
    return C.byref(C.c_void_p(a.ctypes.data))


DEFAULT_PARAMS = dict(remap_alg=1, limiter_option=8, rsplit=3, qsplit=1, time_step_type=5, energy_fixer=-1, qsize=4,
                      state_frequency=9999, nu=1e15, nu_p=1e15, nu_q=1e15, nu_s=1e15, nu_div=1e15, nu_top=2.5e5,
                      hypervis_order=2, hypervis_subcycle=3, hypervis_scaling=0.0, ftype=0, prescribed_wind=False,
                      moisture=False, disable_diagnostics=True, use_cpstar=False, use_semi_lagrangian_transport=False)


def init_simulation_params(lib, **kw):
    p = dict(DEFAULT_PARAMS)
    p.update(kw)
    lib.init_simulation_params_c(
        _ri(p["remap_alg"]), _ri(p["limiter_option"]), _ri(p["rsplit"]), _ri(p["qsplit"]), _ri(p["time_step_type"]),
        _ri(p["energy_fixer"]), _ri(p["qsize"]), _ri(p["state_frequency"]), _rd(p["nu"]), _rd(p["nu_p"]),
        _rd(p["nu_q"]), _rd(p["nu_s"]), _rd(p["nu_div"]), _rd(p["nu_top"]), _ri(p["hypervis_order"]),
        _ri(p["hypervis_subcycle"]), _rd(p["hypervis_scaling"]), _ri(p["ftype"]), _rb(p["prescribed_wind"]),
        _rb(p["moisture"]), _rb(p["disable_diagnostics"]), _rb(p["use_cpstar"]),
        _rb(p["use_semi_lagrangian_transport"]))


def isolated_elements_session(lib, nelem, nlev, dvv, D, Dinv, metdet, metinv=None, mp=None, spheremp=None,
                              rspheremp=None, fcor=None, phis=None, hyai=None, hybi=None, ps0=100000.0, **params):
    """A session of `nelem` elements with NO connections (every neighbour MISSING) and the given
    Fortran-layout geometry: the setting of the reference's functor unit tests."""
    f8 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), shape)).copy()
    n = nelem
    D, Dinv = f8(D, (n, 2, 2, 4, 4)), f8(Dinv, (n, 2, 2, 4, 4))
    metdet = f8(metdet, (n, 4, 4))
    metinv = f8(metinv if metinv is not None else 0.0, (n, 2, 2, 4, 4))
    mp = f8(mp if mp is not None else 1.0, (n, 4, 4))
    spheremp = f8(spheremp if spheremp is not None else mp * metdet, (n, 4, 4))
    rspheremp = f8(rspheremp if rspheremp is not None else 1.0 / spheremp, (n, 4, 4))
    fcor = f8(fcor if fcor is not None else 0.0, (n, 4, 4))
    phis = f8(phis if phis is not None else 0.0, (n, 4, 4))
    tv = np.zeros((n, 2, 2, 4, 4)); vs = np.zeros((n, 2, 3, 4, 4))
    dvv = f8(dvv, (4, 4))
    lib.reset_cxx_comm(_ri(0))
    lib.initialize_hommexx_session()
    lib.init_connectivity(_ri(n))
    lib.finalize_connectivity()
    lib.init_derivative_c(_rp(dvv))
    init_simulation_params(lib, **params)
    lib.init_elements_2d_c(_ri(n), _rp(D), _rp(Dinv), _rp(fcor), _rp(mp), _rp(spheremp), _rp(rspheremp), _rp(metdet),
                           _rp(metinv), _rp(phis), _rp(tv), _rp(vs), _rb(True))
    if hyai is None:
        hyai = np.linspace(0.002, 0.0, nlev + 1); hybi = np.linspace(0.0, 1.0, nlev + 1)
    hyai, hybi = f8(hyai, (nlev + 1,)), f8(hybi, (nlev + 1,))
    hyam, hybm = 0.5 * (hyai[1:] + hyai[:-1]), 0.5 * (hybi[1:] + hybi[:-1])
    lib.init_hvcoord_c(_rd(ps0), _rp(hyam), _rp(hyai), _rp(hybm), _rp(hybi))
    lib.init_boundary_exchanges_c()
    lib.init_time_level_c(_ri(1), _ri(2), _ri(3), _ri(0), _ri(2))
    return dict(D=D, Dinv=Dinv, metdet=metdet, metinv=metinv, mp=mp, spheremp=spheremp, rspheremp=rspheremp)


def sphere_op(lib, op, ie, fields_in, n_out, nlev, nu_ratio=1.0):
    """fields_in: (n_in, 16, nlev) -> (n_out, 16, nlev)"""
    a = np.ascontiguousarray(fields_in, dtype=np.float64)
    out = np.zeros((n_out, 16, nlev))
    lib.hxx_sphere_op(op.encode(), ie, a.ctypes.data, out.ctypes.data, C.c_double(nu_ratio))
    return out
