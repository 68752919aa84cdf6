"""The CUDA product against the REFERENCE'S OWN BUILD, directly: oracle/_ref/libref_hommexx_*.so is the reference's
src/share/cxx compiled from its sources (oracle/Makefile ref_full, serial Kokkos stand-in, VECTOR_SIZE 1); the prebuilt
library travels to the GPU box with the snapshot. Same driver, same inputs, 10-12 dynamics steps: every prognostic
array bit-identical (tol = 0; the north-star bar is 1e-11), as the reference's own C++-vs-Fortran acceptance test
demands of its two builds (cmake/CxxVsF90.cmake.in:28-41)."""
import numpy as np
import pytest

import distinct_tracers
import parity
from hommexx_b200 import homme
from reference_lib import reference_lib

pytestmark = pytest.mark.gpu
PROGNOSTIC = ["v", "T", "dp3d", "ps_v", "Qdp", "Q", "omega_p"]

CASES = {
    "ne4": dict(),                                                          # BASELINE configs[0]
    "prtcA-lim9-alg2": dict(base="prtcA", limiter_option=9, remap_alg=2),   # the reference's prtcA executable sizes
    "prtcA-r0-moist-q3": dict(base="prtcA", rsplit=0, moisture=1, qsplit=3),
    "ne4-tensorhv-nudiv": dict(base="ne4", hypervis_scaling=3.2, nu=5e-8, nu_p=5e-8, nu_q=5e-8, nu_s=5e-8, nu_div=1.25e-7),
    "ne4-q40": dict(base="ne4", qsize=40, qsize_d=40),                      # the benchmarked binary, 40 distinct tracers
    "ne8-q40": dict(base="ne8", qsize=40, qsize_d=40),
}


def run(cfg, lib):
    h = homme.Homme(cfg, lib)
    if cfg.qsize > 4:
        distinct_tracers.install(h)
    h.init_dycore()
    nstep = 0
    while nstep < 10:
        nstep = h.run_subcycle()
    h.push_results()
    out = {k: v.copy() for k, v in h.state().items()}
    backend = h.lib.hommexx_b200_backend()
    h.close()
    return out, backend


@pytest.mark.parametrize("case", list(CASES))
def test_cuda_is_bit_identical_to_the_reference_build(case):
    parity.need_gpu()
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    ref, rb = run(cfg, reference_lib(cfg.nlev, cfg.qsize_d))
    cu, cb = run(cfg, parity.cuda_lib(cfg.nlev, cfg.qsize_d))
    assert rb == b"reference-serial" and cb == b"cuda-sm100a"
    for k in PROGNOSTIC:
        assert not np.isnan(ref[k]).any(), k
        assert np.array_equal(cu[k], ref[k]), (case, k, float(np.abs(cu[k] - ref[k]).max()))


# ---- per functor: each CUDA phase against the reference's own functor, from the same state -------------------------
FUNCTOR_CASES = {
    "ne4": dict(),
    "prtcA-lim9-alg2": dict(base="prtcA", limiter_option=9, remap_alg=2),
    "ne4-r0-moist": dict(base="ne4", rsplit=0, moisture=1),
    "prtcA-tensorhv": dict(base="prtcA", hypervis_scaling=3.2, nu=1e-9, nu_p=1e-9, nu_q=1e-9, nu_s=1e-9, nu_div=2.5e-9,
                           hypervis_subcycle=2),
    "ne4-q40": dict(base="ne4", qsize=40, qsize_d=40),
}


@pytest.mark.parametrize("case", list(FUNCTOR_CASES))
def test_every_cuda_phase_against_the_reference_functor(case):
    """hxx_caar_run / hxx_hypervis_run / hxx_euler_* / hxx_vertical_remap / hxx_update_q of the CUDA library against
    CaarFunctor::run, HyperviscosityFunctor::run, EulerStepFunctor::*, VerticalRemapManager::run_remap and update_q of
    the reference build (oracle/ref_hommexx_api.cpp binds the same hooks to the reference's objects): every named
    array bit-identical after every call."""
    parity.need_gpu()
    from functor_pair import Pair
    over = dict(FUNCTOR_CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    p = Pair(cfg, parity.cuda_lib(cfg.nlev, cfg.qsize_d), b"cuda-sm100a", fields=parity.STATE_FIELDS + ["phi"])
    try:
        moist = 1 if cfg.moisture else -1
        for (nm1, n0, np1, frac, w) in [(1, 1, 0, 0.2, 0.25), (1, 0, 2, 0.2, 0.0), (1, 2, 2, 1.0 / 3.0, 0.0),
                                        (0, 2, 2, 0.75, 0.75)]:
            p.reset()
            p.call("hxx_caar_run", nm1, n0, np1, frac * cfg.tstep, w, moist, 1)
            p.same(f"{case} caar {(nm1, n0, np1)}", changed=("v", "t", "dp3d"))
        p.call("hxx_hypervis_run", 2, cfg.tstep, 1.0)
        # vtens / ttens are scratch: the reference leaves dt * tens * rspheremp there, the CUDA kernel keeps it in registers
        p.same(case + " hypervis", changed=("v", "t"), skip=("vtens", "ttens"))
        p.reset()
        p.call("hxx_euler_reset")
        p.call("hxx_euler_precompute_divdp")
        p.same(case + " precompute_divdp")
        dtq = cfg.tstep * cfg.qsplit
        for (np1q, n0q, rhs, opt) in [(1, 0, 0.0, 2), (1, 1, 1.0, 0), (1, 1, 2.0, 1)]:
            p.call("hxx_euler_step", np1q, n0q, dtq / 2.0, rhs, opt)
            p.same(f"{case} euler_step rhs_multiplier={rhs}", changed=("qdp",) if cfg.qsize else ())
        p.call("hxx_euler_qdp_time_avg", 0, 1)
        p.same(case + " qdp_time_avg")
        p.reset()
        p.call("hxx_caar_run", 1, 1, 2, cfg.tstep, 1.0, moist, 1)
        p.call("hxx_vertical_remap", 2, 1, dtq * max(cfg.rsplit, 1))
        p.call("hxx_update_q", 1, 2)
        p.same(case + " remap + update_q")
    finally:
        p.close()
