"""THE PIN OF THE ORACLE: oracle/oracle.c against the reference's own C++ sources.

oracle/_ref/libref_hommexx_<PLEV>_<QSIZE_D>.so is every file of /root/reference/src/share/cxx on the
prim_run_subcycle_c path — CaarFunctorImpl.hpp, SphereOperators.hpp, HyperviscosityFunctorImpl.{hpp,cpp},
EulerStepFunctorImpl.hpp, RemapFunctor.hpp / PpmRemap.hpp, mpi/BoundaryExchange.cpp, CamForcing.cpp, Diagnostics.cpp,
prim_driver.cpp, cxx_f90_interface.cpp ... — compiled where it lies (oracle/Makefile, ref_full) against a serial
stand-in for Kokkos (oracle/ref_shim), VECTOR_SIZE 1, -ffp-contract=off. Both libraries are driven through the same
C ABI by the same host driver on the same inputs; after 10-12 dynamics steps every prognostic array must agree
BIT FOR BIT (the reference's own acceptance test compares its C++ and Fortran builds the same way,
cmake/CxxVsF90.cmake.in:28-41). With that, "CUDA bit-identical to the oracle" (tests/test_cuda_*.py) is
"CUDA bit-identical to the reference's own build"."""
import numpy as np
import pytest

import distinct_tracers
from forcing_inputs import fill_smooth_forcing
from hommexx_b200 import homme
from oracle import oraclelib
from reference_lib import reference_lib

PROGNOSTIC = ["v", "T", "dp3d", "ps_v", "Qdp", "Q", "omega_p"]

CASES = {
    # BASELINE configs[0]: ne4, nlev 72, qsize 4
    "ne4": dict(),
    # the reference's prtcA executable sizes (PLEV 26, QSIZE_D 4) and the option variants of its test-list.cmake
    "prtcA": dict(),
    "prtcA-lim9-alg2": dict(base="prtcA", limiter_option=9, remap_alg=2),
    "prtcA-moist-nudiv-q2": dict(base="prtcA", moisture=1, nu_div=1.75e16, qsplit=2, rsplit=2),
    "prtcA-r0-moist-q3": dict(base="prtcA", rsplit=0, moisture=1, qsplit=3),
    "ne4-r0": dict(base="ne4", rsplit=0),
    "prtcA-tensorhv": dict(base="prtcA", hypervis_scaling=3.2, nu=1e-9, nu_p=1e-9, nu_q=1e-9, nu_s=1e-9, nu_div=1e-9,
                           hypervis_subcycle=2),
    "ne4-tensorhv-nudiv": dict(base="ne4", hypervis_scaling=3.2, nu=5e-8, nu_p=5e-8, nu_q=5e-8, nu_s=5e-8, nu_div=1.25e-7),
    "prtcA-q0": dict(base="prtcA", qsize=0),
    # the benchmarked dimensions (PLEV 72, QSIZE_D 40) with 40 pairwise-distinct tracers
    "ne4-q40": dict(base="ne4", qsize=40, qsize_d=40),
    "ne4-q35of40-moist-q2": dict(base="ne4", qsize=35, qsize_d=40, moisture=1, qsplit=2, rsplit=2),
}


def run(cfg, lib, ncalls_min_steps=10, forcing=False, diagnostics=False):
    h = homme.Homme(cfg, lib)
    if cfg.qsize > 4:
        distinct_tracers.install(h)
    h.init_dycore()
    nstep = 0
    while nstep < ncalls_min_steps:
        if forcing:
            fill_smooth_forcing(h)
            h.push_forcing()
        nstep = h.run_subcycle()
    h.push_results()
    out = {k: v.copy() for k, v in h.state().items()}
    if diagnostics:
        out.update({"accum_" + k: v.copy() for k, v in h.accum().items()})
    out["_tl"] = np.array(h.time_levels())
    h.close()
    return out


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_is_bit_identical_to_the_reference_build(case):
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    ref = run(cfg, reference_lib(cfg.nlev, cfg.qsize_d))
    ora = run(cfg, oraclelib.ORACLE_LIB)
    assert np.array_equal(ref["_tl"], ora["_tl"])
    for k in PROGNOSTIC:
        assert not np.isnan(ref[k]).any(), k
        assert np.array_equal(ref[k], ora[k]), (case, k, float(np.abs(ref[k] - ora[k]).max()))


@pytest.mark.parametrize("moist,ftype", [(0, 0), (1, 0), (0, 2)])
def test_forced_runs_match_the_reference_build(moist, ftype):
    """CAM forcing through f90_push_forcing_to_cxx (CamForcing.cpp), pushed before every call."""
    cfg = homme.preset("ne4", moisture=moist, ftype=ftype)
    ref = run(cfg, reference_lib(cfg.nlev, cfg.qsize_d), 6, forcing=True)
    ora = run(cfg, oraclelib.ORACLE_LIB, 6, forcing=True)
    for k in PROGNOSTIC:
        assert np.array_equal(ref[k], ora[k]), (k, float(np.abs(ref[k] - ora[k]).max()))


def test_diagnostics_match_the_reference_build():
    """Diagnostics.cpp: the energy and tracer-mass accumulators written into the arrays of init_diagnostics_c."""
    cfg = homme.preset("prtcA", disable_diagnostics=0, state_frequency=3)
    ref = run(cfg, reference_lib(cfg.nlev, cfg.qsize_d), 9, diagnostics=True)
    ora = run(cfg, oraclelib.ORACLE_LIB, 9, diagnostics=True)
    for k in ref:
        assert np.array_equal(ref[k], ora[k]), (k, float(np.abs(ref[k] - ora[k]).max()))


def test_reference_timing_build_agrees_with_its_serial_build():
    """oracle/_ref/libref_hommexx_72_40_omp.so is what bench.py's CPU arm times: the same sources with AVX2 vector packs,
    FMA contraction and OpenMP over elements. It must agree with the serial scalar build to round-off (north-star
    tolerance 1e-11) — which also checks that the stand-in's thread-parallel league loop races on nothing."""
    import os
    from reference_lib import REF_DIR
    omp = REF_DIR / "libref_hommexx_72_40_omp.so"
    if not omp.exists() or " avx2" not in open("/proc/cpuinfo").read():
        pytest.skip("timing build of the reference absent (or no AVX2 on this host)")
    os.environ.setdefault("OMP_NUM_THREADS", str(min(8, os.cpu_count() or 1)))
    cfg = homme.preset("ne4", qsize=40, qsize_d=40)
    a = run(cfg, omp)
    b = run(cfg, reference_lib(72, 40))
    for k in ("v", "T", "dp3d", "ps_v", "Qdp", "Q"):
        err = np.sqrt(((a[k] - b[k]) ** 2).sum()) / np.sqrt((b[k] ** 2).sum())
        assert err <= 1e-11, (k, err)


def test_long_run_stays_bit_identical():
    """120 dynamics steps (30 calls) of the moist, limiter-9, qsplit 2 / rsplit 2 variant: two and a half simulated hours
    past the ten-step cases, through many more limiter iterations and remap configurations, still bit for bit."""
    cfg = homme.preset("prtcA", moisture=1, qsplit=2, rsplit=2, limiter_option=9)
    ref = run(cfg, reference_lib(cfg.nlev, cfg.qsize_d), 120)
    ora = run(cfg, oraclelib.ORACLE_LIB, 120)
    assert np.array_equal(ref["_tl"], ora["_tl"]) and ref["_tl"][0] == 120
    for k in PROGNOSTIC:
        assert np.array_equal(ref[k], ora[k]), (k, float(np.abs(ref[k] - ora[k]).max()))
