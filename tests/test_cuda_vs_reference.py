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
