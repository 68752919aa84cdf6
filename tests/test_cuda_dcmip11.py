"""BASELINE configs[2] on the GPU: DCMIP 2012 test 1-1 (3-D deformational flow) with its ANALYTIC winds
(src/test_src/dcmip2012_test1_2_3.F90:87-270) driving the CUDA library through the functors' run methods
(hommexx_b200/dcmip.py: EulerStep x 3 + time average + PPM remap per step), checked against the CPU oracle on the
same inputs: every tracer bit-identical, masses conserved to round-off. ne30 is the configuration BASELINE names."""
import numpy as np
import pytest

import parity
from hommexx_b200 import dcmip
from oracle import oraclelib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ne,nsteps,tstep", [(8, 24, 600.0), (30, 6, 150.0)])
def test_dcmip11_analytic_winds_cuda_vs_oracle(ne, nsteps, tstep):
    parity.need_gpu()
    dc = dcmip.Dcmip11(ne, 26, parity.cuda_lib(26, 4), tstep=tstep)
    do = dcmip.Dcmip11(ne, 26, oraclelib.ORACLE_LIB, tstep=tstep)
    assert dc.h.lib.hommexx_b200_backend() == b"cuda-sm100a"
    m0, q0 = dc.masses(), dc.q()
    for _ in range(nsteps):
        dc.step(); do.step()
    qc, qo = dc.q(), do.q()
    assert np.isfinite(qo).all()
    assert np.array_equal(qc, qo), float(np.abs(qc - qo).max())          # tol = 0: bit-identical
    for name in ("qdp", "Q", "dp3d", "ps_v", "qlim", "divdp_proj"):
        assert np.array_equal(dc.h.get_field(name), do.h.get_field(name)), name
    assert np.abs(dc.masses() / m0 - 1.0).max() <= 1e-13                  # tracer mass: round-off
    for i in range(4):                                                     # the limiter keeps the initial global ranges
        lo, hi = q0[i].min(), q0[i].max()
        assert qc[i].min() >= lo - 1e-2 * (hi - lo) and qc[i].max() <= hi + 1e-2 * (hi - lo), i
    assert np.abs(qc[0] - q0[0]).max() > 0.02                              # and the flow moved the bells
    dc.close(); do.close()
