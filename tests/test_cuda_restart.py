"""Checkpoint / resume of the CUDA dycore through the C ABI (tests/restart_check.py): a new session initialised from
cxx_push_results_to_f90's arrays and the time levels continues bit-identically — and a run checkpointed on the
REFERENCE's own build continues on the GPU to the reference's own final state (switching libraries mid-run)."""
import numpy as np
import pytest

import distinct_tracers
import parity
import restart_check
from hommexx_b200 import homme
from reference_lib import reference_lib

pytestmark = pytest.mark.gpu

CASES = {
    "ne4": dict(),
    "prtcA-moist-q2": dict(base="prtcA", moisture=1, qsplit=2, rsplit=2),
    "prtcA-r0": dict(base="prtcA", rsplit=0),
    "ne4-q40": dict(base="ne4", qsize=40, qsize_d=40),
}


@pytest.mark.parametrize("case", list(CASES))
def test_cuda_restart_run_is_bit_identical(tmp_path, case):
    parity.need_gpu()
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)
    restart_check.restart_is_bit_identical(cfg, parity.cuda_lib(cfg.nlev, cfg.qsize_d), tmp_path / "R.npz",
                                           first=3 if cfg.rsplit == 0 else 1, more=2)


@pytest.mark.parametrize("case", ["ne4", "ne4-q40"])
def test_reference_checkpoint_resumed_on_the_gpu(tmp_path, case):
    parity.need_gpu()
    over = dict(CASES[case])
    cfg = homme.preset(over.pop("base", case), **over)

    def new(lib):
        h = homme.Homme(cfg, lib)
        if cfg.qsize > 4:
            distinct_tracers.install(h)
        return h

    ref = new(reference_lib(cfg.nlev, cfg.qsize_d))
    ref.init_dycore()
    ref.run_subcycle()
    ref.write_restart(tmp_path / "R.npz")
    for _ in range(3):
        ref.run_subcycle()
    ref.push_results()
    want = {k: v.copy() for k, v in ref.state().items()}
    tl = ref.time_levels()
    ref.close()
    cu = new(parity.cuda_lib(cfg.nlev, cfg.qsize_d))
    cu.read_restart(tmp_path / "R.npz")
    cu.init_dycore()
    assert cu.lib.hommexx_b200_backend() == b"cuda-sm100a"
    for _ in range(3):
        cu.run_subcycle()
    cu.push_results()
    got = {k: v.copy() for k, v in cu.state().items()}
    assert cu.time_levels() == tl
    cu.close()
    for k in restart_check.PROGNOSTIC:
        assert np.array_equal(got[k], want[k]), (case, k, float(np.abs(got[k] - want[k]).max()))
