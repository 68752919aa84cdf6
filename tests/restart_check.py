"""A restart run against the uninterrupted one (WriteRestart / ReadRestart, restart_io_mod.F90:626-698, as prim_main
drives them): checkpoint after `first` calls, resume in a NEW session from the file, compare after `more` calls."""
import numpy as np

import distinct_tracers
from hommexx_b200 import homme

PROGNOSTIC = ["v", "T", "dp3d", "ps_v", "Qdp", "Q", "omega_p"]


def _new(cfg, lib):
    h = homme.Homme(cfg, lib)
    if cfg.qsize > 4:
        distinct_tracers.install(h)
    return h


def restart_is_bit_identical(cfg, lib, path, first=2, more=2):
    # the uninterrupted run, writing its restart file on the way
    h = _new(cfg, lib)
    h.init_dycore()
    for _ in range(first):
        h.run_subcycle()
    h.write_restart(path)
    tl_ck = h.time_levels()
    for _ in range(more):
        h.run_subcycle()
    h.push_results()
    want = {k: v.copy() for k, v in h.state().items()}
    tl_want = h.time_levels()
    h.close()
    # the restart run: a new session initialised from the file
    r = _new(cfg, lib)
    r.read_restart(path)
    assert r.time_levels() == tl_ck and tl_ck[0] > 0
    r.init_dycore()
    for _ in range(more):
        r.run_subcycle()
    r.push_results()
    got = {k: v.copy() for k, v in r.state().items()}
    tl_got = r.time_levels()
    r.close()
    assert tl_got == tl_want
    for k in PROGNOSTIC:
        assert np.isfinite(want[k]).all(), k
        assert np.array_equal(got[k], want[k]), (k, float(np.abs(got[k] - want[k]).max()))
    return tl_got
