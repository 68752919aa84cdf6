"""Shared helpers of the CUDA parity tests: run the CUDA dycore and the CPU oracle side by side
through the same C ABI and compare device-layout fields."""
import numpy as np
import pytest

from hommexx_b200 import homme
from oracle import oraclelib

# every named device array whose meaning is identical in both libraries
STATE_FIELDS = ["v", "t", "dp3d", "ps_v", "omega_p", "eta_dot_dpdn", "derived_vn0", "derived_dp", "divdp",
                "divdp_proj", "dpdiss_ave", "dpdiss_biharmonic", "qdp", "qlim", "vtens", "ttens", "dptens", "Q"]
PROGNOSTIC = ["v", "T", "dp3d", "ps_v", "Qdp", "Q", "omega_p"]


def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container (run with -m gpu on the B200 box)")


def cuda_lib(nlev, qsize_d):
    p = homme.cuda_lib_path(nlev, qsize_d)
    if not p.exists():
        raise RuntimeError(f"{p} is not built; the CUDA dycore has no fallback. Run __graft_entry__.build().")
    return p


def pair(cfg, init="jw"):
    """(cuda, oracle) Homme objects on identical inputs, both initialised."""
    need_gpu()
    hc = homme.Homme(cfg, cuda_lib(cfg.nlev, cfg.qsize_d), init=init)
    ho = homme.Homme(cfg, oraclelib.ORACLE_LIB, init=init)
    hc.init_dycore()
    ho.init_dycore()
    assert hc.lib.hommexx_b200_backend() == b"cuda-sm100a"
    assert ho.lib.hommexx_b200_backend() == b"cpu-oracle"
    return hc, ho


def copy_state(src, dst, names=STATE_FIELDS):
    for n in names:
        dst.set_field(n, src.get_field(n))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    d = np.sqrt(((a - b) ** 2).sum())
    n = np.sqrt((b ** 2).sum())
    return float(d / n) if n > 0 else float(d)


def compare_fields(hc, ho, names=STATE_FIELDS, tol=0.0, what=""):
    """Normalised L2 difference of each named device array; tol=0 demands bit-identical results."""
    worst = {}
    for n in names:
        a, b = hc.get_field(n), ho.get_field(n)
        assert not np.isnan(b).any(), f"{what}: oracle {n} has NaN"
        e = 0.0 if np.array_equal(a, b) else rel_l2(a, b)
        worst[n] = e
    bad = {k: v for k, v in worst.items() if not (v <= tol)}
    assert not bad, f"{what}: fields differ beyond {tol:g}: {bad}"
    return worst
