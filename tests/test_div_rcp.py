"""The kernels' reciprocal division (hxx.cuh div_rcp / div_rcp_plane: q = x r, e = fma(-d, q, x), fma(e, r, q) with
r = 1/d) must be the IEEE quotient bit for bit — it is what lets 16 values share one divisor's reciprocal
without leaving the oracle's arithmetic. oracle/div_rcp_check.c restates the formula in C and compares it
with x / d on operands built to hit the hard cases (mantissas next to powers of two, long runs of ones,
quotients perturbed around a rounding boundary); DESIGN.md quotes the 10^8-pair run, this test repeats 2 x 10^7."""
import pathlib
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
EXE = ROOT / "oracle" / "div_rcp_check"


def test_reciprocal_division_is_ieee_division():
    if "fma" not in pathlib.Path("/proc/cpuinfo").read_text():
        pytest.skip("host CPU without FMA")
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "div_rcp_check"], check=True)
    for seed in (1, 2):
        out = subprocess.run([str(EXE), "10", str(seed)], capture_output=True, text=True, check=True)
        pairs, bad = map(int, out.stdout.split())
        assert pairs == 10_000_000 and bad == 0, out.stderr
