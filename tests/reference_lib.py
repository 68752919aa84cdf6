"""The REFERENCE's own preqx C++ library, built by oracle/Makefile (target ref_full) from the sources under
/root/reference/src/share/cxx against the serial Kokkos stand-in of oracle/ref_shim: oracle/_ref/libref_hommexx_*.so.
TEST INFRASTRUCTURE — the pin of the oracle, and (on the GPU box, where the prebuilt .so travels with the snapshot)
a second checker of the CUDA product."""
import pathlib

import pytest

REF_DIR = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_ref"


def reference_lib(nlev: int, qsize_d: int) -> pathlib.Path:
    p = REF_DIR / f"libref_hommexx_{nlev}_{qsize_d}.so"
    if not p.exists():
        pytest.skip(f"{p.name} not built (needs the reference tree: make -C oracle ref_full)")
    return p
