"""The C-ABI shared libraries load on a machine without a GPU and export every symbol that
include/hommexx_b200.h declares (no compute call is made here)."""
import ctypes as C
import pathlib
import re

import pytest

import __graft_entry__ as g
from hommexx_b200 import homme
from oracle import oraclelib

HEADER = pathlib.Path(__file__).resolve().parents[1] / "include" / "hommexx_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b([a-z][a-z0-9_]*(?:_c|_f90|_cxx|_session|_comm|_connectivity|_connection)?)\s*\(", text))
                  - {"defined", "extern"})


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for s in ("reset_cxx_comm", "initialize_hommexx_session", "finalize_hommexx_session", "init_connectivity",
              "add_connection", "finalize_connectivity", "init_derivative_c", "init_simulation_params_c",
              "init_elements_2d_c", "init_elements_states_c", "init_diagnostics_c", "init_hvcoord_c",
              "init_boundary_exchanges_c", "init_time_level_c", "prim_run_subcycle_c", "cxx_push_results_to_f90",
              "f90_push_forcing_to_cxx", "cxx_push_forcing_to_f90"):
        assert s in syms, s


@pytest.mark.parametrize("variant", g.CUDA_VARIANTS)
def test_cuda_library_exports_every_declared_symbol(variant):
    path = homme.cuda_lib_path(*variant)
    if not path.exists():
        g.build_cuda([variant])
    lib = C.CDLL(str(path))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{path.name} lacks {s}"
    lib.hommexx_b200_backend.restype = C.c_char_p
    assert lib.hommexx_b200_backend() == b"cuda-sm100a"
    assert (lib.hommexx_b200_nlev(), lib.hommexx_b200_qsize_d()) == variant


def test_oracle_exports_the_same_abi():
    lib = C.CDLL(str(oraclelib.ORACLE_LIB))
    for s in declared_symbols():
        if s == "hommexx_b200_nccl_unique_id":
            continue  # multi-GPU wiring exists only in the product
        assert hasattr(lib, s), s
