"""Where the CPU oracle lives and how to load it. TEST INFRASTRUCTURE: imported by tests/, by
__graft_entry__.smoke() and by bench.py's CPU arm only — never by the product package hommexx_b200/."""
import ctypes as C
import pathlib

ORACLE_LIB = pathlib.Path(__file__).resolve().parent / "liboracle.so"


def load_oracle(nlev: int, qsize_d: int) -> C.CDLL:
    """dlopen oracle/liboracle.so with the product's section B/C signatures and set its run-time dimensions."""
    from hommexx_b200 import homme
    lib = homme.load_dycore(ORACLE_LIB)
    lib.oracle_set_dims.argtypes = [C.c_int, C.c_int]
    lib.oracle_set_dims(nlev, qsize_d)
    return lib
