import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """CPU-side artefacts (driver + oracle) are built on demand; the CUDA libs by build()."""
    from hommexx_b200 import homme
    from oracle import oraclelib
    if not homme.DRIVER_LIB.exists() or not oraclelib.ORACLE_LIB.exists():
        import __graft_entry__ as g
        g.build_host()
    yield
