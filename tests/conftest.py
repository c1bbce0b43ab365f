import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# parity tolerances of BASELINE.json:north_star (relative L2 over every output element)
TOL = {"float64": 1e-12, "float32": 1e-5}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU checker; builds oracle/liboracle.so (and oracle/_ref where /root/reference exists)."""
    from oracle import oracle

    oracle.build(("oracle",))
    if os.path.exists("/root/reference/kronmult_omp/kronmult.hpp"):
        oracle.build(("ref",))
    return oracle


@pytest.fixture(scope="session")
def kron():
    """The product library through its Python mirror (fails loudly if the .so cannot be loaded)."""
    from kronmult993_b200 import api

    api.load_library()
    return api
