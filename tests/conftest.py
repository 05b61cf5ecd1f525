import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMUL_DIR = os.path.join(ROOT, "tests", "host_emul")
EMUL_LIB = os.path.join(EMUL_DIR, "libfluidsolver_hostemul.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("3dfluidsimulation_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def emul_lib():
    """Host-emulation build of the solver core (test scaffolding, see tests/host_emul)."""
    subprocess.run(["make", "-C", EMUL_DIR, "-s"], check=True)
    return EMUL_LIB


@pytest.fixture(scope="session")
def cuda_lib(pkg):
    """The product library.  GPU tests fail (not skip) if it is missing."""
    bld = importlib.import_module("3dfluidsimulation_b200.build")
    if not os.path.exists(bld.LIB):
        bld.build()
    return bld.LIB
