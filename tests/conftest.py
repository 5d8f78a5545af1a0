"""pytest configuration: markers, paths and shared fixtures.

-m "not gpu"  covers the oracle against the committed golden vectors, the host logic of libhadi.so, the
              CPU emulation of the kernel phases and the C-ABI surface (no compute calls);
-m gpu        are the parity tests proper: the CUDA path through the C ABI against the oracle.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "emu")
if EMU not in sys.path:
    sys.path.insert(0, EMU)

BASE = dict(S0=100.0, V0=0.04, r_d=0.025, r_f=0.0, rho=-0.9, sigma=0.3, kappa=1.5, eta=0.04)
DIVS = ([0.2, 0.4, 0.6, 0.8], [0.5, 0.3, 0.2, 0.1], [0.02, 0.02, 0.02, 0.02])
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    import __graft_entry__ as ge

    hadi = ge.load_hadi()
    from oracle.reflib import oracle_path

    if not os.path.exists(hadi.LIB_PATH) or not os.path.exists(oracle_path()):
        ge.build()
    return ge, hadi


@pytest.fixture(scope="session")
def hadi():
    return _ensure_built()[1]


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    from oracle.reflib import OracleLib

    return OracleLib()


@pytest.fixture(scope="session")
def reflib():
    """The real reference (oracle/_ref); only present where it was built from /root/reference."""
    from oracle.reflib import RefLib, have_ref

    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefLib()


@pytest.fixture(scope="session")
def ctx(hadi):
    return hadi.Context(0)


def golden(name):
    import json

    return json.load(open(os.path.join(GOLDEN, name)))
