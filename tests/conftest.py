import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "integrating-diagenetic-equations-using-python_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for both CPU (symbol checks) and GPU tests."""
    import __graft_entry__ as entry
    entry.build()


@pytest.fixture(scope="session")
def rhs_golden():
    g = np.load(os.path.join(GOLDEN, "rhs_reference.npz"))
    meta = json.loads(str(g["__meta__"]))
    return g, meta


@pytest.fixture(scope="session")
def rhs_golden_vardphi():
    """Outputs of the reference's pde_rhs / fun with its commented-out time-varying dPhi line switched on
    (tests/golden/make_golden.py vardphi)."""
    g = np.load(os.path.join(GOLDEN, "rhs_reference_vardphi.npz"))
    meta = json.loads(str(g["__meta__"]))
    return g, meta


@pytest.fixture(scope="session")
def stepper_golden():
    g = np.load(os.path.join(GOLDEN, "stepper_reference.npz"))
    cases = json.loads(str(g["__cases__"]))
    return g, cases


@pytest.fixture(scope="session")
def fixtures_reference():
    return np.load(os.path.join(GOLDEN, "fixtures_reference.npz"))


@pytest.fixture(scope="session")
def scenario_reference():
    with open(os.path.join(GOLDEN, "scenario_reference.json")) as fh:
        return json.load(fh)


def rhs_states(g, name):
    """Yield (state_name, y, rhs_numba, rhs_numpy, events) for one golden parameter set."""
    for key in sorted(k for k in g.files if k.startswith(name + "/") and k.endswith("/y")):
        stem = key[:-2]
        yield stem.split("/", 1)[1], g[key], g[stem + "/rhs_numba"], g[stem + "/rhs_numpy"], g[stem + "/events"]
