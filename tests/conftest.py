import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pmo():
    """The CPU oracle (test infrastructure)."""
    import pmo as _pmo
    _pmo.build()
    return _pmo


@pytest.fixture(scope="session")
def pkg():
    """The product package `ocean-perception_b200` (hyphenated: importlib only)."""
    return importlib.import_module("ocean-perception_b200")


@pytest.fixture(scope="session")
def built_lib(pkg):
    build = importlib.import_module("ocean-perception_b200.build")
    return build.build()


@pytest.fixture(scope="session")
def c1():
    """fsl1/fsr1 at 376x240 with the cv2-literal seeds (oracle/gen_goldens.py)."""
    return dict(np.load(os.path.join(GOLDEN, "c1_inputs.npz")))


@pytest.fixture(scope="session")
def c1_cpu():
    return dict(np.load(os.path.join(GOLDEN, "c1_cpu.npz")))


@pytest.fixture(scope="session")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "kat.npz")))


def make_engine(pkg, built_lib, **kw):
    P = pkg.PatchmatchGpu.Params()
    for k, v in kw.items():
        if not hasattr(P, k):
            raise AttributeError(k)
        setattr(P, k, v)
    return pkg.PatchmatchGpu(P, device=0)


@pytest.fixture()
def engine_factory(pkg, built_lib):
    made = []

    def factory(**kw):
        e = make_engine(pkg, built_lib, **kw)
        made.append(e)
        return e

    yield factory
    for e in made:
        e.close()
