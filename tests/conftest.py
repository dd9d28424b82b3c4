import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# relative-L2 tolerance of the parity contract (BASELINE.json north_star): 1e-5 in float32
TOL = 1e-5


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box only)")


def rel_l2(a, b) -> float:
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64).ravel()
    b = np.asarray(b).astype(np.complex128 if np.iscomplexobj(b) else np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def port():
    import oracle

    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle

    r = oracle.ref()
    if r is None:
        pytest.skip("compiled reference (oracle/_ref/libclfft_ref.so) not present")
    return r


@pytest.fixture(scope="session")
def eng():
    """the product package, on a box with a GPU"""
    import opencl_fft_b200 as e

    if e.device_count() < 1:
        pytest.fail("no CUDA device visible: gpu-marked tests must run on the GPU box")
    return e


@pytest.fixture
def options(eng):
    """eng.set_option() with the defaults restored afterwards (options are process-wide defaults that a handle
    copies when it is created; include/b200fft.h lists them)"""
    saved = {}

    def setter(name, value):
        saved.setdefault(name, eng.get_option(name))
        eng.set_option(name, value)

    yield setter
    for name, value in saved.items():
        eng.set_option(name, value)
