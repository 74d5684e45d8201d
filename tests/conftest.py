import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sp():
    """The product package (ctypes over libspmv_b200.so); builds the library if absent."""
    from _load_pkg import load_pkg
    lib = os.path.join(ROOT, "gpu-spmv_b200", "lib", "libspmv_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    return load_pkg()


@pytest.fixture(scope="session")
def orc():
    from oracle_binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle_binding import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libspmv_ref.so not built (reference tree absent)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        spmv = np.load(os.path.join(GOLDEN, "spmv_property_cases.npz"))
        c1 = np.load(os.path.join(GOLDEN, "config1_random10k.npz"))
        pr = np.load(os.path.join(GOLDEN, "pagerank_cases.npz"))
        known = np.load(os.path.join(GOLDEN, "known_answers.npz"))
    return G


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
