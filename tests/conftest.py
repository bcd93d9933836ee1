import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "consts_golden.npz"))


@pytest.fixture(scope="session")
def halo():
    import shutil

    import halo_accumulation_b200 as H

    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        H.build()  # no-op when the in-tree libraries are newer than their sources
    return H


@pytest.fixture(scope="session")
def ctx(halo):
    """A device context with 2^16 derived generators.  No fallback: with a GPU present a missing or
    broken libhalo_b200.so fails the test run."""
    if not _have_cuda():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    c = halo.Context(0, 1 << 20)
    c.derive_generators(1 << 16)
    yield c
    bad, live = halo.check_canaries()  # no kernel of the whole session wrote past the end of a device buffer
    c.close()
    assert bad == 0 and live > 10, (bad, live)
