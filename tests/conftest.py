import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_usable():
    """(ok, why): a CUDA device is visible and the in-tree library has been built."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device visible"
    except Exception as exc:                                  # pragma: no cover
        return False, "torch unusable: %r" % (exc,)
    if not os.path.exists(os.path.join(ROOT, "onekapy_b200", "liboneka_b200.so")):
        return False, "onekapy_b200/liboneka_b200.so has not been built (python __graft_entry__.py)"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU box SKIPS the gpu-marked tests instead of erroring in the Engine fixture.  On a GPU box
    nothing is skipped: a missing library there is an error the tests must show (there is no CPU fallback)."""
    ok, why = _gpu_usable()
    if ok:
        return
    try:
        import torch
        if torch.cuda.is_available():
            return                                            # GPU present but library missing: let the tests fail loudly
    except Exception:
        pass
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load
