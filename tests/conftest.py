import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hockey-vision-analytics_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "tcgen05: exercises the tcgen05 Gram kernel")


@pytest.fixture(scope="session")
def ctx():
    """A libhvb context on cuda:0 — fails loudly (no CPU fallback) if the library or GPU is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a GPU"
    from hvb.runtime import get_context
    return get_context(0)
