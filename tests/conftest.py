import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "rust-llkv_b200"), ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One llkv_gpu context for the whole GPU session.  Fails loudly (no skip, no fallback) without a device."""
    from llkv_b200 import gpu
    ctx = gpu.Context(0)
    yield ctx
    ctx.close()
