import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "rust-llkv_b200"), ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device.  On a machine without one (this container, CPU-only CI) they are skipped, so a
    plain `pytest tests` is green; LLKV_REQUIRE_GPU=1 keeps the loud failure (a missing device is then an error, never a
    silent pass).  A missing CUDA library is always an error: gpu.load() raises."""
    markexpr = (config.getoption("-m") or "").strip()
    if os.environ.get("LLKV_REQUIRE_GPU") == "1" or markexpr == "gpu" or not any("gpu" in it.keywords for it in items):
        return  # an explicit `-m gpu` run asks for the device: fail loudly without one
    from llkv_b200 import gpu
    if gpu.device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (set LLKV_REQUIRE_GPU=1 to fail instead)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One llkv_gpu context for the whole GPU session.  Fails loudly (no skip, no fallback) without a device."""
    from llkv_b200 import gpu
    ctx = gpu.Context(0)
    yield ctx
    ctx.close()
