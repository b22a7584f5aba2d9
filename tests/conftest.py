import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "ibm-cbc-genomic-tools_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference binaries built into oracle/_ref/")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _device_kept_initialised():
    """The command-line drivers are separate processes, each of which creates a CUDA context.  Without a client the driver tears the
    device's state down when a process exits and the next one pays for bringing it up again (about a second per invocation, several
    hundred invocations in the suite); one context held by the test session itself keeps the device initialised, as
    `nvidia-smi -pm 1` would.  Nothing is computed on it."""
    ctx = None
    if _have_gpu():
        try:
            import gtb200
            ctx = gtb200.Context(0)
        except Exception:
            ctx = None
    yield
    if ctx is not None:
        ctx.close()
