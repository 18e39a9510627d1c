import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with `-m gpu` on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a B200 and the built engine: skip them (instead of 100+ 'no NVIDIA driver' failures that hide
    real regressions) on a CPU box.  On a GPU box nothing is skipped: a missing library fails loudly."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The CPU oracle is test infrastructure; compile it once per session (gcc, ~1 s)."""
    from oracle import oracle
    oracle.load()
