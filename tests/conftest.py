import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are the parity tests proper: they need a CUDA device AND the built C-ABI library.  On a box without
    either they are skipped with the reason, instead of failing inside torch (`-m "not gpu"` deselects them anyway)."""
    import torch

    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device (run on the B200 box: pytest -m gpu)"
    else:
        from d3pm_b200 import _lib
        if not os.path.isfile(_lib.library_path()):
            reason = f"{_lib.library_path()} is not built (python __graft_entry__.py build)"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
