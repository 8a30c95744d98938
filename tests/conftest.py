import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the oracle (gcc) and the CUDA library (nvcc, cross-compiles without a GPU) exist."""
    from oracle import oracle as O
    O.build()
    from mila_b200 import _lib
    # always run make (a no-op when the library is newer than every source): tests must never exercise a stale binary
    # after csrc edits.  If the rebuild fails but a library exists (e.g. a box without nvcc), say so and carry on with it.
    try:
        _lib.build()
    except Exception as e:                                                  # pragma: no cover
        if not _lib.LIB_PATH.exists():
            raise
        import warnings
        warnings.warn(f"could not rebuild libmila_b200_linear.so, testing the existing binary: {e}")
    yield
