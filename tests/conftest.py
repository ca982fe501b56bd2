import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:       # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped — not failed — on a box without a CUDA device (plain `pytest tests` on a laptop)."""
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _built(request):
    """The native library and the CPU oracle are built ahead of the tests (idempotent, seconds). A box with neither a
    prebuilt library nor nvcc can still run the tests that need no native code: those that do fail at their import."""
    from megastep_b200 import build
    try:
        build.build()
    except RuntimeError as e:
        if 'nvcc not found' not in str(e):
            raise
        print(f'conftest: {e}; tests that load libmegastep_b200.so will fail')
    from oracle import oracle
    oracle.lib()
