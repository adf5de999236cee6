import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def xi_stats():
    z = np.load(os.path.join(ROOT, 'deepxi_b200', 'data', 'xi_stats.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def built_lib():
    """Builds libdeepxi_b200.so if needed (nvcc cross-compiles without a GPU) and returns its path."""
    from deepxi_b200 import build
    return build.build()
