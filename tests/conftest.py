import os
import sys

import pytest

# fresh device arenas are filled with garbage in the tests: a block that is not properly initialised must
# fail parity instead of hiding behind zero pages from cudaMalloc (read once, when libnvbx.so is loaded)
os.environ.setdefault('NVBX_POISON_ARENAS', '1')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
    """GPU tests fail loudly (not skip) on a box whose GPU is missing when -m gpu is requested."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    wants_gpu = 'gpu' in (config.getoption('-m') or '') and 'not gpu' not in (config.getoption('-m') or '')
    if wants_gpu:
        return    # let them run and fail: a GPU run without a GPU is an error, not a skip
    skip = pytest.mark.skip(reason='no CUDA device in this container (run under gpurun with -m gpu)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)

# The reference's own test files are staged (unmodified, git-ignored) under tests/ref_tests/_ref/ and are run THROUGH
# tests/test_gpu_reference_suite.py (which supplies the package layout and stubs they need), never collected directly.
collect_ignore_glob = ['ref_tests/_ref/*']
