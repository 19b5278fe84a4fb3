import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the reference checkout at /root/reference')
    config.addinivalue_line('markers', 'halotools: parity with a real halotools install (skipped '
                            'where it is not importable)')


@pytest.fixture(scope='session')
def golden():
    with np.load(os.path.join(GOLDEN_DIR, 'reference_outputs.npz')) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN_DIR
