import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """Accessor for tests/golden/<name>.npz: g('case.key') -> torch tensor."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'))

    def __call__(self, key):
        a = self.z[key]
        return torch.from_numpy(a) if a.dtype.kind in 'fiu' and a.shape != () else a

    def keys(self):
        return list(self.z.keys())


@pytest.fixture(scope='session')
def golden():
    return Golden
