import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        self.l1 = json.load(open(os.path.join(GOLDEN, 'l1_cases.json')))
        self.l1_arr = np.load(os.path.join(GOLDEN, 'l1_cases.npz'))
        self.scripts = json.load(open(os.path.join(GOLDEN, 'scripts.json')))
        self.scripts_arr = np.load(os.path.join(GOLDEN, 'scripts.npz'))
        self.scripts_fuzz = json.load(open(os.path.join(GOLDEN, 'scripts_fuzz.json')))
        self.scripts_fuzz_arr = np.load(os.path.join(GOLDEN, 'scripts_fuzz.npz'))
        self.scripts_fuzz_big = json.load(open(os.path.join(GOLDEN, 'scripts_fuzz_big.json')))
        self.scripts_fuzz_big_arr = np.load(os.path.join(GOLDEN, 'scripts_fuzz_big.npz'))
        self.probval = json.load(open(os.path.join(GOLDEN, 'probval.json')))
        self.rc = np.load(os.path.join(GOLDEN, 'rc_small.npz'))

    def cases(self, kind):
        return [c for c in self.l1 if c['kind'] == kind]

    def arr(self, key):
        return self.l1_arr[key]


@pytest.fixture(scope='session')
def golden():
    return Golden()


def close(a, b, rtol=1e-12):
    """|a - b| <= rtol * max|b| elementwise bound -- the fp64 tolerance BASELINE.json states
    (1e-12 relative), taken relative to the largest entry of the expected array."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return False
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    return bool(np.max(np.abs(a - b)) <= rtol * scale) if b.size else True
