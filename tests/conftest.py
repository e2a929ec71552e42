import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


def _have_b200():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a B200 runs the CPU suite and reports the GPU tests as skipped
    (libgpp has no CPU fallback, so they could only fail in gpp_create)."""
    if _have_b200():
        return
    skip = pytest.mark.skip(reason='needs a B200 (sm_100) device; libgpp has no CPU fallback')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_planes(tag):
    return np.load(os.path.join(ROOT, 'road_planes_database', 'road_planes_database_%s.npy' % tag))


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and not f.startswith(('detect_', 'driver_', 'pose_')))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    g['planes_raw'] = load_planes(str(g['planes_db'])) if 'planes_db' in g else g['planes']
    return g


@pytest.fixture(scope='session')
def built():
    """Build the native pieces once per session (nvcc cross-compiles on CPU; on the GPU box the prebuilt
    .so files that travelled with the snapshot are up to date and this is a no-op)."""
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture(scope='session')
def gpp(built):
    import gpp_b200
    return gpp_b200


@pytest.fixture(scope='session')
def poller(gpp):
    return gpp.get_poller(0)
