import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def imdb_fixture():
    """graph-tool output shipped by the reference (re-packed by scripts/make_golden.py)"""
    z = np.load(os.path.join(GOLDEN, 'imdb_k5_edge_counts.npz'))
    return {'node_ptr': z['node_ptr'].astype(np.int64), 'edge_ptr': z['edge_ptr'].astype(np.int64),
            'edge_index': z['edge_index'].astype(np.int64), 'identifiers': z['identifiers'].astype(np.int64)}


@pytest.fixture(scope='session')
def sr_fixture():
    z = np.load(os.path.join(GOLDEN, 'sr251256.npz'))
    return z['edge_index'].astype(np.int64)      # [15, 2, 300]


@pytest.fixture(scope='session')
def graphlet_patterns():
    z = np.load(os.path.join(GOLDEN, 'graphlets.npz'))
    out = {}
    for k in range(2, 7):
        ptr, ed = z[f'k{k}_ptr'], z[f'k{k}_edges']
        out[k] = [ed[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]
    return out
