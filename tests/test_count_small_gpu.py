"""One-launch COUNT for batches of small graphs (csrc/count_small.cu): bit-exact vs the C oracle and vs the general
path (gsn_graph_build + gsn_count_pattern), including the inputs the reference's tests of this path care about:
self loops, duplicate columns, asymmetric edge_index, edgeless graphs, columns that are not grouped by graph."""
import numpy as np
import pytest
import torch

from oracle import count_c, count_vf2
from tests.test_count_gpu import FAMILIES, _dicts
from tests.util import batch_graphs, random_graph

pytestmark = pytest.mark.gpu


def _ids(node_ptr, ei, sds, induced, scope_name, small, **kw):
    from gsn_b200 import _lib, counting
    l0 = _lib.launch_count()
    old = counting.SMALL_PATH
    counting.SMALL_PATH = small
    try:
        out = counting.count_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(node_ptr), sds, induced, scope_name, **kw)
    finally:
        counting.SMALL_PATH = old
    return out.cpu().numpy(), _lib.launch_count() - l0


@pytest.mark.parametrize('family', list(FAMILIES))
@pytest.mark.parametrize('scope_name', ['global', 'local'])
@pytest.mark.parametrize('induced', [False, True])
def test_small_path_vs_oracle_and_general_path(family, scope_name, induced, graphlet_patterns):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f'{family}/{scope_name}/{induced}/7'.encode()))      # reproducible across processes
    els = FAMILIES[family](graphlet_patterns)
    graphs = []
    for _ in range(60):
        n = int(rng.integers(1, 30))
        graphs.append((random_graph(rng, n, float(rng.uniform(0.05, 0.5))), n))
    graphs.insert(5, (np.zeros((2, 0), np.int64), 3))          # edgeless graph inside the batch
    graphs.append((random_graph(rng, 64, 0.08), 64))           # largest graph the path takes
    graphs.append((np.zeros((2, 0), np.int64), 2))             # edgeless graph at the end
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    scope = 1 if scope_name == 'local' else 0
    sds = _dicts(els, scope_name)
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), induced, scope)
    got, n_small = _ids(node_ptr, ei, sds, induced, scope_name, True)
    ref, n_general = _ids(node_ptr, ei, sds, induced, scope_name, False)
    assert got.shape == exp.shape and got.dtype == np.int64
    assert np.array_equal(got, exp)
    assert np.array_equal(ref, exp)
    assert n_small < n_general                                  # one launch per plan instead of build + count + write-out


@pytest.mark.parametrize('scope_name', ['global', 'local'])
def test_dense_graphs_take_the_global_accumulator_pass(scope_name):
    """64-node graphs at density 0.6: ~2400 slots per graph x 4 columns do not fit the shared-memory accumulators"""
    rng = np.random.default_rng(11)
    graphs = [(random_graph(rng, 64, 0.6), 64) for _ in range(3)] + [(random_graph(rng, 20, 0.3), 20) for _ in range(10)]
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    els = count_vf2.pattern_edge_lists('cycle_graph', 5)
    scope = 1 if scope_name == 'local' else 0
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), False, scope)
    got, _ = _ids(node_ptr, ei, _dicts(els, scope_name), False, scope_name, True)
    assert np.array_equal(got, exp)


def test_columns_not_grouped_by_graph_fall_back():
    """arbitrary column order: GSN_S_NOT_GROUPED -> the general path, same identifiers (rows follow the columns)"""
    from gsn_b200 import _lib, counting
    rng = np.random.default_rng(3)
    graphs = [(random_graph(rng, int(rng.integers(4, 20)), 0.4), 0) for _ in range(30)]
    graphs = [(g, int(g.max()) + 1 if g.size else 3) for g, _ in graphs]
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    perm = rng.permutation(ei.shape[1])
    ei_p = np.ascontiguousarray(ei[:, perm])
    els = count_vf2.pattern_edge_lists('cycle_graph', 6)
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, 'local'), False, 1)
    got, _ = _ids(node_ptr, ei_p, _dicts(els, 'local'), False, 'local', True)
    assert np.array_equal(got, exp[perm])
    # check=False: the caller sees the status bit
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    counting.count_batch(torch.from_numpy(ei_p).cuda(), torch.from_numpy(node_ptr), _dicts(els, 'local'), False, 'local',
                         check=False, status=st, max_nodes_per_graph=32)
    assert int(st.item()) & _lib.S_NOT_GROUPED


def test_self_loops_duplicates_and_asymmetric_columns():
    rng = np.random.default_rng(9)
    g = random_graph(rng, 14, 0.4)
    g2 = np.concatenate([g, np.array([[3, 5], [3, 5]]), g[:, :4]], 1)       # self loops + duplicated columns
    node_ptr, edge_ptr, ei = batch_graphs([(g2, 14), (random_graph(rng, 9, 0.5), 9)])
    els = count_vf2.pattern_edge_lists('cycle_graph', 6)
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, 'local'), False, 1)
    got, _ = _ids(node_ptr, ei, _dicts(els, 'local'), False, 'local', True)
    assert np.array_equal(got, exp)
    # asymmetric: a triangle with two directions missing -> the reference's KeyError (utils_graph_processing.py:173)
    tri = np.array([[0, 1, 2, 1], [1, 2, 0, 0]], np.int64)
    with pytest.raises(KeyError):
        _ids(np.array([0, 3]), tri, _dicts([[(0, 1), (1, 2), (2, 0)]], 'local'), False, 'local', True)
