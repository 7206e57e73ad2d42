"""COUNT parity: CUDA path (through the C ABI) vs the oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import count_c, count_vf2
from tests.util import batch_graphs, random_graph

pytestmark = pytest.mark.gpu


def _cuda_ids(node_ptr, ei, sds, induced, scope_name, **kw):
    from gsn_b200 import counting
    out = counting.count_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(node_ptr), sds, induced, scope_name, **kw)
    return out.cpu().numpy()


def _dicts(edge_lists, scope_name):
    from gsn_b200 import patterns
    return patterns.make_subgraph_dicts(edge_lists, scope_name)


def test_imdb_fixture_bit_exact(imdb_fixture):
    """the reference's own graph-tool output: K3/K4/K5 per directed edge, 1000 graphs"""
    f = imdb_fixture
    ei = f['edge_index'].copy()
    for g in range(len(f['node_ptr']) - 1):
        ei[:, f['edge_ptr'][g]:f['edge_ptr'][g + 1]] += f['node_ptr'][g]
    sds = _dicts(count_vf2.pattern_edge_lists('complete_graph', 5), 'local')
    got = _cuda_ids(f['node_ptr'], ei, sds, False, 'local')
    assert got.dtype == np.int64
    assert np.array_equal(got, f['identifiers'])
    # vertex scope through the invariant  count_v(K_k) = sum_out-edges count_e / (k-1)   
    sdv = _dicts(count_vf2.pattern_edge_lists('complete_graph', 5), 'global')
    gv = _cuda_ids(f['node_ptr'], ei, sdv, False, 'global')
    acc = np.zeros_like(gv)
    np.add.at(acc, ei[0], f['identifiers'])
    assert np.array_equal(gv * np.array([2, 3, 4]), acc)


def test_sr25_known_answers(sr_fixture):
    """SR(25,12,5,6), induced cycles k<=6, edge scope (README.md:84)"""
    graphs = [(sr_fixture[i], 25) for i in range(15)]
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    sds = _dicts(count_vf2.pattern_edge_lists('cycle_graph', 6), 'local')
    ids = _cuda_ids(node_ptr, ei, sds, True, 'local')
    assert ids.shape == (15 * 300, 4)
    assert (ids[:, 0] == 5).all()                                   # lambda = 5 triangles per edge
    g0 = ids[:300]
    assert sorted(np.unique(g0[:, 1], return_counts=True)[1].tolist()) == [48, 252]
    assert set(np.unique(g0[:, 1]).tolist()) == {16, 18}
    vals, cnts = np.unique(g0[:, 2], return_counts=True)
    assert dict(zip(vals.tolist(), cnts.tolist())) == {30: 24, 33: 24, 39: 144, 40: 36, 41: 72}
    g1 = ids[300:600]
    assert (g1[:, 1] == 16).all() and (g1[:, 2] == 40).all() and (g1[:, 3] == 24).all()
    n_c6 = [int(ids[i * 300:(i + 1) * 300, 3].sum()) // 12 for i in range(15)]
    assert n_c6 == [570, 600, 678, 769, 787, 519, 797, 845, 816, 764, 528, 796, 803, 815, 786]
    # oracle on two graphs
    exp = count_c.count_batch(node_ptr[:3], edge_ptr[:3], ei[:, :600], count_vf2.make_subgraph_dicts(
        count_vf2.pattern_edge_lists('cycle_graph', 6), 'local'), True, 1)
    assert np.array_equal(ids[:600], exp)


FAMILIES = {
    'cycles8': lambda gl: count_vf2.pattern_edge_lists('cycle_graph', 8),
    'cliques5': lambda gl: count_vf2.pattern_edge_lists('complete_graph', 5),
    'paths5': lambda gl: count_vf2.pattern_edge_lists('path_graph', 5),
    'stars4': lambda gl: count_vf2.pattern_edge_lists('star_graph', 4),
    'graphlets5': lambda gl: gl[3] + gl[4] + gl[5],
}


@pytest.mark.parametrize('family', list(FAMILIES))
@pytest.mark.parametrize('scope_name', ['global', 'local'])
@pytest.mark.parametrize('induced', [False, True])
def test_random_batches_vs_oracle(family, scope_name, induced, graphlet_patterns):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f'{family}/{scope_name}/{induced}'.encode()))      # reproducible across processes
    els = FAMILIES[family](graphlet_patterns)
    graphs = []
    for _ in range(40):
        n = int(rng.integers(1, 26))
        graphs.append((random_graph(rng, n, float(rng.uniform(0.05, 0.5))), n))
    graphs.append((np.zeros((2, 0), np.int64), 3))          # edgeless graph inside the batch
    graphs.append((random_graph(rng, 70, 0.06), 70))        # forces W = 2
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    scope = 1 if scope_name == 'local' else 0
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), induced, scope)
    got = _cuda_ids(node_ptr, ei, _dicts(els, scope_name), induced, scope_name)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp)


@pytest.mark.parametrize('family', ['cliques5', 'cycles5', 'graphlets4'])
@pytest.mark.parametrize('scope_name', ['global', 'local'])
def test_dense_batches_vs_oracle(family, scope_name, graphlet_patterns):
    """average degree >= 8: heavy items are deferred to count_heavy_kernel (one warp per clique item, 128 shares per item
    otherwise), graphs of <= 64 vertices run with one-word sets inside a batch laid out for a 70-vertex graph"""
    import zlib
    rng = np.random.default_rng(zlib.crc32(f'dense/{family}/{scope_name}'.encode()))
    # k <= 5 / graphlets up to 4 vertices: the all-maps oracle on dense graphs
    els = {'cliques5': lambda: FAMILIES['cliques5'](graphlet_patterns), 'cycles5': lambda: FAMILIES['cycles8'](graphlet_patterns)[:3],
           'graphlets4': lambda: graphlet_patterns[3] + graphlet_patterns[4]}[family]()
    graphs = []
    for _ in range(14):
        n = int(rng.integers(12, 23))
        graphs.append((random_graph(rng, n, float(rng.uniform(0.6, 0.95))), n))
    graphs.append((random_graph(rng, 70, 0.25), 70))         # forces W = 2, itself dense enough for heavy items
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    assert ei.shape[1] / node_ptr[-1] >= 8
    scope = 1 if scope_name == 'local' else 0
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), False, scope)
    got = _cuda_ids(node_ptr, ei, _dicts(els, scope_name), False, scope_name)
    assert np.array_equal(got, exp)


def test_single_graph_api_matches_reference_semantics():
    """count_fn(edge_index, subgraph_dict=, induced=, num_nodes=) as called at utils_ids.py:24"""
    from gsn_b200 import counting, patterns
    rng = np.random.default_rng(5)
    ei = random_graph(rng, 12, 0.4)
    # self loops, a duplicated column and a trailing isolated vertex
    ei = np.concatenate([ei, np.array([[3, 3], [3, 3]]), ei[:, :2]], 1)
    els = count_vf2.pattern_edge_lists('cycle_graph', 5)
    for scope_name, fn, ofn in (('global', counting.subgraph_isomorphism_vertex_counts, count_vf2.subgraph_isomorphism_vertex_counts),
                                ('local', counting.subgraph_isomorphism_edge_counts, count_vf2.subgraph_isomorphism_edge_counts)):
        for sd, osd in zip(patterns.make_subgraph_dicts(els, scope_name), count_vf2.make_subgraph_dicts(els, scope_name)):
            got = fn(torch.from_numpy(ei), subgraph_dict=sd, induced=False, num_nodes=14, directed=False)
            exp = ofn(ei, subgraph_dict=osd, induced=False, num_nodes=14)
            assert got.dtype == torch.float64 and got.device.type == 'cpu'
            assert np.array_equal(got.numpy(), exp)


def test_asymmetric_edge_index_raises_keyerror():
    """utils_graph_processing.py:173: edge_dict[mapped_edge] -> KeyError"""
    from gsn_b200 import counting, patterns
    ei = np.array([[0, 1, 2, 1], [1, 2, 0, 0]], np.int64)       # triangle, (2,1) and (0,2) missing
    sd = patterns.make_subgraph_dicts([[(0, 1), (1, 2), (2, 0)]], 'local')[0]
    with pytest.raises(KeyError):
        counting.subgraph_isomorphism_edge_counts(torch.from_numpy(ei), subgraph_dict=sd, induced=False)


def test_large_batch_properties():
    """BASELINE config-5 shape at reduced count: invariants instead of the oracle"""
    from gsn_b200.synthetic import zinc_like_batch
    from gsn_b200 import counting, patterns
    b = zinc_like_batch(20000, seed=3, mean_nodes=30.0, max_nodes=64)
    els = count_vf2.pattern_edge_lists('cycle_graph', 10)
    ei = torch.from_numpy(b['edge_index']).cuda()
    ptr = torch.from_numpy(b['node_ptr'])
    ide = counting.count_batch(ei, ptr, patterns.make_subgraph_dicts(els, 'local'), False, 'local').cpu().numpy()
    idv = counting.count_batch(ei, ptr, patterns.make_subgraph_dicts(els, 'global'), False, 'global').cpu().numpy()
    acc = np.zeros_like(idv)
    np.add.at(acc, b['edge_index'][0], ide)
    assert np.array_equal(2 * idv, acc)              # every cycle through v uses two of v's out-edges
    # both directions of an edge carry the same counts
    E = b['edge_index'].shape[1]
    key = b['edge_index'][0] * (b['node_ptr'][-1] + 1) + b['edge_index'][1]
    rkey = b['edge_index'][1] * (b['node_ptr'][-1] + 1) + b['edge_index'][0]
    order, rorder = np.argsort(key), np.argsort(rkey)
    assert np.array_equal(ide[order], ide[rorder])
    # sample vs the C oracle
    sub = 200
    n_sub, e_sub = b['node_ptr'][sub], b['edge_ptr'][sub]
    exp = count_c.count_batch(b['node_ptr'][:sub + 1], b['edge_ptr'][:sub + 1], b['edge_index'][:, :e_sub],
                              count_vf2.make_subgraph_dicts(els, 'global'), False, 0)
    assert np.array_equal(idv[:n_sub], exp)
