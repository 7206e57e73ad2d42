"""Edge cases of both hot paths: empty and ragged inputs, degenerate graphs, limits, deeper msg_fn."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import count_c, count_vf2, mp_ref
from tests.util import batch_graphs, random_graph

pytestmark = pytest.mark.gpu


def _count(node_ptr, ei, els, induced, scope_name):
    from gsn_b200 import counting, patterns
    return counting.count_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(node_ptr),
                                patterns.make_subgraph_dicts(els, scope_name), induced, scope_name).cpu().numpy()


def test_count_degenerate_batches():
    els = count_vf2.pattern_edge_lists('cycle_graph', 5)
    # no edges at all
    node_ptr = np.array([0, 3, 3, 7], np.int64)         # includes an empty graph
    ei = np.zeros((2, 0), np.int64)
    assert _count(node_ptr, ei, els, False, 'global').tolist() == np.zeros((7, 3), np.int64).tolist()
    assert _count(node_ptr, ei, els, False, 'local').shape == (0, 3)
    # single-vertex graphs, a lone edge, self loops only
    graphs = [(np.zeros((2, 0), np.int64), 1), (np.array([[0, 1], [1, 0]], np.int64), 2),
              (np.array([[0, 1, 2], [0, 1, 2]], np.int64), 3)]
    rng = np.random.default_rng(0)
    graphs += [(random_graph(rng, 6, 0.7), 6)]
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    for scope_name, scope in (('global', 0), ('local', 1)):
        exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), False, scope)
        assert np.array_equal(_count(node_ptr, ei, els, False, scope_name), exp)


def test_count_largest_supported_graph_and_limit():
    """W = 16 words: up to 1,024 vertices per graph; beyond that the host refuses loudly"""
    from gsn_b200 import counting, patterns
    rng = np.random.default_rng(1)
    n = 1000
    ei = random_graph(rng, n, 3.0 / n)
    els = count_vf2.pattern_edge_lists('cycle_graph', 4)
    node_ptr = np.array([0, n], np.int64)
    edge_ptr = np.array([0, ei.shape[1]], np.int64)
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, 'global'), False, 0)
    assert np.array_equal(_count(node_ptr, ei, els, False, 'global'), exp)
    with pytest.raises(NotImplementedError):
        counting.count_batch(torch.zeros((2, 0), dtype=torch.int64).cuda(), torch.tensor([0, 1025]),
                             patterns.make_subgraph_dicts(els, 'global'), False, 'global')


def test_count_rejects_cross_graph_edges_and_bad_indices():
    from gsn_b200 import counting, patterns
    sds = patterns.make_subgraph_dicts([[(0, 1), (1, 2), (2, 0)]], 'global')
    ei = torch.tensor([[0, 3], [3, 0]]).cuda()           # joins graph 0 and graph 1
    with pytest.raises(ValueError):
        counting.count_batch(ei, torch.tensor([0, 2, 4]), sds, False, 'global')
    with pytest.raises(ValueError):
        counting.count_batch(torch.tensor([[0, 9], [9, 0]]).cuda(), torch.tensor([0, 4]), sds, False, 'global')


def test_layers_with_no_edges_and_isolated_nodes():
    from gsn_b200.graph_filters import GSN_edge_sparse, GSN_sparse
    kw = dict(d_in=8, d_id=4, d_degree=1, degree_as_tag=False, retain_features=True, id_scope='global', d_msg=8, d_up=8,
              d_h=[8], seed=0, activation_name='relu', bn=True, edge_embedding='one_hot_encoder',
              id_embedding='one_hot_encoder', extend_dims=True)
    torch.manual_seed(0)
    for cls, extra in ((GSN_sparse, {}), (GSN_edge_sparse, {'d_ef': 3})):
        for kind in ('gin', 'general'):
            with contextlib.redirect_stdout(io.StringIO()):
                layer = cls(msg_kind=kind, **kw, **extra).eval()
            x, ids = torch.randn(5, 8), torch.randn(5, 4)
            for ei in (torch.zeros((2, 0), dtype=torch.int64), torch.tensor([[0, 1], [1, 0]])):
                ef = torch.randn(ei.shape[1], 3) if extra else None
                cfg = dict(uses_ids=True, uses_ef=bool(extra), msg_kind=kind, id_scope='global', flow='source_to_target',
                           activation_name='relu', bn=True, degree_as_tag=False, retain_features=True,
                           edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder', extend_dims=True)
                ref = mp_ref.layer_forward(cfg, layer.state_dict(), x, ei, ids, torch.zeros(5, 1), ef)
                lc = layer.cuda()
                with torch.no_grad():
                    out = lc(x.cuda(), ei.cuda(), identifiers=ids.cuda(), degrees=torch.zeros(5, 1).cuda(),
                             edge_features=None if ef is None else ef.cuda())
                torch.testing.assert_close(out.cpu(), ref, atol=1e-5, rtol=1e-5)
                layer = lc.cpu()


def test_general_layer_with_three_layer_msg_fn():
    """--num_mlp_layers 3: the extra message layers act per edge (no N-row split possible)"""
    from gsn_b200.graph_filters import GSN_edge_sparse
    kw = dict(d_in=8, d_ef=3, d_id=4, d_degree=1, degree_as_tag=False, retain_features=True, id_scope='local', d_msg=8,
              d_up=8, d_h=[12, 10], seed=0, activation_name='relu', bn=True, msg_kind='general',
              edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder', extend_dims=True)
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        layer = GSN_edge_sparse(**kw).eval()
    rng = np.random.default_rng(2)
    ei = torch.from_numpy(random_graph(rng, 20, 0.2))
    E = ei.shape[1]
    x, ids, ef = torch.randn(20, 8), torch.randn(E, 4), torch.randn(E, 3)
    cfg = dict(uses_ids=True, uses_ef=True, msg_kind='general', id_scope='local', flow='source_to_target',
               activation_name='relu', bn=True, degree_as_tag=False, retain_features=True,
               edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder', extend_dims=True)
    ref = mp_ref.layer_forward(cfg, layer.state_dict(), x, ei, ids, torch.zeros(20, 1), ef)
    layer = layer.cuda()
    with torch.no_grad():
        out = layer(x.cuda(), ei.cuda(), identifiers=ids.cuda(), degrees=torch.zeros(20, 1).cuda(), edge_features=ef.cuda())
    torch.testing.assert_close(out.cpu(), ref, atol=1e-5, rtol=1e-5)


def test_mean_aggregation_is_refused_like_the_reference():
    """aggr='mean' raises NameError in every reference layer; here NotImplementedError"""
    from gsn_b200.graph_filters import MPNN_sparse
    with contextlib.redirect_stdout(io.StringIO()):
        layer = MPNN_sparse(d_in=4, d_degree=1, degree_as_tag=False, retain_features=True, d_msg=4, d_up=4, d_h=[4], seed=0,
                            activation_name='relu', bn=False, aggr='mean', msg_kind='gin').cuda()
    with pytest.raises(NotImplementedError):
        layer(torch.randn(3, 4).cuda(), torch.tensor([[0, 1], [1, 0]]).cuda(), degrees=torch.zeros(3, 1).cuda())
