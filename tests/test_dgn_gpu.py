"""CUDA DGN aggregation (csrc/dgn_kernels.cu through the C ABI) and the DGNLayerSimple / DGNNet mirrors vs the
reference-generated golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from tests.tolerance import FWD, GRAD, close

from oracle import dgn_ref
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DGN_GOLDEN = torch.load(os.path.join(GOLDEN, 'dgn.pt'))


def _cuda(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize('name', sorted(DGN_GOLDEN))
def test_aggregate_and_layer_vs_reference_golden(name):
    from gsn_b200 import directional
    c = DGN_GOLDEN[name]
    g = directional.DirectionalBatch(c['edge_index'].cuda(), c['num_nodes'], ndata_eig=_cuda(c['node_field']),
                                     edata_eig=_cuda(c['edge_field']))
    agg = directional.dgn_aggregate(g.plan, c['h'].cuda(), g.ndata_eig, g.edata_eig, c['aggregators'], c['scalers'], c['avg_d'])
    torch.testing.assert_close(agg.cpu(), c['agg'], atol=1e-5, rtol=1e-5)
    layer = directional.DGNLayer(in_dim=c['d_in'], out_dim=c['d_out'], dropout=0.3, graph_norm=c['graph_norm'],
                                 batch_norm=True, aggregators=c['aggregators'], scalers=c['scalers'], avg_d=c['avg_d'],
                                 type_net='simple', residual=c['residual']).model
    layer.load_state_dict(c['state_dict'])
    layer = layer.cuda().eval()
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        y = layer(g, c['h'].cuda(), None, c['snorm_n'].cuda())
    close(y, c['out'])


def test_count_fields_feed_the_aggregation_zinc_sized():
    """COUNT (edge scope, cycles k<=6 as molhiv_10_runs.sh) -> 'eig' edge field -> aggregation, vs the oracle"""
    from gsn_b200 import directional, patterns
    from gsn_b200.synthetic import zinc_like_batch
    from oracle import count_c, count_vf2
    b = zinc_like_batch(128, seed=3)
    ei = torch.from_numpy(b['edge_index'])
    N = int(b['node_ptr'][-1])
    els = count_vf2.pattern_edge_lists('cycle_graph', 6)
    nd, ed = directional.prepare_subgraph_fields(ei.cuda(), torch.from_numpy(b['node_ptr']),
                                                 patterns.make_subgraph_dicts(els, 'local'), {'induced': False}, 'local')
    assert nd is None and ed.dtype == torch.float32
    exp = count_c.count_batch(b['node_ptr'], b['edge_ptr'], b['edge_index'], count_vf2.make_subgraph_dicts(els, 'local'), False, 1)
    assert np.array_equal(ed.cpu().numpy(), exp.astype(np.float32))
    h = torch.randn((N, 60), generator=torch.Generator().manual_seed(0))
    names = 'mean max min dir0-av dir1-av dir2-av dir3-av'
    g = directional.DirectionalBatch(ei.cuda(), N, edata_eig=ed)
    got = directional.dgn_aggregate(g.plan, h.cuda(), None, ed, names, 'identity', {'log': 1.2})
    ref = dgn_ref.aggregate(ei, N, h, None, ed.cpu(), directional.parse_aggregators(names), [0], 1.2)
    torch.testing.assert_close(got.cpu(), ref, atol=1e-5, rtol=1e-5)
    # vertex scope -> node field -> eig[src] - eig[dst]
    nd, ed2 = directional.prepare_subgraph_fields(ei.cuda(), torch.from_numpy(b['node_ptr']),
                                                  patterns.make_subgraph_dicts(els, 'global'), {'induced': False}, 'global')
    assert ed2 is None and nd.shape == (N, 4)
    names = 'sum std dir3-dx dir2-dx-balanced dir1-0.1'
    got = directional.dgn_aggregate(g.plan, h.cuda(), nd, None, names, 'identity amplification', {'log': 1.2})
    ref = dgn_ref.aggregate(ei, N, h, nd.cpu(), None, directional.parse_aggregators(names), [0, 1], 1.2)
    torch.testing.assert_close(got.cpu(), ref, atol=1e-5, rtol=1e-5)


def test_edge_cases_and_errors():
    from gsn_b200 import directional, ops
    ei = torch.zeros((2, 0), dtype=torch.int64).cuda()
    g = directional.DirectionalBatch(ei, 5)
    out = directional.dgn_aggregate(g.plan, torch.randn(5, 8).cuda(), None, None, 'mean max', 'identity')
    assert out.shape == (5, 16) and bool((out == 0).all())            # no edges: every row keeps zeros
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]]).cuda()
    g = directional.DirectionalBatch(ei, 3, edata_eig=torch.ones(3, 2).cuda())
    with pytest.raises(IndexError):
        directional.dgn_aggregate(g.plan, torch.randn(3, 4).cuda(), None, g.edata_eig, 'dir2-av', 'identity')
    with pytest.raises(RuntimeError):
        directional.dgn_aggregate(g.plan, torch.randn(3, 4), None, None, 'mean', 'identity')       # CPU tensor: no fallback
    # no edges at all: backward is all zeros
    x = torch.randn(5, 8, requires_grad=True, device='cuda')
    g0 = directional.DirectionalBatch(torch.zeros((2, 0), dtype=torch.int64).cuda(), 5)
    directional.dgn_aggregate(g0.plan, x, None, None, 'mean max', 'identity').sum().backward()
    assert bool((x.grad == 0).all())


@pytest.mark.parametrize('name', sorted(DGN_GOLDEN))
def test_aggregate_backward_vs_reference_autograd(name):
    """gsn_dgn_aggregate_bwd + segment-sum vs autograd through the reference's own aggregator code (golden h_grad)"""
    from gsn_b200 import directional
    c = DGN_GOLDEN[name]
    g = directional.DirectionalBatch(c['edge_index'].cuda(), c['num_nodes'], ndata_eig=_cuda(c['node_field']),
                                     edata_eig=_cuda(c['edge_field']))
    h = c['h'].cuda().requires_grad_(True)
    agg = directional.dgn_aggregate(g.plan, h, g.ndata_eig, g.edata_eig, c['aggregators'], c['scalers'], c['avg_d'])
    (agg * c['cot'].cuda()).sum().backward()
    close(h.grad, c['h_grad'], GRAD)


def test_layer_training_step_gradients_vs_oracle():
    """DGNLayerSimple in train mode: loss + parameter / input gradients vs the same modules on the CPU with the
    aggregation done by the oracle restatement (autograd)"""
    import copy
    from gsn_b200 import directional
    c = DGN_GOLDEN['molhiv_recipe']
    torch.manual_seed(3)
    layer = directional.DGNLayer(in_dim=c['d_in'], out_dim=c['d_out'], dropout=0.0, graph_norm=False, batch_norm=True,
                                 aggregators=c['aggregators'], scalers=c['scalers'], avg_d=c['avg_d'], type_net='simple',
                                 residual=True).model.train()
    ref = copy.deepcopy(layer)
    h0 = c['h'].clone().requires_grad_(True)
    a = dgn_ref.aggregate(c['edge_index'], c['num_nodes'], h0, None, c['edge_field'], ref.aggregators, ref.scalers,
                          c['avg_d']['log'])
    y0 = h0 + torch.relu(ref.batchnorm_h(ref.posttrans(a)))
    (y0 ** 2).mean().backward()
    torch.backends.cuda.matmul.allow_tf32 = False
    layer = layer.cuda()
    g = directional.DirectionalBatch(c['edge_index'].cuda(), c['num_nodes'], edata_eig=c['edge_field'].cuda())
    h1 = c['h'].cuda().requires_grad_(True)
    y1 = layer(g, h1, None, None)
    (y1 ** 2).mean().backward()
    close(y1, y0)
    close(h1.grad, h0.grad, GRAD)
    for (k, p1), (_, p0) in zip(layer.named_parameters(), ref.named_parameters()):
        close(p1.grad, p0.grad, GRAD, msg=k)


def test_dgn_net_forward_matches_layerwise_oracle():
    from gsn_b200 import directional
    from gsn_b200.synthetic import zinc_like_batch
    b = zinc_like_batch(16, seed=5)
    ei, node_ptr = torch.from_numpy(b['edge_index']), torch.from_numpy(b['node_ptr'])
    N, E = int(node_ptr[-1]), ei.shape[1]
    gen = torch.Generator().manual_seed(1)
    dims = [119, 4, 12, 12, 10, 6, 6, 2, 2]
    x = torch.stack([torch.randint(0, d, (N,), generator=gen) for d in dims], 1)
    ef = torch.randint(0, 3, (E, 4), generator=gen).float()
    net_params = dict(hidden_dim=32, out_dim=32, in_feat_dropout=0.0, dropout=0.3, L=3, type_net='simple', pos_enc_dim=0,
                      readout='mean', graph_norm=False, batch_norm=True, aggregators='mean max min dir0-av dir1-av',
                      scalers='identity', avg_d={'log': 1.0}, residual=True, edge_feat=False, edge_dim=0,
                      pretrans_layers=1, posttrans_layers=1, device='cuda')
    torch.manual_seed(0)
    net = directional.DGNNet(net_params).eval()
    # oracle forward: same modules on CPU, aggregation by the restatement
    with torch.no_grad():
        h = net.embedding_h(x)
        for conv in net.layers:
            a = dgn_ref.aggregate(ei, N, h, None, ef, conv.aggregators, conv.scalers, 1.0)
            hh = torch.relu(conv.batchnorm_h(conv.posttrans(a)))
            h = h + hh if conv.residual else hh
        sizes = (node_ptr[1:] - node_ptr[:-1]).tolist()
        exp = net.MLP_layer(torch.stack([t.mean(0) for t in h.split(sizes)]))
    net = net.cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = directional.DirectionalBatch(ei.cuda(), N, node_ptr=node_ptr.cuda(), edata_eig=ef.cuda())
    with torch.no_grad():
        got = net(g, x.cuda(), None)
    close(got, exp)
