"""MP parity: CUDA layers (through the C ABI) vs the reference's own outputs
(tests/golden/mp_*.pt) and vs the fp32 CPU oracle (oracle/mp_ref.py).
Tolerance: tests/tolerance.py (1e-5 relative to the scale of the compared tensor; BASELINE.json north_star)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from tests.tolerance import FWD, GRAD, close

from oracle import mp_ref
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = dict(atol=1e-5, rtol=1e-5)


def _build_layer(cls_name, ctor):
    import gsn_b200.graph_filters as gf
    with contextlib.redirect_stdout(io.StringIO()):
        return getattr(gf, cls_name)(**ctor)


def _cfg(cls, kw):
    return dict(uses_ids=cls.startswith('GSN'), uses_ef='edge' in cls, msg_kind=kw['msg_kind'],
                id_scope=kw.get('id_scope'), flow=kw.get('flow', 'source_to_target'),
                activation_name=kw['activation_name'], bn=kw['bn'], degree_as_tag=kw['degree_as_tag'],
                retain_features=kw['retain_features'], edge_embedding=kw['edge_embedding'],
                id_embedding=kw['id_embedding'], extend_dims=kw['extend_dims'])


def _to(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


LAYER_GOLDEN = torch.load(os.path.join(GOLDEN, 'mp_layers.pt'))
MODEL_GOLDEN = torch.load(os.path.join(GOLDEN, 'mp_models.pt'))


@pytest.mark.parametrize('name', list(LAYER_GOLDEN))
def test_layer_matches_reference_output(name):
    c = LAYER_GOLDEN[name]
    layer = _build_layer(c['cls'], c['ctor'])
    layer.load_state_dict(c['state_dict'], strict=True)            # same parameter names as the reference
    layer = layer.cuda().eval()
    i = _to(c['inputs'], 'cuda')
    with torch.no_grad():
        out = layer(i['x'], i['edge_index'], identifiers=i['identifiers'], degrees=i['degrees'],
                    edge_features=i.get('edge_features'))
        out2 = layer(i['x'], i['edge_index'], identifiers=i['identifiers'], degrees=i['degrees'],
                     edge_features=i.get('edge_features'))
    assert torch.equal(out, out2)                                  # deterministic
    torch.testing.assert_close(out.cpu(), c['out'], **TOL)


@pytest.mark.parametrize('name', list(LAYER_GOLDEN))
def test_layer_training_mode_and_gradients(name):
    """training-mode BatchNorm (batch statistics over E message rows, models_misc.py:54-55)
    and gradients of the autograd path vs the oracle differentiated on the CPU"""
    c = LAYER_GOLDEN[name]
    kw = c['ctor']
    layer = _build_layer(c['cls'], kw)
    layer.load_state_dict(c['state_dict'], strict=True)
    layer = layer.cuda().train()
    inp = c['inputs']
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ('x', 'identifiers', 'edge_features')
              if inp.get(k) is not None and inp[k].is_floating_point()}
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in c['state_dict'].items()}
    ref = mp_ref.layer_forward(_cfg(c['cls'], kw), sd, leaves['x'], inp['edge_index'], leaves.get('identifiers', inp['identifiers']),
                               inp['degrees'], leaves.get('edge_features'), training=True)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(0))
    (ref * w).sum().backward()

    g = {k: v.detach().clone().cuda().requires_grad_(True) for k, v in leaves.items()}
    out = layer(g['x'], inp['edge_index'].cuda(), identifiers=g.get('identifiers', None if inp['identifiers'] is None else inp['identifiers'].cuda()),
                degrees=inp['degrees'].cuda(), edge_features=g.get('edge_features'))
    close(out, ref)
    (out * w.cuda()).sum().backward()
    for k in leaves:
        if kw.get('degree_as_tag') and not kw.get('retain_features') and k == 'x':
            continue
        close(g[k].grad, leaves[k].grad, GRAD, msg=k)
    for pname, p in layer.named_parameters():
        if sd[pname].grad is not None:
            close(p.grad, sd[pname].grad, GRAD, msg=pname)
    # no-grad training-mode forward (fused kernels with batch statistics) agrees as well
    layer.load_state_dict(c['state_dict'], strict=True)
    with torch.no_grad():
        out_ng = layer(g['x'].detach(), inp['edge_index'].cuda(), identifiers=None if inp['identifiers'] is None else inp['identifiers'].cuda(),
                       degrees=inp['degrees'].cuda(), edge_features=None if inp.get('edge_features') is None else inp['edge_features'].cuda())
    close(out_ng, ref)


@pytest.mark.parametrize('name', list(MODEL_GOLDEN))
def test_model_matches_reference_output(name):
    from gsn_b200.network import GNNSubstructures
    c = MODEL_GOLDEN[name]
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**c['ctor'], **c['args'])
    model.load_state_dict(c['state_dict'], strict=True)
    model = model.cuda().eval()

    class Batch:
        pass
    b = Batch()
    for k, v in c['data'].items():
        setattr(b, k, v.cuda())
    with torch.no_grad():
        out = model(b)
    close(out, c['out'])


@pytest.mark.parametrize('name', ['gsne_general_local', 'gsne_general_global', 'gsn_gin_local_onehot', 'gsn_ogb_local',
                                  'mpnne_general', 'gsne_gin_local_embedding'])
def test_layer_vs_oracle_zinc_sized_batch(name):
    """same layers at BASELINE config-2 size (ZINC-shaped, B=128) against the CPU oracle"""
    from gsn_b200.synthetic import zinc_like_batch
    c = LAYER_GOLDEN[name]
    kw = dict(c['ctor'])
    wide = {'d_in': 128, 'd_msg': 128, 'd_up': 128, 'd_h': [128]}
    if kw['msg_kind'] == 'ogb':
        wide.update(d_in=300, d_ef=300, d_id=300, d_up=300, d_h=[600])
    kw.update({k: v for k, v in wide.items() if k in kw})
    torch.manual_seed(7)
    layer = _build_layer(c['cls'], kw)
    g = torch.Generator().manual_seed(3)
    for m in layer.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    layer.eval()
    b = zinc_like_batch(128, seed=11)
    ei = torch.from_numpy(b['edge_index'])
    N, E = int(b['node_ptr'][-1]), ei.shape[1]
    x = torch.randn((N, kw['d_in']), generator=g)
    ids = None
    if 'd_id' in kw:
        rows = E if kw['id_scope'] == 'local' else N
        ids = torch.randn((rows, kw['d_id']), generator=g) if kw['msg_kind'] == 'ogb' or kw['id_embedding'] == 'embedding' \
            else torch.nn.functional.one_hot(torch.randint(0, kw['d_id'], (rows,), generator=g), kw['d_id']).float()
    ef = torch.randn((E, kw['d_ef']), generator=g) if 'd_ef' in kw else None
    deg = torch.randn((N, kw['d_degree']), generator=g)
    ref = mp_ref.layer_forward(_cfg(c['cls'], kw), layer.state_dict(), x, ei, ids, deg, ef)
    layer = layer.cuda()
    with torch.no_grad():
        out = layer(x.cuda(), ei.cuda(), identifiers=None if ids is None else ids.cuda(), degrees=deg.cuda(),
                    edge_features=None if ef is None else ef.cuda())
    # outputs are O(1); 1e-5 relative to the layer's output scale
    scale = float(ref.abs().max())
    assert float((out.cpu() - ref).abs().max()) <= 1e-5 * max(scale, 1.0) + 1e-5 * 0
    close(out, ref)


def test_edge_plan_rows_are_sorted_and_complete():
    from gsn_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, E = 50, 400
    ei = torch.randint(0, N, (2, E), generator=g)
    plan = ops.EdgePlan(ei.cuda(), N)
    rowptr, eid, nbr = plan.rowptr.cpu(), plan.eid.cpu()[:E], plan.nbr.cpu()[:E]
    assert rowptr[0] == 0 and rowptr[-1] == E
    assert sorted(eid.tolist()) == list(range(E))
    for r in range(N):
        seg = eid[rowptr[r]:rowptr[r + 1]]
        assert (ei[1][seg.long()] == r).all()
        assert (seg[1:] > seg[:-1]).all()
    assert torch.equal(nbr.long(), ei[0][eid.long()])
    # empty graph / isolated nodes
    plan0 = ops.EdgePlan(torch.zeros((2, 0), dtype=torch.int64).cuda(), 5)
    out = ops.segment_sum(plan0, torch.zeros((0, 4)).cuda())
    assert out.shape == (5, 4) and float(out.abs().sum()) == 0.0


def test_readout_pools():
    from gsn_b200.encoders import global_add_pool_sparse, global_mean_pool_sparse
    g = torch.Generator().manual_seed(1)
    batch = torch.sort(torch.randint(0, 9, (100,), generator=g))[0]
    x = torch.randn((100, 24), generator=g)
    ng = int(batch.max()) + 1
    torch.testing.assert_close(global_add_pool_sparse(x.cuda(), batch.cuda()).cpu(), mp_ref.pool(x, batch, 'sum', ng), **TOL)
    torch.testing.assert_close(global_mean_pool_sparse(x.cuda(), batch.cuda()).cpu(), mp_ref.pool(x, batch, 'mean', ng), **TOL)
