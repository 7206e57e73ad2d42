"""Pins the MP oracle (oracle/mp_ref.py, test infrastructure) to the outputs of the
reference's own modules (tests/golden/mp_*.pt, made by scripts/make_golden_mp.py), and,
when /root/reference is present (build container), re-runs the reference live."""
import contextlib
import io
import os
import warnings

import pytest
import torch

from oracle import mp_ref, ref_import
from tests.conftest import GOLDEN

LAYERS = torch.load(os.path.join(GOLDEN, 'mp_layers.pt'))
MODELS = torch.load(os.path.join(GOLDEN, 'mp_models.pt'))


def layer_cfg(cls, kw):
    return dict(uses_ids=cls.startswith('GSN'), uses_ef='edge' in cls, msg_kind=kw['msg_kind'],
                id_scope=kw.get('id_scope'), flow=kw.get('flow', 'source_to_target'),
                activation_name=kw['activation_name'], bn=kw['bn'], degree_as_tag=kw['degree_as_tag'],
                retain_features=kw['retain_features'], edge_embedding=kw['edge_embedding'],
                id_embedding=kw['id_embedding'], extend_dims=kw['extend_dims'])


def model_layer_cfgs(args):
    cfgs = []
    for i in range(len(args['d_out'])):
        gsn = args['model_name'] in ('GSN_sparse', 'GSN_edge_sparse')
        edge = args['model_name'] in ('GSN_edge_sparse', 'MPNN_edge_sparse')
        cfgs.append(dict(uses_ids=gsn and (i == 0 or args['inject_ids']),
                         uses_ef=edge and (i == 0 or args['inject_edge_features']), msg_kind=args['msg_kind'],
                         id_scope=args['id_scope'], flow=args['flow'], activation_name=args['activation_mlp'],
                         bn=args['bn_mlp'], degree_as_tag=args['degree_as_tag'][i],
                         retain_features=args['retain_features'][i], edge_embedding=args['edge_encoder'],
                         id_embedding=args['id_embedding'], extend_dims=args['extend_dims']))
    return cfgs


@pytest.mark.parametrize('name', list(LAYERS))
def test_layer_oracle_vs_reference_golden(name):
    c = LAYERS[name]
    i = c['inputs']
    out = mp_ref.layer_forward(layer_cfg(c['cls'], c['ctor']), c['state_dict'], i['x'], i['edge_index'],
                               i['identifiers'], i['degrees'], i.get('edge_features'))
    torch.testing.assert_close(out, c['out'], atol=1e-6, rtol=1e-6)


@pytest.mark.parametrize('name', list(MODELS))
def test_model_oracle_vs_reference_golden(name):
    c = MODELS[name]
    args = dict(c['args'])
    args.update(d_in_id=c['ctor']['d_in_id'], d_in_node_encoder=c['ctor']['d_in_node_encoder'],
                d_in_edge_encoder=c['ctor']['d_in_edge_encoder'], d_degree=c['ctor']['d_degree'])
    out = mp_ref.gnn_substructures_forward(args, c['state_dict'], c['data'], model_layer_cfgs(c['args']))
    torch.testing.assert_close(out, c['out'], atol=2e-6, rtol=2e-6)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present (GPU box)')
def test_golden_is_reproducible_from_the_reference():
    """the committed vectors really are what the unmodified reference computes"""
    L = ref_import.layers()
    for name in ('gsne_general_local', 'gsn_gin_local_onehot', 'gsn_ogb_global'):
        c = LAYERS[name]
        with contextlib.redirect_stdout(io.StringIO()):
            layer = L[c['cls']](**c['ctor'])
        layer.load_state_dict(c['state_dict'])
        layer.eval()
        i = c['inputs']
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            out = layer(i['x'], i['edge_index'], identifiers=i['identifiers'], degrees=i['degrees'],
                        edge_features=i.get('edge_features'))
        assert torch.equal(out, c['out'])


def test_state_dict_keys_match_reference():
    """drop-in layers must load reference checkpoints: identical parameter names"""
    import gsn_b200.graph_filters as gf
    for name, c in LAYERS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            mine = getattr(gf, c['cls'])(**c['ctor'])
        assert list(mine.state_dict().keys()) == list(c['state_dict'].keys()), name
        for k, v in mine.state_dict().items():
            assert v.shape == c['state_dict'][k].shape, (name, k)
    from gsn_b200.network import GNNSubstructures
    for name, c in MODELS.items():
        with contextlib.redirect_stdout(io.StringIO()):
            m = GNNSubstructures(**c['ctor'], **c['args'])
        assert set(m.state_dict().keys()) == set(c['state_dict'].keys()), name
