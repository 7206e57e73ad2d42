"""DGN consumer oracle (oracle/dgn_ref.py) pinned to outputs of the unmodified reference layer
(tests/golden/dgn.pt, scripts/make_golden_dgn.py) and, where /root/reference is present, to the
reference's aggregator functions one by one."""
import os

import pytest
import torch

from gsn_b200 import directional
from oracle import dgn_ref
from tests.conftest import GOLDEN

DGN_GOLDEN = torch.load(os.path.join(GOLDEN, 'dgn.pt'))


@pytest.mark.parametrize('name', sorted(DGN_GOLDEN))
def test_restatement_matches_reference_reduce(name):
    c = DGN_GOLDEN[name]
    aggr = directional.parse_aggregators(c['aggregators'])
    sc = [directional.SCALER[s] for s in c['scalers'].split()]
    got = dgn_ref.aggregate(c['edge_index'], c['num_nodes'], c['h'], c['node_field'], c['edge_field'], aggr, sc,
                            c['avg_d']['log'])
    assert got.shape == c['agg'].shape
    torch.testing.assert_close(got, c['agg'], atol=1e-6, rtol=1e-6)
    # rows of nodes that receive no message stay zero (DGL update_all semantics)
    deg = torch.bincount(c['edge_index'][1], minlength=c['num_nodes'])
    assert (deg == 0).any() and bool((c['agg'][deg == 0] == 0).all())


def test_parse_aggregators_names():
    p = directional.parse_aggregators('mean sum max min std var dir0-av dir6-av dir1-0.1 dir3-neg-0.1 dir2-dx '
                                      'dir1-dx-no-abs dir3-dx-balanced')
    assert [k for k, _, _ in p] == [0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 9, 10]
    assert [i for _, i, _ in p][6:] == [0, 6, 1, 3, 2, 1, 3]
    assert p[8][2] == pytest.approx(0.1) and p[9][2] == pytest.approx(-0.1)
    with pytest.raises(KeyError):
        directional.parse_aggregators('median')


@pytest.mark.skipif(not dgn_ref.available(), reason='/root/reference not present')
def test_every_reference_aggregator_name_parses_and_matches():
    ref = dgn_ref.import_reference()
    g = torch.Generator().manual_seed(3)
    h, vf, h_in = torch.randn((5, 4, 6), generator=g), torch.randn((5, 4, 7), generator=g), torch.randn((5, 6), generator=g)
    for name, fn in ref.AGGREGATORS.items():
        (kind, idx, alpha), = directional.parse_aggregators(name)
        got = dgn_ref._aggr(dgn_ref._KIND_NAMES[kind], idx, alpha, h, vf, h_in)
        torch.testing.assert_close(got, fn(h, vf, h_in), atol=1e-6, rtol=1e-6, msg=name)
    assert set(ref.SCALERS) == set(directional.SCALER)


def test_state_dict_keys_match_reference_layer():
    c = DGN_GOLDEN['molhiv_recipe']
    layer = directional.DGNLayer(in_dim=c['d_in'], out_dim=c['d_out'], dropout=0.3, graph_norm=False, batch_norm=True,
                                 aggregators=c['aggregators'], scalers=c['scalers'], avg_d=c['avg_d'], type_net='simple',
                                 residual=True).model
    assert set(layer.state_dict()) == set(c['state_dict'])
    layer.load_state_dict(c['state_dict'])
    with pytest.raises(NotImplementedError):
        directional.DGNLayer(in_dim=4, out_dim=4, dropout=0., graph_norm=False, batch_norm=True, aggregators='mean',
                             scalers='identity', avg_d=None, type_net='towers', residual=True)
