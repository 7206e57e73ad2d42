"""BASELINE config 3 (ogbg-molhiv recipe, README.md:121): GNN_OGB with GSN_edge_sparse_ogb layers --
eval forward and one training step (loss + gradients of every parameter) against the reference's own
model (tests/golden/mp_ogb.pt, scripts/make_golden_mp.py ogb)."""
import contextlib
import io
import os

import pytest
import torch

from tests.tolerance import FWD, GRAD, close

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
OGB = torch.load(os.path.join(GOLDEN, 'mp_ogb.pt'))


def _model(c):
    from gsn_b200.network import GNN_OGB
    with contextlib.redirect_stdout(io.StringIO()):
        m = GNN_OGB(**c['ctor'], **c['args'])
    m.load_state_dict(c['state_dict'], strict=True)
    return m.cuda()


def _batch(c):
    class B:
        pass
    b = B()
    for k, v in c['data'].items():
        setattr(b, k, v.cuda())
    return b


@pytest.mark.parametrize('name', list(OGB))
def test_gnn_ogb_eval_forward(name):
    c = OGB[name]
    m = _model(c).eval()
    with torch.no_grad():
        y = m(_batch(c))
    close(y, c['y_eval'])


@pytest.mark.parametrize('name', [n for n in OGB if 'grads' in OGB[n]])
def test_gnn_ogb_train_step_gradients(name):
    c = OGB[name]
    m = _model(c).train()
    y = m(_batch(c))
    close(y, c['y_train'])
    loss = torch.nn.functional.binary_cross_entropy_with_logits(y, c['target'].cuda())
    close(loss, c['loss'])
    loss.backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(c['grads'])
    for k, gref in c['grads'].items():
        close(got[k], gref, GRAD, msg=k)
    # one optimiser step runs end to end
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    opt.step()
