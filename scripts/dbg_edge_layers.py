import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import mp_ref
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
            print(cls.__name__, kind, 'E', ei.shape[1], 'maxdiff', float((out.cpu() - ref).abs().max()), flush=True)
            layer = lc.cpu()
