"""DGN golden vectors: runs the UNMODIFIED reference directional_gsn/nets/{dgn_layer,aggregators,scalers,layers}.py
on CPU over oracle.dgn_ref.FakeDGLGraph (dgl itself is absent from this image) and stores inputs, the aggregated
mailbox reduction, the layer state_dict and the layer output in tests/golden/dgn.pt.

    python scripts/make_golden_dgn.py
"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_mp import rand_graph, randomize          # noqa: E402
from oracle import dgn_ref                                # noqa: E402

CASES = [
    # name, aggregators, scalers, d_in, d_out, node fields, edge fields, residual, graph_norm
    ('molhiv_recipe', 'mean max min dir0-av dir1-av dir2-av dir3-av', 'identity', 60, 60, 0, 4, True, False),
    ('config_json', 'mean max min dir1-dx dir1-av', 'identity', 70, 70, 3, 0, True, False),
    ('all_kinds', 'mean sum max min std var dir0-av dir1-0.1 dir2-neg-0.1 dir0-dx dir1-dx-no-abs dir2-dx-balanced',
     'identity amplification attenuation', 8, 12, 2, 1, True, True),
    ('odd_width', 'sum dir0-dx dir3-av std', 'amplification', 7, 7, 2, 2, True, False),
    ('two_scalers', 'mean dir1-dx-balanced', 'identity attenuation', 16, 16, 0, 2, False, False),
]


def main():
    ref = dgn_ref.import_reference()
    out = {}
    for ci, (name, aggr, scal, d_in, d_out, Fn, Fe, residual, graph_norm) in enumerate(CASES):
        g = torch.Generator().manual_seed(100 + ci)
        ei, batch, N = rand_graph(g, n_graphs=6, lo=3, hi=11, p=0.35)
        E = ei.shape[1]
        h = torch.randn((N, d_in), generator=g)
        nf = torch.randint(0, 5, (N, Fn), generator=g).float() if Fn else None
        ef = torch.randint(0, 4, (E, Fe), generator=g).float() if Fe else None
        avg_d = {'log': 1.37, 'lin': 2.2, 'exp': 9.0}
        torch.manual_seed(ci)
        layer = ref.DGNLayer(in_dim=d_in, out_dim=d_out, dropout=0.3, graph_norm=graph_norm, batch_norm=True,
                             aggregators=aggr, scalers=scal, avg_d=avg_d, type_net='simple', residual=residual).model
        randomize(layer, g)
        layer.eval()
        snorm = torch.rand((N, 1), generator=g) + 0.5

        def graph():
            fg = dgn_ref.FakeDGLGraph(ei, N)
            if nf is not None:
                fg.ndata['eig'] = nf
            if ef is not None:
                fg.edata['eig'] = ef
            return fg
        with torch.no_grad():
            fg = graph()
            fg.ndata['h'] = h
            fg.apply_edges(layer.pretrans_edges)
            fg.update_all(layer.message_func, layer.reduce_func)
            agg = fg.ndata['h'].clone()
            y = layer(graph(), h, None, snorm)
        # gradient of the mailbox reduction w.r.t. h through the reference's own aggregator code (autograd)
        hr = h.clone().requires_grad_(True)
        fg = graph()
        fg.ndata['h'] = hr
        fg.apply_edges(layer.pretrans_edges)
        fg.update_all(layer.message_func, layer.reduce_func)
        cot = torch.randn(agg.shape, generator=g)
        (fg.ndata['h'] * cot).sum().backward()
        out[name] = dict(cot=cot, h_grad=hr.grad.clone(), aggregators=aggr, scalers=scal, d_in=d_in, d_out=d_out, residual=residual, graph_norm=graph_norm,
                         avg_d=avg_d, edge_index=ei, batch=batch, num_nodes=N, h=h, node_field=nf, edge_field=ef,
                         snorm_n=snorm, agg=agg, state_dict={k: v.clone() for k, v in layer.state_dict().items()}, out=y)
        deg0 = int((torch.bincount(ei[1], minlength=N) == 0).sum())
        print(f'{name}: N={N} E={E} agg {tuple(agg.shape)} out {tuple(y.shape)} zero-in-degree nodes {deg0}')
    path = os.path.join(ROOT, 'tests', 'golden', 'dgn.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
