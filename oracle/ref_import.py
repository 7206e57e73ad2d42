"""Imports the UNMODIFIED reference MP code from /root/reference (build container
only; TEST INFRASTRUCTURE).  Installs the stubs of SURVEY.md A.1 for the
third-party modules that are absent from this image and only touched at import
time (torch_geometric.utils.degree / is_undirected, ogb encoders)."""
from __future__ import annotations

import os
import sys
import types

REF = '/root/reference'


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'graph_filters'))


def install():
    if not available():
        raise RuntimeError('/root/reference is not present (GPU box?)')
    import torch
    import torch.nn as nn
    if 'torch_geometric' not in sys.modules or not hasattr(sys.modules['torch_geometric'], '__gsn_stub__'):
        tg = types.ModuleType('torch_geometric')
        tg.__gsn_stub__ = True
        tgu = types.ModuleType('torch_geometric.utils')

        def degree(index, num_nodes=None, dtype=None):
            n = int(index.max()) + 1 if num_nodes is None else num_nodes
            out = torch.zeros(n, dtype=dtype or torch.float32, device=index.device)
            return out.scatter_add_(0, index, torch.ones_like(index, dtype=out.dtype))
        tgu.degree = degree
        tgu.is_undirected = lambda *a, **k: True
        tg.utils = tgu
        sys.modules['torch_geometric'] = tg
        sys.modules['torch_geometric.utils'] = tgu
    if 'ogb' not in sys.modules:
        atom_dims, bond_dims = [119, 4, 12, 12, 10, 6, 6, 2, 2], [5, 6, 2]

        class _Enc(nn.Module):   # ogb.graphproppred.mol_encoder (ogb>=1.1.1): summed xavier-uniform embeddings
            def __init__(self, emb_dim, dims, name):
                super().__init__()
                lst = nn.ModuleList()
                for d in dims:
                    e = nn.Embedding(d, emb_dim)
                    nn.init.xavier_uniform_(e.weight.data)
                    lst.append(e)
                setattr(self, name, lst)
                self._n = name

            def forward(self, x):
                out = 0
                for i in range(x.shape[1]):
                    out = out + getattr(self, self._n)[i](x[:, i])
                return out

        class AtomEncoder(_Enc):
            def __init__(self, emb_dim):
                super().__init__(emb_dim, atom_dims, 'atom_embedding_list')

        class BondEncoder(_Enc):
            def __init__(self, emb_dim):
                super().__init__(emb_dim, bond_dims, 'bond_embedding_list')
        ogb = types.ModuleType('ogb')
        gpp = types.ModuleType('ogb.graphproppred')
        me = types.ModuleType('ogb.graphproppred.mol_encoder')
        me.AtomEncoder, me.BondEncoder = AtomEncoder, BondEncoder
        ou = types.ModuleType('ogb.utils')
        of = types.ModuleType('ogb.utils.features')
        of.get_atom_feature_dims = lambda: list(atom_dims)
        of.get_bond_feature_dims = lambda: list(bond_dims)
        ogb.graphproppred, gpp.mol_encoder, ogb.utils, ou.features = gpp, me, ou, of
        sys.modules.update({'ogb': ogb, 'ogb.graphproppred': gpp, 'ogb.graphproppred.mol_encoder': me,
                            'ogb.utils': ou, 'ogb.utils.features': of})
    if REF not in sys.path:
        sys.path.insert(0, REF)


def layers():
    install()
    import importlib
    out = {}
    for name in ('GSN_sparse', 'GSN_edge_sparse', 'GSN_edge_sparse_ogb', 'MPNN_sparse', 'MPNN_edge_sparse',
                 'MPNN_edge_sparse_ogb'):
        out[name] = getattr(importlib.import_module(f'graph_filters.{name}'), name)
    return out


def models():
    install()
    import importlib
    return {'GNNSubstructures': importlib.import_module('models_graph_classification').GNNSubstructures,
            'GNN_OGB': importlib.import_module('models_graph_classification_ogb_original').GNN_OGB}
