"""Encoders on the way into the layers (mirror of
/root/reference/utils_graph_learning.py:44-260): DiscreteEmbedding,
multi_embedding, one_hot_encoder, zero_encoder, central_encoder, and the sparse
readouts global_add_pool_sparse / global_mean_pool_sparse (:23-41).

Same class names, constructor arguments and parameter names as the reference
(state_dict compatible).  The readouts use the library's segment-sum instead of
the reference's COO tensor + Python list(range(N)).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .models_misc import mlp

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]     # ogb.utils.features.get_atom_feature_dims()
BOND_FEATURE_DIMS = [5, 6, 2]                            # ogb.utils.features.get_bond_feature_dims()


def get_atom_feature_dims():
    return list(ATOM_FEATURE_DIMS)


def get_bond_feature_dims():
    return list(BOND_FEATURE_DIMS)


class _OgbSumEmbedding(nn.Module):
    """ogb.graphproppred.mol_encoder.AtomEncoder / BondEncoder (ogb >= 1.1.1, not
    vendored by the reference): one nn.Embedding per categorical column,
    xavier-uniform initialised, summed."""

    def __init__(self, dims, emb_dim, list_name):
        super().__init__()
        embs = nn.ModuleList()
        for d in dims:
            e = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(e.weight.data)
            embs.append(e)
        setattr(self, list_name, embs)
        self._list_name = list_name

    def forward(self, x):
        embs = getattr(self, self._list_name)
        from . import ops
        tables = [embs[i].weight for i in range(x.shape[1])]
        if ops.embedding_bag_ok(x, tables):
            return ops.embedding_bag(x, tables)          # one launch (and one for the backward) instead of one per column
        out = 0
        for i in range(x.shape[1]):
            out = out + embs[i](x[:, i])
        return out


class AtomEncoder(_OgbSumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(ATOM_FEATURE_DIMS, emb_dim, 'atom_embedding_list')


class BondEncoder(_OgbSumEmbedding):
    def __init__(self, emb_dim):
        super().__init__(BOND_FEATURE_DIMS, emb_dim, 'bond_embedding_list')


class one_hot_encoder(nn.Module):
    """utils_graph_learning.py:170-187: concatenated one-hot of every column."""

    def __init__(self, d_in):
        super().__init__()
        self.d_in = d_in
        off = torch.tensor([0] + [int(d) for d in d_in[:-1]], dtype=torch.int64).cumsum(0)
        self.register_buffer('_offsets', off.unsqueeze(0), persistent=False)   # not part of the state_dict

    def forward(self, tensor):
        n, c = tensor.shape[0], tensor.shape[1]
        out = torch.zeros((n, int(sum(self.d_in[:c]))), device=tensor.device)
        out.scatter_(1, tensor + self._offsets[:, :c].to(tensor.device), 1.0)
        return out

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.d_in)


class zero_encoder(nn.Module):
    """utils_graph_learning.py:193-208"""

    def __init__(self, d_out):
        super().__init__()
        self.d_out = d_out

    def forward(self, tensor):
        return torch.zeros((tensor.shape[0], self.d_out), device=tensor.device)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.d_out)


class multi_embedding(nn.Module):
    """utils_graph_learning.py:134-167"""

    def __init__(self, d_in, d_out, aggr='concat', init=None):
        super().__init__()
        self.d_in, self.aggr = d_in, aggr
        tables = []
        for d in d_in:
            e = nn.Embedding(d, d_out)
            if init == 'zeros':
                nn.init.constant_(e.weight.data, 0)
            else:
                nn.init.xavier_uniform_(e.weight.data)
            tables.append(e)
        self.encoder = nn.ModuleList(tables)

    def forward(self, tensor):
        if self.aggr == 'sum':
            from . import ops
            tables = [self.encoder[i].weight for i in range(tensor.shape[1])]
            if ops.embedding_bag_ok(tensor, tables):
                return ops.embedding_bag(tensor, tables)
        parts = [self.encoder[i](tensor[:, i]) for i in range(tensor.shape[1])]
        if self.aggr == 'concat':
            return torch.cat(parts, 1)
        if self.aggr == 'sum':
            out = parts[0]
            for p in parts[1:]:
                out = out + p
            return out
        raise NotImplementedError('multi embedding aggregation {} is not currently supported.'.format(self.aggr))


class DiscreteEmbedding(nn.Module):
    """utils_graph_learning.py:44-131"""

    def __init__(self, encoder_name, d_in_features, d_in_encoder, d_out_encoder, **kwargs):
        super().__init__()
        kwargs.setdefault('init', None)
        self.encoder_name = encoder_name
        if encoder_name == 'zero_encoder':
            self.encoder, d_out = zero_encoder(d_out_encoder), d_out_encoder
        elif encoder_name == 'linear':
            self.encoder, d_out = nn.Linear(d_in_features, d_out_encoder, bias=True), d_out_encoder
        elif encoder_name == 'mlp':
            self.encoder = mlp(d_in_features, d_out_encoder, d_out_encoder, kwargs['seed'],
                               kwargs['activation_mlp'], kwargs['bn_mlp'])
            d_out = d_out_encoder
        elif encoder_name == 'one_hot_encoder':
            self.encoder, d_out = one_hot_encoder(d_in_encoder), sum(d_in_encoder)
        elif encoder_name == 'embedding':
            self.encoder = multi_embedding(d_in_encoder, d_out_encoder, kwargs['aggr'], kwargs['init'])
            d_out = len(d_in_encoder) * d_out_encoder if kwargs['aggr'] == 'concat' else d_out_encoder
        elif encoder_name in ('atom_one_hot_encoder', 'bond_one_hot_encoder'):
            dims = get_atom_feature_dims() if encoder_name.startswith('atom') else get_bond_feature_dims()
            dims = dims if kwargs['features_scope'] == 'full' else dims[:2]
            self.encoder, d_out = one_hot_encoder(dims), sum(dims)
        elif encoder_name == 'atom_encoder':
            self.encoder, d_out = AtomEncoder(d_out_encoder), d_out_encoder
        elif encoder_name == 'bond_encoder':
            self.encoder, d_out = BondEncoder(emb_dim=d_out_encoder), d_out_encoder
        elif encoder_name == 'None':
            self.encoder, d_out = None, d_in_features
        else:
            raise NotImplementedError('Encoder {} is not currently supported.'.format(encoder_name))
        self.d_out = d_out

    def forward(self, x):
        x = x.unsqueeze(-1) if x.dim() == 1 else x
        if self.encoder is None:
            return x.float()
        x = x.float() if self.encoder_name in ('linear', 'mlp') else x.long()
        return self.encoder(x)


class central_encoder(nn.Module):
    """utils_graph_learning.py:211-260: the dummy "self loop" category of the
    central node for per-edge quantities (edge features, GSN-e identifiers).

    `forward` is the reference's dense form.  `segments()` describes the same
    thing to the fused gin kernel without building the [E, d+1] tensor."""

    def __init__(self, nb_encoder, d_ef, extend=True):
        super().__init__()
        self.extend, self.nb_encoder = extend, nb_encoder
        self.one_hot = 'one_hot_encoder' in nb_encoder
        self.d_in = d_ef
        if self.extend:
            print('##### EXTENDING EDGE FEATURE DIMENSIONS #####')
        if self.one_hot:
            self.d_out = d_ef + 1 if extend else d_ef
            if extend:
                self.encoder = DiscreteEmbedding('one_hot_encoder', 1, [d_ef + 1], None)
        else:
            self.d_out = d_ef
            if extend:
                self.encoder = DiscreteEmbedding('embedding', None, [1], d_ef, aggr='sum')

    def forward(self, x_nb, num_nodes):
        dev = x_nb.device
        if self.one_hot and self.extend:
            x_nb = torch.cat((torch.zeros((x_nb.shape[0], 1), device=dev), x_nb), -1)
        if self.extend:
            x_central = self.encoder(torch.zeros((num_nodes, 1), device=dev).long())
        else:
            x_central = torch.zeros((num_nodes, self.d_out), device=dev)
        return x_central, x_nb

    def central_row(self):
        """[d] learned embedding of the self-loop category (embedding + extend), else None"""
        if self.extend and not self.one_hot:
            return self.encoder.encoder.encoder[0].weight[0]
        return None


_pool_plans = []


def _segment_rows(x, batch, num_graphs=None):
    """sum of the rows of x per graph; the grouping of `batch` is built once per
    batch tensor and reused by every readout of the forward pass"""
    from . import ops
    key = (batch.data_ptr(), batch._version, int(batch.shape[0]), batch.device.index)
    plan = None
    for k, p in _pool_plans:
        if k == key:
            plan = p
    if plan is None:
        if num_graphs is None:
            num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0   # torch.max(batch)+1, :26
        n = x.shape[0]
        ar = torch.arange(n, device=x.device, dtype=torch.int64)
        plan = ops.EdgePlan(torch.stack([ar, batch.to(torch.int64)], 0), num_graphs)
        plan._batch_ref = batch
        _pool_plans.append((key, plan))
        if len(_pool_plans) > 8:
            _pool_plans.pop(0)
    return ops.segment_sum_ad(plan, x.float()), plan


def global_add_pool_sparse(x, batch, num_graphs=None):
    """utils_graph_learning.py:23-29"""
    return _segment_rows(x, batch, num_graphs)[0]


def global_mean_pool_sparse(x, batch, num_graphs=None):
    """utils_graph_learning.py:32-41"""
    s, plan = _segment_rows(x, batch, num_graphs)
    sizes = plan.degree().clone()
    sizes[sizes == 0.0] = 1.0
    return s / sizes.unsqueeze(1)
