"""GNNSubstructures: the caller of the MP layers (boundary of the hot path).

Constructor arguments, sub-module names (hence state_dict keys) and forward
semantics follow /root/reference/models_graph_classification.py:17-247 so that a
reference checkpoint loads with load_state_dict(strict=True).  The reference
file itself also runs unchanged on top of gsn_b200.graph_filters (see
INTEGRATION.md); this class exists because /root/reference is not shipped with
the package and bench.py / the tests need a model to drive the layers.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoders import DiscreteEmbedding, global_add_pool_sparse, global_mean_pool_sparse
from .graph_filters import GSN_edge_sparse, GSN_sparse, MPNN_edge_sparse, MPNN_sparse
from .models_misc import choose_activation, mlp

_GSN_NAMES = {'GSN_sparse', 'GSN_edge_sparse'}
_EDGE_NAMES = {'GSN_edge_sparse', 'MPNN_edge_sparse'}


def _default(v, d):
    return d if v is None else v


class GNNSubstructures(nn.Module):

    def __init__(self, in_features, out_features, encoder_ids, d_in_id, in_edge_features=None,
                 d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None, **kwargs):
        super().__init__()
        kw = kwargs
        seed = kw['seed']
        self.model_name = kw['model_name']
        self.readout = _default(kw['readout'], 'sum')
        self.dropout_features = kw['dropout_features']
        self.bn = kw['bn']
        self.final_projection = kw['final_projection']
        self.inject_ids = kw['inject_ids']
        self.inject_edge_features = kw['inject_edge_features']
        self.random_features = kw['random_features']
        d_out, d_h = kw['d_out'], kw['d_h']
        n_layers = len(d_out)
        train_eps = _default(kw['train_eps'], [False] * n_layers)
        act_mlp, bn_mlp, jk_mlp = kw['activation_mlp'], kw['bn_mlp'], kw['jk_mlp']
        enc_kw = {'seed': seed, 'activation_mlp': act_mlp, 'bn_mlp': bn_mlp, 'aggr': kw['multi_embedding_aggr']}

        # encoders (:66-119)
        self.input_node_encoder = DiscreteEmbedding(kw['input_node_encoder'], in_features, d_in_node_encoder,
                                                    kw['d_out_node_encoder'], **enc_kw)
        d_in = self.input_node_encoder.d_out
        if self.random_features:
            self.r_d_out = d_out[0]
            d_in += self.r_d_out
        self.edge_encoder = nn.ModuleList(
            DiscreteEmbedding(kw['edge_encoder'], in_edge_features, d_in_edge_encoder, kw['d_out_edge_encoder'][i], **enc_kw)
            for i in range(n_layers if self.inject_edge_features else 1))
        d_ef = [e.d_out for e in self.edge_encoder]
        self.id_encoder = nn.ModuleList(
            DiscreteEmbedding(kw['id_embedding'], len(d_in_id), d_in_id, kw['d_out_id_embedding'], **enc_kw)
            for _ in range(n_layers if self.inject_ids else 1))
        d_id = [e.d_out for e in self.id_encoder]
        degree_embedding = kw['degree_embedding'] if kw['degree_as_tag'][0] else 'None'
        self.degree_encoder = DiscreteEmbedding(degree_embedding, 1, d_degree, kw['d_out_degree_embedding'], **enc_kw)

        def projection(width, hidden):
            return mlp(width, out_features, hidden, seed, act_mlp, bn_mlp) if jk_mlp else nn.Linear(width, out_features)

        # layer stack (:122-184); only layer 0 carries identifiers unless inject_ids
        conv, lin_proj, norms = [], [], []
        for i in range(n_layers):
            layer_kw = dict(d_in=d_in, d_degree=self.degree_encoder.d_out, degree_as_tag=kw['degree_as_tag'][i],
                            retain_features=kw['retain_features'][i], d_msg=kw['d_msg'][i], d_up=d_out[i], d_h=d_h[i],
                            d_ef=d_ef[i] if self.inject_edge_features else d_ef[0], seed=seed,
                            activation_name=act_mlp, bn=bn_mlp, aggr=_default(kw['aggr'], 'add'),
                            msg_kind=_default(kw['msg_kind'], 'general'), eps=0, train_eps=train_eps[i],
                            flow=_default(kw['flow'], 'target_to_source'), edge_embedding=kw['edge_encoder'],
                            id_embedding=kw['id_embedding'], extend_dims=kw['extend_dims'])
            first_or_injected = lambda inject: i == 0 or inject
            use_ids = first_or_injected(self.inject_ids) and self.model_name in _GSN_NAMES
            use_efs = first_or_injected(self.inject_edge_features) and self.model_name in _EDGE_NAMES
            if use_ids:
                layer_kw.update(d_id=d_id[i] if self.inject_ids else d_id[0], id_scope=kw['id_scope'])
                cls = GSN_edge_sparse if use_efs else GSN_sparse
            else:
                cls = MPNN_edge_sparse if use_efs else MPNN_sparse
            conv.append(cls(**layer_kw))
            lin_proj.append(projection(d_in, d_h[i]) if self.final_projection[i] else None)
            norms.append(nn.BatchNorm1d(d_out[i]) if self.bn[i] else None)
            d_in = d_out[i]
        lin_proj.append(projection(d_in, d_h[-1]) if self.final_projection[-1] else None)
        self.conv, self.lin_proj, self.batch_norms = nn.ModuleList(conv), nn.ModuleList(lin_proj), nn.ModuleList(norms)

        if self.readout == 'sum':
            self.global_pool = global_add_pool_sparse
        elif self.readout == 'mean':
            self.global_pool = global_mean_pool_sparse
        else:
            raise ValueError('Invalid graph pooling type.')
        self.activation = choose_activation(kw['activation'])

    def forward(self, data, print_flag=False, return_intermediate=False):
        layer_in = {'degrees': self.degree_encoder(data.degrees)}
        x = self.input_node_encoder(data.x)
        if self.random_features:
            x = torch.cat((x, torch.rand((x.shape[0], self.r_d_out), device=x.device)), 1)
        x_interm = [x]
        has_ef = hasattr(data, 'edge_features')
        num_graphs = getattr(data, 'num_graphs', None)
        for i, conv in enumerate(self.conv):
            layer_in['identifiers'] = self.id_encoder[i if self.inject_ids else 0](data.identifiers)
            layer_in['edge_features'] = (self.edge_encoder[i if self.inject_edge_features else 0](data.edge_features)
                                         if has_ef else None)
            x = conv(x, data.edge_index, **layer_in)
            if self.bn[i]:
                x = self.batch_norms[i](x)
            x = self.activation(x)
            x_interm.append(x)
        prediction = 0
        for i, proj in enumerate(self.lin_proj):
            if self.final_projection[i]:
                pooled = self.global_pool(x_interm[i], data.batch, num_graphs)
                prediction = prediction + F.dropout(proj(pooled), p=self.dropout_features[i], training=self.training)
        return (prediction, x_interm) if return_intermediate else prediction


class GNN_OGB(nn.Module):
    """OGB variant (virtual node, residual, dropout) of
    /root/reference/models_graph_classification_ogb_original.py:19-268: same constructor, sub-module
    names and forward; layers are gsn_b200.graph_filters.{GSN,MPNN}_edge_sparse_ogb."""

    def __init__(self, in_features, out_features, encoder_ids, d_in_id, in_edge_features=None,
                 d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None, **kwargs):
        super().__init__()
        from .graph_filters import GSN_edge_sparse_ogb, MPNN_edge_sparse_ogb
        kw = kwargs
        seed = kw['seed']
        self.model_name = kw['model_name']
        self.readout = _default(kw['readout'], 'sum')
        self.dropout_features, self.bn, self.final_projection = kw['dropout_features'], kw['bn'], kw['final_projection']
        self.residual, self.inject_ids, self.vn = kw['residual'], kw['inject_ids'], kw['vn']
        d_out, d_h = kw['d_out'], kw['d_h']
        n_layers = len(d_out)
        train_eps = _default(kw['train_eps'], [False] * n_layers)
        act_mlp, bn_mlp = kw['activation_mlp'], kw['bn_mlp']
        enc_kw = {'seed': seed, 'activation_mlp': act_mlp, 'bn_mlp': bn_mlp, 'aggr': kw['multi_embedding_aggr'],
                  'features_scope': kw['features_scope']}
        self.input_node_encoder = DiscreteEmbedding(kw['input_node_encoder'], in_features, d_in_node_encoder,
                                                    kw['d_out_node_encoder'], **enc_kw)
        d_in = self.input_node_encoder.d_out
        if self.vn:
            self.vn_encoder = DiscreteEmbedding(kw['input_vn_encoder'], 1, [1], kw['d_out_vn_encoder'],
                                                **{**enc_kw, 'init': 'zeros'})
            d_in_vn = self.vn_encoder.d_out
        self.edge_encoder = nn.ModuleList(
            DiscreteEmbedding(kw['edge_encoder'], in_edge_features, d_in_edge_encoder, kw['d_out_edge_encoder'][i], **enc_kw)
            for i in range(n_layers))
        self.id_encoder = nn.ModuleList(
            DiscreteEmbedding(kw['id_embedding'], len(d_in_id), d_in_id, kw['d_out_id_embedding'], **enc_kw)
            for _ in range(n_layers if self.inject_ids else 1))
        degree_embedding = kw['degree_embedding'] if kw['degree_as_tag'][0] else 'None'
        self.degree_encoder = DiscreteEmbedding(degree_embedding, 1, d_degree, kw['d_out_degree_embedding'], **enc_kw)

        conv, norms, mlp_vn = [], [], []
        for i in range(n_layers):
            if i > 0 and self.vn:
                mlp_vn.append(mlp(d_in_vn, kw['d_out_vn'][i - 1], d_h[i], seed, act_mlp, bn_mlp))
                d_in_vn = kw['d_out_vn'][i - 1]
            layer_kw = dict(d_in=d_in, d_degree=self.degree_encoder.d_out, degree_as_tag=kw['degree_as_tag'][i],
                            retain_features=kw['retain_features'][i], d_msg=kw['d_msg'][i], d_up=d_out[i], d_h=d_h[i],
                            seed=seed, activation_name=act_mlp, bn=bn_mlp, aggr=_default(kw['aggr'], 'add'),
                            msg_kind=_default(kw['msg_kind'], 'general'), eps=0, train_eps=train_eps[i],
                            flow=_default(kw['flow'], 'target_to_source'), d_ef=self.edge_encoder[i].d_out,
                            edge_embedding=kw['edge_encoder'], id_embedding=kw['id_embedding'],
                            extend_dims=kw['extend_dims'])
            if (i == 0 or self.inject_ids) and self.model_name == 'GSN_edge_sparse_ogb':
                layer_kw.update(d_id=self.id_encoder[i if self.inject_ids else 0].d_out, id_scope=kw['id_scope'])
                conv.append(GSN_edge_sparse_ogb(**layer_kw))
            else:
                conv.append(MPNN_edge_sparse_ogb(**layer_kw))
            norms.append(nn.BatchNorm1d(d_out[i]) if self.bn[i] else None)
            d_in = d_out[i]
        self.conv, self.batch_norms = nn.ModuleList(conv), nn.ModuleList(norms)
        if self.vn:
            self.mlp_vn = nn.ModuleList(mlp_vn)
        pools = {'sum': global_add_pool_sparse, 'mean': global_mean_pool_sparse}
        if self.readout not in pools:
            raise ValueError('Invalid graph pooling type.')
        self.global_pool = pools[self.readout]
        if self.vn:
            if kw['vn_pooling'] not in pools:
                raise ValueError('Invalid graph virtual node pooling type.')
            self.global_vn_pool = pools[kw['vn_pooling']]
        self.lin_proj = nn.Linear(d_out[-1], out_features)
        self.activation = choose_activation(kw['activation'])

    def forward(self, data, return_intermediate=False):
        layer_in = {'degrees': self.degree_encoder(data.degrees)}
        edge_index, batch = data.edge_index, data.batch
        num_graphs = getattr(data, 'num_graphs', None)
        if num_graphs is None:
            num_graphs = int(batch[-1].item()) + 1                         # :218
        if self.vn:
            vn = self.vn_encoder(torch.zeros(num_graphs, dtype=edge_index.dtype, device=edge_index.device))
        x = self.input_node_encoder(data.x)
        x_interm = [x]
        last = len(self.conv) - 1
        has_ef = hasattr(data, 'edge_features')
        for i, conv in enumerate(self.conv):
            layer_in['identifiers'] = self.id_encoder[i if self.inject_ids else 0](data.identifiers)
            layer_in['edge_features'] = self.edge_encoder[i](data.edge_features) if has_ef else None
            if self.vn:
                x_interm[i] = x_interm[i] + vn[batch]
            x = conv(x_interm[i], edge_index, **layer_in)
            if self.bn[i]:
                x = self.batch_norms[i](x)
            if i != last:
                x = self.activation(x)
            x = F.dropout(x, self.dropout_features[i], training=self.training)
            if self.residual:
                x = x + x_interm[-1]
            x_interm.append(x)
            if i < last and self.vn:
                vn_new = self.mlp_vn[i](self.global_vn_pool(x_interm[i], batch, num_graphs) + vn)
                upd = F.dropout(self.activation(vn_new), self.dropout_features[i], training=self.training)
                vn = vn_new + upd if self.residual else upd
        total = 0
        for i in range(len(self.conv) + 1):
            if self.final_projection[i]:
                total = total + x_interm[i]
        return self.lin_proj(self.global_pool(total, batch, num_graphs))
