"""DGN consumer of COUNT: mirrors of the reference's
directional_gsn/ pieces that sit either side of the counting kernel.

  * prepare_subgraph_fields  <- utils_subgraph_encoding.py:284-303 (_prepare) + data/HIV.py:91-98
    (get_subgraphs): counts of the whole batch in one COUNT launch, returned as the float
    'eig' field DGN reads (node field for id_scope global, edge field for local).
  * dgn_aggregate            <- nets/dgn_layer.py:28-54 + nets/aggregators.py + nets/scalers.py:
    one CUDA kernel (csrc/dgn_kernels.cu) instead of DGL degree-bucketed mailboxes.
  * DGNLayerSimple / DGNNet  <- nets/dgn_layer.py:11-80, nets/HIV_graph_classification/dgn_net.py:8-83
    with the same constructor arguments and state_dict keys; DGL graphs are replaced by
    DirectionalBatch (edge_index + fields + node_ptr).  type_net 'complex' / 'towers' name classes
    the reference never defines (dgn_layer.py:98-108 would raise NameError) and are not built.

The aggregation is differentiable w.r.t. h (gsn_dgn_aggregate_bwd + the deterministic segment-sum), so
DGNLayerSimple trains; the 'eig' fields are data (counts) and get no gradient.
"""
from __future__ import annotations

import ctypes
import re
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .encoders import AtomEncoder, BondEncoder

KIND = {'mean': 0, 'sum': 1, 'max': 2, 'min': 3, 'std': 4, 'var': 5, 'dir-av': 6, 'dir-softmax': 7, 'dir-dx': 8,
        'dir-dx-no-abs': 9, 'dir-dx-balanced': 10}
SCALER = {'identity': 0, 'amplification': 1, 'attenuation': 2}
_DIR = re.compile(r'^dir(\d+)-(av|dx|dx-no-abs|dx-balanced|neg-[0-9.]+|[0-9.]+)$')


class GsnDgnAggr(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('field', ctypes.c_int32), ('alpha', ctypes.c_float), ('_pad', ctypes.c_int32)]


def parse_aggregators(names: str):
    """'mean max min dir0-av dir1-dx dir2-0.1 dir3-neg-0.1 ...' -> [(kind, field, alpha)]; the names of
    AGGREGATORS (aggregators.py:74-99) plus any other eig index / softmax temperature of the same families."""
    out = []
    for n in names.split():
        if n in KIND and not n.startswith('dir'):
            out.append((KIND[n], 0, 0.0))
            continue
        m = _DIR.match(n)
        if not m:
            raise KeyError(n)                       # reference: KeyError from the AGGREGATORS dict
        idx, tail = int(m.group(1)), m.group(2)
        if tail == 'av':
            out.append((KIND['dir-av'], idx, 0.0))
        elif tail in ('dx', 'dx-no-abs', 'dx-balanced'):
            out.append((KIND['dir-' + tail], idx, 0.0))
        elif tail.startswith('neg-'):
            out.append((KIND['dir-softmax'], idx, -float(tail[4:])))
        else:
            out.append((KIND['dir-softmax'], idx, float(tail)))
    return out


def _aggr_array(aggr, Fn, Fe):
    arr = (GsnDgnAggr * len(aggr))()
    for i, (k, f, al) in enumerate(aggr):
        arr[i].kind, arr[i].field, arr[i].alpha = k, f, al
        if k >= KIND['dir-av'] and f >= Fn + Fe:
            raise IndexError(f'eig_idx {f} out of range for a vector field with {Fn + Fe} components')
    return arr


class _DgnAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, plan, node_field, edge_field, aggr, scalers, avg_log):
        _lib.require_cuda(h, 'h')
        h = h.detach().float().contiguous()
        N, d = h.shape
        Fn = 0 if node_field is None else node_field.shape[1]
        Fe = 0 if edge_field is None else edge_field.shape[1]
        nf = None if node_field is None else node_field.detach().float().contiguous()
        ef = None if edge_field is None else edge_field.detach().float().contiguous()
        arr = _aggr_array(aggr, Fn, Fe)
        sc = (ctypes.c_int32 * len(scalers))(*scalers)
        out = torch.empty((N, len(aggr) * len(scalers) * d), dtype=torch.float32, device=h.device)
        # per-edge weights of the directional aggregators (four at a time): workspace of the call, caller-allocated
        n_dir = sum(1 for k, _, _ in aggr if k >= KIND['dir-av'])
        scratch = torch.empty(4 * max(plan.E, 1), dtype=torch.float32, device=h.device) if n_dir > 0 else None
        with torch.cuda.device(h.device):
            _lib.call('dgn_aggregate', 'gsn_dgn_aggregate_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr),
                      N, plan.E, _lib.ptr(h), d, _lib.ptr(nf), Fn, _lib.ptr(ef), Fe, ctypes.cast(arr, ctypes.c_void_p),
                      len(aggr), ctypes.cast(sc, ctypes.c_void_p), len(scalers), ctypes.c_float(avg_log), _lib.ptr(out),
                      _lib.ptr(scratch), _lib.stream_ptr())
        ctx.save_for_backward(h)
        ctx.rest = (plan, nf, ef, list(aggr), list(scalers), avg_log)
        return out

    @staticmethod
    def backward(ctx, g):
        (h,) = ctx.saved_tensors
        plan, nf, ef, aggr, scalers, avg_log = ctx.rest
        N, d = h.shape
        Fn = 0 if nf is None else nf.shape[1]
        Fe = 0 if ef is None else ef.shape[1]
        g = g.float().contiguous()
        arr = _aggr_array(aggr, Fn, Fe)
        sc = (ctypes.c_int32 * len(scalers))(*scalers)
        M = torch.empty((max(plan.E, 1), d), dtype=torch.float32, device=h.device)
        SG = torch.empty((N, d), dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _lib.call('dgn_aggregate_bwd', 'gsn_dgn_aggregate_bwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid),
                      _lib.ptr(plan.nbr), N, plan.E, _lib.ptr(h), d, _lib.ptr(nf), Fn, _lib.ptr(ef), Fe,
                      ctypes.cast(arr, ctypes.c_void_p), len(aggr), ctypes.cast(sc, ctypes.c_void_p), len(scalers),
                      ctypes.c_float(avg_log), _lib.ptr(g), _lib.ptr(M), _lib.ptr(SG), _lib.stream_ptr())
        if plan.E == 0:
            return SG, None, None, None, None, None, None
        # messages were gathered at edge_index[0]: their gradients are summed per source row, in edge-id order
        src_plan = ops.edge_plan(plan.edge_index, N, 'target_to_source' if plan.select == 1 else 'source_to_target')
        return SG + ops.segment_sum(src_plan, M[:plan.E]), None, None, None, None, None, None


def dgn_aggregate(plan, h, node_field=None, edge_field=None, aggregators='mean', scalers='identity', avg_d=None):
    """[N, d] -> [N, A*S*d]: every aggregator (dim-1 concat, dgn_layer.py:49) times every scaler (:50-51)."""
    aggr = parse_aggregators(aggregators) if isinstance(aggregators, str) else list(aggregators)
    sc = [SCALER[s] for s in scalers.split()] if isinstance(scalers, str) else list(scalers)
    avg_log = float(avg_d['log']) if (avg_d is not None and 'log' in avg_d) else 1.0
    return _DgnAggregate.apply(h, plan, node_field, edge_field, aggr, sc, avg_log)


class DirectionalBatch:
    """What the reference keeps in a batched DGLGraph: edges (src -> dst), the 'eig' fields and the graph
    boundaries.  edge_index int64 [2, E] = torch.stack(g.edges()) (utils_subgraph_encoding.py:286)."""

    def __init__(self, edge_index, num_nodes, node_ptr=None, ndata_eig=None, edata_eig=None):
        self.edge_index, self.num_nodes, self.node_ptr = edge_index, int(num_nodes), node_ptr
        self.ndata_eig, self.edata_eig = ndata_eig, edata_eig
        self._plan = None

    @property
    def plan(self):
        if self._plan is None:
            self._plan = ops.EdgePlan(self.edge_index, self.num_nodes)      # grouped by dst = edge_index[1]
        return self._plan


def prepare_subgraph_fields(edge_index, node_ptr, subgraph_dicts, subgraph_params, id_scope):
    """_prepare (utils_subgraph_encoding.py:284-303) for a whole batch + get_subgraphs (data/HIV.py:91-98):
    returns (ndata_eig, edata_eig), one of them None; int64 counts -> float as data/HIV.py:83-86."""
    from . import counting
    ids = counting.count_batch(edge_index, node_ptr, subgraph_dicts, subgraph_params['induced'], id_scope)
    ids = ids.float()
    return (ids, None) if id_scope == 'global' else (None, ids)


# ---------------------------------------------------------------------------- modules (same state_dict keys)
def _activation(name):
    table = {'relu': nn.ReLU, 'sigmoid': nn.Sigmoid, 'tanh': nn.Tanh, 'elu': nn.ELU, 'selu': nn.SELU, 'glu': nn.GLU,
             'leakyrelu': nn.LeakyReLU, 'softplus': nn.Softplus}
    if name is None or str(name).lower() == 'none':
        return None
    return table[str(name).lower()]()


class FCLayer(nn.Module):
    """nets/layers.py:23-121: Linear -> activation -> dropout -> BatchNorm."""

    def __init__(self, in_size, out_size, activation='relu', dropout=0., b_norm=False, bias=True):
        super().__init__()
        self.in_size, self.out_size, self.bias = in_size, out_size, bias
        self.linear = nn.Linear(in_size, out_size, bias=bias)
        self.dropout = nn.Dropout(p=dropout) if dropout else None
        self.b_norm = nn.BatchNorm1d(out_size) if b_norm else None
        self.activation = _activation(activation)
        nn.init.xavier_uniform_(self.linear.weight, 1 / in_size)
        if bias:
            self.linear.bias.data.zero_()

    def forward(self, x):
        h = self.linear(x)
        if self.activation is not None:
            h = self.activation(h)
        if self.dropout is not None:
            h = self.dropout(h)
        if self.b_norm is not None:
            h = self.b_norm(h)
        return h


class MLP(nn.Module):
    """nets/layers.py:124-154."""

    def __init__(self, in_size, hidden_size, out_size, layers, mid_activation='relu', last_activation='none',
                 dropout=0., mid_b_norm=False, last_b_norm=False):
        super().__init__()
        self.fully_connected = nn.ModuleList()
        if layers <= 1:
            self.fully_connected.append(FCLayer(in_size, out_size, last_activation, dropout, last_b_norm))
        else:
            self.fully_connected.append(FCLayer(in_size, hidden_size, mid_activation, dropout, mid_b_norm))
            for _ in range(layers - 2):
                self.fully_connected.append(FCLayer(hidden_size, hidden_size, mid_activation, dropout, mid_b_norm))
            self.fully_connected.append(FCLayer(hidden_size, out_size, last_activation, dropout, last_b_norm))

    def forward(self, x):
        for fc in self.fully_connected:
            x = fc(x)
        return x


class MLPReadout(nn.Module):
    """nets/mlp_readout_layer.py:12-30."""

    def __init__(self, input_dim, output_dim, L=2):
        super().__init__()
        layers = [nn.Linear(input_dim // 2 ** l, input_dim // 2 ** (l + 1), bias=True) for l in range(L)]
        layers.append(nn.Linear(input_dim // 2 ** L, output_dim, bias=True))
        self.FC_layers = nn.ModuleList(layers)
        self.L = L

    def forward(self, x):
        y = x
        for l in range(self.L):
            y = F.relu(self.FC_layers[l](y))
        return self.FC_layers[self.L](y)


class DGNLayerSimple(nn.Module):
    """nets/dgn_layer.py:11-80; `aggregators` / `scalers` are the space-separated names of the reference CLI."""

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, residual, avg_d,
                 posttrans_layers=1):
        super().__init__()
        self.dropout, self.graph_norm, self.batch_norm, self.residual = dropout, graph_norm, batch_norm, residual
        self.aggregators = parse_aggregators(aggregators) if isinstance(aggregators, str) else aggregators
        self.scalers = [SCALER[s] for s in scalers.split()] if isinstance(scalers, str) else scalers
        self.batchnorm_h = nn.BatchNorm1d(out_dim)
        self.posttrans = MLP(in_size=len(self.aggregators) * len(self.scalers) * in_dim, hidden_size=out_dim,
                             out_size=out_dim, layers=posttrans_layers, mid_activation='relu', last_activation='none')
        self.avg_d = avg_d
        if in_dim != out_dim:
            self.residual = False

    def forward(self, g: DirectionalBatch, h, e, snorm_n):
        h_in = h
        if g.ndata_eig is None and g.edata_eig is None and any(k >= KIND['dir-av'] for k, _, _ in self.aggregators):
            raise TypeError("directional aggregator without an 'eig' field")     # reference: indexing None
        h = dgn_aggregate(g.plan, h, g.ndata_eig, g.edata_eig, self.aggregators, self.scalers, self.avg_d)
        h = self.posttrans(h)
        if self.graph_norm:
            h = h * snorm_n
        if self.batch_norm:
            h = self.batchnorm_h(h)
        h = F.relu(h)
        if self.residual:
            h = h_in + h
        return F.dropout(h, self.dropout, training=self.training)


class DGNLayer(nn.Module):
    """nets/dgn_layer.py:83-109: only type_net 'simple' exists in the reference."""

    def __init__(self, in_dim, out_dim, dropout, graph_norm, batch_norm, aggregators, scalers, avg_d, type_net, residual,
                 towers=5, divide_input=True, edge_features=None, edge_dim=None, pretrans_layers=1, posttrans_layers=1):
        super().__init__()
        if type_net != 'simple':
            raise NotImplementedError(f"type_net '{type_net}': the reference names DGNLayerComplex / DGNLayerTower "
                                      'without defining them (NameError there)')
        self.model = DGNLayerSimple(in_dim=in_dim, out_dim=out_dim, dropout=dropout, graph_norm=graph_norm,
                                    batch_norm=batch_norm, residual=residual, aggregators=aggregators, scalers=scalers,
                                    avg_d=avg_d, posttrans_layers=posttrans_layers)


class DGNNet(nn.Module):
    """nets/HIV_graph_classification/dgn_net.py:8-83 (forward; the loss is BCEWithLogits, :82-84)."""

    def __init__(self, net_params):
        super().__init__()
        p = net_params
        hidden_dim, out_dim, n_layers = p['hidden_dim'], p['out_dim'], p['L']
        self.pos_enc_dim, self.readout, self.edge_feat = p['pos_enc_dim'], p['readout'], p['edge_feat']
        if self.pos_enc_dim > 0:
            self.embedding_pos_enc = nn.Linear(self.pos_enc_dim, hidden_dim)
        self.in_feat_dropout = nn.Dropout(p['in_feat_dropout'])
        self.embedding_h = AtomEncoder(emb_dim=hidden_dim)
        if self.edge_feat:
            self.embedding_e = BondEncoder(emb_dim=p['edge_dim'])
        kw = dict(dropout=p['dropout'], graph_norm=p['graph_norm'], batch_norm=p['batch_norm'], residual=p['residual'],
                  aggregators=p['aggregators'], scalers=p['scalers'], avg_d=p['avg_d'], type_net=p['type_net'],
                  edge_features=p['edge_feat'], edge_dim=p['edge_dim'], pretrans_layers=p['pretrans_layers'],
                  posttrans_layers=p['posttrans_layers'])
        self.layers = nn.ModuleList([DGNLayer(in_dim=hidden_dim, out_dim=hidden_dim, **kw).model for _ in range(n_layers - 1)])
        self.layers.append(DGNLayer(in_dim=hidden_dim, out_dim=out_dim, **kw).model)
        self.MLP_layer = MLPReadout(out_dim, 1)

    def forward(self, g: DirectionalBatch, h, e, snorm_n=None, snorm_e=None, pos_enc=None):
        h = self.in_feat_dropout(self.embedding_h(h))
        if self.pos_enc_dim > 0:
            h = h + self.embedding_pos_enc(pos_enc)
        if self.edge_feat:
            e = self.embedding_e(e)
        for conv in self.layers:
            h = conv(g, h, e, snorm_n)
        if self.readout == 'max':
            raise NotImplementedError("readout 'max' is not built (sum / mean are)")
        hg = ops.pool_ptr(h, g.node_ptr, mean=self.readout != 'sum')
        return self.MLP_layer(hg)
