"""TEST INFRASTRUCTURE -- CPU oracle of the DGN consumer path (directional_gsn/).

Two things live here:

1. `aggregate(...)`: a plain-PyTorch restatement of DGNLayerSimple.pretrans_edges /
   message_func / reduce_func (directional_gsn/nets/dgn_layer.py:28-54), of every aggregator
   in nets/aggregators.py:8-69 and of the scalers in nets/scalers.py:7-20, evaluated the way
   DGL evaluates a reduce UDF: nodes are bucketed by in-degree D and each bucket sees a
   mailbox [n, D, ...] whose messages are in edge-id order; nodes without in-edges keep zeros.

2. `FakeDGLGraph` + `import_reference()`: `dgl` is a third-party dependency absent from this
   image (version unpinned, README.md).  The fake graph implements exactly the three calls the
   reference layer makes (ndata / edata dicts, apply_edges, update_all with degree bucketing), so
   the UNMODIFIED reference files nets/dgn_layer.py, nets/aggregators.py, nets/scalers.py and
   nets/layers.py run in the build container and pin (1) -- scripts/make_golden_dgn.py stores
   their outputs in tests/golden/dgn.pt.  Parity pinned by: reference code + degree-bucketing
   restatement of DGL's documented update_all semantics (not by a DGL run).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/directional_gsn'
EPS = 1e-8        # aggregators.py:5


# ------------------------------------------------------------------ (1) restatement
def _buckets(dst, num_nodes):
    """[(D, nodes[n], edge ids [n, D] ascending)] for every in-degree D > 0"""
    E = dst.numel()
    order = torch.argsort(dst * max(E, 1) + torch.arange(E), stable=True)       # by dst, then edge id
    deg = torch.bincount(dst, minlength=num_nodes)
    start = torch.cumsum(deg, 0) - deg
    out = []
    for D in torch.unique(deg).tolist():
        if D == 0:
            continue
        nodes = (deg == D).nonzero(as_tuple=True)[0]
        eids = order[(start[nodes][:, None] + torch.arange(D)[None, :])]
        out.append((D, nodes, eids))
    return out


def _aggr(kind, idx, alpha, h, vf, h_in):
    if kind == 'mean':
        return torch.mean(h, dim=1)
    if kind == 'sum':
        return torch.sum(h, dim=1)
    if kind == 'max':
        return torch.max(h, dim=1)[0]
    if kind == 'min':
        return torch.min(h, dim=1)[0]
    if kind in ('std', 'var'):
        var = torch.relu(torch.mean(h * h, dim=-2) - torch.mean(h, dim=-2) ** 2)
        return var if kind == 'var' else torch.sqrt(var + EPS)
    f = vf[:, :, idx]
    if kind == 'dir-av':
        return torch.sum(h * (f.abs() / (f.abs().sum(1, keepdim=True) + EPS)).unsqueeze(-1), dim=1)
    if kind == 'dir-softmax':
        return torch.sum(h * torch.softmax(alpha * f.abs().unsqueeze(-1), dim=1), dim=1)
    if kind in ('dir-dx', 'dir-dx-no-abs'):
        w = (f / (f.abs().sum(1, keepdim=True) + EPS)).unsqueeze(-1)
        r = torch.sum(h * w, dim=1) - torch.sum(w, dim=1) * h_in
        return r.abs() if kind == 'dir-dx' else r
    if kind == 'dir-dx-balanced':
        fr = torch.relu(f) / (torch.relu(f).abs().sum(1, keepdim=True) + EPS)
        bk = torch.relu(-f) / ((-torch.relu(-f)).abs().sum(1, keepdim=True) + EPS)
        w = ((fr + bk) / 2).unsqueeze(-1)
        return (torch.sum(h * w, dim=1) - torch.sum(w, dim=1) * h_in).abs()
    raise KeyError(kind)


_KIND_NAMES = ['mean', 'sum', 'max', 'min', 'std', 'var', 'dir-av', 'dir-softmax', 'dir-dx', 'dir-dx-no-abs',
               'dir-dx-balanced']


def aggregate(edge_index, num_nodes, h, node_field, edge_field, aggregators, scalers, avg_log):
    """aggregators: [(kind code, eig idx, alpha)] as gsn_b200.directional.parse_aggregators returns;
    scalers: [0 identity | 1 amplification | 2 attenuation]"""
    src, dst = edge_index[0], edge_index[1]
    d = h.shape[1]
    out = torch.zeros((num_nodes, len(aggregators) * len(scalers) * d), dtype=torch.float32)
    vf_e = None
    if node_field is not None:
        vf_e = node_field[src] - node_field[dst]                                   # dgn_layer.py:30
    if edge_field is not None:
        vf_e = edge_field if vf_e is None else torch.cat((vf_e, edge_field), dim=1)   # :33
    for D, nodes, eids in _buckets(dst, num_nodes):
        hm = h[src[eids]]                                                          # mailbox [n, D, d]
        vf = None if vf_e is None else vf_e[eids]
        parts = [_aggr(_KIND_NAMES[k], i, a, hm, vf, h[nodes]) for k, i, a in aggregators]
        r = torch.cat(parts, dim=1)
        if len(scalers) > 1:                                                       # :50-51
            sc = []
            for s in scalers:
                if s == 0:
                    sc.append(r)
                elif s == 1:
                    sc.append(r * (np.log(D + 1) / avg_log))
                else:
                    sc.append(r * (avg_log / np.log(D + 1)))
            r = torch.cat(sc, dim=1)
        out[nodes] = r
    return out


# ------------------------------------------------------------------ (2) the reference's own files on a fake DGL graph
class _Edges:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class _Nodes:
    def __init__(self, data, mailbox):
        self.data, self.mailbox = data, mailbox


class FakeDGLGraph:
    """ndata / edata + apply_edges + update_all(message, reduce) with DGL's degree bucketing"""

    def __init__(self, edge_index, num_nodes):
        self.src, self.dst, self.n = edge_index[0], edge_index[1], int(num_nodes)
        self.ndata, self.edata = {}, {}

    def number_of_nodes(self):
        return self.n

    def edges(self):
        return self.src, self.dst

    def apply_edges(self, fn):
        ed = _Edges({k: v[self.src] for k, v in self.ndata.items()}, {k: v[self.dst] for k, v in self.ndata.items()},
                    self.edata)
        for k, v in fn(ed).items():
            if v is not None:
                self.edata[k] = v

    def update_all(self, message_func, reduce_func):
        ed = _Edges({k: v[self.src] for k, v in self.ndata.items()}, {k: v[self.dst] for k, v in self.ndata.items()},
                    self.edata)
        msgs = message_func(ed)
        results = {}
        for D, nodes, eids in _buckets(self.dst, self.n):
            nb = _Nodes({k: v[nodes] for k, v in self.ndata.items()}, {k: v[eids] for k, v in msgs.items()})
            for k, v in reduce_func(nb).items():
                if k not in results:
                    results[k] = torch.zeros((self.n,) + tuple(v.shape[1:]), dtype=v.dtype)
                results[k][nodes] = v
        self.ndata.update(results)


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'nets'))


def import_reference():
    """-> module namespace with DGNLayerSimple, AGGREGATORS, SCALERS, MLP of the unmodified reference"""
    if not available():
        raise RuntimeError('/root/reference is not present (GPU box?)')
    if 'dgl' not in sys.modules:
        dgl = types.ModuleType('dgl')
        for name in ('dgl.nn', 'dgl.nn.pytorch', 'dgl.nn.pytorch.glob'):
            sys.modules[name] = types.ModuleType(name)
        sys.modules['dgl.nn.pytorch.glob'].mean_nodes = None        # imported by name only (dgn_layer.py:9)
        sys.modules['dgl.nn.pytorch.glob'].sum_nodes = None
        sys.modules['dgl'] = dgl
    pkg = types.ModuleType('gsn_ref_dgn_nets')
    pkg.__path__ = [os.path.join(REF, 'nets')]
    sys.modules['gsn_ref_dgn_nets'] = pkg
    mods = {}
    for name in ('aggregators', 'scalers', 'layers', 'dgn_layer'):
        spec = importlib.util.spec_from_file_location(f'gsn_ref_dgn_nets.{name}', os.path.join(REF, 'nets', name + '.py'))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return types.SimpleNamespace(DGNLayerSimple=mods['dgn_layer'].DGNLayerSimple, DGNLayer=mods['dgn_layer'].DGNLayer,
                                 AGGREGATORS=mods['aggregators'].AGGREGATORS, SCALERS=mods['scalers'].SCALERS,
                                 MLP=mods['layers'].MLP)
