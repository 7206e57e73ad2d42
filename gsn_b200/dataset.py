"""Dataset preparation and the on-disk cache.

  prepare_graphs   <- the per-graph loop of utils_data_gen.py:47-108 (generate_dataset / _prepare + utils_ids.py:7-29):
                      every graph of the dataset goes through COUNT in a few batched launches instead of one
                      graph-tool call per (graph, pattern); the returned objects carry the reference's attributes
                      (edge_index, x, graph_size, degrees, [edge_features], y, identifiers).
  FlatDataset      <- the cache written at utils.py:272-274 (`torch.save((list[Data], num_classes, orbit_partition_sizes))`,
                      one pickled PyG object per graph).  Here the dataset is a handful of flat tensors + row pointers:
                      loads with one read, lives on the GPU as is, and `batch(indices)` (the DataLoader collate of
                      main.py:243-258) is index arithmetic on those tensors -- no per-graph Python objects.
                      `to_list()` / `save_reference_tuple()` give the reference's layout back.

PyG is not a dependency: `Data` below is the attribute bag the reference code actually uses.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import counting
from .collate import Batch, collate
from .patterns import total_columns

FORMAT_VERSION = 1
_NODE_KEYS = ('x', 'degrees')


class Data:
    """torch_geometric.data.Data as far as the reference uses it: a bag of tensors."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]


def _degree(index, num_nodes):
    out = torch.zeros(num_nodes, dtype=torch.float32)
    return out.scatter_add_(0, index, torch.ones_like(index, dtype=torch.float32))


def prepare_graphs(graphs: Sequence, subgraph_dicts, subgraph_params, id_scope: str, regression: bool = False,
                   dataset_name: str = '', graphs_per_launch: int = 65536) -> List[Data]:
    """graphs: objects with .edge_mat int64 [2,e], .node_features [n,*], .label [, .edge_features] -- what
    utils_data_prep.py hands to generate_dataset.  Semantics of utils_data_gen.py:85-108 + utils_ids.py:7-29:
    self loops are dropped from edge_index / edge_features before counting, identifiers are int64, a graph
    without edges gets a [0, C] identifier block in edge scope."""
    C = total_columns(subgraph_dicts)
    out: List[Data] = []
    for g in graphs:
        d = Data()
        ei = g.edge_mat.long()
        d.x = g.node_features
        d.graph_size = int(g.node_features.shape[0])
        d.degrees = torch.zeros((d.graph_size,)) if ei.shape[1] == 0 else _degree(ei[0], d.graph_size)
        if hasattr(g, 'edge_features'):
            ei, d.edge_features = counting.remove_self_loops(ei, g.edge_features)
        else:
            ei = counting.remove_self_loops(ei)[0]
        d.edge_index = ei
        float_y = regression or dataset_name in {'ogbg-molpcba', 'ogbg-molhiv', 'ZINC'}
        d.y = torch.as_tensor(g.label).unsqueeze(0).float() if float_y else torch.as_tensor(g.label).unsqueeze(0).long()
        out.append(d)
    # one COUNT launch per chunk of graphs (the reference: one graph-tool call per graph and pattern)
    for lo in range(0, len(out), graphs_per_launch):
        chunk = out[lo:lo + graphs_per_launch]
        sizes = torch.tensor([d.graph_size for d in chunk], dtype=torch.int64)
        node_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)])
        esizes = [int(d.edge_index.shape[1]) for d in chunk]
        ei = torch.cat([d.edge_index + int(node_ptr[i]) for i, d in enumerate(chunk)], dim=1)
        ids = counting.count_batch(ei.cuda(), node_ptr, subgraph_dicts, subgraph_params['induced'], id_scope).cpu()
        rows = esizes if id_scope == 'local' else sizes.tolist()
        for d, blk in zip(chunk, ids.split(rows)):
            d.identifiers = blk if blk.shape[0] else torch.zeros((0, C), dtype=torch.int64)
    return out


class FlatDataset:
    """Concatenated dataset: node rows, edge rows (edge_index holds LOCAL node ids), graph rows, row pointers."""

    def __init__(self, tensors: dict, node_ptr: torch.Tensor, edge_ptr: torch.Tensor, meta: Optional[dict] = None):
        self.tensors, self.node_ptr, self.edge_ptr, self.meta = tensors, node_ptr, edge_ptr, dict(meta or {})

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_list(cls, graphs: Sequence, num_classes=None, orbit_partition_sizes=None, kinds: Optional[dict] = None) -> 'FlatDataset':
        """kinds: optional {attribute: 'node' | 'edge' | 'graph'} for attributes whose row count is ambiguous
        (a dataset whose total node and edge counts coincide)"""
        b = collate(graphs)
        tensors = {k: v for k, v in vars(b).items() if torch.is_tensor(v) and k not in ('batch', 'node_ptr', 'edge_ptr')}
        tensors['edge_index'] = b.edge_index - b.node_ptr[:-1].repeat_interleave(b.edge_ptr[1:] - b.edge_ptr[:-1])[None, :]
        ds = cls(tensors, b.node_ptr, b.edge_ptr, {'num_classes': num_classes,
                                                   'orbit_partition_sizes': orbit_partition_sizes})
        ds.meta['kinds'] = ds.kinds(kinds)
        return ds

    def __len__(self):
        return int(self.node_ptr.numel() - 1)

    @property
    def num_nodes(self):
        return int(self.node_ptr[-1])

    @property
    def num_edges(self):
        return int(self.edge_ptr[-1])

    def _kind(self, k, v):
        """'edge' | 'node' | 'graph' rows of attribute k"""
        if k == 'edge_index':
            return 'edge'
        n, e, g = self.num_nodes, self.num_edges, len(self)
        r = v.shape[0]
        if k in _NODE_KEYS and r == n:
            return 'node'
        if r == e and r != n:
            return 'edge'
        if r == n:
            return 'node'
        if r == g:
            return 'graph'
        raise ValueError(f'attribute {k}: {r} rows match neither nodes ({n}), edges ({e}) nor graphs ({g})')

    def kinds(self, hint: Optional[dict] = None):
        out = {}
        for k, v in self.tensors.items():
            out[k] = (hint or {}).get(k) or self.meta.get('kinds', {}).get(k) or self._kind(k, v)
        return out

    # ------------------------------------------------------------------ cache file
    def save(self, path: str):
        """one torch.save of plain tensors (no pickled classes): loads with weights_only=True"""
        torch.save({'format': 'gsn_b200.flat', 'version': FORMAT_VERSION, 'node_ptr': self.node_ptr.cpu(),
                    'edge_ptr': self.edge_ptr.cpu(), 'tensors': {k: v.cpu() for k, v in self.tensors.items()},
                    'meta': {**self.meta, 'kinds': self.kinds()}}, path)

    @classmethod
    def load(cls, path: str, device=None) -> 'FlatDataset':
        obj = torch.load(path, map_location='cpu', weights_only=True)
        if not isinstance(obj, dict) or obj.get('format') != 'gsn_b200.flat':
            raise ValueError(f'{path} is not a gsn_b200 flat cache')
        if obj['version'] != FORMAT_VERSION:
            raise ValueError(f'{path}: cache version {obj["version"]} != {FORMAT_VERSION}')
        ds = cls(obj['tensors'], obj['node_ptr'], obj['edge_ptr'], obj['meta'])
        return ds.to(device) if device is not None else ds

    def to(self, device) -> 'FlatDataset':
        return FlatDataset({k: v.to(device) for k, v in self.tensors.items()}, self.node_ptr.to(device),
                           self.edge_ptr.to(device), self.meta)

    # ------------------------------------------------------------------ the reference's layout
    def to_list(self) -> List[Data]:
        kinds = self.kinds()
        ns = (self.node_ptr[1:] - self.node_ptr[:-1]).tolist()
        es = (self.edge_ptr[1:] - self.edge_ptr[:-1]).tolist()
        parts = {}
        for k, v in self.tensors.items():
            if k == 'edge_index':
                parts[k] = v.split(es, dim=1)
            else:
                parts[k] = v.split({'node': ns, 'edge': es, 'graph': [1] * len(self)}[kinds[k]], dim=0)
        return [Data(**{k: parts[k][i] for k in parts}) for i in range(len(self))]

    def save_reference_tuple(self, path: str):
        """(graphs, num_classes, orbit_partition_sizes) as utils.py:272-274 writes it, with gsn_b200.dataset.Data
        standing in for torch_geometric.data.Data (PyG is absent; the attribute names and tensors are the same)"""
        torch.save((self.to_list(), self.meta.get('num_classes'), self.meta.get('orbit_partition_sizes')), path)

    # ------------------------------------------------------------------ multi-GPU: contiguous shards of graphs
    def shard(self, world: int, rank: int, balance: str = 'edges') -> 'FlatDataset':
        """the contiguous range of graphs owned by `rank` (balanced by edge or node count, distributed.shard_ranges);
        both hot paths shard by graph, so nothing is exchanged between shards"""
        from .distributed import shard_ranges
        w = (self.edge_ptr[1:] - self.edge_ptr[:-1]) if balance == 'edges' else (self.node_ptr[1:] - self.node_ptr[:-1])
        g0, g1 = shard_ranges(w.tolist(), world)[rank]
        n0, n1, e0, e1 = int(self.node_ptr[g0]), int(self.node_ptr[g1]), int(self.edge_ptr[g0]), int(self.edge_ptr[g1])
        kinds = self.kinds()
        t = {}
        for k, v in self.tensors.items():
            if k == 'edge_index':
                t[k] = v[:, e0:e1]
            else:
                lo, hi = {'node': (n0, n1), 'edge': (e0, e1), 'graph': (g0, g1)}[kinds[k]]
                t[k] = v[lo:hi]
        return FlatDataset(t, self.node_ptr[g0:g1 + 1] - n0, self.edge_ptr[g0:g1 + 1] - e0, {**self.meta, 'kinds': kinds})

    # ------------------------------------------------------------------ collate by index arithmetic
    def batch(self, indices) -> Batch:
        """PyG DataLoader collate of the graphs `indices` (int64 tensor / list), on the device the dataset is on"""
        dev = self.node_ptr.device
        idx = torch.as_tensor(indices, dtype=torch.int64, device=dev)
        n0, n1 = self.node_ptr[idx], self.node_ptr[idx + 1]
        e0, e1 = self.edge_ptr[idx], self.edge_ptr[idx + 1]
        ns, es = n1 - n0, e1 - e0
        zero = torch.zeros(1, dtype=torch.int64, device=dev)
        node_ptr = torch.cat([zero, ns.cumsum(0)])
        edge_ptr = torch.cat([zero, es.cumsum(0)])
        gid_n = torch.repeat_interleave(torch.arange(idx.numel(), device=dev), ns)
        gid_e = torch.repeat_interleave(torch.arange(idx.numel(), device=dev), es)
        node_rows = torch.arange(int(node_ptr[-1]), device=dev) - node_ptr[gid_n] + n0[gid_n]
        edge_rows = torch.arange(int(edge_ptr[-1]), device=dev) - edge_ptr[gid_e] + e0[gid_e]
        kinds = self.kinds()
        out = {}
        for k, v in self.tensors.items():
            if k == 'edge_index':
                out[k] = v[:, edge_rows] + node_ptr[gid_e][None, :]
            elif kinds[k] == 'node':
                out[k] = v[node_rows]
            elif kinds[k] == 'edge':
                out[k] = v[edge_rows]
            else:
                out[k] = v[idx]
        out['batch'], out['node_ptr'], out['edge_ptr'], out['num_graphs'] = gid_n, node_ptr, edge_ptr, int(idx.numel())
        return Batch(**out)


class FlatLoader:
    """The DataLoader of main.py:243-258 (`DataLoader(dataset, batch_size=, shuffle=, worker_init_fn=...)` with PyG's
    collate) over a FlatDataset: an epoch is a permutation of graph ids cut into batches, every batch one
    `FlatDataset.batch` call on the device the data set lives on -- no worker processes, no per-graph objects.
    shuffle draws `torch.randperm(n, generator=generator)` per epoch (torch's RandomSampler does the same after the
    DataLoader has consumed one draw for its base seed, so the permutations differ for equal seeds); the last short
    batch is kept unless drop_last."""

    def __init__(self, dataset: FlatDataset, batch_size: int = 1, shuffle: bool = False, drop_last: bool = False,
                 generator: Optional[torch.Generator] = None):
        if batch_size < 1:
            raise ValueError('batch_size must be positive')
        self.dataset, self.batch_size, self.shuffle, self.drop_last, self.generator = dataset, int(batch_size), shuffle, drop_last, generator

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for lo in range(0, n, self.batch_size):
            idx = order[lo:lo + self.batch_size]
            if self.drop_last and idx.numel() < self.batch_size:
                return
            yield self.dataset.batch(idx)

