"""Batching of graphs: what torch_geometric's DataLoader collate does for the reference (main.py:243-258,
SURVEY A.5), restated because PyG is not a dependency of this package.

Every tensor attribute is concatenated along dim 0, `edge_index` along dim 1 after adding the cumulative node count;
`batch` (graph id per node), `node_ptr`, `edge_ptr` and `num_graphs` are added.  The result is the block-diagonal
batched graph both hot paths consume (COUNT: edge_index + node_ptr; MP: edge_index + per-node / per-edge features)."""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import torch


class Batch:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, non_blocking=False):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def keys(self):
        return list(self.__dict__.keys())


def collate(graphs: Sequence, keys: Optional[Iterable[str]] = None) -> Batch:
    """graphs: objects with .x [n, *] and .edge_index [2, e] plus any per-node / per-edge / per-graph tensors
    (identifiers, degrees, edge_features, y ...).  Per-edge attributes are recognised by their first dimension
    matching edge_index.shape[1] (and not the node count)."""
    if not graphs:
        raise ValueError('cannot collate an empty list of graphs')
    first = graphs[0]
    names = list(keys) if keys is not None else [k for k, v in vars(first).items() if torch.is_tensor(v)]
    sizes = torch.tensor([int(g.x.shape[0]) for g in graphs], dtype=torch.int64)
    esizes = torch.tensor([int(g.edge_index.shape[1]) for g in graphs], dtype=torch.int64)
    node_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)])
    edge_ptr = torch.cat([torch.zeros(1, dtype=torch.int64), esizes.cumsum(0)])
    out = {}
    for k in names:
        vals = [getattr(g, k) for g in graphs]
        if k == 'edge_index' or 'index' in k or 'face' in k:
            out[k] = torch.cat([v + int(node_ptr[i]) for i, v in enumerate(vals)], dim=-1)
        else:
            vals = [v if v.dim() > 0 else v.reshape(1) for v in vals]
            out[k] = torch.cat(vals, dim=0)
    out['batch'] = torch.repeat_interleave(torch.arange(len(graphs), dtype=torch.int64), sizes)
    out['node_ptr'], out['edge_ptr'], out['num_graphs'] = node_ptr, edge_ptr, len(graphs)
    return Batch(**out)
