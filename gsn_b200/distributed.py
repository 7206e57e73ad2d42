"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink).

Both hot paths shard by graph (graphs are independent units): every rank counts /
runs the forward on its own contiguous range of graphs and NO collective touches
the data path.  The only exchanges are
  * the gradient all-reduce of a training step (one flat fp32 buffer, DDP-equivalent:
    BatchNorm statistics stay per shard), and
  * optionally the union of per-column distinct identifier values that
    one_hot_unique (utils_encoding.py:37-59) needs when COUNT ran sharded.
The reference is single-process (main.py:54-59); its only parallelism is joblib
over graphs in preprocessing (utils_data_gen.py:60-71).
Works with the gloo backend on CPU tensors too (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(weights: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges [g0, g1) of graphs per rank with balanced total weight
    (e.g. edge counts for MP, sum of squared degrees for COUNT).  Every graph is
    assigned exactly once; ranges may be empty when there are fewer graphs than ranks."""
    w = np.asarray(weights, dtype=np.float64)
    G = int(w.shape[0])
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        g = int(np.searchsorted(cum, target, side='left'))
        # choose the closer of the two neighbouring cut points
        if g > 0 and g <= G and abs(cum[g - 1] - target) <= abs(cum[min(g, G)] - target):
            g -= 1
        bounds.append(min(max(g, bounds[-1]), G))
    bounds.append(G)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_batch(batch: Dict[str, np.ndarray], world: int, rank: int, balance: str = 'edges') -> Dict[str, np.ndarray]:
    """Sub-batch of graphs owned by `rank`, node ids re-based to start at 0 (numpy batch as produced
    by gsn_b200.synthetic / a PyG-style collate: edge_index, node_ptr, edge_ptr + per-node/per-edge arrays)."""
    node_ptr, edge_ptr = batch['node_ptr'], batch['edge_ptr']
    G = len(node_ptr) - 1
    if balance == 'edges':
        w = np.diff(edge_ptr)
    elif balance == 'nodes':
        w = np.diff(node_ptr)
    elif balance.startswith('deg_pow'):
        # COUNT of k-vertex patterns: the work of a vertex grows like deg^(k-1) ("balanced by sum deg^2
        # or edge count"); 'deg_pow3' = sum over the graph's vertices of deg^3, ...
        power = float(balance[len('deg_pow'):] or 2)
        deg = np.bincount(batch['edge_index'][0], minlength=int(node_ptr[-1])).astype(np.float64)
        w = np.add.reduceat(deg ** power, node_ptr[:-1].clip(max=max(int(node_ptr[-1]) - 1, 0))) if G else np.zeros(0)
        w = np.where(np.diff(node_ptr) > 0, w, 0.0)
    else:
        raise ValueError(balance)
    g0, g1 = shard_ranges(w, world)[rank]
    n0, n1, e0, e1 = int(node_ptr[g0]), int(node_ptr[g1]), int(edge_ptr[g0]), int(edge_ptr[g1])
    N, E = int(node_ptr[-1]), int(edge_ptr[-1])
    out = {}
    for k, v in batch.items():
        if k == 'edge_index':
            out[k] = v[:, e0:e1] - n0
        elif k == 'node_ptr':
            out[k] = v[g0:g1 + 1] - n0
        elif k == 'edge_ptr':
            out[k] = v[g0:g1 + 1] - e0
        elif k == 'batch':
            out[k] = v[n0:n1] - g0
        elif isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == N:
            out[k] = v[n0:n1]
        elif isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == E:
            out[k] = v[e0:e1]
        elif isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == G:
            out[k] = v[g0:g1]
        elif k == 'num_graphs':
            out[k] = g1 - g0
        else:
            out[k] = v
    out['graph_range'] = (g0, g1)
    return out


def allreduce_gradients(parameters, group=None, average: bool = True):
    """Sum (average) the gradients of all ranks through ONE flat fp32 buffer: the model has ~0.4 M (ZINC) to
    3.3 M (molhiv) parameters, so a single latency-bound all-reduce beats per-tensor calls."""
    # every trainable parameter takes part, a missing gradient (an unused parameter, an empty shard) as zeros: the
    # buffer layout must be identical on every rank or the collective hangs / mixes tensors up
    params = [p for p in parameters if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


class FlatGradients:
    """The gradients of all trainable parameters as views into ONE fp32 buffer, so that the only collective of the
    system -- the gradient all-reduce of a training step -- is a single NCCL call on memory that already holds the
    gradients (no gather into a staging buffer, no copy back): autograd accumulates into an existing .grad in place,
    so the views survive backward; zero() replaces optimizer.zero_grad().  Works under CUDA-graph capture (static
    addresses)."""

    def __init__(self, parameters):
        self.params = [p for p in parameters if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else 'cpu'
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError('FlatGradients: fp32 parameters only')
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None, average: bool = True):
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.mul_(1.0 / dist.get_world_size(group))


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """replicate the weights of rank `src` (start of training)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    # in place on the tensors themselves (not .data): the version counters move, so weight-derived caches
    # (ops.split_weight, FusedForward, mlp.bn_affine) see the new values
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)


def global_unique_per_column(local_ids: torch.Tensor, group=None) -> List[torch.Tensor]:
    """Sorted distinct values of every identifier column over ALL ranks: the vocabulary of
    one_hot_unique when COUNT ran sharded.  Exchanges only the per-shard uniques (KBs)."""
    cols = [torch.unique(local_ids[:, c]) for c in range(local_ids.shape[1])]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return cols
    world = dist.get_world_size(group)
    out = []
    for u in cols:
        n = torch.tensor([u.numel()], dtype=torch.int64, device=u.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        m = int(max(int(s) for s in sizes))
        pad = torch.full((m,), -1, dtype=u.dtype, device=u.device)
        pad[:u.numel()] = u
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        allv = torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)])
        out.append(torch.unique(allv))
    return out
