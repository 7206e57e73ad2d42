"""COUNT: structural identifiers (subgraph-isomorphism orbit counts) on the GPU.

Host-side mirror of the reference's counting interface:

  subgraph_isomorphism_vertex_counts(edge_index, subgraph_dict=, induced=, num_nodes=, directed=)
  subgraph_isomorphism_edge_counts(edge_index, subgraph_dict=, induced=, directed=)
        -> float64 tensor, same meaning as utils_graph_processing.py:103-179
  subgraph_counts2ids(count_fn, data, subgraph_dicts, subgraph_params)
        -> utils_ids.py:7-29
  count_batch(...)   the batched entry the reference's per-graph loop
        (utils_data_gen.py:60-78) cannot express: every graph of a PyG-style
        batch and every pattern, int64 identifiers out.

All arithmetic runs in libgsn_b200.so (csrc/graph_build.cu, csrc/count_kernels.cu).
There is no CPU path: CPU tensors are copied to the current CUDA device and the
result is copied back.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _lib
from .patterns import GsnPlan, compile_plans, total_columns

_W_CLASSES = (1, 2, 4, 8, 16)


def remove_self_loops(edge_index, edge_attr=None):
    """torch_geometric.utils.remove_self_loops as used at utils_ids.py:11-15."""
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def _pick_W(max_nodes: int) -> int:
    need = max(1, (int(max_nodes) + 63) // 64)
    for w in _W_CLASSES:
        if w >= need:
            return w
    raise NotImplementedError(f'graphs with more than {64 * _W_CLASSES[-1]} nodes are not supported '
                              f'(largest graph in batch: {max_nodes})')


_plan_cache = {}


def _plans_for(subgraph_dicts, induced: bool, scope: int) -> List[GsnPlan]:
    key = (tuple((tuple(map(tuple, sd['subgraph'].get_edges().tolist())),
                  tuple(sorted(sd['orbit_membership'].items())), int(sd['aut_count']))
                 for sd in subgraph_dicts), bool(induced), int(scope))
    plans = _plan_cache.get(key)
    if plans is None:
        plans = compile_plans(subgraph_dicts, induced, scope)
        _plan_cache[key] = plans
    return plans


class BatchedGraph:
    """Device-resident simple undirected graph of a whole batch (adjacency
    bitmasks + slot CSR + edge_dict), the GPU counterpart of the gt.Graph built
    at utils_graph_processing.py:110-113 / :150-153 for every single graph."""

    def __init__(self, edge_index: torch.Tensor, node_ptr: torch.Tensor, num_nodes: Optional[int] = None,
                 max_nodes_per_graph: Optional[int] = None):
        _lib.require_cuda(edge_index, 'edge_index')
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError('edge_index must be int64 [2, E]')
        self.edge_index = edge_index.contiguous()
        dev = edge_index.device
        if max_nodes_per_graph is None:
            sizes = node_ptr[1:] - node_ptr[:-1]
            max_nodes_per_graph = int(sizes.max().item()) if sizes.numel() else 0
        if num_nodes is None:
            num_nodes = int(node_ptr[-1].item())
        self.node_ptr = node_ptr.to(device=dev, dtype=torch.int64).contiguous()
        self.N, self.E, self.G = int(num_nodes), int(edge_index.shape[1]), int(node_ptr.numel() - 1)
        self.W = _pick_W(max_nodes_per_graph)
        L = _lib.lib()
        nbytes = ctypes.c_size_t(0)
        _lib.check(L.gsn_graph_workspace_bytes(self.N, self.E, self.W, ctypes.byref(nbytes)), 'gsn_graph_workspace_bytes')
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call('graph_build', 'gsn_graph_build', _lib.ptr(self.edge_index), self.E, _lib.ptr(self.node_ptr),
                      self.G, self.N, self.W, _lib.ptr(self.ws), nbytes.value, _lib.ptr(self.status),
                      _lib.stream_ptr())

    def count(self, plans: Sequence[GsnPlan], n_cols: int, scope: int, out: Optional[torch.Tensor] = None):
        rows = self.N if scope == 0 else self.E
        dev = self.edge_index.device
        if out is None:
            out = torch.empty((rows, n_cols), dtype=torch.int64, device=dev)
        L = _lib.lib()
        scratch, scratch_bytes = None, 0
        if rows == 0:
            return out
        with torch.cuda.device(dev):
            for P in plans:
                nb = ctypes.c_size_t(0)
                _lib.check(L.gsn_count_scratch_bytes(self.N, self.E, ctypes.byref(P), ctypes.byref(nb)), 'gsn_count_scratch_bytes')
                if scratch is None or nb.value > scratch_bytes:
                    scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
                    scratch_bytes = nb.value
                _lib.call('count_pattern', 'gsn_count_pattern', _lib.ptr(self.ws), self.N, self.E, self.W,
                          _lib.ptr(self.edge_index), _lib.ptr(self.node_ptr), self.G, ctypes.byref(P), _lib.ptr(out),
                          n_cols, _lib.ptr(scratch), scratch_bytes, _lib.ptr(self.status), _lib.stream_ptr())
        return out

    def raise_on_status(self):
        """Synchronises; turns device-side status bits into the reference's errors."""
        raise_on_status_bits(int(self.status.item()))


SMALL_GRAPH_NODES = 64          # gsn_count_small: one 64-bit adjacency word per vertex
SMALL_PATH = True               # False: always gsn_graph_build + gsn_count_pattern (tests compare the two)


def _small_path_ok(plans: Sequence[GsnPlan], max_nodes_per_graph: Optional[int]) -> bool:
    if not SMALL_PATH or max_nodes_per_graph is None or max_nodes_per_graph > SMALL_GRAPH_NODES:
        return False
    return all(not (P.family == 1 and P.kmax > 12) for P in plans)


def _count_small(edge_index, node_ptr, plans, n_cols, scope, N, status):
    """one launch per plan, no workspace (csrc/count_small.cu)"""
    E, G = int(edge_index.shape[1]), int(node_ptr.numel() - 1)
    dev = edge_index.device
    rows = N if scope == 0 else E
    out = torch.empty((rows, n_cols), dtype=torch.int64, device=dev)
    if rows == 0 or G == 0:
        return out.zero_()
    with torch.cuda.device(dev):
        for P in plans:
            _lib.call('count_pattern', 'gsn_count_small', _lib.ptr(edge_index), E, _lib.ptr(node_ptr), G, N, ctypes.byref(P),
                      _lib.ptr(out), n_cols, _lib.ptr(status), _lib.stream_ptr())
    return out


def raise_on_status_bits(bits: int):
    """device-side status bits -> the reference's errors"""
    if bits:
        msg = _lib.status_message(bits)
        if bits & _lib.S_MISSING_EDGE:
            raise KeyError(msg)          # utils_graph_processing.py:173 raises KeyError
        if bits & _lib.S_COUNT_OVERFLOW:
            raise OverflowError(msg)
        raise ValueError(msg)


def count_batch(edge_index: torch.Tensor, node_ptr: torch.Tensor, subgraph_dicts, induced: bool, id_scope: str,
                num_nodes: Optional[int] = None, max_nodes_per_graph: Optional[int] = None,
                check: bool = True, graph: Optional[BatchedGraph] = None, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """identifiers int64 [N, C] (id_scope='global') or [E, C] ('local') for a whole
    batch: the concatenated result of utils_ids.py:19-27 over every graph.
    `edge_index` holds batched (global) node ids, `node_ptr` the first node of
    every graph (PyG `batch.ptr`).  Self loops get zero rows (callers strip them
    first like utils_ids.py:11-15 if they want the reference's row set).

    Batches of small graphs (<= 64 nodes each, known from `max_nodes_per_graph`) take the one-launch path
    (gsn_count_small); it needs edge_index grouped by graph, and a batch that is not falls back to the general path
    (with check=True; with check=False the caller reads the GSN_S_NOT_GROUPED bit from `status`, an int32 [1] device
    tensor it passes in)."""
    scope = 1 if id_scope == 'local' else 0
    dev_in = edge_index.device
    if not edge_index.is_cuda:
        edge_index = edge_index.cuda()
    plans = _plans_for(subgraph_dicts, induced, scope)
    n_cols = total_columns(subgraph_dicts)
    if graph is None and max_nodes_per_graph is None and SMALL_PATH:
        sizes = node_ptr[1:] - node_ptr[:-1]
        max_nodes_per_graph = int(sizes.max().item()) if sizes.numel() else 0
    if graph is None and _small_path_ok(plans, max_nodes_per_graph):
        ei = edge_index.contiguous()
        dev = ei.device
        nptr = node_ptr.to(device=dev, dtype=torch.int64).contiguous()
        N = int(num_nodes) if num_nodes is not None else int(node_ptr[-1].item())
        st = status if status is not None else torch.zeros(1, dtype=torch.int32, device=dev)
        out = _count_small(ei, nptr, plans, n_cols, scope, N, st)
        if check:
            bits = int(st.item())
            if bits & _lib.S_NOT_GROUPED and not (bits & ~_lib.S_NOT_GROUPED & ~_lib.S_CROSS_GRAPH_EDGE & ~_lib.S_MISSING_EDGE):
                st.zero_()
                graph = BatchedGraph(ei, nptr, N, max_nodes_per_graph)      # arbitrary column order: general path
                out = graph.count(plans, n_cols, scope)
                graph.raise_on_status()
            else:
                raise_on_status_bits(bits)
        return out if dev_in.type == 'cuda' else out.to(dev_in)
    if graph is None:
        graph = BatchedGraph(edge_index, node_ptr, num_nodes, max_nodes_per_graph)
    out = graph.count(plans, n_cols, scope)
    if check:
        graph.raise_on_status()
    elif status is not None:
        status.bitwise_or_(graph.status)
    return out if dev_in.type == 'cuda' else out.to(dev_in)


def _single(edge_index, subgraph_dict, induced, num_nodes, scope):
    ei = edge_index if torch.is_tensor(edge_index) else torch.as_tensor(edge_index)
    ei = ei.long()
    dev_in = ei.device
    if ei.numel():
        n_match = int(ei.max().item()) + 1        # graph-tool creates vertices 0..max id
    else:
        n_match = 0
    n = max(n_match, int(num_nodes) if num_nodes is not None else 0)
    if scope == 0 and num_nodes is not None and n_match > num_nodes:
        raise IndexError('edge_index refers to a vertex >= num_nodes')
    node_ptr = torch.tensor([0, n], dtype=torch.int64)
    out = count_batch(ei.cuda() if not ei.is_cuda else ei, node_ptr, [subgraph_dict], induced,
                      'local' if scope else 'global', num_nodes=n, max_nodes_per_graph=n)
    if scope == 0:
        out = out[:num_nodes] if num_nodes is not None else out
    return out.to(torch.float64).to(dev_in)       # count_fn returns float64 (torch.tensor(np.float64), :129/:177)


def subgraph_isomorphism_vertex_counts(edge_index, **kwargs):
    """Drop-in for utils_graph_processing.py:103-131."""
    if kwargs.get('directed', False):
        raise NotImplementedError('directed=True is not supported')
    return _single(edge_index, kwargs['subgraph_dict'], kwargs['induced'], kwargs['num_nodes'], 0)


def subgraph_isomorphism_edge_counts(edge_index, **kwargs):
    """Drop-in for utils_graph_processing.py:134-179."""
    if kwargs.get('directed', False):
        raise NotImplementedError('directed=True is broken in the reference (SURVEY A.3) and not supported')
    return _single(edge_index, kwargs['subgraph_dict'], kwargs['induced'], kwargs.get('num_nodes'), 1)


def subgraph_counts2ids(count_fn, data, subgraph_dicts, subgraph_params):
    """Drop-in for utils_ids.py:7-29 (data: any object with .x, .edge_index
    [, .edge_features]).  When count_fn is one of this module's two functions all
    patterns are counted in one batched call; any other callable is looped over
    exactly like the reference."""
    if hasattr(data, 'edge_features'):
        edge_index, edge_features = remove_self_loops(data.edge_index, data.edge_features)
        setattr(data, 'edge_features', edge_features)
    else:
        edge_index = remove_self_loops(data.edge_index)[0]
    num_nodes = data.x.shape[0]
    if count_fn in (subgraph_isomorphism_vertex_counts, subgraph_isomorphism_edge_counts):
        scope = 'local' if count_fn is subgraph_isomorphism_edge_counts else 'global'
        n_match = int(edge_index.max().item()) + 1 if edge_index.numel() else 0
        n = max(num_nodes, n_match)
        ids = count_batch(edge_index, torch.tensor([0, n], dtype=torch.int64), subgraph_dicts,
                          subgraph_params['induced'], scope, num_nodes=n, max_nodes_per_graph=n)
        identifiers = ids[:num_nodes] if scope == 'global' else ids
    else:
        identifiers = None
        for subgraph_dict in subgraph_dicts:
            kwargs = {'subgraph_dict': subgraph_dict, 'induced': subgraph_params['induced'],
                      'num_nodes': num_nodes, 'directed': subgraph_params['directed']}
            counts = count_fn(edge_index, **kwargs)
            identifiers = counts if identifiers is None else torch.cat((identifiers, counts), 1)
    setattr(data, 'edge_index', edge_index)
    setattr(data, 'identifiers', identifiers.long())
    return data
