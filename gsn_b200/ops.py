"""MP operators: thin PyTorch-tensor front-ends of the C ABI (include/gsn_b200.h).

`EdgePlan` is the once-per-batch grouping of edge_index by aggregation index; it
replaces the COO tensor + torch.sparse.sum coalesce the reference redoes in every
layer (graph_filters/GSN_sparse.py:140-143).  The three fused forward operators
replace propagate()+message() of the reference layers; see csrc/mp_kernels.cu.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib


class GsnSegment(ctypes.Structure):
    """ctypes image of `struct GsnSegment` (include/gsn_b200.h)."""
    _fields_ = [('src', ctypes.c_void_p), ('self_', ctypes.c_void_p), ('width', ctypes.c_int32),
                ('src_ld', ctypes.c_int32), ('self_ld', ctypes.c_int32), ('index_mode', ctypes.c_int32),
                ('self_const', ctypes.c_float), ('_pad', ctypes.c_int32)]


MODE_NONE, MODE_NBR, MODE_EDGE = 0, 1, 2


def _f32c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class EdgePlan:
    """CSR of edge_index grouped by the aggregation index.

    flow='source_to_target' (the CLI default, main.py:618): select = 1, i.e.
    messages are summed at edge_index[1] and x_j is gathered at edge_index[0]
    (GSN_sparse.py:125-129)."""

    def __init__(self, edge_index: torch.Tensor, num_nodes: int, flow: str = 'source_to_target',
                 status: Optional[torch.Tensor] = None):
        _lib.require_cuda(edge_index, 'edge_index')
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError('edge_index must be int64 [2, E]')
        ei = edge_index.contiguous()
        select = 0 if flow == 'target_to_source' else 1
        self.select = select
        self.N, self.E = int(num_nodes), int(ei.shape[1])
        dev = ei.device
        self.device = dev
        self.edge_index = ei
        self.rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=dev)
        self.eid = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)
        self.nbr = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)
        # device-side status word (GSN_S_INDEX_RANGE for ids outside [0, N)): the caller's, so that one check at the
        # end of a step covers the plan too, or a private one read by raise_on_status()
        self.status = status if status is not None else torch.zeros(1, dtype=torch.int32, device=dev)
        L = _lib.lib()
        nb = ctypes.c_size_t(0)
        _lib.check(L.gsn_csr_workspace_bytes(self.N, self.E, ctypes.byref(nb)), 'gsn_csr_workspace_bytes')
        ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        key, other = ei[select], ei[1 - select]
        with torch.cuda.device(dev):
            _lib.call('csr_build', 'gsn_csr_build', _lib.ptr(key), _lib.ptr(other), self.E, self.N, _lib.ptr(self.rowptr), _lib.ptr(self.eid), _lib.ptr(self.nbr), _lib.ptr(ws), nb.value, _lib.ptr(self.status), _lib.stream_ptr())
        self._ws = ws           # keep alive until the stream has consumed it
        self._transposed: Optional['EdgePlan'] = None
        self._deg: Optional[torch.Tensor] = None

    def degree(self) -> torch.Tensor:
        """number of aggregated messages per node, float32 [N]"""
        if self._deg is None:
            self._deg = torch.diff(self.rowptr).to(torch.float32)
        return self._deg

    def raise_on_status(self):
        bits = int(self.status.item())
        if bits:
            raise IndexError(_lib.status_message(bits))


_plan_cache: List[Tuple[tuple, EdgePlan]] = []
_PLAN_CACHE_SIZE = 8


CHECK_PLANS = False      # True (tests): synchronise and raise IndexError on a malformed edge_index when a plan is built


def edge_plan(edge_index: torch.Tensor, num_nodes: int, flow: str = 'source_to_target',
              status: Optional[torch.Tensor] = None) -> EdgePlan:
    """Cached EdgePlan: the layers of one model see the same edge_index tensor, so
    the CSR is built once per batch, not once per layer.  status: see EdgePlan."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes), flow,
           edge_index.device.index)
    for k, p in _plan_cache:
        if k == key and p.edge_index.data_ptr() == edge_index.data_ptr():
            return p
    p = EdgePlan(edge_index, num_nodes, flow, status)
    if CHECK_PLANS:
        p.raise_on_status()
    _plan_cache.append((key, p))
    if len(_plan_cache) > _PLAN_CACHE_SIZE:
        _plan_cache.pop(0)
    return p


def clear_plan_cache():
    _plan_cache.clear()


# ----------------------------------------------------------------------------
def gin_aggregate(plan: EdgePlan, segments: Sequence[dict], eps: Optional[torch.Tensor]) -> torch.Tensor:
    """out[i] = (1+eps) * cat_s(self_s[i] + const_s) + sum_{e -> i} cat_s(src_s[index_s(e)]).

    segments: dicts with keys width, src (tensor or None), mode (MODE_*), self
    (tensor [N,w], [1,w]/[w] broadcast, or None), const (float)."""
    segs = (GsnSegment * len(segments))()
    keep = []
    D = 0
    for i, s in enumerate(segments):
        w = int(s['width'])
        src = _f32c(s.get('src'), 'segment src')
        slf = _f32c(s.get('self'), 'segment self')
        mode = int(s.get('mode', MODE_NONE))
        if src is None:
            mode = MODE_NONE
        segs[i].width = w
        segs[i].index_mode = mode
        segs[i].self_const = float(s.get('const', 0.0))
        if src is not None:
            if src.dim() != 2 or src.shape[1] != w:
                raise ValueError(f'segment {i}: src must be [rows, {w}]')
            want = plan.N if mode == MODE_NBR else plan.E
            if src.shape[0] != want:
                raise ValueError(f'segment {i}: src has {src.shape[0]} rows, expected {want}')
            segs[i].src = src.data_ptr()
            segs[i].src_ld = w
            keep.append(src)
        if slf is not None:
            if slf.numel() == w:
                segs[i].self_ld = 0
            elif slf.dim() == 2 and slf.shape == (plan.N, w):
                segs[i].self_ld = w
            else:
                raise ValueError(f'segment {i}: self must be [N,{w}] or [{w}]')
            segs[i].self_ = slf.data_ptr()
            keep.append(slf)
        D += w
    out = torch.empty((plan.N, D), dtype=torch.float32, device=plan.device)
    eps_t = _f32c(eps, 'eps')
    with torch.cuda.device(plan.device):
        _lib.call('mp_gin', 'gsn_mp_gin_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr), plan.N, plan.E, ctypes.cast(segs, ctypes.c_void_p), len(segments), _lib.ptr(eps_t), _lib.ptr(out), _lib.stream_ptr())
    return out


def ogb_aggregate(plan: EdgePlan, x: torch.Tensor, identifiers: Optional[torch.Tensor], id_per_edge: bool,
                  edge_features: torch.Tensor, eps: Optional[torch.Tensor]) -> torch.Tensor:
    """(1+eps)*(x [+ id]) + sum_e relu(x_j + id + e_ij)  (GSN_edge_sparse_ogb.py:75-84,119-126)."""
    x = _f32c(x, 'x')
    idt = _f32c(identifiers, 'identifiers')
    ef = _f32c(edge_features, 'edge_features')
    d = x.shape[1]
    if ef.shape != (plan.E, d) or (idt is not None and idt.shape != ((plan.E if id_per_edge else plan.N), d)):
        raise ValueError('ogb message kind needs x, identifiers and edge_features of equal width')
    out = torch.empty((plan.N, d), dtype=torch.float32, device=plan.device)
    eps_t = _f32c(eps, 'eps')
    with torch.cuda.device(plan.device):
        _lib.call('mp_ogb', 'gsn_mp_ogb_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr), plan.N, plan.E, _lib.ptr(x), _lib.ptr(idt), int(bool(id_per_edge)), _lib.ptr(ef), d, _lib.ptr(eps_t), _lib.ptr(out), _lib.stream_ptr())
    return out


def segment_sum(plan: EdgePlan, rows: torch.Tensor, gather_neighbour: bool = False) -> torch.Tensor:
    """out[i] = sum_{e -> i} rows[e]   (or rows[nbr(e)] with gather_neighbour)."""
    rows = _f32c(rows, 'rows')
    d = rows.shape[1]
    out = torch.empty((plan.N, d), dtype=torch.float32, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.call('segment_sum', 'gsn_mp_segment_sum', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr), plan.N, plan.E, _lib.ptr(rows), d, int(bool(gather_neighbour)), _lib.ptr(out), _lib.stream_ptr())
    return out


ACTIVATIONS = {'relu': 0, 'elu': 1, 'tanh': 2, 'identity': 3}


def general_edge(plan: EdgePlan, P: torch.Tensor, Q: Optional[torch.Tensor], scale: Optional[torch.Tensor],
                 shift: Optional[torch.Tensor], activation: str = 'relu') -> torch.Tensor:
    """S[i] = sum_{e -> i} act((P[i,:dh] + P[nbr(e),dh:] + Q[e]) * scale + shift)."""
    P = _f32c(P, 'P')
    Q = _f32c(Q, 'Q')
    dh = P.shape[1] // 2
    S = torch.empty((plan.N, dh), dtype=torch.float32, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.call('general_edge', 'gsn_mp_general_edge_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr), plan.N, plan.E, _lib.ptr(P), _lib.ptr(Q), dh, _lib.ptr(_f32c(scale, 'scale')), _lib.ptr(_f32c(shift, 'shift')), ACTIVATIONS[activation], _lib.ptr(S), None, _lib.stream_ptr())
    return S


def general_edge_stats(plan: EdgePlan, P: torch.Tensor, Q: Optional[torch.Tensor]) -> torch.Tensor:
    """Per-channel [sum, sum of squares] of h_e = P_i + P_j + Q_e over all edges
    (float64 [2, dh]): the batch statistics BatchNorm1d needs in training mode
    (models_misc.py:54-55 runs BN over the E message rows)."""
    P = _f32c(P, 'P')
    Q = _f32c(Q, 'Q')
    dh = P.shape[1] // 2
    stats = torch.zeros((2, dh), dtype=torch.float64, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.call('general_edge_stats', 'gsn_mp_general_edge_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid), _lib.ptr(plan.nbr), plan.N, plan.E, _lib.ptr(P), _lib.ptr(Q), dh, None, None, 3, None, _lib.ptr(stats), _lib.stream_ptr())
    return stats



# ----------------------------------------------------------------------------
# dense tail + index fast path
# ----------------------------------------------------------------------------
class GsnLinear(ctypes.Structure):
    """ctypes image of `struct GsnLinear` (include/gsn_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ('A1', 'A2', 'W', 'bias', 'row_scale', 'row_vec', 'tab_idx', 'tab',
                                               'scale', 'shift', 'C')] + \
               [(n, ctypes.c_int32) for n in ('M', 'Nout', 'K1', 'K2', 'lda1', 'lda2', 'ldw', 'ldc', 'tab_ld', 'act',
                                              'vec_ok', 'accumulate', 'tc_path')]


class GsnEncodeCol(ctypes.Structure):
    """ctypes image of `struct GsnEncodeCol`."""
    _fields_ = [('src', ctypes.c_void_p), ('stride', ctypes.c_int64), ('vocab_begin', ctypes.c_int32),
                ('vocab_end', ctypes.c_int32), ('table_off', ctypes.c_int32), ('rows', ctypes.c_int32)]


def _dp(t):
    return None if t is None else t.data_ptr()


# Tensor-core dense tail (tcgen05 3xTF32).  "auto": use it whenever the shapes allow (K % 4 == 0, aligned).
TENSOR_CORES = 'auto'       # 'auto' | 'off'
TC_PATH = 0                  # GsnLinear.tc_path for every call (GSN_TC_PATH_*: 0 auto, 1 pre-split pass, 2 one tile per CTA)
import weakref

_wsplit_cache = {}          # id(tensor) -> (weakref to the tensor, stamp, hi, lo); evicted when the tensor dies


def split_weight(W: torch.Tensor):
    """(W_hi, W_lo) tf32 split of a weight matrix, cached ON the tensor object: an entry is valid for the same Python
    tensor with the same storage address and version only.  (Keying on data_ptr alone is wrong: nn.Module.cpu() / .cuda()
    swap a Parameter's storage in place, the freed address is recycled by another layer's weight of the same shape and
    version, and the stale split of the OLD weight would be returned.)"""
    stamp = (W.data_ptr(), W._version, tuple(W.shape), W.stride(0))
    key = id(W)
    hit = _wsplit_cache.get(key)
    if hit is None or hit[0]() is not W or hit[1] != stamp:
        hi = torch.empty((W.shape[0], W.shape[1]), dtype=torch.float32, device=W.device)
        lo = torch.empty_like(hi)
        with torch.cuda.device(W.device):
            _lib.call('split_tf32', 'gsn_split_tf32', _lib.ptr(W), W.shape[0], W.shape[1], W.stride(0), _lib.ptr(hi),
                      _lib.ptr(lo), _lib.stream_ptr())
        hit = (weakref.ref(W, lambda _r, k=key: _wsplit_cache.pop(k, None)), stamp, hi, lo)
        _wsplit_cache[key] = hit
    return hit[2], hit[3]


def _tc_eligible(A1, A2, W):
    if TENSOR_CORES == 'off' or W is None or A1 is None:
        return False
    K1 = A1.shape[1]
    K2 = 0 if A2 is None else A2.shape[1]
    if K1 % 4 or K2 % 4 or (K1 + K2) < 8 or A1.stride(0) % 4 or (A2 is not None and A2.stride(0) % 4) or W.stride(0) % 4:
        return False
    if A1.data_ptr() % 16 or (A2 is not None and A2.data_ptr() % 16) or W.data_ptr() % 16:
        return False
    return A1.shape[0] >= 1


def linear(A1, W, bias=None, A2=None, row_scale=None, row_vec=None, tab_idx=None, tab=None, scale=None, shift=None,
           activation='identity', out=None, accumulate=False):
    """C = act((cat(A1,A2) @ W^T + row_scale (x) row_vec + tab[tab_idx] + bias) * scale + shift); see gsn_linear_fwd.
    All tensors fp32 CUDA, last dim contiguous; W is [Nout, K1+K2] (nn.Linear layout)."""
    M = (A1 if A1 is not None else (A2 if A2 is not None else tab_idx)).shape[0]
    Nout = W.shape[0] if W is not None else tab.shape[1]
    dev = (W if W is not None else tab).device
    if out is None:
        out = torch.empty((M, Nout), dtype=torch.float32, device=dev)
    p = GsnLinear()
    p.A1, p.A2, p.W = _dp(A1), _dp(A2), _dp(W)
    p.K1 = 0 if A1 is None else A1.shape[1]
    p.K2 = 0 if A2 is None else A2.shape[1]
    p.lda1 = 0 if A1 is None else A1.stride(0)
    p.lda2 = 0 if A2 is None else A2.stride(0)
    p.ldw = 0 if W is None else W.stride(0)
    if W is not None and W.shape[1] != p.K1 + p.K2:
        raise ValueError(f'W has {W.shape[1]} columns, inputs have {p.K1}+{p.K2}')
    p.bias, p.row_scale, p.row_vec = _dp(bias), _dp(row_scale), _dp(row_vec)
    p.tab_idx, p.tab, p.tab_ld = _dp(tab_idx), _dp(tab), (0 if tab is None else tab.stride(0))
    p.scale, p.shift, p.C, p.ldc = _dp(scale), _dp(shift), out.data_ptr(), out.stride(0)
    p.M, p.Nout, p.act, p.accumulate, p.tc_path = M, Nout, ACTIVATIONS[activation], int(accumulate), int(TC_PATH)
    with torch.cuda.device(dev):
        if _tc_eligible(A1, A2, W):
            whi, wlo = split_weight(W)
            ws, nbytes = None, 0
            if TC_PATH == 1 or (p.K2 > 0 and p.K1 % 32 != 0):       # only then the activations need a pre-pass
                nb = ctypes.c_size_t(0)
                _lib.check(_lib.lib().gsn_tc_linear_workspace_bytes(M, p.K1 + p.K2, ctypes.byref(nb)), 'gsn_tc_linear_workspace_bytes')
                ws, nbytes = torch.empty(nb.value, dtype=torch.uint8, device=dev), nb.value
            _lib.call('tc_linear', 'gsn_tc_linear_fwd', ctypes.byref(p), _lib.ptr(whi), _lib.ptr(wlo), _lib.ptr(ws),
                      nbytes, _lib.stream_ptr())
        else:
            _lib.call('linear', 'gsn_linear_fwd', ctypes.byref(p), _lib.stream_ptr())
    return out


def pool_ptr(x, node_ptr, mean=False):
    """sum / mean of the rows of every graph of a PyG batch (node_ptr = batch.ptr, int64 [G+1])"""
    G = node_ptr.numel() - 1
    x = _f32c(x, 'x')
    out = torch.empty((G, x.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call('pool_ptr', 'gsn_pool_ptr', _lib.ptr(x), _lib.ptr(node_ptr), G, x.shape[1], x.stride(0), int(mean),
                  _lib.ptr(out), _lib.stream_ptr())
    return out


def _encode_cols(columns):
    """columns: (int64 view [R], (vocab_begin, vocab_end) | None, table_off[, rows]); rows = number of categories of an
    identity column (0 / absent: unchecked)"""
    cols = (GsnEncodeCol * len(columns))()
    for i, col in enumerate(columns):
        src, vr, off = col[0], col[1], col[2]
        if src.dtype != torch.int64:
            raise ValueError('categorical columns must be int64')
        cols[i].src, cols[i].stride = src.data_ptr(), (src.stride(0) if src.dim() else 1)
        cols[i].vocab_begin, cols[i].vocab_end = (0, 0) if vr is None else (int(vr[0]), int(vr[1]))
        cols[i].table_off = int(off)
        cols[i].rows = int(col[3]) if len(col) > 3 and col[3] else 0
    return cols


def encode_rows(columns, vocab, num_rows, device, perm=None, status=None):
    """columns: list of (int64 tensor view [R] (any stride), (vocab_begin, vocab_end) or None, table_off[, rows]).
    Returns int32 [R, len(columns)] rows into a concatenated embedding table; with perm (int32 [R], e.g.
    EdgePlan.eid) output row r encodes source row perm[r].  status (int32 [1], optional): out-of-range / unseen values."""
    cols = _encode_cols(columns)
    out = torch.empty((num_rows, len(columns)), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.call('encode_rows', 'gsn_encode_rows', ctypes.cast(cols, ctypes.c_void_p), len(columns), _lib.ptr(vocab),
                  _lib.ptr(perm), num_rows, _lib.ptr(out), _lib.ptr(status), _lib.stream_ptr())
    return out


def encode_rows_grouped(columns, groups, mults, n_groups, vocab, num_rows, device, perm=None, status=None):
    """encode_rows with the columns of a group folded into one mixed-radix row index (gsn_encode_rows_grouped):
    out[r, g] = sum_{c in group g} (table_off_c + rank_c * mult_c)"""
    cols = _encode_cols(columns)
    gr = (ctypes.c_int32 * len(columns))(*[int(g) for g in groups])
    mu = (ctypes.c_int32 * len(columns))(*[int(m) for m in mults])
    out = torch.empty((num_rows, n_groups), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.call('encode_rows', 'gsn_encode_rows_grouped', ctypes.cast(cols, ctypes.c_void_p), len(columns),
                  ctypes.cast(gr, ctypes.c_void_p), ctypes.cast(mu, ctypes.c_void_p), n_groups, _lib.ptr(vocab),
                  _lib.ptr(perm), num_rows, _lib.ptr(out), _lib.ptr(status), _lib.stream_ptr())
    return out


def general_edge_idx(plan, dh, P=None, Q=None, node_rows=None, Tn=None, edge_rows=None, Te=None, scale=None, shift=None,
                     activation='relu', edge_rows_csr=False):
    """S[i] = sum_e act((P_i + P_j + sum Tn[node rows] + Q_e + sum Te[edge rows]) * scale + shift)"""
    S = torch.empty((plan.N, dh), dtype=torch.float32, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.call('general_edge', 'gsn_mp_general_edge_idx_fwd', _lib.ptr(plan.rowptr), _lib.ptr(plan.eid),
                  _lib.ptr(plan.nbr), plan.N, plan.E, _lib.ptr(P), _lib.ptr(Q), _lib.ptr(node_rows),
                  0 if node_rows is None else node_rows.shape[1], _lib.ptr(Tn), _lib.ptr(edge_rows),
                  0 if edge_rows is None else edge_rows.shape[1], _lib.ptr(Te), 0 if Te is None else Te.shape[0],
                  int(bool(edge_rows_csr)), dh,
                  _lib.ptr(scale), _lib.ptr(shift),
                  ACTIVATIONS[activation], _lib.ptr(S), _lib.stream_ptr())
    return S


# ----------------------------------------------------------------------------
# differentiable wrappers (training): backward of a segment-sum is a gather,
# backward of a gather is a segment-sum over the grouping by the gather index
# ----------------------------------------------------------------------------
class _SegmentSumFn(torch.autograd.Function):
    """out[i] = sum_{e: key(e)=i} rows[e]   (torch.sparse.sum(...).to_dense(), GSN_sparse.py:143)"""

    @staticmethod
    def forward(ctx, rows, plan):
        ctx.plan = plan
        return segment_sum(plan, rows)

    @staticmethod
    def backward(ctx, grad_out):
        plan = ctx.plan
        return grad_out.contiguous().index_select(0, plan.edge_index[plan.select]), None


class _GatherRowsFn(torch.autograd.Function):
    """rows[e] = x[edge_index[row][e]]; backward = deterministic segment-sum grouped by that index"""

    @staticmethod
    def forward(ctx, x, edge_index, row, num_nodes):
        ctx.args = (edge_index, row, num_nodes)
        return x.index_select(0, edge_index[row])

    @staticmethod
    def backward(ctx, grad_rows):
        edge_index, row, num_nodes = ctx.args
        plan = edge_plan(edge_index, num_nodes, 'target_to_source' if row == 0 else 'source_to_target')
        return segment_sum(plan, grad_rows.contiguous()), None, None, None


def segment_sum_ad(plan: EdgePlan, rows: torch.Tensor) -> torch.Tensor:
    if torch.is_grad_enabled() and rows.requires_grad:
        return _SegmentSumFn.apply(rows, plan)
    return segment_sum(plan, rows)


def gather_rows_ad(x: torch.Tensor, edge_index: torch.Tensor, row: int) -> torch.Tensor:
    if torch.is_grad_enabled() and x.requires_grad:
        return _GatherRowsFn.apply(x, edge_index, row, x.shape[0])
    return x.index_select(0, edge_index[row])


# ----------------------------------------------------------------------------
# training-step operators (ogbg-molhiv recipe): fused ogb message with its own backward, embedding bag
# ----------------------------------------------------------------------------
class _OgbAggregateFn(torch.autograd.Function):
    """(1+eps)*(x [+ id]) + sum_e relu(x_j + id + e_ij)  (GSN_edge_sparse_ogb.py:75-84,119-126) with the fused forward
    kernel and a fused backward that recomputes the relu mask: no [E, d] tensor is saved"""

    @staticmethod
    def forward(ctx, x, identifiers, ef, eps, plan, plan_src, id_per_edge):
        ctx.plan_src, ctx.id_per_edge = plan_src, bool(id_per_edge)
        x, ef = _f32c(x, 'x'), _f32c(ef, 'edge_features')
        idt = _f32c(identifiers, 'identifiers')
        ctx.save_for_backward(x, idt if idt is not None else x.new_empty(0), ef, eps)
        ctx.has_id = idt is not None
        return ogb_aggregate(plan, x, idt, id_per_edge, ef, eps)

    @staticmethod
    def backward(ctx, grad_out):
        x, idt, ef, eps = ctx.saved_tensors
        idt = idt if ctx.has_id else None
        p = ctx.plan_src
        g = _f32c(grad_out, 'grad_out')
        d = x.shape[1]
        grad_x = torch.empty_like(x)
        grad_ef = torch.empty_like(ef)
        with torch.cuda.device(x.device):
            _lib.call('mp_ogb_bwd', 'gsn_mp_ogb_bwd', _lib.ptr(p.rowptr), _lib.ptr(p.eid), _lib.ptr(p.nbr), p.N, p.E, _lib.ptr(x),
                      _lib.ptr(idt), int(ctx.id_per_edge), _lib.ptr(ef), d, _lib.ptr(eps), _lib.ptr(g), _lib.ptr(grad_x),
                      _lib.ptr(grad_ef), _lib.stream_ptr())
        grad_id = None
        if idt is not None:
            grad_id = grad_ef if ctx.id_per_edge else grad_x
        return grad_x, grad_id, grad_ef, None, None, None, None


def ogb_aggregate_ad(edge_index, num_nodes, flow, x, identifiers, id_per_edge, ef, eps):
    """differentiable ogb message (w.r.t. x, identifiers, edge features; eps must not require grad)"""
    plan = edge_plan(edge_index, num_nodes, flow)
    plan_src = edge_plan(edge_index, num_nodes, 'target_to_source' if plan.select == 1 else 'source_to_target')
    return _OgbAggregateFn.apply(x, identifiers, ef, eps, plan, plan_src, id_per_edge)


class GsnBagCol(ctypes.Structure):
    """ctypes image of `struct GsnBagCol`."""
    _fields_ = [('table', ctypes.c_void_p), ('rows', ctypes.c_int32), ('_pad', ctypes.c_int32)]


MAX_BAG_COLS = 96
_bag_status = {}


def _bag_cols(tables):
    cols = (GsnBagCol * len(tables))()
    for i, t in enumerate(tables):
        cols[i].table, cols[i].rows = t.data_ptr(), int(t.shape[0])
    return cols


def _bag_status_word(dev):
    st = _bag_status.get(dev)
    if st is None:
        st = _bag_status[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    return st


class _EmbeddingBagFn(torch.autograd.Function):
    """sum_c table_c[idx[:, c]] (multi_embedding aggr 'sum', AtomEncoder, BondEncoder) in one launch each way"""

    @staticmethod
    def forward(ctx, idx, *tables):
        ctx.save_for_backward(idx)
        ctx.shapes = [tuple(t.shape) for t in tables]
        ws = [_f32c(t.detach(), 'embedding table') for t in tables]
        R, d = int(idx.shape[0]), int(ws[0].shape[1])
        out = torch.empty((R, d), dtype=torch.float32, device=idx.device)
        with torch.cuda.device(idx.device):
            _lib.call('embedding_bag', 'gsn_embedding_bag_fwd', ctypes.cast(_bag_cols(ws), ctypes.c_void_p), len(ws), _lib.ptr(idx),
                      idx.stride(0), R, d, _lib.ptr(out), _lib.ptr(_bag_status_word(idx.device)), _lib.stream_ptr())
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = _f32c(grad_out, 'grad_out')
        grads = [torch.zeros(s, dtype=torch.float32, device=g.device) for s in ctx.shapes]
        R, d = int(idx.shape[0]), int(g.shape[1])
        with torch.cuda.device(g.device):
            _lib.call('embedding_bag_bwd', 'gsn_embedding_bag_bwd', ctypes.cast(_bag_cols(grads), ctypes.c_void_p), len(grads),
                      _lib.ptr(idx), idx.stride(0), R, d, _lib.ptr(g), _lib.ptr(_bag_status_word(g.device)), _lib.stream_ptr())
        return (None,) + tuple(grads)


EMBEDDING_BAG = True        # False: per-column nn.Embedding lookups (measurement aid, see graph_filters/autograd.py)


def embedding_bag_ok(idx, tables) -> bool:
    return (EMBEDDING_BAG and idx.is_cuda and idx.dtype == torch.int64 and idx.dim() == 2 and idx.stride(1) == 1 and 1 <= len(tables) <= MAX_BAG_COLS
            and idx.shape[1] == len(tables) and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                                                    and t.shape[1] == tables[0].shape[1] for t in tables))


def embedding_bag(idx, tables):
    """idx int64 [R, C] (row stride arbitrary), tables: C weight matrices [rows_c, d] -> fp32 [R, d]"""
    return _EmbeddingBagFn.apply(idx, *tables)


class _TcLinearFn(torch.autograd.Function):
    """y = x W^T + b with the forward and the input gradient on the tcgen05 3xTF32 kernel (both are row-parallel GEMMs
    over the N or E rows); the weight gradient g^T x reduces over those rows and stays a library SGEMM (fp32)"""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return linear(x, W, bias=b)

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = _f32c(g, 'grad_out')
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            Wt = W.detach().t().contiguous()               # [in, out]: dX = g W  ==  linear(g, weight = W^T)
            gx = linear(g, Wt) if _tc_eligible(g, None, Wt) else g @ W
        if ctx.needs_input_grad[1]:
            gw = g.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g.sum(0)
        return gx, gw, gb


TC_TRAINING = True          # False: nn.Linear (cuBLAS fp32 SGEMM) in the training path


def linear_ad(x, W, b):
    """differentiable Linear: tensor-core forward / input gradient when the shapes allow, else torch"""
    if (TC_TRAINING and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[0] >= 128
            and _tc_eligible(x if x.is_contiguous() else x.contiguous(), None, W)):
        return _TcLinearFn.apply(x.contiguous(), W, b)
    return torch.nn.functional.linear(x, W, b)
