"""ctypes binding of libgsn_b200.so (include/gsn_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call
fails, the product path raises.  PyTorch is only used for device memory and
streams; every tensor crosses the boundary as a raw pointer.
"""
from __future__ import annotations

import ctypes
import os

import torch

from .patterns import GsnPlan

# GSN_B200_LIB: developer override (e.g. the profiling build libgsn_b200_prof.so); the package itself never sets it
_SO = os.environ.get('GSN_B200_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libgsn_b200.so')

GSN_OK = 0
_ERRORS = {-1: 'GSN_E_INVALID (bad argument)', -2: 'GSN_E_UNSUPPORTED (shape outside the built kernels)',
           -3: 'GSN_E_WORKSPACE (workspace too small)', -4: 'GSN_E_CUDA'}

S_GRAPH_TOO_LARGE, S_CROSS_GRAPH_EDGE, S_MISSING_EDGE, S_INDEX_RANGE, S_COUNT_OVERFLOW, S_NOT_GROUPED = 1, 2, 4, 8, 16, 32
S_UNSEEN_VALUE = 64          # informative: an identifier value outside the fitted one_hot_unique vocabulary
S_FATAL = 63

_vp, _i64, _i32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t
_szp = ctypes.POINTER(ctypes.c_size_t)
_planp = ctypes.POINTER(GsnPlan)

_SIGNATURES = {
    'gsn_abi_version': (ctypes.c_int, []),
    'gsn_last_cuda_error': (ctypes.c_char_p, []),
    'gsn_launch_count': (ctypes.c_uint64, []),
    'gsn_graph_workspace_bytes': (ctypes.c_int, [_i64, _i64, _i32, _szp]),
    'gsn_graph_build': (ctypes.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _sz, _vp, _vp]),
    'gsn_count_scratch_bytes': (ctypes.c_int, [_i64, _i64, _planp, _szp]),
    'gsn_count_pattern': (ctypes.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _planp, _vp, _i64, _vp, _sz, _vp, _vp]),
    'gsn_count_small': (ctypes.c_int, [_vp, _i64, _vp, _i64, _i64, _planp, _vp, _i64, _vp, _vp]),
    'gsn_csr_workspace_bytes': (ctypes.c_int, [_i64, _i64, _szp]),
    'gsn_csr_build': (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    'gsn_mp_gin_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp, _vp]),
    'gsn_mp_ogb_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    'gsn_mp_segment_sum': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _vp, _vp]),
    'gsn_mp_general_edge_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    'gsn_mp_general_edge_idx_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    'gsn_linear_fwd': (ctypes.c_int, [_vp, _vp]),
    'gsn_split_tf32': (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'gsn_tc_linear_workspace_bytes': (ctypes.c_int, [_i64, _i32, _szp]),
    'gsn_tc_linear_fwd': (ctypes.c_int, [_vp, _vp, _vp, _vp, _sz, _vp]),
    'gsn_embedding_bag_fwd': (ctypes.c_int, [_vp, _i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    'gsn_embedding_bag_bwd': (ctypes.c_int, [_vp, _i32, _vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    'gsn_mp_ogb_bwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    'gsn_fused_model_fwd': (ctypes.c_int, [_vp, _vp]),
    'gsn_tile_plan': (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int32, _vp, _vp]),
    'gsn_submit_step': (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_size_t, _vp, _vp, ctypes.c_size_t, _vp]),
    'gsn_pool_ptr': (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    'gsn_encode_rows': (ctypes.c_int, [_vp, _i32, _vp, _vp, _i64, _vp, _vp, _vp]),
    'gsn_encode_rows_grouped': (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _vp, _vp]),
    'gsn_dgn_aggregate_fwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32,
                                             ctypes.c_float, _vp, _vp, _vp]),
    'gsn_dgn_aggregate_bwd': (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32,
                                             ctypes.c_float, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def so_path() -> str:
    return _SO


def lib():
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                f'{_SO} is missing: build the CUDA extension first (python -m gsn_b200.build). '
                'gsn_b200 has no CPU or PyTorch fallback.')
        L = ctypes.CDLL(_SO)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.gsn_abi_version() != 1:
            raise RuntimeError('libgsn_b200.so ABI version mismatch; rebuild')
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().gsn_launch_count())


# optional per-call device timing (bench.py): TIMER = list -> (name, start_event, end_event) appended per call
TIMER = None


def call(tag: str, fname: str, *args):
    """invoke one C-ABI entry point, raise on a non-zero return; when TIMER is a
    list, bracket the call with CUDA events on the current stream"""
    fn = getattr(lib(), fname)
    if TIMER is None:
        check(fn(*args), fname)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    TIMER.append((tag, e0, e1))
    check(rc, fname)


def check(rc: int, what: str):
    if rc != GSN_OK:
        msg = _ERRORS.get(rc, str(rc))
        if rc == -4:
            msg += ': ' + (lib().gsn_last_cuda_error() or b'').decode()
        raise RuntimeError(f'{what} failed: {msg}')


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: gsn_b200 has no CPU path')


def status_message(bits: int) -> str:
    out = []
    if bits & S_GRAPH_TOO_LARGE:
        out.append('a graph has more vertices than the adjacency bitmask width allows')
    if bits & S_CROSS_GRAPH_EDGE:
        out.append('an edge joins two different graphs of the batch')
    if bits & S_MISSING_EDGE:
        out.append('a match used an edge (a,b) that is not a column of edge_index (asymmetric edge_index)')
    if bits & S_INDEX_RANGE:
        out.append('a node index is outside [0, num_nodes)')
    if bits & S_COUNT_OVERFLOW:
        out.append('a per-vertex / per-edge occurrence count exceeded 2^32 - 1')
    if bits & S_NOT_GROUPED:
        out.append('edge_index columns are not grouped by graph (batch order)')
    if bits & S_UNSEEN_VALUE:
        out.append('an identifier value is not in the fitted vocabulary (encoded as the next larger known value)')
    return '; '.join(out)
