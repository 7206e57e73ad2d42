"""ctypes front-end of oracle/count_enum.c (TEST INFRASTRUCTURE ONLY).

Users: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference arm.  gsn_b200/ never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'count_enum.c')
_OUT_DIR = os.path.join(_HERE, '_build')
_SO = os.path.join(_OUT_DIR, 'libgsn_oracle.so')

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(['gcc', '-O3', '-march=x86-64-v2', '-fopenmp', '-shared', '-fPIC',
                               _SRC, '-o', _SO])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.gsn_oracle_aut_count.restype = ctypes.c_int64
        L.gsn_oracle_aut_count.argtypes = [ctypes.c_int, ctypes.c_int, _i32p]
        L.gsn_oracle_count_graph.restype = ctypes.c_int
        L.gsn_oracle_count_graph.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _i64p, _i64p,
                                             ctypes.c_int, ctypes.c_int, _i32p, _i32p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int64, _f64p]
        L.gsn_oracle_count_batch.restype = ctypes.c_int
        L.gsn_oracle_count_batch.argtypes = [ctypes.c_int64, _i64p, _i64p, _i64p, _i64p,
                                             ctypes.c_int, ctypes.c_int, _i32p, _i32p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int64, _f64p, ctypes.c_int]
        L.gsn_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _pattern_arrays(subgraph_dict, scope):
    """Directed, coalesced pattern edge list (utils_graph_processing.py:74,147)
    and the orbit table in the layout count_enum.c wants."""
    from .count_vf2 import to_undirected
    und = np.ascontiguousarray(to_undirected(subgraph_dict['subgraph'].get_edges().T).T.astype(np.int32))
    k = int(subgraph_dict['subgraph'].num_vertices())
    memb = subgraph_dict['orbit_membership']
    n_orb = len(subgraph_dict['orbit_partition'])
    size = k if scope == 0 else und.shape[0]
    orbit = np.array([memb[i] for i in range(size)], dtype=np.int32)
    return k, und, orbit, n_orb


def aut_count(edge_list) -> int:
    from .count_vf2 import to_undirected
    el = np.asarray(edge_list, dtype=np.int64).reshape(-1, 2)
    und = np.ascontiguousarray(to_undirected(el.T).T.astype(np.int32))
    k = int(el.max()) + 1
    return int(lib().gsn_oracle_aut_count(k, und.shape[0], _p(und, _i32p)))


def count_graph(edge_index, subgraph_dict, induced, num_nodes, scope):
    """scope 0 -> subgraph_isomorphism_vertex_counts, 1 -> ..._edge_counts;
    returns float64 like the reference's count_fn."""
    ei = np.ascontiguousarray(np.asarray(edge_index, dtype=np.int64).reshape(2, -1))
    E = ei.shape[1]
    k, und, orbit, n_orb = _pattern_arrays(subgraph_dict, scope)
    rows = num_nodes if scope == 0 else E
    out = np.zeros((rows, n_orb), dtype=np.float64)
    src, dst = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
    rc = lib().gsn_oracle_count_graph(num_nodes, 0, E, _p(src, _i64p), _p(dst, _i64p), k, und.shape[0],
                                      _p(und, _i32p), _p(orbit, _i32p), n_orb, int(induced), scope,
                                      int(subgraph_dict['aut_count']), _p(out, _f64p))
    if rc == -1:
        raise KeyError('mapped edge missing from edge_index (asymmetric input)')
    if rc:
        raise RuntimeError(f'oracle error {rc}')
    return out


def count_batch(node_ptr, edge_ptr, edge_index, subgraph_dicts, induced, scope, nthreads=0):
    """Whole batch, all patterns; returns int64 identifiers [N|E, sum orbits]
    (the per-graph loop of utils_data_gen.py:60-78 + utils_ids.py:19-27)."""
    node_ptr = np.ascontiguousarray(node_ptr, dtype=np.int64)
    edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int64)
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    src, dst = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
    rows = int(node_ptr[-1]) if scope == 0 else ei.shape[1]
    cols = []
    for sd in subgraph_dicts:
        k, und, orbit, n_orb = _pattern_arrays(sd, scope)
        out = np.zeros((rows, n_orb), dtype=np.float64)
        rc = lib().gsn_oracle_count_batch(len(node_ptr) - 1, _p(node_ptr, _i64p), _p(edge_ptr, _i64p),
                                          _p(src, _i64p), _p(dst, _i64p), k, und.shape[0], _p(und, _i32p),
                                          _p(orbit, _i32p), n_orb, int(induced), scope,
                                          int(sd['aut_count']), _p(out, _f64p), int(nthreads))
        if rc == -1:
            raise KeyError('mapped edge missing from edge_index (asymmetric input)')
        if rc:
            raise RuntimeError(f'oracle error {rc}')
        cols.append(out)
    return np.concatenate(cols, 1).astype(np.int64)


def max_threads() -> int:
    return int(lib().gsn_oracle_max_threads())
