"""Host logic of the COUNT path: orbit numbering, plan compilation, and the
__host__ __device__ enumeration cores run on the CPU (tests/host_sim) against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from gsn_b200 import patterns
from oracle import count_c, count_vf2
from tests.util import random_graph

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def host_sim():
    src = os.path.join(HERE, 'host_sim', 'host_sim.cpp')
    out = os.path.join(HERE, 'host_sim', '_build')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libhost_sim.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', src, '-o', so])
    L = ctypes.CDLL(so)
    i64p = ctypes.POINTER(ctypes.c_int64)

    def run(ei, n, plans, rows, ld, parts=1):
        L.gsn_host_sim_set_parts(parts)
        ei = np.ascontiguousarray(ei, dtype=np.int64)
        src_, dst_ = np.ascontiguousarray(ei[0]), np.ascontiguousarray(ei[1])
        o = np.zeros((rows, ld), dtype=np.int64)
        for P in plans:
            rc = L.gsn_host_sim_count(n, ei.shape[1], src_.ctypes.data_as(i64p), dst_.ctypes.data_as(i64p),
                                      ctypes.byref(P), o.ctypes.data_as(i64p), ld)
            assert rc == 0
        return o
    return run


def _all_patterns(graphlet_patterns):
    pats = [el for k in (2, 3, 4, 5, 6) for el in graphlet_patterns[k]]
    for fam, kk in (('cycle_graph', 10), ('path_graph', 7), ('complete_graph', 7), ('star_graph', 6), ('binomial_tree', 3)):
        pats += count_vf2.pattern_edge_lists(fam, kk)
    return pats


def test_orbits_match_reference_numbering(graphlet_patterns):
    """automorphism_orbits / induced_edge_automorphism_orbits: same partition, membership and
    aut_count as the restated reference (which enumerates the full group, utils_graph_processing.py:22)"""
    for el in _all_patterns(graphlet_patterns):
        for scope in ('global', 'local'):
            a = count_vf2.make_subgraph_dicts([el], scope)[0]
            b = patterns.make_subgraph_dicts([el], scope)[0]
            assert a['orbit_membership'] == b['orbit_membership'], el
            assert a['orbit_partition'] == b['orbit_partition'], el
            assert a['aut_count'] == b['aut_count'], el
            assert np.array_equal(a['subgraph'].get_edges(), b['subgraph'].get_edges())


def test_directed_orbits_flag():
    el = [(0, 1), (1, 2), (2, 3)]
    a = count_vf2.induced_edge_automorphism_orbits(el, directed_orbits=True)
    b = patterns.induced_edge_automorphism_orbits(el, directed_orbits=True, print_msgs=False)
    assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3]


def test_large_groups_without_enumeration():
    import math
    import networkx as nx
    assert patterns.automorphism_orbits(list(nx.complete_graph(12).edges), print_msgs=False)[3] == math.factorial(12)
    assert patterns.automorphism_orbits(list(nx.star_graph(11).edges), print_msgs=False)[3] == math.factorial(11)
    assert patterns.automorphism_orbits(list(nx.cycle_graph(12).edges), print_msgs=False)[3] == 24


def test_plan_struct_layout_matches_header():
    assert ctypes.sizeof(patterns.GsnPlan) == 8 * 4 + 3 * 16 * 4 + 16 + 2 * 16 * 16


def test_family_fusion():
    sds = patterns.make_subgraph_dicts(count_vf2.pattern_edge_lists('cycle_graph', 8), 'local')
    plans = patterns.compile_plans(sds, False, 1)
    assert len(plans) == 1 and plans[0].family == patterns.FAMILY_CYCLES and (plans[0].kmin, plans[0].kmax) == (3, 8)
    sds = patterns.make_subgraph_dicts(count_vf2.pattern_edge_lists('complete_graph', 5), 'global')
    plans = patterns.compile_plans(sds, False, 0)
    assert len(plans) == 1 and plans[0].family == patterns.FAMILY_CLIQUES and plans[0].n_cols == 3
    sds = patterns.make_subgraph_dicts(count_vf2.pattern_edge_lists('path_graph', 5), 'global')
    plans = patterns.compile_plans(sds, False, 0)
    assert [p.family for p in plans] == [0, 0, 0] and [p.col0 for p in plans] == [0, 2, 4]


def test_graph6_reader(graphlet_patterns):
    # "D?{" is the first line of datasets/all_simple_graphs/graph5c.g6
    n, edges = patterns.parse_graph6('D?{')
    assert n == 5 and sorted(edges) == sorted(map(tuple, graphlet_patterns[5][0]))


FAMS = {'cycles8': ('cycle_graph', 8), 'cliques5': ('complete_graph', 5), 'paths5': ('path_graph', 5),
        'stars4': ('star_graph', 4)}


@pytest.mark.parametrize('family', list(FAMS) + ['graphlets5'])
@pytest.mark.parametrize('scope_name', ['global', 'local'])
@pytest.mark.parametrize('induced', [False, True])
def test_enumeration_cores_vs_oracle(host_sim, family, scope_name, induced, graphlet_patterns):
    """symmetry-broken enumerate-once == all maps / |Aut| (utils_graph_processing.py:127,175), incl.
    the fused cycle / clique families, W = 1 and W = 2 words per adjacency row"""
    rng = np.random.default_rng(abs(hash((family, scope_name, induced))) % 2**32)
    els = (graphlet_patterns[3] + graphlet_patterns[4] + graphlet_patterns[5]) if family == 'graphlets5' \
        else count_vf2.pattern_edge_lists(*FAMS[family])
    scope = 1 if scope_name == 'local' else 0
    sds = patterns.make_subgraph_dicts(els, scope_name)
    osds = count_vf2.make_subgraph_dicts(els, scope_name)
    ld = patterns.total_columns(sds)
    for fuse in (True, False):
        plans = patterns.compile_plans(sds, induced, scope, fuse_families=fuse)
        for n, p in ((8, 0.5), (14, 0.3), (24, 0.12), (70, 0.05)):
            ei = random_graph(rng, n, p)
            if ei.shape[1] == 0:
                continue
            rows = n if scope == 0 else ei.shape[1]
            exp = np.concatenate([count_c.count_graph(ei, sd, induced, n, scope) for sd in osds], 1).astype(np.int64)
            for parts in (1, 4):        # sub-items per directed edge (small-batch work splitting)
                got = host_sim(ei, n, plans, rows, ld, parts)
                assert np.array_equal(got, exp), (family, scope_name, induced, fuse, n, parts)
