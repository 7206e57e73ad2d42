"""Pins the COUNT oracles (test infrastructure) to the reference's own artifact and
to each other.  No GPU needed."""
import numpy as np
import pytest

from oracle import count_c, count_vf2


def _imdb_graph(f, g):
    e0, e1 = f['edge_ptr'][g], f['edge_ptr'][g + 1]
    return f['edge_index'][:, e0:e1], int(f['node_ptr'][g + 1] - f['node_ptr'][g]), f['identifiers'][e0:e1]


def test_c_oracle_reproduces_graph_tool_fixture(imdb_fixture):
    """edge scope, non-induced, complete_graph k=3..5 on IMDB-BINARY: the only configuration the
    reference pins with an artifact (datasets/social/IMDBBINARY/processed/local/complete_graph_5.pt).
    K3/K4 on every graph; K5 on the 700 graphs with the fewest edges (the rest costs minutes of
    all-maps enumeration on the CPU; the CUDA path is checked on all 1000 in test_count_gpu.py)."""
    f = imdb_fixture
    sds = count_vf2.make_subgraph_dicts(count_vf2.pattern_edge_lists('complete_graph', 5), 'local')
    assert [sd['aut_count'] for sd in sds] == [6, 24, 120]
    ei = f['edge_index'].copy()
    G = len(f['node_ptr']) - 1
    for g in range(G):
        ei[:, f['edge_ptr'][g]:f['edge_ptr'][g + 1]] += f['node_ptr'][g]
    got = count_c.count_batch(f['node_ptr'], f['edge_ptr'], ei, sds[:2], False, 1)
    assert np.array_equal(got, f['identifiers'][:, :2])
    order = np.argsort(f['edge_ptr'][1:] - f['edge_ptr'][:-1])[:700]
    for g in order:
        e, n, ids = _imdb_graph(f, int(g))
        out = count_c.count_graph(e, sds[2], False, n, 1)
        assert np.array_equal(out[:, 0].astype(np.int64), ids[:, 2]), g


def test_vf2_oracle_reproduces_graph_tool_fixture(imdb_fixture):
    f = imdb_fixture
    sds = count_vf2.make_subgraph_dicts(count_vf2.pattern_edge_lists('complete_graph', 5), 'local')
    order = np.argsort(f['edge_ptr'][1:] - f['edge_ptr'][:-1])
    for g in list(order[:12]) + [int(order[300])]:
        e, n, ids = _imdb_graph(f, int(g))
        _, _, got = count_vf2.subgraph_counts2ids(count_vf2.subgraph_isomorphism_edge_counts, e, n, sds, False)
        assert np.array_equal(got, ids)


def test_fixture_totals(imdb_fixture):
    """BASELINE.md section 2"""
    ids = imdb_fixture['identifiers']
    assert ids.shape == (193062, 3)
    assert ids.sum(0).tolist() == [2351946, 20334156, 141614300]
    assert ids.max(0).tolist() == [85, 678, 5576]


@pytest.mark.parametrize('family,kmax', [('cycle_graph', 6), ('path_graph', 4), ('star_graph', 3), ('complete_graph', 4)])
@pytest.mark.parametrize('induced', [False, True])
def test_c_and_vf2_oracles_agree(family, kmax, induced):
    rng = np.random.default_rng(1)
    from tests.util import random_graph
    els = count_vf2.pattern_edge_lists(family, kmax)
    for scope_name, scope, fn in (('global', 0, count_vf2.subgraph_isomorphism_vertex_counts),
                                  ('local', 1, count_vf2.subgraph_isomorphism_edge_counts)):
        sds = count_vf2.make_subgraph_dicts(els, scope_name)
        for n, p in ((7, 0.6), (11, 0.35)):
            ei = random_graph(rng, n, p)
            for sd in sds:
                a = fn(ei, subgraph_dict=sd, induced=induced, num_nodes=n)
                b = count_c.count_graph(ei, sd, induced, n, scope)
                assert np.array_equal(a, b)


def test_sr25_known_answers_c_oracle(sr_fixture):
    """SURVEY A.4 on two graphs (the CUDA test covers all 15)"""
    sds = count_vf2.make_subgraph_dicts(count_vf2.pattern_edge_lists('cycle_graph', 5), 'local')
    for g in (0, 1):
        ids = np.concatenate([count_c.count_graph(sr_fixture[g], sd, True, 25, 1) for sd in sds], 1)
        assert (ids[:, 0] == 5).all()
        if g == 1:
            assert (ids[:, 1] == 16).all() and (ids[:, 2] == 40).all()


def test_aut_counts_and_orbit_totals(graphlet_patterns):
    """|Aut| and orbit counts of the all_simple_graphs families"""
    import networkx as nx
    assert count_c.aut_count(list(nx.cycle_graph(6).edges)) == 12
    assert count_c.aut_count(list(nx.path_graph(4).edges)) == 2
    assert count_c.aut_count(list(nx.complete_graph(5).edges)) == 120
    assert count_c.aut_count(list(nx.star_graph(4).edges)) == 24
    assert count_c.aut_count(list(nx.diamond_graph().edges)) == 4
    assert {k: len(v) for k, v in graphlet_patterns.items()} == {2: 1, 3: 2, 4: 6, 5: 21, 6: 112}
    tot_v = {k: sum(len(count_vf2.automorphism_orbits(el)[1]) for el in graphlet_patterns[k]) for k in (3, 4, 5)}
    tot_e = {k: sum(len(count_vf2.induced_edge_automorphism_orbits(el)[1]) for el in graphlet_patterns[k]) for k in (3, 4, 5)}
    assert tot_v == {3: 3, 4: 11, 5: 58} and tot_e == {3: 2, 4: 10, 5: 56}
