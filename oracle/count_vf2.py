"""COUNT oracle (TEST INFRASTRUCTURE ONLY -- never imported by gsn_b200/).

CPU restatement of the reference's subgraph-isomorphism counting path,
/root/reference/utils_graph_processing.py:10-179 and utils_ids.py:7-29, with
networkx's VF2 (`GraphMatcher.subgraph_monomorphisms_iter` for induced=False,
`subgraph_isomorphisms_iter` for induced=True) standing in for graph-tool's
`subgraph_isomorphism` (graph-tool: conda-forge, version unpinned by the
reference README.md:39, not vendored under /root/reference and not importable
in this image).  Both libraries enumerate *every* injective edge-preserving map
H -> G, so the integer counts are identical by definition.

Parity pin: this file is checked against the reference's own shipped
graph-tool output (datasets/social/IMDBBINARY/processed/local/
complete_graph_5.pt, re-packed as tests/golden/imdb_k5_edge_counts.npz by
scripts/make_golden.py) in tests/test_oracle_count.py.  Every configuration
other than "edge scope, non-induced, complete_graph k=3..5" is *unpinned by a
reference artifact* and is pinned instead by the invariants / known answers of
SURVEY.md A.4.

Only numpy + networkx; no torch.
"""
from __future__ import annotations

import numpy as np
import networkx as nx
from networkx.algorithms.isomorphism import GraphMatcher


# --------------------------------------------------------------------------
# third-party semantics restated (PyG is absent): SURVEY.md A.5
# --------------------------------------------------------------------------
def remove_self_loops(edge_index, edge_attr=None):
    """torch_geometric.utils.remove_self_loops (utils_ids.py:11-15)."""
    edge_index = np.asarray(edge_index)
    mask = edge_index[0] != edge_index[1]
    ea = None if edge_attr is None else np.asarray(edge_attr)[mask]
    return edge_index[:, mask], ea


def to_undirected(edge_index):
    """torch_geometric.utils.to_undirected: add reverse edges, then coalesce
    (lexicographic sort by (row, col), duplicates dropped).  Used at
    utils_graph_processing.py:74,147 and utils_data_prep.py:207."""
    edge_index = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    row = np.concatenate([edge_index[0], edge_index[1]])
    col = np.concatenate([edge_index[1], edge_index[0]])
    if row.size == 0:
        return np.zeros((2, 0), dtype=np.int64)
    n = int(max(row.max(), col.max())) + 1
    key = np.unique(row * n + col)
    return np.stack([key // n, key % n])


def _simple_graph(edge_list, num_vertices=None):
    """gt.Graph(directed=False); add_edge_list; remove_self_loops;
    remove_parallel_edges  (utils_graph_processing.py:16-19, 110-113, 150-153).
    Vertices are 0..max id (graph-tool creates every index up to the maximum)."""
    edge_list = [(int(a), int(b)) for a, b in edge_list]
    n = (max(max(a, b) for a, b in edge_list) + 1) if edge_list else 0
    if num_vertices is not None:
        n = max(n, num_vertices)
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from((a, b) for a, b in edge_list if a != b)
    return g


class PatternGraph:
    """Minimal stand-in for the gt.Graph stored in subgraph_dict['subgraph']:
    the only method the reference calls on it is .get_edges()
    (utils_graph_processing.py:74,147)."""

    def __init__(self, g: nx.Graph):
        self.g = g

    def get_edges(self):
        return np.array(sorted(self.g.edges()), dtype=np.int64).reshape(-1, 2)

    def num_vertices(self):
        return self.g.number_of_nodes()


# --------------------------------------------------------------------------
# utils_graph_processing.py:10-56
# --------------------------------------------------------------------------
def automorphism_orbits(edge_list, print_msgs=False, **kwargs):
    directed = kwargs.get('directed', False)
    if directed:
        raise NotImplementedError('directed patterns: see SURVEY.md A.3')
    graph = _simple_graph(edge_list)

    # :22  all automorphisms = all monomorphisms H -> H
    gm = GraphMatcher(graph, graph)
    aut_group = []
    n = graph.number_of_nodes()
    for m in gm.subgraph_monomorphisms_iter():
        # networkx yields {G_node: H_node}; graph-tool yields an array indexed
        # by pattern vertex holding the target vertex.
        inv = {h: g for g, h in m.items()}
        aut_group.append([inv[i] for i in range(n)])

    # :24-32  orbit id = min vertex that can be mapped onto it
    orbit_membership = {v: v for v in range(n)}
    for aut in aut_group:
        for original, vertex in enumerate(aut):
            role = min(original, orbit_membership[vertex])
            orbit_membership[vertex] = role

    # :34-42 make contiguous
    verts = list(orbit_membership.keys())
    roles = [orbit_membership[v] for v in verts]
    _, contiguous = np.unique(roles, return_inverse=True)
    orbit_membership = {v: int(contiguous[i]) for i, v in enumerate(verts)}

    orbit_partition = {}
    for vertex, orbit in orbit_membership.items():
        orbit_partition.setdefault(orbit, []).append(vertex)

    aut_count = len(aut_group)
    return PatternGraph(graph), orbit_partition, orbit_membership, aut_count


# --------------------------------------------------------------------------
# utils_graph_processing.py:58-100
# --------------------------------------------------------------------------
def induced_edge_automorphism_orbits(edge_list, **kwargs):
    directed = kwargs.get('directed', False)
    directed_orbits = kwargs.get('directed_orbits', False)
    graph, orbit_partition, orbit_membership, aut_count = automorphism_orbits(
        edge_list=edge_list, directed=directed, print_msgs=False)

    edge_orbit_partition, edge_orbit_membership, edge_orbits2inds = {}, {}, {}
    ind = 0
    # :73-74
    edges = to_undirected(graph.get_edges().T).T.tolist()
    for i, edge in enumerate(edges):
        if directed_orbits:
            edge_orbit = (orbit_membership[edge[0]], orbit_membership[edge[1]])
        else:
            edge_orbit = frozenset([orbit_membership[edge[0]], orbit_membership[edge[1]]])
        if edge_orbit not in edge_orbits2inds:
            edge_orbits2inds[edge_orbit] = ind
            ind_edge_orbit = ind
            ind += 1
        else:
            ind_edge_orbit = edge_orbits2inds[edge_orbit]
        edge_orbit_partition.setdefault(ind_edge_orbit, []).append(tuple(edge))
        edge_orbit_membership[i] = ind_edge_orbit
    return graph, edge_orbit_partition, edge_orbit_membership, aut_count


def _all_maps(pattern: nx.Graph, G: nx.Graph, induced: bool):
    """graph-tool subgraph_isomorphism(sub, g, induced, subgraph=True,
    generator=True): yields arrays map[pattern vertex] = target vertex."""
    gm = GraphMatcher(G, pattern)
    it = gm.subgraph_isomorphisms_iter() if induced else gm.subgraph_monomorphisms_iter()
    k = pattern.number_of_nodes()
    for m in it:
        arr = [0] * k
        for gnode, hnode in m.items():
            arr[hnode] = gnode
        yield arr


# --------------------------------------------------------------------------
# utils_graph_processing.py:103-131
# --------------------------------------------------------------------------
def subgraph_isomorphism_vertex_counts(edge_index, **kwargs):
    subgraph_dict, induced, num_nodes = kwargs['subgraph_dict'], kwargs['induced'], kwargs['num_nodes']
    edge_index = np.asarray(edge_index)
    G = _simple_graph(edge_index.T.tolist())
    counts = np.zeros((num_nodes, len(subgraph_dict['orbit_partition'])))
    for sub_iso_curr in _all_maps(subgraph_dict['subgraph'].g, G, induced):
        for i, node in enumerate(sub_iso_curr):
            counts[node, subgraph_dict['orbit_membership'][i]] += 1
    return counts / subgraph_dict['aut_count']


# --------------------------------------------------------------------------
# utils_graph_processing.py:134-179
# --------------------------------------------------------------------------
def subgraph_isomorphism_edge_counts(edge_index, **kwargs):
    subgraph_dict, induced = kwargs['subgraph_dict'], kwargs['induced']
    edge_index = np.asarray(edge_index).T
    edge_dict = {}
    for i, edge in enumerate(edge_index):
        edge_dict[(int(edge[0]), int(edge[1]))] = i          # :142-144 last wins
    subgraph_edges = to_undirected(subgraph_dict['subgraph'].get_edges().T).T.tolist()   # :147
    G = _simple_graph(edge_index.tolist())
    counts = np.zeros((edge_index.shape[0], len(subgraph_dict['orbit_partition'])))
    for mapping in _all_maps(subgraph_dict['subgraph'].g, G, induced):
        for i, edge in enumerate(subgraph_edges):
            edge_orbit = subgraph_dict['orbit_membership'][i]
            mapped_edge = (mapping[edge[0]], mapping[edge[1]])
            counts[edge_dict[mapped_edge], edge_orbit] += 1   # KeyError if asymmetric (A.3)
    return counts / subgraph_dict['aut_count']


# --------------------------------------------------------------------------
# utils_data_gen.py:31-42 (pattern set-up loop) and utils_ids.py:7-29
# --------------------------------------------------------------------------
def make_subgraph_dicts(edge_lists, id_scope, directed_orbits=False):
    fn = induced_edge_automorphism_orbits if id_scope == 'local' else automorphism_orbits
    dicts = []
    for el in edge_lists:
        subgraph, part, memb, aut = fn(edge_list=el, directed=False, directed_orbits=directed_orbits)
        dicts.append({'subgraph': subgraph, 'orbit_partition': part,
                      'orbit_membership': memb, 'aut_count': aut})
    return dicts


def subgraph_counts2ids(count_fn, edge_index, num_nodes, subgraph_dicts, induced, edge_features=None):
    """utils_ids.py:7-29 on plain arrays.  Returns (edge_index, edge_features,
    identifiers int64)."""
    edge_index, edge_features = remove_self_loops(edge_index, edge_features)
    identifiers = None
    for sd in subgraph_dicts:
        counts = count_fn(edge_index, subgraph_dict=sd, induced=induced,
                          num_nodes=num_nodes, directed=False)
        identifiers = counts if identifiers is None else np.concatenate([identifiers, counts], 1)
    return edge_index, edge_features, identifiers.astype(np.int64)   # .long() :27


def pattern_edge_lists(id_type, k_max, k_min=None):
    """utils.get_custom_edge_list (utils.py:16-33) + process_arguments k range
    (utils.py:53-62) for the networkx generator families."""
    if k_min is None:
        k_min = 2 if id_type == 'star_graph' else 3
    out = []
    for k in range(k_min, k_max + 1):
        g = getattr(nx, id_type)(k)
        if isinstance(g, nx.Graph):
            out.append(list(g.edges))
        else:
            out += [list(h.edges) for h in g]
    return out
