"""Pattern (substructure) set-up on the host: automorphism orbits with the
reference's numbering, and compilation of the GPU matching plan.

Mirrors /root/reference/utils_graph_processing.py:10-100
(`automorphism_orbits`, `induced_edge_automorphism_orbits`): same arguments,
same return tuple `(subgraph, orbit_partition, orbit_membership, aut_count)`,
so `utils_data_gen.generate_dataset` (:35-42) can call them unchanged.

The reference enumerates the whole automorphism group with graph-tool
(:22, |Aut(K_12)| = 479,001,600 maps); here orbits and |Aut| come from a
stabiliser chain (one small backtracking search per orbit candidate), which
also yields the symmetry-breaking constraints of the plan
(include/gsn_b200.h: GsnPlan).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence, Tuple

import numpy as np

GSN_MAXK = 16
FAMILY_GENERIC, FAMILY_CYCLES, FAMILY_CLIQUES = 0, 1, 2


class GsnPlan(ctypes.Structure):
    """ctypes image of `struct GsnPlan` (include/gsn_b200.h)."""
    _fields_ = [
        ('k', ctypes.c_int32), ('induced', ctypes.c_int32), ('scope', ctypes.c_int32),
        ('n_cols', ctypes.c_int32), ('col0', ctypes.c_int32), ('family', ctypes.c_int32),
        ('kmin', ctypes.c_int32), ('kmax', ctypes.c_int32),
        ('nbr_mask', ctypes.c_uint32 * GSN_MAXK),
        ('non_mask', ctypes.c_uint32 * GSN_MAXK),
        ('gt_mask', ctypes.c_uint32 * GSN_MAXK),
        ('vorbit', ctypes.c_int8 * GSN_MAXK),
        ('e_fwd', (ctypes.c_int8 * GSN_MAXK) * GSN_MAXK),
        ('e_bwd', (ctypes.c_int8 * GSN_MAXK) * GSN_MAXK),
    ]


class Pattern:
    """The simple undirected pattern graph H; stands in for the `gt.Graph` the
    reference keeps in subgraph_dict['subgraph'] -- the only method the reference
    calls on it is `.get_edges()` (utils_graph_processing.py:74,147)."""

    def __init__(self, edge_list):
        pairs = set()
        kmax = -1
        for a, b in edge_list:
            a, b = int(a), int(b)
            kmax = max(kmax, a, b)
            if a != b:                                  # remove_self_loops (:18)
                pairs.add((min(a, b), max(a, b)))       # remove_parallel_edges (:19)
        self.k = kmax + 1                               # graph-tool creates vertices 0..max id
        self.edges: List[Tuple[int, int]] = sorted(pairs)
        self.adj = [[False] * self.k for _ in range(self.k)]
        for a, b in self.edges:
            self.adj[a][b] = self.adj[b][a] = True
        self.deg = [sum(r) for r in self.adj]

    def get_edges(self):
        return np.array(self.edges, dtype=np.int64).reshape(-1, 2)

    def num_vertices(self):
        return self.k

    def get_vertices(self):
        return np.arange(self.k)

    def directed_edges(self):
        """to_undirected(get_edges()) of :74 / :147 -- both directions, sorted
        lexicographically by (row, col), duplicates dropped."""
        return sorted({(a, b) for a, b in self.edges} | {(b, a) for a, b in self.edges})

    def is_connected(self):
        if self.k == 0:
            return False
        seen, todo = {0}, [0]
        while todo:
            u = todo.pop()
            for v in range(self.k):
                if self.adj[u][v] and v not in seen:
                    seen.add(v)
                    todo.append(v)
        return len(seen) == self.k

    def is_cycle(self):
        return self.k >= 3 and len(self.edges) == self.k and all(d == 2 for d in self.deg) and self.is_connected()

    def is_clique(self):
        return self.k >= 3 and len(self.edges) == self.k * (self.k - 1) // 2

    # -- automorphisms ---------------------------------------------------
    def _extend(self, img: List[int], used: List[bool], v: int) -> bool:
        """Backtracking: is there an automorphism extending the partial
        assignment img (img[u] = -1: free)?"""
        k = self.k
        while v < k and img[v] >= 0:
            v += 1
        if v == k:
            return True
        for c in range(k):
            if used[c] or self.deg[c] != self.deg[v]:
                continue
            ok = True
            for u in range(k):
                if img[u] >= 0 and self.adj[v][u] != self.adj[c][img[u]]:
                    ok = False
                    break
            if ok:
                img[v] = c
                used[c] = True
                if self._extend(img, used, v + 1):
                    img[v] = -1
                    used[c] = False
                    return True
                img[v] = -1
                used[c] = False
        return False

    def has_automorphism(self, assign: Dict[int, int]) -> bool:
        img = [-1] * self.k
        used = [False] * self.k
        for u, c in assign.items():
            if used[c] or self.deg[u] != self.deg[c]:
                return False
            img[u] = c
            used[c] = True
        for u, c in assign.items():
            for w, d in assign.items():
                if self.adj[u][w] != self.adj[c][d]:
                    return False
        return self._extend(img, used, 0)

    def vertex_orbits(self) -> List[int]:
        """orbit id per vertex = contiguous rank of the smallest vertex of its
        orbit (what :24-42 computes from the full group)."""
        rep = list(range(self.k))
        for v in range(self.k):
            if rep[v] != v:
                continue
            for u in range(v + 1, self.k):
                if rep[u] == u and self.has_automorphism({v: u}):
                    rep[u] = v
        uniq = sorted(set(rep))
        rank = {r: i for i, r in enumerate(uniq)}
        return [rank[r] for r in rep]

    def matching_order(self) -> List[int]:
        k = self.k
        start = max(range(k), key=lambda v: (self.deg[v], -v))
        order, inset = [start], {start}
        while len(order) < k:
            best = max((v for v in range(k) if v not in inset),
                       key=lambda v: (sum(self.adj[v][u] for u in order), self.deg[v], -v))
            if not any(self.adj[best][u] for u in order):
                raise NotImplementedError('disconnected patterns are not supported by the CUDA matcher')
            order.append(best)
            inset.add(best)
        return order

    def stabiliser_chain(self, base: Sequence[int]):
        """|Aut(H)| and symmetry-breaking constraints [(u, v)]: f(u) < f(v)."""
        fixed: Dict[int, int] = {}
        aut = 1
        cons = []
        for b in base:
            orbit = [u for u in range(self.k) if u not in fixed and self.has_automorphism({**fixed, b: u})]
            aut *= len(orbit)
            cons += [(b, u) for u in orbit if u != b]
            fixed[b] = b
        return aut, cons


def automorphism_orbits(edge_list, print_msgs=True, **kwargs):
    """utils_graph_processing.py:10-56."""
    if kwargs.get('directed', False):
        raise NotImplementedError('directed substructures are not supported (reference quirk, SURVEY A.3)')
    graph = Pattern(edge_list)
    orb = graph.vertex_orbits()
    orbit_membership = {v: int(orb[v]) for v in range(graph.k)}
    orbit_partition: Dict[int, List[int]] = {}
    for vertex, orbit in orbit_membership.items():
        orbit_partition.setdefault(orbit, []).append(vertex)
    aut_count, _ = graph.stabiliser_chain(range(graph.k))
    if print_msgs:
        print('Orbit partition of given substructure: {}'.format(orbit_partition))
        print('Number of orbits: {}'.format(len(orbit_partition)))
        print('Automorphism count: {}'.format(aut_count))
    return graph, orbit_partition, orbit_membership, aut_count


def induced_edge_automorphism_orbits(edge_list, **kwargs):
    """utils_graph_processing.py:58-100: edge orbit = (un)ordered pair of the
    end points' vertex orbits, numbered in first-seen order over the coalesced
    bidirectional edge list."""
    directed = kwargs.get('directed', False)
    directed_orbits = kwargs.get('directed_orbits', False)
    graph, _, orbit_membership, aut_count = automorphism_orbits(edge_list=edge_list, directed=directed,
                                                                print_msgs=False)
    edge_orbit_partition: Dict[int, List[Tuple[int, int]]] = {}
    edge_orbit_membership: Dict[int, int] = {}
    seen: Dict[object, int] = {}
    for i, edge in enumerate(graph.directed_edges()):
        key = ((orbit_membership[edge[0]], orbit_membership[edge[1]]) if directed_orbits
               else frozenset([orbit_membership[edge[0]], orbit_membership[edge[1]]]))
        if key not in seen:
            seen[key] = len(seen)
        edge_orbit_partition.setdefault(seen[key], []).append(tuple(edge))
        edge_orbit_membership[i] = seen[key]
    if kwargs.get('print_msgs', True):
        print('Edge orbit partition of given substructure: {}'.format(edge_orbit_partition))
        print('Number of edge orbits: {}'.format(len(edge_orbit_partition)))
        print('Graph (vertex) automorphism count: {}'.format(aut_count))
    return graph, edge_orbit_partition, edge_orbit_membership, aut_count


def make_subgraph_dicts(edge_lists, id_scope, directed_orbits=False, print_msgs=False):
    """The pattern set-up loop of utils_data_gen.py:31-42."""
    dicts = []
    for el in edge_lists:
        if id_scope == 'local':
            sub, part, memb, aut = induced_edge_automorphism_orbits(edge_list=el, directed=False,
                                                                    directed_orbits=directed_orbits,
                                                                    print_msgs=print_msgs)
        else:
            sub, part, memb, aut = automorphism_orbits(edge_list=el, directed=False, print_msgs=print_msgs)
        dicts.append({'subgraph': sub, 'orbit_partition': part, 'orbit_membership': memb, 'aut_count': aut})
    return dicts


def _as_pattern(subgraph) -> Pattern:
    if isinstance(subgraph, Pattern):
        return subgraph
    return Pattern(np.asarray(subgraph.get_edges()).reshape(-1, 2)[:, :2].tolist())


def compile_plan(subgraph_dict, induced: bool, scope: int, col0: int = 0) -> GsnPlan:
    """subgraph_dict (as built by utils_data_gen.py:40-41) -> GsnPlan."""
    H = _as_pattern(subgraph_dict['subgraph'])
    k = H.k
    if k < 2 or k > GSN_MAXK:
        raise NotImplementedError(f'pattern size {k} outside 2..{GSN_MAXK}')
    memb = subgraph_dict['orbit_membership']
    n_cols = len(subgraph_dict['orbit_partition'])
    if n_cols > 127:
        raise NotImplementedError('more than 127 orbits in one pattern')
    order = H.matching_order()
    pos = {v: p for p, v in enumerate(order)}
    aut, cons = H.stabiliser_chain(order)
    if int(subgraph_dict['aut_count']) != aut:
        raise ValueError(f"aut_count {subgraph_dict['aut_count']} does not match the pattern (|Aut| = {aut})")

    P = GsnPlan()
    P.k, P.induced, P.scope, P.n_cols, P.col0 = k, int(bool(induced)), int(scope), n_cols, int(col0)
    P.family, P.kmin, P.kmax = FAMILY_GENERIC, k, k
    for p in range(k):
        for q in range(p):
            if H.adj[order[p]][order[q]]:
                P.nbr_mask[p] |= 1 << q
            else:
                P.non_mask[p] |= 1 << q
        P.vorbit[p] = -1
        for q in range(GSN_MAXK):
            P.e_fwd[p][q] = -1
            P.e_bwd[p][q] = -1
    for p in range(k, GSN_MAXK):
        P.vorbit[p] = -1
        for q in range(GSN_MAXK):
            P.e_fwd[p][q] = -1
            P.e_bwd[p][q] = -1
    for u, v in cons:                                   # f(u) < f(v); u precedes v in `order`
        assert pos[u] < pos[v]
        P.gt_mask[pos[v]] |= 1 << pos[u]
    if scope == 0:
        for p in range(k):
            P.vorbit[p] = int(memb[order[p]])
    else:
        for i, (u, v) in enumerate(H.directed_edges()):
            o = int(memb[i])
            pu, pv = pos[u], pos[v]
            if pu < pv:
                P.e_fwd[pv][pu] = o                     # directed pattern edge earlier -> later
            else:
                P.e_bwd[pu][pv] = o                     # later -> earlier
    return P


def family_plan(family: int, kmin: int, kmax: int, induced: bool, scope: int, col0: int) -> GsnPlan:
    """All cycle lengths / clique sizes kmin..kmax in one traversal; one column
    per size (cycles and cliques have a single vertex orbit and a single edge
    orbit)."""
    if kmax > GSN_MAXK or kmin < 3 or kmax < kmin:
        raise NotImplementedError(f'family sizes {kmin}..{kmax} outside 3..{GSN_MAXK}')
    P = GsnPlan()
    P.k, P.induced, P.scope = kmax, int(bool(induced)), int(scope)
    P.n_cols, P.col0, P.family, P.kmin, P.kmax = kmax - kmin + 1, int(col0), family, kmin, kmax
    return P


def compile_plans(subgraph_dicts, induced: bool, scope: int, fuse_families: bool = True) -> List[GsnPlan]:
    """One plan per pattern, except that runs of consecutive cycle graphs (or
    complete graphs) of consecutive sizes -- what `--id_type cycle_graph --k K`
    produces, utils.py:53-62 -- are fused into one family plan."""
    plans: List[GsnPlan] = []
    pats = [_as_pattern(sd['subgraph']) for sd in subgraph_dicts]
    i, col = 0, 0
    while i < len(pats):
        fam = FAMILY_CYCLES if pats[i].is_cycle() else FAMILY_CLIQUES if pats[i].is_clique() else FAMILY_GENERIC
        # K3 is both a cycle and a clique: decide by the next pattern
        if fam != FAMILY_GENERIC and pats[i].k == 3 and i + 1 < len(pats) and pats[i + 1].k == 4:
            fam = FAMILY_CYCLES if pats[i + 1].is_cycle() else FAMILY_CLIQUES if pats[i + 1].is_clique() else fam
        if fuse_families and fam != FAMILY_GENERIC and len(subgraph_dicts[i]['orbit_partition']) == 1:
            test = Pattern.is_cycle if fam == FAMILY_CYCLES else Pattern.is_clique
            j = i
            while (j + 1 < len(pats) and test(pats[j + 1]) and pats[j + 1].k == pats[j].k + 1
                   and len(subgraph_dicts[j + 1]['orbit_partition']) == 1):
                j += 1
            plans.append(family_plan(fam, pats[i].k, pats[j].k, induced, scope, col))
            col += j - i + 1
            i = j + 1
        else:
            plans.append(compile_plan(subgraph_dicts[i], induced, scope, col))
            col += len(subgraph_dicts[i]['orbit_partition'])
            i += 1
    return plans


def total_columns(subgraph_dicts) -> int:
    return sum(len(sd['orbit_partition']) for sd in subgraph_dicts)


# ---------------------------------------------------------------------------
# substructure families (utils.py:16-33 uses networkx generators / .g6 files)
# ---------------------------------------------------------------------------
def parse_graph6(line: str):
    """Minimal graph6 reader (n <= 62), enough for datasets/all_simple_graphs."""
    data = [ord(c) - 63 for c in line.strip()]
    n = data[0]
    bits = []
    for d in data[1:]:
        bits += [(d >> s) & 1 for s in range(5, -1, -1)]
    edges, idx = [], 0
    for j in range(1, n):
        for i in range(j):
            if bits[idx]:
                edges.append((i, j))
            idx += 1
    return n, edges


def get_custom_edge_list(ks, substructure_type=None, filename=None):
    """utils.get_custom_edge_list (utils.py:16-33)."""
    import os
    if substructure_type is None and filename is None:
        raise ValueError('You must specify either a type or a filename where to read substructures from.')
    edge_lists = []
    for k in ks:
        if substructure_type is not None:
            import networkx as nx
            graphs_nx = getattr(nx, substructure_type)(k)
            if isinstance(graphs_nx, nx.Graph):
                edge_lists.append(list(graphs_nx.edges))
            else:
                edge_lists += [list(g.edges) for g in graphs_nx]
        else:
            with open(os.path.join(filename, 'graph{}c.g6'.format(k))) as fh:
                for line in fh:
                    if line.strip():
                        edge_lists.append(parse_graph6(line)[1])
    return edge_lists
