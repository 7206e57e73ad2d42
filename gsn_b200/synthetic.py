"""Synthetic "ZINC-shaped" / "molhiv-shaped" molecule batches.

ZINC itself is not shipped with the reference (README.md:104-109) and there is
no network, so benchmarks and large parity tests use this generator:
  n ~ clip(round(N(mean, sd)), lo, hi); random recursive tree with max degree 3;
  r ~ Poisson(2.75) ring closures joining vertices at tree distance 4 or 5
  (5- and 6-rings) keeping degree <= 4; symmetric edge_index in row-major sorted
  order (ZINC's (adj != 0).nonzero() order, utils_data_prep.py:151);
  x ~ U{0..27}, edge_features ~ U{1..3}.
Batches follow PyG's collate: node ids offset per graph, `batch`
vector, node_ptr / edge_ptr.
"""
from __future__ import annotations

import numpy as np


def _molecule(rng, n, ring_lambda):
    nbrs = [[] for _ in range(n)]
    for i in range(1, n):
        while True:
            p = int(rng.integers(0, i))
            if len(nbrs[p]) < 3:
                break
        nbrs[p].append(i)
        nbrs[i].append(p)
    tree = [list(a) for a in nbrs]
    for _ in range(int(rng.poisson(ring_lambda))):
        u = int(rng.integers(0, n))
        if len(nbrs[u]) >= 4:
            continue
        want = 4 + int(rng.integers(0, 2))
        dist = {u: 0}
        frontier = [u]
        for d in range(1, want + 1):
            nxt = []
            for a in frontier:
                for b in tree[a]:
                    if b not in dist:
                        dist[b] = d
                        nxt.append(b)
            frontier = nxt
        cands = [v for v in frontier if len(nbrs[v]) < 4 and v not in nbrs[u]]
        if cands:
            v = cands[int(rng.integers(0, len(cands)))]
            nbrs[u].append(v)
            nbrs[v].append(u)
    src = [a for a in range(n) for _ in nbrs[a]]
    dst = [b for a in range(n) for b in sorted(nbrs[a])]
    return np.array([src, dst], dtype=np.int64)


def zinc_like_batch(num_graphs, seed=0, mean_nodes=23.15, sd_nodes=4.5, min_nodes=9, max_nodes=37,
                    ring_lambda=2.75, distinct=None):
    """dict(edge_index [2,E] global ids, node_ptr, edge_ptr, batch, x [N,1] int64,
    edge_features [E,1] int64, degrees [N]).  distinct=K generates only K distinct
    molecules and tiles them (fast path for very large batches)."""
    rng = np.random.default_rng(seed)
    k = num_graphs if distinct is None else min(distinct, num_graphs)
    mols = []
    for _ in range(k):
        n = int(np.clip(np.rint(rng.normal(mean_nodes, sd_nodes)), min_nodes, max_nodes))
        mols.append((_molecule(rng, n, ring_lambda), n))
    pick = np.arange(num_graphs) % k
    sizes = np.array([mols[i][1] for i in pick], dtype=np.int64)
    esizes = np.array([mols[i][0].shape[1] for i in pick], dtype=np.int64)
    node_ptr = np.concatenate([[0], np.cumsum(sizes)])
    edge_ptr = np.concatenate([[0], np.cumsum(esizes)])
    if k == num_graphs:
        ei = np.concatenate([m[0] + node_ptr[i] for i, m in enumerate(mols)], 1)
    else:
        base = np.concatenate([m[0] for m in mols], 1)
        bptr = np.concatenate([[0], np.cumsum([m[0].shape[1] for m in mols])])
        reps = num_graphs // k
        tile_nodes = int(sum(m[1] for m in mols))
        local_off = np.repeat(np.concatenate([[0], np.cumsum([m[1] for m in mols])])[:-1],
                              [m[0].shape[1] for m in mols])
        tile = base + local_off
        parts = [tile + r * tile_nodes for r in range(reps)]
        rem = num_graphs - reps * k
        if rem:
            parts.append(tile[:, :bptr[rem]] + reps * tile_nodes)
        ei = np.concatenate(parts, 1)
    N, E = int(node_ptr[-1]), int(edge_ptr[-1])
    assert ei.shape[1] == E
    batch = np.repeat(np.arange(num_graphs, dtype=np.int64), sizes)
    return {'edge_index': ei, 'node_ptr': node_ptr.astype(np.int64), 'edge_ptr': edge_ptr.astype(np.int64),
            'batch': batch, 'x': rng.integers(0, 28, size=(N, 1), dtype=np.int64),
            'edge_features': rng.integers(1, 4, size=(E, 1), dtype=np.int64),
            'degrees': np.bincount(ei[0], minlength=N).astype(np.float32), 'num_graphs': num_graphs}
