"""Host-side mirrors that need no GPU: data-set encoding (utils_encoding.py) and batching (PyG collate)."""
import numpy as np
import torch

from gsn_b200 import collate as gc
from gsn_b200 import encoding as ge


class G:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _graphs():
    g = torch.Generator().manual_seed(0)
    out = []
    for n, e in ((3, 4), (5, 6), (2, 2)):
        out.append(G(x=torch.ones(n, 1), edge_index=torch.randint(0, n, (2, e), generator=g),
                     identifiers=torch.randint(0, 50, (e, 3), generator=g) * 7,
                     degrees=torch.randint(0, 4, (n,), generator=g).float(), y=torch.tensor([1])))
    return out


def test_one_hot_unique_matches_numpy_reference_semantics():
    graphs = _graphs()
    raw = [g.identifiers.clone() for g in graphs]
    graphs, enc_ids, d_id, enc_deg, d_deg = ge.encode(graphs, 'one_hot_unique', 'one_hot_unique', ids={}, degree={})
    cat = torch.cat(raw, 0).numpy()
    exp_d, exp_cols = [], []
    for c in range(cat.shape[1]):           # utils_encoding.py:41-46
        u, inv = np.unique(cat[:, c], return_inverse=True)
        exp_d.append(len(u))
        exp_cols.append(inv)
    assert d_id == exp_d
    got = torch.cat([g.identifiers for g in graphs], 0).numpy()
    assert np.array_equal(got, np.stack(exp_cols, 1))
    assert graphs[0].identifiers.dtype == torch.int64
    assert d_deg == [len(np.unique(np.concatenate([[0]])))] or len(d_deg) == 1


def test_one_hot_max():
    graphs = _graphs()
    _, enc, d_id, _, _ = ge.encode(graphs, 'one_hot_max', None, ids={})
    assert d_id == [int(torch.cat([g.identifiers for g in graphs])[:, c].max()) + 1 for c in range(3)]


def test_collate_matches_pyg_semantics():
    graphs = _graphs()
    b = gc.collate(graphs)
    assert b.num_graphs == 3 and b.node_ptr.tolist() == [0, 3, 8, 10] and b.edge_ptr.tolist() == [0, 4, 10, 12]
    assert b.x.shape == (10, 1) and b.identifiers.shape == (12, 3) and b.y.tolist() == [1, 1, 1]
    assert b.batch.tolist() == [0] * 3 + [1] * 5 + [2] * 2
    off = 0
    for i, g in enumerate(graphs):
        e0, e1 = int(b.edge_ptr[i]), int(b.edge_ptr[i + 1])
        assert torch.equal(b.edge_index[:, e0:e1], g.edge_index + int(b.node_ptr[i]))


def test_merge_edge_columns_mixed_radix_tables():
    """fused.py: groups of categorical edge columns -> one pre-summed table per group; the mixed-radix row index of a
    tuple of ranks must address the sum of the per-column rows"""
    import itertools
    import torch
    from gsn_b200.fused import _merge_edge_columns
    dims = [2, 3, 5, 7, 8, 10, 4]
    cols, o = [], 0
    for d in dims:
        cols.append((d, o))
        o += d
    Te = torch.randn((o, 8), generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    tab, eg = _merge_edge_columns(Te, cols, 64)
    assert eg['n_groups'] == 3 and eg['group'] == [0, 0, 0, 1, 1, 2, 2]          # 2*3*5=30 | 7*8=56 | 10*4=40
    assert tab.shape[0] == 30 + 56 + 40
    g = torch.Generator().manual_seed(1)
    for _ in range(50):
        ranks = [int(torch.randint(0, d, (1,), generator=g)) for d in dims]
        rows = [0] * eg['n_groups']
        for c, r in enumerate(ranks):
            rows[eg['group'][c]] += eg['off'][c] + r * eg['mult'][c]
        got = sum(tab[r] for r in rows)
        exp = sum(Te[cols[c][1] + r] for c, r in enumerate(ranks))
        torch.testing.assert_close(got, exp, atol=1e-12, rtol=0)
    # a column larger than the cap stands alone; single column -> identity table
    tab1, eg1 = _merge_edge_columns(Te[:100 if o >= 100 else o], [(o, 0)], 16)
    assert eg1['n_groups'] == 1 and eg1['mult'] == [1] and torch.equal(tab1, Te)
