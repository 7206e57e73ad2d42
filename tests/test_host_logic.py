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
