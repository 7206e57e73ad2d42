"""Flat dataset cache + index-arithmetic collate (host logic, CPU) and the batched generate_dataset drop-in (GPU)."""
import os

import numpy as np
import pytest
import torch

from gsn_b200.collate import collate
from gsn_b200.dataset import Data, FlatDataset


def _graphs(seed=0, n_graphs=9, edge_ids=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_graphs):
        n = int(torch.randint(1, 9, (1,), generator=g))
        a = torch.triu(torch.rand((n, n), generator=g) < 0.5, 1)
        r, c = a.nonzero(as_tuple=True)
        ei = torch.stack((torch.cat((r, c)), torch.cat((c, r)))).long()
        e = ei.shape[1]
        out.append(Data(x=torch.randint(0, 5, (n, 2), generator=g), edge_index=ei,
                        degrees=torch.rand(n, generator=g), edge_features=torch.rand((e, 3), generator=g),
                        identifiers=torch.randint(0, 7, ((e if edge_ids else n), 4), generator=g),
                        y=torch.tensor([float(i)])))
    return out


@pytest.mark.parametrize('edge_ids', [True, False])
def test_flat_round_trip_and_batch_equals_collate(tmp_path, edge_ids):
    graphs = _graphs(1, edge_ids=edge_ids)
    kinds = {'identifiers': 'edge' if edge_ids else 'node'}
    ds = FlatDataset.from_list(graphs, num_classes=2, orbit_partition_sizes=[1, 1, 2], kinds=kinds)
    assert len(ds) == len(graphs) and ds.kinds()['edge_features'] == 'edge' and ds.kinds()['x'] == 'node' and ds.kinds()['y'] == 'graph'
    path = os.path.join(tmp_path, 'cache.pt')
    ds.save(path)
    ds2 = FlatDataset.load(path)
    assert ds2.meta['num_classes'] == 2 and ds2.meta['orbit_partition_sizes'] == [1, 1, 2]
    back = ds2.to_list()
    for a, b in zip(graphs, back):
        for k in ('x', 'edge_index', 'degrees', 'edge_features', 'identifiers', 'y'):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
    for sel in ([0], [3, 1, 4], list(range(len(graphs))), [8, 8, 2]):
        exp = collate([graphs[i] for i in sel])
        got = ds2.batch(sel)
        for k in ('x', 'edge_index', 'degrees', 'edge_features', 'identifiers', 'y', 'batch', 'node_ptr', 'edge_ptr'):
            assert torch.equal(getattr(exp, k), getattr(got, k)), (k, sel)
        assert got.num_graphs == len(sel)
    # the reference's tuple layout
    ref_path = os.path.join(tmp_path, 'ref.pt')
    ds2.save_reference_tuple(ref_path)
    lst, nc, ops_ = torch.load(ref_path, weights_only=False)
    assert nc == 2 and ops_ == [1, 1, 2] and torch.equal(lst[3].edge_index, graphs[3].edge_index)


def test_load_rejects_foreign_files(tmp_path):
    p = os.path.join(tmp_path, 'x.pt')
    torch.save({'format': 'other'}, p)
    with pytest.raises(ValueError):
        FlatDataset.load(p)


class _Raw:
    def __init__(self, edge_mat, node_features, label, edge_features=None):
        self.edge_mat, self.node_features, self.label = edge_mat, node_features, label
        if edge_features is not None:
            self.edge_features = edge_features


@pytest.mark.gpu
@pytest.mark.parametrize('scope', ['local', 'global'])
def test_prepare_graphs_matches_per_graph_oracle(scope):
    """generate_dataset drop-in: one batched COUNT launch == the reference's per-graph loop (oracle restatement),
    incl. a graph with self loops (stripped, utils_ids.py:11-15) and a graph without edges"""
    from gsn_b200 import patterns
    from gsn_b200.dataset import prepare_graphs
    from gsn_b200.synthetic import zinc_like_batch
    from oracle import count_vf2
    b = zinc_like_batch(12, seed=11)
    raws = []
    for i in range(12):
        n0, n1, e0, e1 = b['node_ptr'][i], b['node_ptr'][i + 1], b['edge_ptr'][i], b['edge_ptr'][i + 1]
        ei = torch.from_numpy(b['edge_index'][:, e0:e1] - n0)
        raws.append(_Raw(ei, torch.zeros((int(n1 - n0), 1)), 0.5, torch.arange(ei.shape[1]).float().unsqueeze(1)))
    loops = torch.tensor([[0, 1], [0, 1]])
    raws[2].edge_mat = torch.cat((raws[2].edge_mat, loops), 1)
    raws[2].edge_features = torch.cat((raws[2].edge_features, torch.tensor([[-1.], [-2.]])), 0)
    raws.append(_Raw(torch.zeros((2, 0), dtype=torch.int64), torch.zeros((3, 1)), 1.0, torch.zeros((0, 1))))
    els = count_vf2.pattern_edge_lists('cycle_graph', 6)
    sds = patterns.make_subgraph_dicts(els, scope)
    out = prepare_graphs(raws, sds, {'induced': False, 'directed': False}, scope, dataset_name='ZINC')
    sds_o = count_vf2.make_subgraph_dicts(els, scope)
    fn = count_vf2.subgraph_isomorphism_edge_counts if scope == 'local' else count_vf2.subgraph_isomorphism_vertex_counts
    for r, d in zip(raws[:6] + raws[-1:], out[:6] + out[-1:]):
        n = r.node_features.shape[0]
        assert d.graph_size == n and d.y.dtype == torch.float32
        assert not bool((d.edge_index[0] == d.edge_index[1]).any())
        assert d.edge_features.shape[0] == d.edge_index.shape[1]
        if d.edge_index.shape[1] == 0:
            assert d.identifiers.shape == ((0, 4) if scope == 'local' else (n, 4)) and d.identifiers.dtype == torch.int64
            continue
        exp = torch.cat([torch.as_tensor(np.asarray(fn(d.edge_index.numpy(), subgraph_dict=sd, induced=False, num_nodes=n,
                                                        directed=False))) for sd in sds_o], 1).long()
        assert torch.equal(d.identifiers, exp)
    assert float(out[2].edge_features.min()) >= 0          # the self-loop rows are gone


def test_shards_partition_the_dataset():
    graphs = _graphs(4, n_graphs=11)
    ds = FlatDataset.from_list(graphs, kinds={'identifiers': 'edge'})
    for world in (1, 2, 3, 16):
        seen = []
        for rank in range(world):
            sh = ds.shard(world, rank)
            seen += sh.to_list()
            assert int(sh.node_ptr[0]) == 0 and int(sh.edge_ptr[0]) == 0
        assert len(seen) == len(graphs)
        for a, b in zip(graphs, seen):
            for k in ('x', 'edge_index', 'edge_features', 'identifiers', 'y'):
                assert torch.equal(getattr(a, k), getattr(b, k)), (world, k)


def test_bucket_padding_and_packing():
    """BucketedPipeline.pad (host side of the shape-bucketed capture): sentinel graphs hold the padding nodes, padding
    edge columns are self loops on sentinel nodes (<= 4 per node), the real batch is untouched, shapes depend on the
    bucket only, and the single-buffer packing round-trips bit for bit"""
    import bench
    from gsn_b200.pipeline import BucketedPipeline, FIELDS
    bp = BucketedPipeline(None, None, False, 'local', None, 64, node_step=64, edge_step=128)
    seen = {}
    for seed in range(8):
        b = bench.build_batches(16, 1, seed0=seed)[0]
        N, E, G = int(b['node_ptr'][-1]), b['edge_index'].shape[1], 16
        N_cap, E_cap, G_cap = bp.bucket(N, E, G)
        assert N_cap % 64 == 0 and E_cap % 128 == 0 and N_cap > N and E_cap >= E and G_cap > G
        p = bp.pad(b)
        assert p['x'].shape[0] == N_cap and p['edge_index'].shape == (2, E_cap) and p['node_ptr'].numel() == G_cap + 1
        assert p['batch'].shape[0] == N_cap and p['degrees'].shape[0] == N_cap and p['edge_features'].shape[0] == E_cap
        for k in FIELDS:
            n = {'edge_index': E, 'node_ptr': G + 1, 'edge_features': E}.get(k, N)
            real = p[k][:, :n] if k == 'edge_index' else p[k][:n]
            assert np.array_equal(real.numpy(), b[k]), k
        ptr = p['node_ptr'].numpy()
        assert ptr[-1] == N_cap and (np.diff(ptr) >= 0).all() and np.diff(ptr)[G:].max() <= 64
        pad_e = p['edge_index'][:, E:].numpy()
        assert (pad_e[0] == pad_e[1]).all() and (pad_e >= N).all() and (pad_e < N_cap).all()
        if pad_e.shape[1]:
            assert np.bincount(pad_e[0] - N).max() <= 4
        assert np.array_equal(p['batch'].numpy(), np.repeat(np.arange(G_cap), np.diff(ptr)))
        pk = bp.packing(p)
        buf = pk.pack(p)
        views = pk.views(buf)
        for k in FIELDS:
            assert torch.equal(views[k], p[k]), k
        seen.setdefault((N_cap, E_cap, G_cap), pk.nbytes)
        assert seen[(N_cap, E_cap, G_cap)] == pk.nbytes              # layout is a function of the bucket


def test_flat_loader_batches_equal_collate_over_the_same_order():
    """sequential order == torch DataLoader(list, collate_fn=collate); shuffled order == collate over randperm(seed)"""
    from torch.utils.data import DataLoader
    from gsn_b200.dataset import FlatLoader
    graphs = _graphs(5, n_graphs=10)
    ds = FlatDataset.from_list(graphs, kinds={'identifiers': 'edge'})
    keys = ('x', 'edge_index', 'edge_features', 'identifiers', 'y', 'batch', 'node_ptr')
    for drop_last, bs in ((False, 4), (True, 4), (False, 10), (False, 1)):
        ref = list(DataLoader(graphs, batch_size=bs, shuffle=False, drop_last=drop_last, collate_fn=collate))
        got = list(FlatLoader(ds, batch_size=bs, drop_last=drop_last))
        assert len(ref) == len(got) == len(FlatLoader(ds, bs, False, drop_last))
        for a, b in zip(ref, got):
            for k in keys:
                assert torch.equal(getattr(a, k), getattr(b, k)), (k, drop_last, bs)
    order = torch.randperm(10, generator=torch.Generator().manual_seed(9)).tolist()
    got = list(FlatLoader(ds, batch_size=3, shuffle=True, generator=torch.Generator().manual_seed(9)))
    assert len(got) == 4
    for i, b in enumerate(got):
        exp = collate([graphs[j] for j in order[3 * i:3 * i + 3]])
        for k in keys:
            assert torch.equal(getattr(exp, k), getattr(b, k)), (k, i)


def test_tile_plan_capacity_bound():
    """gsn_b200.fused_model.tile_plan sizes its buffer with min(G, 2N/128 + G/32 + 2); the greedy packing the kernel
    performs (<= 128 rows, <= 32 graphs per tile, oversized graphs alone) never needs more, whatever the graph sizes"""
    import numpy as np
    rng = np.random.default_rng(11)
    for trial in range(300):
        G = int(rng.integers(1, 400))
        kind = trial % 4
        sizes = (rng.integers(0, 3, G) if kind == 0 else rng.integers(1, 129, G) if kind == 1
                 else rng.integers(9, 38, G) if kind == 2 else rng.integers(0, 300, G))
        N = int(sizes.sum())
        tiles, g0 = 0, 0
        while g0 < G:
            g1, rows = g0, 0
            while g1 < G and g1 - g0 < 32 and rows + sizes[g1] <= 128:
                rows += sizes[g1]
                g1 += 1
            g0 = max(g1, g0 + 1)
            tiles += 1
        assert tiles <= min(G, 2 * N // 128 + G // 32 + 2), (trial, G, N, tiles)
