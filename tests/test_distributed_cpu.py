"""Multi-rank host logic on CPU (gloo, world_size 2): batch sharding, flat gradient all-reduce,
vocabulary union.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsn_b200 import distributed as gd
from gsn_b200.synthetic import zinc_like_batch


def test_shard_ranges_partition_and_balance():
    rng = np.random.default_rng(0)
    for G, world in ((1000, 4), (7, 8), (128, 2), (1, 3), (0, 2)):
        w = rng.integers(1, 100, size=G)
        ranges = gd.shard_ranges(w, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == G
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        if G >= 8 * world:
            loads = [w[a:b].sum() for a, b in ranges]
            assert max(loads) <= 1.1 * w.sum() / world + w.max()


def test_shard_batch_is_a_partition_of_the_batch():
    b = zinc_like_batch(37, seed=2)
    parts = [gd.shard_batch(b, 4, r) for r in range(4)]
    assert sum(p['num_graphs'] for p in parts) == 37
    ei = np.concatenate([p['edge_index'] + b['node_ptr'][p['graph_range'][0]] for p in parts], 1)
    assert np.array_equal(ei, b['edge_index'])
    assert np.array_equal(np.concatenate([p['x'] for p in parts]), b['x'])
    assert np.array_equal(np.concatenate([p['edge_features'] for p in parts]), b['edge_features'])
    for p in parts:
        assert p['node_ptr'][0] == 0 and p['edge_ptr'][0] == 0
        if p['edge_index'].size:
            assert p['edge_index'].min() >= 0 and p['edge_index'].max() < p['node_ptr'][-1]
        assert np.array_equal(p['batch'], np.repeat(np.arange(p['num_graphs']), np.diff(p['node_ptr'])))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 2))
        if rank == 1:
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        gd.broadcast_parameters(model, src=0)
        # every rank: loss on its own shard; averaged gradients must equal the mean of the per-rank gradients
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn((16, 5), generator=g)
        model(x).square().mean().backward()
        local = torch.cat([p.grad.reshape(-1).clone() for p in model.parameters()])
        gd.allreduce_gradients(model.parameters())
        avg = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ok_grad = torch.allclose(avg, torch.stack(gathered).mean(0), atol=1e-6)
        w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        ws = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(ws, w0)
        ok_bcast = all(torch.equal(ws[0], w) for w in ws)
        # vocabulary union
        ids = torch.tensor([[1, 10], [3, 10], [5, 30]]) if rank == 0 else torch.tensor([[3, 20], [7, 20]])
        voc = gd.global_unique_per_column(ids)
        ok_voc = voc[0].tolist() == [1, 3, 5, 7] and voc[1].tolist() == [10, 20, 30]
        q.put((rank, ok_grad, ok_bcast, ok_voc))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_broadcast_and_vocabulary():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_grad, ok_bcast, ok_voc in res:
        assert ok_grad and ok_bcast and ok_voc, (rank, ok_grad, ok_bcast, ok_voc)


def _dataset_worker(rank, world, port, path, q):
    """N>1 host path of the dataset cache: every rank loads the flat cache, takes its shard, collates its own
    mini-batches and contributes to the data-set-wide identifier vocabulary (the one exchange COUNT sharding needs)"""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gsn_b200.dataset import FlatDataset
        ds = FlatDataset.load(path)
        sh = ds.shard(world, rank)
        counts = torch.tensor([len(sh), sh.num_nodes, sh.num_edges])
        dist.all_reduce(counts)
        ok_partition = counts.tolist() == [len(ds), ds.num_nodes, ds.num_edges]
        voc = gd.global_unique_per_column(sh.tensors['identifiers'])
        full = [torch.unique(ds.tensors['identifiers'][:, c]) for c in range(ds.tensors['identifiers'].shape[1])]
        ok_vocab = all(torch.equal(a, b) for a, b in zip(voc, full))
        b = sh.batch(list(range(min(3, len(sh)))))
        ok_batch = int(b.node_ptr[-1]) == b.x.shape[0] and (b.edge_index.numel() == 0 or int(b.edge_index.max()) < b.x.shape[0])
        q.put((rank, ok_partition, ok_vocab, ok_batch))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_dataset_cache(tmp_path):
    from gsn_b200.dataset import FlatDataset
    from tests.test_dataset import _graphs
    path = os.path.join(tmp_path, 'cache.pt')
    FlatDataset.from_list(_graphs(7, n_graphs=13), kinds={'identifiers': 'edge'}).save(path)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dataset_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *oks in res:
        assert all(oks), (rank, oks)


def _flat_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gsn_b200.distributed import FlatGradients
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    unused = torch.nn.Parameter(torch.ones(2))               # never receives a gradient: stays zero, layout identical
    fg = FlatGradients(list(lin.parameters()) + [unused])
    for step in range(2):
        fg.zero()
        x = torch.full((5, 4), float(rank + 1 + step))
        lin(x).sum().backward()                              # accumulates into the views, in place
        assert lin.weight.grad.data_ptr() == fg.flat.data_ptr()
        fg.allreduce()
        q.put((rank, step, lin.weight.grad.clone(), unused.grad.clone()))
    dist.destroy_process_group()


def test_flat_gradients_allreduce_world2():
    """FlatGradients: .grad views into one buffer survive backward, one all-reduce averages them over the ranks"""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    ps = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in range(4)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(2):
        ws = [g for g in got if g[1] == step]
        exp = torch.full((3, 4), 5.0 * ((1 + step) + (2 + step)) / 2)          # d/dW sum(W x) = sum over 5 rows of x
        for _, _, w, u in ws:
            torch.testing.assert_close(w, exp)
            assert float(u.abs().max()) == 0.0
