"""BASELINE config 4: IMDB-BINARY (the reference's shipped graph-tool fixture, 1000 graphs), cliques k<=5, edge
scope; README.md:99 recipe (GSN_sparse, gin, local ids, 4 layers, d_out 64, mean readout).  COUNT must reproduce
the fixture bit-exactly; forward = one 1000-graph batch and B=32 batches.  Under torchrun the graphs are sharded over
the ranks by edge count (no data-path collective)."""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def imdb_args(L=4, d=64):
    return dict(seed=0, model_name='GSN_sparse', readout='mean', dropout_features=[0.0] * (L + 1), bn=[True] * L,
                final_projection=[True] * (L + 1), inject_ids=False, inject_edge_features=True, random_features=False,
                id_scope='local', d_msg=[d] * L, d_out=[d] * L, d_h=[[d]] * L, aggr='add', flow='source_to_target',
                msg_kind='gin', train_eps=[False] * L, activation_mlp='relu', bn_mlp=True, jk_mlp=False,
                degree_embedding='one_hot_encoder', degree_as_tag=[False] * L, retain_features=[False] + [True] * (L - 1),
                multi_embedding_aggr='sum', input_node_encoder='None', d_out_node_encoder=d, edge_encoder='None',
                d_out_edge_encoder=[d] * L, id_embedding='one_hot_encoder', d_out_id_embedding=d,
                d_out_degree_embedding=d, extend_dims=True, activation='relu')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=10)
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        torch.distributed.init_process_group('nccl', device_id=dev)
    import networkx as nx
    from gsn_b200 import counting, distributed as gd, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import UniqueEncoder
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'imdb_k5_edge_counts.npz'))
    node_ptr, edge_ptr = z['node_ptr'].astype(np.int64), z['edge_ptr'].astype(np.int64)
    ei = z['edge_index'].astype(np.int64)
    for g in range(len(node_ptr) - 1):
        ei[:, edge_ptr[g]:edge_ptr[g + 1]] += node_ptr[g]
    gold = z['identifiers'].astype(np.int64)
    full = {'edge_index': ei, 'node_ptr': node_ptr, 'edge_ptr': edge_ptr,
            'batch': np.repeat(np.arange(1000), np.diff(node_ptr)), 'gold': gold, 'num_graphs': 1000}
    shard = gd.shard_batch(full, world, rank, balance='deg_pow4')       # K5: work per vertex ~ deg^4
    sds = patterns.make_subgraph_dicts([list(nx.complete_graph(k).edges) for k in (3, 4, 5)], 'local')
    ei_t, ptr_t = torch.from_numpy(shard['edge_index']).to(dev), torch.from_numpy(shard['node_ptr'])
    max_n = int(np.diff(node_ptr).max())

    def timed(fn, reps):
        ts = []
        for _ in range(reps):
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t[0])

    import bench
    clk = bench.Clocks(int(os.environ.get('LOCAL_RANK', 0)))
    clk.__enter__()
    ids = counting.count_batch(ei_t, ptr_t, sds, False, 'local', max_nodes_per_graph=max_n)
    ok = bool(np.array_equal(ids.cpu().numpy(), shard['gold']))
    t_count = timed(lambda: counting.count_batch(ei_t, ptr_t, sds, False, 'local', max_nodes_per_graph=max_n, check=False), a.reps)
    # one_hot_unique over the whole data set (vocabulary union across ranks)
    enc = UniqueEncoder(gd.global_unique_per_column(ids))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(in_features=1, out_features=2, encoder_ids=None, d_in_id=enc.d, in_edge_features=None,
                                 d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None,
                                 **imdb_args()).to(dev).eval()

    class B:
        pass

    def batch_of(g0, g1):
        n0, n1, e0, e1 = shard['node_ptr'][g0], shard['node_ptr'][g1], shard['edge_ptr'][g0], shard['edge_ptr'][g1]
        b = B()
        b.edge_index = (ei_t[:, e0:e1] - int(n0)).contiguous()
        b.x = torch.ones((int(n1 - n0), 1), device=dev)
        b.identifiers = enc(ids[e0:e1])
        b.degrees = torch.zeros(int(n1 - n0), device=dev)
        b.batch = torch.from_numpy(shard['batch'][n0:n1] - g0).to(dev)
        b.num_graphs = g1 - g0
        return b
    G = shard['num_graphs']
    whole = batch_of(0, G)
    with torch.no_grad():
        t_full = timed(lambda: model(whole), a.reps)
        small = [batch_of(g, min(g + 32, G)) for g in range(0, min(G, 32 * 8), 32)]

        def run_small():
            for b in small:
                model(b)
        t_small = timed(run_small, a.reps) / max(len(small), 1)
    clk.__exit__()
    if rank == 0:
        E_tot = int(edge_ptr[-1])
        print(json.dumps({'config': 'IMDB-BINARY fixture, 1000 graphs, cliques k<=5 edge scope + README.md:99 model (gin, local)',
                          'n_gpus': world, 'count_bit_exact_vs_graph_tool_fixture': ok, 'id_vocab': enc.d,
                          'count_seconds': t_count, 'count_graphs_per_s': 1000 / t_count, 'count_edges_per_s': E_tot / t_count,
                          'forward_all_graphs_seconds': t_full, 'forward_all_graphs_per_s': 1000 / t_full,
                          'forward_b32_seconds_per_batch': t_small, 'forward_b32_graphs_per_s': 32 * world / t_small,
                          'clocks': clk.summary()}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
