"""BASELINE config 5: counting throughput on synthetic molecule-like graphs (avg 30 nodes), cycles k <= K
(default 12), vertex and edge scope, non-induced; inputs device-resident.  One process per GPU; under torchrun every
rank counts its own shard of the graphs (no collective on the data path) and rank 0 prints one JSON line.

  python scripts/bench_count.py --graphs 1000000 --k 12
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_count.py --graphs 1000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--graphs', type=int, default=1000000)
    ap.add_argument('--k', type=int, default=12)
    ap.add_argument('--distinct', type=int, default=32768)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--check', type=int, default=300, help='graphs verified against the C oracle')
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        torch.distributed.init_process_group('nccl', device_id=dev)
    import networkx as nx
    from gsn_b200 import counting, patterns
    from gsn_b200.synthetic import zinc_like_batch
    per_rank = a.graphs // world
    b = zinc_like_batch(per_rank, seed=100 + rank, mean_nodes=30.0, sd_nodes=6.0, min_nodes=9, max_nodes=64,
                        distinct=min(a.distinct, per_rank))
    els = [list(nx.cycle_graph(k).edges) for k in range(3, a.k + 1)]
    ei = torch.from_numpy(b['edge_index']).to(dev)
    ptr = torch.from_numpy(b['node_ptr'])
    N, E = int(b['node_ptr'][-1]), int(ei.shape[1])
    import bench
    out = {}
    clk = bench.Clocks(local)
    clk.__enter__()
    for scope in ('global', 'local'):
        sds = patterns.make_subgraph_dicts(els, scope)
        ids = counting.count_batch(ei, ptr, sds, False, scope, max_nodes_per_graph=64)          # warm-up + result
        if rank == 0 and a.check:
            from oracle import count_c, count_vf2
            g = min(a.check, per_rank)
            exp = count_c.count_batch(b['node_ptr'][:g + 1], b['edge_ptr'][:g + 1], b['edge_index'][:, :b['edge_ptr'][g]],
                                      count_vf2.make_subgraph_dicts(els, scope), False, 1 if scope == 'local' else 0)
            rows = b['edge_ptr'][g] if scope == 'local' else b['node_ptr'][g]
            assert np.array_equal(ids[:rows].cpu().numpy(), exp), 'COUNT mismatch vs the C oracle'
        ts_small = []
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(a.reps):          # the one-launch path (gsn_count_small): build + count + write-out in one kernel
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            counting.count_batch(ei, ptr, sds, False, scope, num_nodes=N, max_nodes_per_graph=64, check=False, status=st)
            e1.record()
            torch.cuda.synchronize()
            ts_small.append(e0.elapsed_time(e1) * 1e-3)
        assert int(st.item()) == 0
        ts_total, ts_kernel = [], []
        for _ in range(a.reps):          # the general path (gsn_graph_build + gsn_count_pattern), for comparison
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            graph = counting.BatchedGraph(ei, ptr, num_nodes=N, max_nodes_per_graph=64)
            e1.record()
            counting.count_batch(ei, ptr, sds, False, scope, num_nodes=N, max_nodes_per_graph=64, check=False, graph=graph)
            e2.record()
            torch.cuda.synchronize()
            ts_total.append(e0.elapsed_time(e2) * 1e-3)
            ts_kernel.append(e1.elapsed_time(e2) * 1e-3)
        t = torch.tensor([min(ts_total), min(ts_kernel), min(ts_small)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        out[scope] = {'seconds': float(t[2]), 'graphs_per_s': per_rank * world / float(t[2]),
                      'edges_per_s': E * world / float(t[2]),
                      'general_path_seconds_build_plus_count': float(t[0]), 'general_path_seconds_count_kernels': float(t[1]),
                      'general_path_graphs_per_s': per_rank * world / float(t[0]),
                      'columns': a.k - 2, 'checksum': int(ids.sum().item())}
    clk.__exit__()
    if rank == 0:
        bytes_alg = 16 * E + 8 * N * (a.k - 2)
        print(json.dumps({'config': f'{per_rank * world} synthetic graphs (mean 30 nodes, {per_rank} per GPU, '
                                    f'{min(a.distinct, per_rank)} distinct molecules tiled), cycles k<={a.k}, non-induced',
                          'n_gpus': world, 'N_per_gpu': N, 'E_per_gpu': E, 'vertex_scope': out['global'],
                          'edge_scope': out['local'],
                          'hbm_fraction_vertex_scope': bytes_alg / out['global']['seconds'] / 1e9 / 6539.5,
                          'oracle_checked_graphs': a.check, 'clocks': clk.summary()}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
