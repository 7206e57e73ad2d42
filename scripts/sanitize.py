"""Small invocations of every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize.py"""
import contextlib, io, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gsn_b200 import counting, directional, ops, patterns
from gsn_b200.synthetic import zinc_like_batch
import bench

dev = torch.device('cuda', 0)
b = zinc_like_batch(24, seed=0)
ei, nptr = torch.from_numpy(b['edge_index']).to(dev), torch.from_numpy(b['node_ptr'])
N, E = int(nptr[-1]), ei.shape[1]
sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
ids = counting.count_batch(ei, nptr, sds, False, 'local')                                  # graph build + count
ids_v = counting.count_batch(ei, nptr, patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'global'), False, 'global')
plan = ops.EdgePlan(ei, N)                                                                  # CSR build
g = torch.Generator(device=dev).manual_seed(0)
dh = 128
P = torch.randn((N, 2 * dh), device=dev, generator=g)
Q = torch.randn((E, dh), device=dev, generator=g)
ops.general_edge(plan, P, Q, None, None)                                                    # dense message kernel
er1 = torch.randint(0, 4, (E, 1), device=dev, dtype=torch.int32, generator=g)
ops.general_edge_idx(plan, dh, P=P, edge_rows=er1, Te=torch.randn((4, dh), device=dev), edge_rows_csr=True)   # p1_tight
nr = torch.randint(0, 28, (N, 1), device=dev, dtype=torch.int32, generator=g)
er3 = torch.randint(0, 60, (E, 3), device=dev, dtype=torch.int32, generator=g)
ops.general_edge_idx(plan, dh, node_rows=nr, Tn=torch.randn((28, 2 * dh), device=dev), edge_rows=er3,
                     Te=torch.randn((60, dh), device=dev), edge_rows_csr=True)              # tab_tight
ops.segment_sum(plan, Q)
x = torch.randn((N, dh), device=dev, generator=g)
ops.ogb_aggregate(plan, x, Q, True, Q, None)
for M in (N, 40000):                                                                        # one-tile and persistent GEMM
    A = torch.randn((M, 128), device=dev, generator=g)
    W = torch.randn((128, 256), device=dev, generator=g)                                      # [Nout, K1 + K2]
    out = ops.linear(A, W, A2=A, bias=torch.randn(128, device=dev), activation='relu')
    ops.linear(A, W[:, :128].contiguous(), out=out, accumulate=True)
    ops.linear(A, torch.randn((256, 128), device=dev, generator=g), scale=torch.rand(256, device=dev), shift=torch.randn(256, device=dev))
ops.pool_ptr(x, nptr.to(dev))
h = torch.randn((N, 60), device=dev, generator=g, requires_grad=True)                        # DGN forward + backward
ef = ids[:, :4].float()
y = directional.dgn_aggregate(plan, h, ids_v[:, :2].float(), ef, 'mean max min std dir0-av dir1-dx dir5-0.1 dir2-dx-balanced',
                              'identity amplification', {'log': 1.1})
y.square().sum().backward()
torch.cuda.synchronize()
print('sanitize: all kernels ran', float(y.sum()), float(h.grad.abs().sum()))
