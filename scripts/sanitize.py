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
# ---- round 2: general COUNT path (graph build + count + heavy-item kernel) on a dense batch, one-launch COUNT (above),
#      one-kernel forward through the pipeline, training kernels (embedding bag fwd/bwd, ogb fwd/bwd, tensor-core Linear)
from tests.util import batch_graphs, random_graph
rng = np.random.default_rng(0)
gs = [(random_graph(rng, 30, 0.6), 30) for _ in range(4)] + [(random_graph(rng, 70, 0.3), 70)]
dptr, _, dei = batch_graphs(gs)
import networkx as nx
cl = patterns.make_subgraph_dicts([list(nx.complete_graph(k).edges) for k in (3, 4, 5)], 'local')
counting.count_batch(torch.from_numpy(dei).to(dev), torch.from_numpy(dptr), cl, False, 'local')
from gsn_b200.network import GNNSubstructures
from gsn_b200.pipeline import BucketedPipeline, GSNPipeline, UniqueEncoder
enc = UniqueEncoder.fit(ids)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = GNNSubstructures(**bench.model_ctor(enc.d), **bench.model_args(enc.d)).to(dev).eval()
t = bench.to_tensors(b, device=dev)
with torch.no_grad():
    o1 = GSNPipeline(model, sds, False, 'local', enc, 64, fused='model').step(t)
bp = BucketedPipeline(model, sds, False, 'local', enc, 64)
key, packed, G = bp.prepare(b, dev)
o2 = bp.run(key, packed, G)
tabs = [torch.randn((v, 300), device=dev, requires_grad=True) for v in (64, 5, 2, 119)]
idx = torch.stack([torch.randint(0, v, (N,), device=dev) for v in (64, 5, 2, 119)], 1)
eb = ops.embedding_bag(idx, tabs)
xx = torch.randn((N, 300), device=dev, requires_grad=True)
eff = torch.randn((E, 300), device=dev, requires_grad=True)
agg = ops.ogb_aggregate_ad(ei, N, 'source_to_target', xx, eb, False, eff, torch.zeros(1, device=dev))
lin = torch.nn.Linear(300, 600).to(dev)
(ops.linear_ad(agg.repeat(8, 1), lin.weight, lin.bias).square().sum()).backward()
torch.cuda.synchronize()
print('sanitize: all kernels ran', float(y.sum()), float(h.grad.abs().sum()), float((o1 - o2).abs().max()), float(xx.grad.abs().sum()))
