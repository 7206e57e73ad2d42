"""Runs the bench step eagerly (no CUDA graph) a few times at a given batch size: the command
ncu wraps (B200_PROFILING.md).  python scripts/profile_step.py --batch 32768 --iters 3"""
import argparse
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32768)
    ap.add_argument('--iters', type=int, default=3)
    a = ap.parse_args()
    from gsn_b200 import counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    dev = torch.device('cuda', 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']), sds,
                               False, 'local', max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **bench.model_args(enc.d)).to(dev).eval()
    pipe = GSNPipeline(model, sds, False, 'local', enc, 64)
    t = bench.to_tensors(bench.build_batches(a.batch, 1, seed0=5)[0], device=dev)
    with torch.no_grad():
        for _ in range(a.iters):
            pipe.step(t)
    torch.cuda.synchronize()
    print('done', a.batch, a.iters)


if __name__ == '__main__':
    main()
