"""One eager COUNT + encode + forward step of the bench workload between cudaProfilerStart/Stop, for
  ncu --profile-from-start off [--set full -k regex:fused_model_kernel -c 1 | --metrics gpu__time_duration.sum] \
      python scripts/ncu_step.py --batch 131072
(the numbers printed by a run under ncu are never bench values)."""
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
    ap.add_argument('--batch', type=int, default=131072)
    a = ap.parse_args()
    from gsn_b200 import counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    dev = torch.device('cuda', 0)
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']), sds, False,
                               'local', max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **bench.model_args(enc.d)).to(dev).eval()
    pipe = GSNPipeline(model, sds, False, 'local', enc, 64)
    b = bench.build_batches(a.batch, 1, seed0=5)[0]
    t = bench.to_tensors(b, device=dev)
    with torch.no_grad():
        for _ in range(2):
            pipe.step(t)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        pipe.step(t)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print('N', int(b['node_ptr'][-1]), 'E', int(b['edge_index'].shape[1]))


if __name__ == '__main__':
    main()
