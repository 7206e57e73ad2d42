"""Per-call device time of one bench step (eager, CUDA events around every C-ABI call) and the latency of the
captured step, at several batch sizes.  python scripts/step_breakdown.py [--batches 128,4096,131072] [--fused model|layers]"""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='128,4096,131072')
    ap.add_argument('--fused', default='model')
    ap.add_argument('--scope', default='local')
    a = ap.parse_args()
    from gsn_b200 import _lib, counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    dev = torch.device('cuda', 0)
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), a.scope)
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']), sds,
                               False, a.scope, max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids)
    torch.manual_seed(0)
    args = bench.model_args(enc.d)
    args['id_scope'] = a.scope
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **args).to(dev).eval()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for B in [int(x) for x in a.batches.split(',')]:
        pipe = GSNPipeline(model, sds, False, a.scope, enc, 64, fused=a.fused)
        t = bench.to_tensors(bench.build_batches(B, 1, seed0=5)[0], device=dev)
        with torch.no_grad():
            for _ in range(3):
                pipe.step(t)
            torch.cuda.synchronize()
            _lib.TIMER = []
            for _ in range(5):
                flush.zero_()
                pipe.step(t)
            torch.cuda.synchronize()
            agg = {}
            for tag, e0, e1 in _lib.TIMER:
                agg.setdefault(tag, []).append(e0.elapsed_time(e1) * 1e3)
            _lib.TIMER = None
        rows = {k: (round(float(np.median(v)), 1), len(v) // 5) for k, v in agg.items()}
        l0 = _lib.launch_count()
        with torch.no_grad():
            pipe.step(t)
        launches = _lib.launch_count() - l0
        res = {'batch': B, 'N': int(t['x'].shape[0]), 'E': int(t['edge_index'].shape[1]), 'own_launches': launches,
               'per_call_us(median, calls/step)': rows}
        try:
            pipe.capture(t, warmup=3)
            for _ in range(3):
                pipe.replay()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pipe.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            res['captured_step_us'] = round(float(np.median(ts)), 1)
            res['graphs_per_s'] = B / (np.median(ts) * 1e-6)
        except Exception as ex:
            res['capture_error'] = repr(ex)[:200]
        print(json.dumps(res), flush=True)
        del pipe, t
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
