"""DGN aggregation kernel (molhiv_10_runs.sh recipe: 7 aggregators, edge-scope cycle counts k<=6 as the vector field)
on a large ZINC-shaped batch: device time (CUDA events, L2 flushed) against the HBM roofline.
python scripts/bench_dgn.py --batch 131072"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from gsn_b200 import directional, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=131072)
    ap.add_argument('--d', type=int, default=60)
    a = ap.parse_args()
    dev = torch.device('cuda')
    b = bench.build_batches(a.batch, 1, seed0=5)[0]
    ei = torch.from_numpy(b['edge_index']).to(dev)
    N, E = int(b['node_ptr'][-1]), ei.shape[1]
    plan = ops.EdgePlan(ei, N)
    g = torch.Generator(device=dev).manual_seed(0)
    h = torch.randn((N, a.d), device=dev, generator=g)
    ef = torch.randint(0, 4, (E, 4), device=dev, generator=g).float()
    names = 'mean max min dir0-av dir1-av dir2-av dir3-av'
    A = len(names.split())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, _ = bench.peaks()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        directional.dgn_aggregate(plan, h, None, ef, names, 'identity')
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts[2:])[2]
    by = 4 * N * a.d + 4 * E * 4 + 8 * E + 4 * (N + 1) + 4 * N * A * a.d
    print(json.dumps({'kernel': 'dgn_aggregate_kernel', 'batch': a.batch, 'N': N, 'E': E, 'd': a.d, 'aggregators': names,
                      'us': t * 1e6, 'algorithmic_bytes': by, 'GBps': by / t / 1e9, 'frac_of_hbm_peak': by / t / 1e9 / peak}))


if __name__ == '__main__':
    main()
