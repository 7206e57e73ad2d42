"""Micro-benchmark of the scatter kernels at a given batch size (CUDA events, L2 flushed between launches).
python scripts/bench_scatter.py --batch 131072"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
from gsn_b200 import ops


def timeit(fn, flush, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts[2:])[len(ts[2:]) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=131072)
    a = ap.parse_args()
    dev = torch.device('cuda')
    b = bench.build_batches(a.batch, 1, seed0=5)[0]
    ei = torch.from_numpy(b['edge_index']).to(dev)
    N, E, dh = int(b['node_ptr'][-1]), ei.shape[1], 128
    plan = ops.EdgePlan(ei, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, _ = bench.peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn((N, 2 * dh), device=dev, generator=g)
    Q = torch.randn((E, dh), device=dev, generator=g)
    sc, sf = torch.rand(dh, device=dev) + 0.5, torch.randn(dh, device=dev)
    csr = 8 * E + 4 * (N + 1)
    t = timeit(lambda: ops.general_edge(plan, P, Q, sc, sf), flush)
    by = 4 * 2 * dh * N + 4 * dh * E + 4 * dh * N + csr
    print(f'general_edge dense P+Q      : {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}% of {peak}')
    # layer >= 1 of the fused path: dense P, ef as 1 index column
    er1 = torch.randint(0, 4, (E, 1), device=dev, dtype=torch.int32)
    Te1 = torch.randn((4, dh), device=dev)
    t = timeit(lambda: ops.general_edge_idx(plan, dh, P=P, edge_rows=er1, Te=Te1, scale=sc, shift=sf, edge_rows_csr=True), flush)
    by = 4 * 2 * dh * N + 4 * E + 4 * dh * N + csr
    print(f'general_edge_idx P + 1 col  : {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}%')
    # layer 0: x index + 7 edge index columns
    nr = torch.randint(0, 28, (N, 1), device=dev, dtype=torch.int32)
    Tn = torch.randn((28, 2 * dh), device=dev)
    er7 = torch.randint(0, 39, (E, 7), device=dev, dtype=torch.int32)
    Te7 = torch.randn((39, dh), device=dev)
    t = timeit(lambda: ops.general_edge_idx(plan, dh, node_rows=nr, Tn=Tn, edge_rows=er7, Te=Te7, scale=sc, shift=sf, edge_rows_csr=True), flush)
    by = 4 * N + 28 * E + 4 * dh * N + csr
    print(f'general_edge_idx idx-only   : {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}%')
    x = torch.randn((N, dh), device=dev)
    ef = torch.randn((E, dh), device=dev)
    t = timeit(lambda: ops.ogb_aggregate(plan, x, ef, True, ef, None), flush)
    by = 4 * dh * (2 * N + 2 * E) + csr
    print(f'ogb_aggregate (local ids)   : {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}%')
    t = timeit(lambda: ops.segment_sum(plan, ef), flush)
    by = 4 * dh * (N + E) + 4 * E + 4 * (N + 1)
    print(f'segment_sum [E,128]         : {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}%')


if __name__ == '__main__':
    main()
