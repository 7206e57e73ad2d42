"""A/B of the tight message kernels against the lean ones they replaced (suffix x = GSN_NO_TIGHT, n = no scale/shift
operands, as fused.py calls them) at a given batch: time + checksum of the output bits.
python scripts/p1_variants.py --batch 131072 --variants 0nx,0n"""
import argparse, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)


def child(batch):
    import torch
    import bench
    from gsn_b200 import ops
    from bench_scatter import timeit
    dev = torch.device('cuda')
    b = bench.build_batches(batch, 1, seed0=5)[0]
    ei = torch.from_numpy(b['edge_index']).to(dev)
    N, E, dh = int(b['node_ptr'][-1]), ei.shape[1], 128
    plan = ops.EdgePlan(ei, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, _ = bench.peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn((N, 2 * dh), device=dev, generator=g)
    sc, sf = torch.rand(dh, device=dev, generator=g) + 0.5, torch.randn(dh, device=dev, generator=g)
    er1 = torch.randint(0, 4, (E, 1), device=dev, dtype=torch.int32, generator=g)
    Te1 = torch.randn((4, dh), device=dev, generator=g)
    if os.environ.get('GSN_P1_NOAFFINE'):
        sc = sf = None
    fn = lambda: ops.general_edge_idx(plan, dh, P=P, edge_rows=er1, Te=Te1, scale=sc, shift=sf, edge_rows_csr=True)
    S = fn()
    t = timeit(fn, flush)
    by = 4 * 2 * dh * N + 4 * E + 4 * dh * N + 8 * E + 4 * (N + 1)
    chk = S.view(torch.int32).to(torch.int64).sum().item()
    tag = ' no-tight' if os.environ.get('GSN_NO_TIGHT') else ''
    nr = torch.randint(0, 28, (N, 1), device=dev, dtype=torch.int32, generator=g)
    Tn = torch.randn((28, 2 * dh), device=dev, generator=g)
    er7 = torch.randint(0, 39, (E, 7), device=dev, dtype=torch.int32, generator=g)
    Te7 = torch.randn((39, dh), device=dev, generator=g)
    fn0 = lambda: ops.general_edge_idx(plan, dh, node_rows=nr, Tn=Tn, edge_rows=er7, Te=Te7, scale=sc, shift=sf, edge_rows_csr=True)
    S0 = fn0()
    t0 = timeit(fn0, flush)
    by0 = 4 * N + 28 * E + 4 * dh * N + 8 * E + 4 * (N + 1)
    print(f'layer-0 tab{tag}{" noaffine" if sc is None else ""}: {t0*1e6:8.1f} us  {by0/t0/1e9:7.1f} GB/s  bits-checksum {S0.view(torch.int32).to(torch.int64).sum().item()}')
    print(f'variant {os.environ.get("GSN_P1_VARIANT", "0")}{tag}{" noaffine" if sc is None else ""}: {t*1e6:8.1f} us  {by/t/1e9:7.1f} GB/s  {by/t/1e9/peak*100:5.1f}%  bits-checksum {chk}', flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=131072)
    ap.add_argument('--child', action='store_true')
    ap.add_argument('--variants', default='0nx,0n')
    a = ap.parse_args()
    if a.child:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        child(a.batch)
    else:
        for v in a.variants.split(','):
            subprocess.run([sys.executable, os.path.abspath(__file__), '--child', '--batch', str(a.batch)],
                           env=dict(os.environ, GSN_P1_VARIANT=v.rstrip('nx'), **({'GSN_P1_NOAFFINE': '1'} if 'n' in v else {}),
                                    **({'GSN_NO_TIGHT': '1'} if 'x' in v else {})))
