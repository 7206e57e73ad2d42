"""experiment: time the persistent GEMM with parts disabled (GSN_TC_MODE bits: 1 epilogue, 2 MMA, 4 split)"""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    from gsn_b200 import ops
    M = 3040896
    for K, N, two in ((128, 128, False), (128, 256, False), (256, 128, True)):
        A = torch.randn(M, 128, device='cuda'); A2 = torch.randn(M, 128, device='cuda') if two else None
        W = torch.randn(N, K, device='cuda'); b = torch.randn(N, device='cuda'); sc = torch.rand(N, device='cuda'); sf = torch.randn(N, device='cuda')
        out = torch.empty(M, N, device='cuda')
        fn = lambda: ops.linear(A, W, A2=A2, bias=b, scale=sc, shift=sf, activation='relu', out=out)
        for _ in range(2): fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        by = 4 * M * (K + N)
        t = sorted(ts)[2]
        print(f'mode {os.environ.get("GSN_TC_MODE","0")}: K={K} N={N}  {t*1e3:8.1f} us  {by/t/1e6:7.1f} GB/s', flush=True)
else:
    for m in sys.argv[1].split(','):
        subprocess.run([sys.executable, os.path.abspath(__file__), 'child'], env=dict(os.environ, GSN_TC_MODE=m))
