import sys, os, ctypes, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gsn_b200 import ops, _lib
torch.manual_seed(0)
for (M, K, N) in [(2924, 256, 128), (128, 128, 128), (2924, 128, 256), (128, 128, 64), (300000, 256, 128)]:
    A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda')
    for _ in range(3): ops.linear(A, W)
    buf = torch.zeros((4096, 8), dtype=torch.int64, device='cuda')
    _lib.lib().gsn_tc_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ops.linear(A, W); e1.record(); torch.cuda.synchronize()
    if os.environ.get('NOSTORE'):
        import gsn_b200.ops as _o
        buf.zero_(); out2 = torch.zeros_like(out); _o.linear(A, W, out=out2, accumulate=2); torch.cuda.synchronize()
    _lib.lib().gsn_tc_debug_buffer(None)
    b = buf.cpu()[:min(4096, (M + 127) // 128)]
    d = (b[:, 1:7] - b[:, 0:1]).float()
    print('   chunk0: tmem_full', (b[:,4]-b[:,0]).float().mean().item(), 'after ld', (b[:,7]-b[:,0]).float().mean().item(), 'after transpose', (b[:,1]-b[:,0]).float().mean().item(), 'after stores', (b[:,2]-b[:,0]).float().mean().item())
    print(M, K, N, 'event us', e0.elapsed_time(e1) * 1e3, 'cycles: alloc', d[:, 0].mean().item(), 'tma_issued', d[:, 1].mean().item(),
          'mma_issued', d[:, 2].mean().item(), 'tmem_full', d[:, 3].mean().item(), 'epi_done', d[:, 4].mean().item(), 'sync', d[:, 5].mean().item())
    ref = A.double() @ W.double().t()
    print('   maxerr', (out.double() - ref).abs().max().item())
