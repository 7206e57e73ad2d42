"""Phase timing inside count_small_kernel (every CTA) from the profiling build:
    python -m gsn_b200.build --profile && python scripts/cs_stamps.py [--batch 128]
Stamps (clock64, thread 0 of each CTA): 0 start, 1 searches + graph offsets, 2 adjacency + slot offsets, 3 pass set-up + edge_dict,
4 search done, 5 write-out done."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
os.environ['GSN_B200_LIB'] = os.path.join(ROOT, 'gsn_b200', 'libgsn_b200_prof.so')
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402

NAMES = ['searches + offsets', 'adjacency + scan', 'pass set-up + edge_dict', 'search', 'write-out']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--flush', default='write', choices=['write', 'write+read', 'none'],
                    help='L2 state before the launch: full of DIRTY lines (memset), clean (memset, then read another buffer), warm')
    a = ap.parse_args()
    from gsn_b200 import _lib, counting, patterns
    dev = torch.device('cuda', 0)
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
    b = bench.build_batches(a.batch, 1, seed0=5)[0]
    ei, ptr = torch.from_numpy(b['edge_index']).to(dev), torch.from_numpy(b['node_ptr'])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if a.flush == 'write+read' else None
    for _ in range(3):
        if a.flush != 'none':
            flush.zero_()
        if flush2 is not None:
            flush2.view(torch.int64).sum()
        counting.count_batch(ei, ptr, sds, False, 'local', max_nodes_per_graph=64, check=False)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (256 * 8))()
    fn = _lib.lib().gsn_cs_profile_read
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
    assert fn(ctypes.cast(buf, ctypes.c_void_p)) == 0
    st = np.array(buf, dtype=np.int64).reshape(256, 8)
    st = st[st[:, 5] > 0]
    d = np.diff(st[:, :6], axis=1)
    tot = st[:, 5] - st[:, 0]
    print(f'{len(st)} CTAs; total cycles per CTA: mean {tot.mean():.0f}, max {tot.max()} (CTA {int(tot.argmax())})')
    for i, n in enumerate(NAMES):
        print(f'  {n:26s} mean {d[:, i].mean():8.0f}  max {d[:, i].max():8d}  in the slowest CTA {d[tot.argmax(), i]:8d}')
    print('kernel span (min start -> max end):', int(st[:, 5].max() - st[:, 0].min()), 'cycles (SM clocks are not synchronised: indicative)')


if __name__ == '__main__':
    main()
