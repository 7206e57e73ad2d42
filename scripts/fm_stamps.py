"""Phase timing inside fused_model_kernel (first tile of CTA 0) from the profiling build:
    python -m gsn_b200.build --profile && python scripts/fm_stamps.py [--batch 128]
Stamps (clock64): 0 layer start, 1 constants staged, 2 P_j in smem, 3 message loop done, 4 Ux GEMM done,
5 S operand written, 6 S Wf^T GEMM done, 7 H operand written, 8 H U2^T GEMM done, 9 layer end."""
import argparse
import contextlib
import ctypes
import io
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
os.environ['GSN_B200_LIB'] = os.path.join(ROOT, 'gsn_b200', 'libgsn_b200_prof.so')
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

NAMES = ['stage consts', 'P_j epilogue (+GEMM wait)', 'message loop (+P_i wait)', 'wait Ux GEMM', 'write S', 'wait S.Wf GEMM',
         'H epilogue + write', 'wait H.U2 GEMM', 'x epilogue + pool + write']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    a = ap.parse_args()
    from gsn_b200 import _lib, counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    dev = torch.device('cuda', 0)
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']), sds,
                               False, 'local', max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **bench.model_args(enc.d)).to(dev).eval()
    pipe = GSNPipeline(model, sds, False, 'local', enc, 64, fused='model')
    t = bench.to_tensors(bench.build_batches(a.batch, 1, seed0=5)[0], device=dev)
    with torch.no_grad():
        for _ in range(3):
            pipe.step(t)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (8 * 16))()
    fn = _lib.lib().gsn_fm_profile_read
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p]
    assert fn(ctypes.cast(buf, ctypes.c_void_p)) == 0
    st = [[buf[l * 16 + i] for i in range(10)] for l in range(len(model.conv))]
    for l, s in enumerate(st):
        print(f'layer {l}: total {s[9] - s[0]} cycles')
        for i in range(9):
            print(f'   {NAMES[i]:34s} {s[i + 1] - s[i]:8d}')
        print(f'   (of the message phase: wait for the P_i GEMM {buf[l * 16 + 10] - s[2]})')
    print('tile total', st[-1][9] - st[0][0])


if __name__ == '__main__':
    main()
