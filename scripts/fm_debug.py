"""Debug aid for csrc/fused_model.cu: per-layer max |diff| of the one-kernel forward against the per-layer fused
path on golden models and ZINC-shaped batches.  Run on a GPU box:  python scripts/fm_debug.py"""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_fused_model_gpu as T      # noqa: E402


def report(tag, ref_interm, fm, out_ref, out_one):
    for i, (xr, xo) in enumerate(zip(ref_interm[1:], fm.last_x_out)):
        d = (xo[:, :xr.shape[1]] - xr).abs()
        print(f'  {tag} layer {i}: max|ref| {float(xr.abs().max()):.3e}  max|diff| {float(d.max()):.3e}  '
              f'rows>1e-4: {int((d.amax(1) > 1e-4).sum())}/{xr.shape[0]}  nan {int(torch.isnan(xo).sum())}', flush=True)
        if float(d.max()) > 1e-3:
            bad = torch.nonzero(d.amax(1) > 1e-3).flatten()[:8].tolist()
            print('    first bad rows', bad, 'cols', torch.nonzero(d[bad[0]] > 1e-3).flatten()[:8].tolist())
            print('    ref', xr[bad[0], :6].tolist(), 'got', xo[bad[0], :6].tolist())
    print(f'  {tag} out: max|ref| {float(out_ref.abs().max()):.3e} max|diff| {float((out_one - out_ref).abs().max()):.3e}', flush=True)


def main():
    from gsn_b200 import fused, fused_model
    from gsn_b200.pipeline import GSNPipeline
    for name in ['zinc_gsne_general', 'zinc_gsnv_general', 'sr_general_local_nobn', 'mpnn_general']:
        try:
            model, b, ref = T._golden_model(name)
            r = fused.FusedForward(model)
            out_ref = r(b)
            fm = fused_model.FusedModel(model)
            fm.debug_x_out = True
            out = fm(b)
            torch.cuda.synchronize()
            print(name, 'status', int(fm.status.item()), 'golden diff', float((out.cpu() - ref).abs().max()))
            report(name, r.last_x_interm, fm, out_ref, out)
        except Exception:
            traceback.print_exc()
            return 1
    for B, d_out, scope in [(128, 128, 'local'), (128, 64, 'local'), (128, 128, 'global'), (700, 128, 'local')]:
        try:
            model, sds, enc, b, t, _ = T._zinc_setup(B, 3, d_out, scope)
            with torch.no_grad():
                p_ref = GSNPipeline(model, sds, False, scope, enc, 64, fused='layers')
                p_one = GSNPipeline(model, sds, False, scope, enc, 64, fused='model')
                p_one.fused.debug_x_out = True
                out_ref = p_ref.step(t)
                out_one = p_one.step(t)
                torch.cuda.synchronize()
            print(f'zinc B={B} d={d_out} {scope}: status', int(p_one.fused.status.item()))
            report(f'zinc{B}/{d_out}/{scope}', p_ref.fused.last_x_interm, p_one.fused, out_ref, out_one)
            # timing
            p_one.fused.debug_x_out = False
            for pipe, tag in ((p_ref, 'layers'), (p_one, 'one-kernel')):
                data = None
                with torch.no_grad():
                    for _ in range(3):
                        pipe.step(t)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        pipe.step(t)
                    e1.record()
                    torch.cuda.synchronize()
                print(f'  eager step ({tag}): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us', flush=True)
        except Exception:
            traceback.print_exc()
            return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
