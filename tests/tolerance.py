"""ONE tolerance policy for every fp32 comparison in the GPU tests.

BASELINE.json (north_star): "forward activations within 1e-5 fp32 of the reference PyG path".  fp32 results of the same
formula evaluated in two summation orders differ by a few ulp of the LARGEST magnitude involved, so the bound is stated
relative to the scale of the compared tensor:

    forward values   |got - ref| <= FWD  * max(1, max|ref|)  +  FWD  * |ref|        FWD  = 1e-5
    gradients        |got - ref| <= GRAD * max(1, max|ref|)  +  GRAD * |ref|        GRAD = 1e-4

(gradients: every element is a sum over up to E rows of products of two fp32-accurate factors, and the golden gradients
come from autograd through the reference's own CPU code in yet another summation order).  Integer results (COUNT,
encoders, CSR) are always compared bit-exactly.  No test uses any other constant.
"""
import torch

FWD = 1e-5
GRAD = 1e-4


def close(got, ref, tol=FWD, msg=None):
    got, ref = got.detach(), ref.detach()
    if got.device != ref.device:
        got, ref = got.cpu(), ref.cpu()
    scale = max(float(ref.abs().max()) if ref.numel() else 0.0, 1.0)
    kw = {} if msg is None else {'msg': (lambda m: f'{msg}: {m}')}
    torch.testing.assert_close(got, ref, atol=tol * scale, rtol=tol, **kw)
