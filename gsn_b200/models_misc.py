"""Dense tail of every layer: `mlp` = Linear -> [BatchNorm1d] -> act, ..., Linear.

Same constructor and the same parameter names (fc.{i}.*, bn.{i}.*) as
/root/reference/models_misc.py:18-59, so reference checkpoints load unchanged.
Inference on CUDA tensors runs on the library's own dense-tail kernel
(gsn_tc_linear_fwd: tcgen05 3xTF32 with bias, eval-mode BatchNorm as scale/shift
and the activation in the epilogue; `forward(x, x2)` multiplies cat(x, x2)
without materialising the cat).  Training (autograd) and CPU tensors use the
torch modules (cuBLAS SGEMM, TF32 off).  The class also exposes the accessors
the fused message kernels use (first-layer split, BatchNorm as scale/shift).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

_ACTS = {'elu': nn.ELU, 'relu': nn.ReLU, 'tanh': nn.Tanh}


def choose_activation(activation):
    """models_misc.py:5-15"""
    if activation in _ACTS:
        return _ACTS[activation]()
    if activation == 'identity':
        return lambda x: x
    raise NotImplementedError


class mlp(nn.Module):

    def __init__(self, in_features, out_features, d_k, seed, activation='elu', batch_norm=False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.d_k, self.seed = d_k, seed
        self.activation_name, self.batch_norm = activation, batch_norm
        widths = [in_features] + list(d_k) + [out_features]
        n_lin = len(widths) - 1
        self.fc = nn.ModuleList(nn.Linear(widths[i], widths[i + 1], bias=True) for i in range(n_lin))
        self.bn = nn.ModuleList(nn.BatchNorm1d(widths[i + 1]) for i in range(n_lin - 1) if batch_norm)
        self.activation = choose_activation(activation)

    @property
    def depth(self):
        return len(self.fc)

    def hidden(self, x, i):
        """activation(bn_i(fc_i(x)))"""
        x = self._fc(i, x)
        if self.batch_norm:
            x = self.bn[i](x)
        return self.activation(x)

    def _fc(self, i, x):
        """Linear i on the autograd path: forward + input gradient on the library's tensor-core kernel for CUDA rows"""
        if x.is_cuda and x.dim() == 2:
            from . import ops
            return ops.linear_ad(x, self.fc[i].weight, self.fc[i].bias)
        return self.fc[i](x)

    def _own_kernels(self, x) -> bool:
        """eval / no-grad forward on a CUDA tensor: nothing to differentiate, BatchNorm is a fixed affine"""
        if not x.is_cuda or x.dim() != 2 or x.shape[0] == 0:
            return False
        if self.batch_norm and self.training:
            return False
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def forward(self, x, x2=None):
        """x2: optional second input block; the result equals forward(cat(x, x2, -1)) (models_misc.py:52-59)"""
        if self._own_kernels(x):
            from . import ops
            x = x.float().contiguous()
            x2 = None if x2 is None else x2.float().contiguous()
            with torch.no_grad():
                for i in range(len(self.fc) - 1):
                    s, t = self.bn_affine(i)
                    x = ops.linear(x, self.fc[i].weight, bias=self.fc[i].bias, A2=x2, scale=s, shift=t,
                                   activation=self.activation_name)
                    x2 = None
                return ops.linear(x, self.fc[-1].weight, bias=self.fc[-1].bias, A2=x2)
        if x2 is not None:
            x = torch.cat((x, x2), -1)
        for i in range(len(self.fc) - 1):
            x = self.hidden(x, i)
        return self._fc(len(self.fc) - 1, x)

    def bn_affine(self, i, batch_stats=None):
        """BatchNorm i as per-channel (scale, shift): y = h*scale + shift.
        batch_stats = (mean, biased var, count) switches to training-mode
        statistics and updates the running buffers like nn.BatchNorm1d."""
        if not self.batch_norm:
            return None, None
        bn = self.bn[i]
        if batch_stats is None:
            # eval: a fixed affine, cached until the module's tensors change (no elementwise launches per forward)
            stamp = (bn.running_mean.data_ptr(), bn.running_mean._version, bn.running_var._version,
                     bn.weight._version if bn.affine else 0, bn.bias._version if bn.affine else 0)
            cache = self.__dict__.setdefault('_affine_cache', {})
            hit = cache.get(i)
            if hit is not None and hit[0] == stamp:
                return hit[1], hit[2]
            with torch.no_grad():
                inv = torch.rsqrt(bn.running_var.float() + bn.eps)
                scale = (bn.weight * inv if bn.affine else inv).contiguous()
                shift = ((bn.bias if bn.affine else 0) - bn.running_mean.float() * scale).contiguous()
            if not (torch.is_grad_enabled() and bn.affine and bn.weight.requires_grad):
                cache[i] = (stamp, scale, shift)
                return scale, shift
            mean, var = bn.running_mean, bn.running_var
        else:
            mean, var, count = batch_stats
            with torch.no_grad():
                m = bn.momentum
                if bn.track_running_stats:
                    bn.num_batches_tracked += 1
                    if m is None:
                        m = 1.0 / float(bn.num_batches_tracked)
                    unbiased = var * (count / max(count - 1, 1))
                    bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
                    bn.running_var.mul_(1 - m).add_(unbiased.to(bn.running_var.dtype), alpha=m)
        inv = torch.rsqrt(var.float() + bn.eps)
        scale = bn.weight * inv if bn.affine else inv
        shift = (bn.bias if bn.affine else 0) - mean.float() * scale
        return scale, shift
