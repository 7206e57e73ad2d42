"""Dense tail of every layer: `mlp` = Linear -> [BatchNorm1d] -> act, ..., Linear.

Same constructor and the same parameter names (fc.{i}.*, bn.{i}.*) as
/root/reference/models_misc.py:18-59, so reference checkpoints load unchanged.
The GEMMs themselves are plain library GEMMs (cuBLAS through torch, TF32 off);
what this class adds over the reference are the accessors the fused message
kernels use (first-layer split, BatchNorm as scale/shift).
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
        x = self.fc[i](x)
        if self.batch_norm:
            x = self.bn[i](x)
        return self.activation(x)

    def forward(self, x):
        for i in range(len(self.fc) - 1):
            x = self.hidden(x, i)
        return self.fc[-1](x)

    def bn_affine(self, i, batch_stats=None):
        """BatchNorm i as per-channel (scale, shift): y = h*scale + shift.
        batch_stats = (mean, biased var, count) switches to training-mode
        statistics and updates the running buffers like nn.BatchNorm1d."""
        if not self.batch_norm:
            return None, None
        bn = self.bn[i]
        if batch_stats is None:
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
