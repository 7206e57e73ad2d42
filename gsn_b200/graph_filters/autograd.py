"""Differentiable (training) path of the sparse layers.

The inference path (base.py) fuses gather+concat+transform+sum into single
kernels.  Training needs gradients w.r.t. x, identifiers, edge features and the
MLP weights; here the scatter-add and the row gather are autograd Functions
backed by the library's deterministic segment-sum (forward AND backward: the
backward of a gather is a segment-sum over the transposed grouping, the backward
of a segment-sum is a gather), and the message is formed as in the reference
(graph_filters/GSN_edge_sparse.py:152-170 etc.) so that BatchNorm inside msg_fn
sees the same E rows.
"""
from __future__ import annotations

import torch

from .. import ops


FUSED_OGB = True        # 'ogb' message kind: fused forward + fused backward (gsn_mp_ogb_fwd / gsn_mp_ogb_bwd)
PURE_TORCH = False      # measurement aid (scripts/bench_ogb.py --impl eager_torch): the reference's formulation op by op
#                         in eager PyTorch (index_select gathers, index_add_ scatter) -- the incumbent on the same GPU


def segment_sum(msgs, plan):
    if PURE_TORCH:
        out = torch.zeros((plan.N, msgs.shape[1]), dtype=msgs.dtype, device=msgs.device)
        return out.index_add_(0, plan.edge_index[plan.select], msgs)
    return ops.segment_sum_ad(plan, msgs)


def gather_rows(x, edge_index, row):
    if PURE_TORCH:
        return x.index_select(0, edge_index[row])
    return ops.gather_rows_ad(x, edge_index, row)


def forward_with_grad(layer, x, edge_index, identifiers, ef):
    n = x.shape[0]
    plan = ops.edge_plan(edge_index, n, layer.flow)
    sel = plan.select
    x = x.float()
    x_j = gather_rows(x, edge_index, 1 - sel)
    local = layer.id_scope == 'local'
    if layer.msg_kind == 'gin':
        self_parts, msg_parts = [x], [x_j]
        if layer.uses_ids:
            if local:
                id_ii, id_nb = layer.central_node_id_encoder(identifiers, n)
                self_parts.append(id_ii)
                msg_parts.append(id_nb)
            else:
                self_parts.append(identifiers)
                msg_parts.append(gather_rows(identifiers.float(), edge_index, 1 - sel))
        if layer.uses_ef:
            ef_ii, ef_nb = layer.central_node_edge_encoder(ef, n)
            self_parts.append(ef_ii)
            msg_parts.append(ef_nb)
        agg = segment_sum(torch.cat(msg_parts, -1).float(), plan)
        return layer.update_fn((1 + layer.eps) * torch.cat(self_parts, -1) + agg)
    if (layer.msg_kind == 'ogb' and FUSED_OGB and not PURE_TORCH and x.is_cuda and not layer.eps.requires_grad
            and ef is not None):
        # fused forward + fused backward (relu mask recomputed): gsn_mp_ogb_fwd / gsn_mp_ogb_bwd
        agg = ops.ogb_aggregate_ad(edge_index, n, layer.flow, x, identifiers if layer.uses_ids else None, local, ef.float(),
                                   layer.eps)
        return layer.update_fn(agg)
    if layer.msg_kind == 'ogb':
        self_msg = x
        m = x_j
        if layer.uses_ids:
            if local:
                m = m + identifiers
            else:
                self_msg = x + identifiers
                m = m + gather_rows(identifiers.float(), edge_index, 1 - sel)
        m = torch.relu(m + ef)
        return layer.update_fn((1 + layer.eps) * self_msg + segment_sum(m, plan))
    # general
    x_i = gather_rows(x, edge_index, sel)
    parts = [x_i, x_j]
    if layer.uses_ids:
        if local:
            parts.append(identifiers.float())
        else:
            idf = identifiers.float()
            parts += [gather_rows(idf, edge_index, sel), gather_rows(idf, edge_index, 1 - sel)]
    if layer.uses_ef:
        parts.append(ef.float())
    msgs = layer.msg_fn(torch.cat(parts, -1))
    return layer.update_fn(torch.cat((x, segment_sum(msgs, plan)), -1))
