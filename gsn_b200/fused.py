"""Fused inference forward of GNNSubstructures ('general' message kind).

Same arithmetic as /root/reference/models_graph_classification.py:204-247 in eval
mode, re-associated so that the whole forward is a handful of library kernels:

  * categorical inputs stay indices: a one-hot block times a weight matrix is a
    table row (one_hot_encoder, utils_graph_learning.py:170-187, followed by the
    first Linear of msg_fn / update_fn), and one_hot_unique
    (utils_encoding.py:37-59) is a binary search over the sorted vocabulary, so
    raw identifier counts go straight from COUNT into the message kernel;
  * BatchNorm1d (eval) is a per-channel scale/shift: applied in the GEMM epilogues of
    update_fn / the model, folded into the weights and tables of msg_fn's first Linear;
    the second Linear of msg_fn commutes with the neighbour sum and is
    pre-multiplied into update_fn's first Linear:
        cat(x, sum_e(W2 h_e + b2)) U1^T = x U1x^T + S (U1a W2)^T + deg (U1a b2)
  * per layer: [P GEMM] -> message kernel -> update GEMM -> output GEMM (+ model BN + act).

Only inference (no autograd); training and every other configuration use the
per-layer path in graph_filters/.  Parity: tests/test_fused_gpu.py.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from .encoders import one_hot_encoder


MERGE_MAX_ROWS = 64          # joint vocabulary of a column group (rows of its pre-summed table)


def _merge_edge_columns(Te, cols, max_rows):
    """cols: [(vocabulary size, first row in Te)] in column order -> (merged table, dict(group, mult, off, n_groups)).
    Greedy packing of consecutive columns while the product of their vocabulary sizes stays <= max_rows; row
    sum_c digit_c * mult_c of a group's table holds sum_c Te[first_c + digit_c]."""
    groups, cur, prod = [], [], 1
    for ci, (dcol, _) in enumerate(cols):
        if cur and prod * dcol > max_rows:
            groups.append(cur)
            cur, prod = [], 1
        cur.append(ci)
        prod *= dcol
    groups.append(cur)
    tables, group_of, mult, off, base = [], [0] * len(cols), [1] * len(cols), [0] * len(cols), 0
    for gi, members in enumerate(groups):
        rows = 1
        for ci in members:
            rows *= cols[ci][0]
        idx = torch.arange(rows, device=Te.device)
        tab, m = None, 1
        for k, ci in enumerate(members):
            dcol, first = cols[ci]
            part = Te[first + (idx // m) % dcol]
            tab = part if tab is None else tab + part
            group_of[ci], mult[ci], off[ci] = gi, m, (base if k == 0 else 0)
            m *= dcol
        tables.append(tab)
        base += rows
    return torch.cat(tables, 0).contiguous(), {'group': group_of, 'mult': mult, 'off': off, 'n_groups': len(groups)}


def _is_onehot(enc) -> bool:
    return enc.encoder_name == 'one_hot_encoder'


def supported(model) -> bool:
    from .network import GNNSubstructures
    if not isinstance(model, GNNSubstructures) or model.random_features:
        return False
    ok_enc = lambda e: e.encoder_name in ('one_hot_encoder', 'None')
    if not ok_enc(model.input_node_encoder) or not all(ok_enc(e) for e in model.edge_encoder):
        return False
    if not all(_is_onehot(e) for e in model.id_encoder):
        return False
    for conv in model.conv:
        if conv.msg_kind != 'general' or conv.degree_as_tag or conv.aggr != 'add' or conv.msg_fn.depth != 2 \
                or conv.update_fn.depth != 2:
            return False
    if model.final_projection[0] and _is_onehot(model.input_node_encoder):
        return False
    return True


class FusedForward:

    def __init__(self, model):
        if not supported(model):
            raise NotImplementedError('fused forward: unsupported model configuration')
        self.model = model
        self._stamp = None
        self.layers: List[dict] = []

    # ------------------------------------------------------------------ weight preparation
    def _version_stamp(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.model.parameters()) + list(self.model.buffers()))

    @torch.no_grad()
    def prepare(self):
        m = self.model
        self.layers = []
        x_cat = _is_onehot(m.input_node_encoder)
        for i, conv in enumerate(m.conv):
            f, u = conv.msg_fn, conv.update_fn
            dh = f.fc[0].weight.shape[0]
            d_in = conv._dims[0]
            ee = m.edge_encoder[i if m.inject_edge_features else 0] if conv.uses_ef else None
            ie = m.id_encoder[i if m.inject_ids else 0] if conv.uses_ids else None
            ef_cat = ee is not None and _is_onehot(ee)
            d_id = ie.d_out if ie is not None else 0
            d_ef = ee.d_out if ee is not None else 0
            Wxi, Wxj, Wii, Wij, Wq_id, Wq_ef = conv._split_first_linear(d_in, d_id, d_ef)
            # BatchNorm (eval) of msg_fn + bias of its first Linear: ((P_i + P_j + Q) + b1) * s0 + t0.  The scale is folded
            # into every block of the first Linear and the shift into the P_i operand (GEMM bias / x-table rows), so the
            # message kernel is a bare act(P_i + P_j + rows): no per-edge scale / shift traffic on the L1 data pipe.
            s0, t0 = f.bn_affine(0)
            b1 = f.fc[0].bias
            shift = (b1 if s0 is None else b1 * s0 + t0).contiguous()
            if s0 is not None:
                Wxi, Wxj, Wii, Wij, Wq_id, Wq_ef = [None if W is None else W * s0[:, None]
                                                    for W in (Wxi, Wxj, Wii, Wij, Wq_id, Wq_ef)]
            L = {'dh': dh, 'd_in': d_in, 'act_mlp': conv.activation_name, 'x_cat': x_cat and i == 0, 'ef_cat': ef_cat,
                 'uses_ids': conv.uses_ids, 'uses_ef': conv.uses_ef, 'local': conv.id_scope == 'local', 'flow': conv.flow}
            # node side of the first Linear
            Wp = torch.cat((Wxi, Wxj), 0).contiguous()                  # [2dh, d_in]
            node_tables = []
            if L['x_cat']:
                tab = Wp.t().contiguous()                                # [d_in rows, 2dh]; exactly one row per node
                tab[:, :dh] += shift
                node_tables.append(tab)
                L['Wp'], L['bp'] = None, None
            else:
                L['Wp'], L['bp'] = Wp, torch.cat((shift, torch.zeros_like(shift))).contiguous()
            if conv.uses_ids and not L['local']:
                node_tables.append(torch.cat((Wii, Wij), 0).t().contiguous())   # [d_id rows, 2dh]
            L['Tn'] = torch.cat(node_tables, 0).contiguous() if node_tables else None
            L['Tn_off_ids'] = node_tables[0].shape[0] if (L['x_cat'] and len(node_tables) == 2) else 0
            # edge side
            edge_tables = []
            if conv.uses_ids and L['local']:
                edge_tables.append(Wq_id.t().contiguous())               # [d_id rows, dh]
            L['Te_off_ef'] = edge_tables[0].shape[0] if edge_tables else 0
            if conv.uses_ef:
                if ef_cat:
                    edge_tables.append(Wq_ef.t().contiguous())
                    L['Wq_ef'] = None
                    L['ef_rows'] = int(Wq_ef.shape[1])
                else:
                    L['Wq_ef'] = Wq_ef.contiguous()
            L['Te'] = torch.cat(edge_tables, 0).contiguous() if edge_tables else None
            # several categorical edge columns (identifier ranks, bond type): fold columns with a small joint vocabulary
            # into one pre-summed table per group, so the message kernel gathers one row per group instead of one per
            # column (ZINC layer 0: 7 columns -> 3 lookups per edge); gsn_encode_rows_grouped emits the mixed-radix rows
            L['edge_groups'] = None
            ecols = []
            if conv.uses_ids and L['local']:
                o = 0
                for dcol in ie.encoder.d_in:
                    ecols.append((int(dcol), o))
                    o += int(dcol)
            if conv.uses_ef and ef_cat:
                ecols.append((int(Wq_ef.shape[1]), L['Te_off_ef']))
            if len(ecols) > 1:
                L['Te'], L['edge_groups'] = _merge_edge_columns(L['Te'], ecols, MERGE_MAX_ROWS)
            # update_fn with the second message Linear folded in
            W2, b2 = f.fc[1].weight, f.fc[1].bias
            U1, c1 = u.fc[0].weight, u.fc[0].bias
            U1x, U1a = U1[:, :d_in], U1[:, d_in:]
            Wf = (U1a.double() @ W2.double()).float()                     # [dh_u, dh]
            L['vf'] = (U1a.double() @ b2.double()).float().contiguous()
            if L['x_cat']:
                L['Wu'] = Wf.contiguous()
                L['Tu'] = U1x.t().contiguous()                            # [d_in rows, dh_u]
            else:
                L['Wu'] = torch.cat((U1x, Wf), 1).contiguous()            # [dh_u, d_in + dh]
                L['Tu'] = None
            L['c1'] = c1.contiguous()
            su, tu = u.bn_affine(0)
            L['su'], L['tu'] = (None, None) if su is None else (su.contiguous(), tu.contiguous())
            L['U2'], L['c2'] = u.fc[1].weight.contiguous(), u.fc[1].bias.contiguous()
            if m.bn[i]:
                bn = m.batch_norms[i]
                inv = torch.rsqrt(bn.running_var + bn.eps)
                sm = bn.weight * inv
                L['sm'], L['tm'] = sm.contiguous(), (bn.bias - bn.running_mean * sm).contiguous()
            else:
                L['sm'], L['tm'] = None, None
            self.layers.append(L)
        self._act_model = {torch.nn.ReLU: 'relu', torch.nn.ELU: 'elu', torch.nn.Tanh: 'tanh'}.get(type(m.activation), 'identity')
        # JK projections
        self.proj = []
        for i, pr in enumerate(m.lin_proj):
            if not m.final_projection[i]:
                self.proj.append(None)
            elif isinstance(pr, torch.nn.Linear):
                self.proj.append(('linear', pr.weight.contiguous(), pr.bias.contiguous()))
            else:
                if pr.depth != 2:
                    raise NotImplementedError('JK mlp depth != 2')
                s, t = pr.bn_affine(0)
                self.proj.append(('mlp', pr.fc[0].weight.contiguous(), pr.fc[0].bias.contiguous(),
                                  None if s is None else s.contiguous(), None if t is None else t.contiguous(),
                                  pr.fc[1].weight.contiguous(), pr.fc[1].bias.contiguous(), pr.activation_name))
        self._stamp = self._version_stamp()

    # ------------------------------------------------------------------ categorical inputs -> table rows
    def _context(self, data, raw_identifiers, vocab, need_ids=True):
        """per-call state shared by the layers: dense x (or None), index columns, the concatenated vocabulary"""
        m = self.model
        if m.training:
            raise RuntimeError('FusedForward is inference-only')
        if self._stamp != self._version_stamp():
            self.prepare()
        # identifier rows into the one-hot table of the id encoder (offset inside the table added later)
        ids = (raw_identifiers if raw_identifiers is not None else data.identifiers) if need_ids else None
        id_dims = m.id_encoder[0].encoder.d_in
        id_off, o = [], 0
        for d in id_dims:
            id_off.append(o)
            o += int(d)
        if vocab is not None:
            key = tuple((v.data_ptr(), v.numel()) for v in vocab)
            if getattr(self, '_vocab_key', None) != key:           # concatenated once, not once per step
                if [int(v.numel()) for v in vocab] != [int(d) for d in id_dims]:
                    raise ValueError('vocabulary sizes differ from the d_in of the model\'s identifier encoder')
                self._vocab_key, self._vcat = key, torch.cat(vocab)
                self._vocab_keep = list(vocab)       # pins the tensors whose data_ptr is the cache key
                self._vptr, o2 = [], 0
                for v in vocab:
                    self._vptr.append((o2, o2 + v.numel()))
                    o2 += v.numel()
            vcat, vptr = self._vcat, self._vptr
        else:
            vcat, vptr = None, [None] * len(id_dims)
        x_cat = _is_onehot(m.input_node_encoder)
        xi = data.x if data.x.dim() == 2 else data.x.unsqueeze(-1)
        x = None if (x_cat or not need_ids) else m.input_node_encoder(data.x)
        efi = None
        if getattr(data, 'edge_features', None) is not None:
            efi = data.edge_features if data.edge_features.dim() == 2 else data.edge_features.unsqueeze(-1)
        dev0 = data.edge_index.device
        if getattr(self, 'status', None) is None or self.status.device != dev0:
            self.status = torch.zeros(1, dtype=torch.int32, device=dev0)
        return {'ids': ids, 'id_off': id_off, 'vcat': vcat, 'vptr': vptr, 'xi': xi, 'x': x, 'efi': efi, 'status': self.status,
                'N': data.x.shape[0], 'E': data.edge_index.shape[1], 'dev': data.edge_index.device, 'ef_rows_cache': {},
                'pre': self._take_prefetched(data) if need_ids else {}}

    # index rows that do not depend on the identifiers can be produced before COUNT has finished
    def prefetch_rows(self, data):
        """Call on a side stream while COUNT runs: encodes the rows of the categorical inputs that do not involve the
        identifiers (atom types; bond types of layers without edge identifiers).  The next __call__ on the same tensors
        picks them up; the caller joins the streams in between."""
        ctx = self._context(data, None, None, need_ids=False)
        rows = {}
        for li, L in enumerate(self.layers):
            ids_nodes, ids_edges = L['uses_ids'] and not L['local'], L['uses_ids'] and L['local']
            if L['x_cat'] and not ids_nodes:
                rows[('n', li)] = self._node_rows(L, ctx)
            if L['uses_ef'] and L['ef_cat'] and not ids_edges:
                plan = ops.edge_plan(data.edge_index, ctx['N'], L['flow'])
                rows[('e', L['Te_off_ef'], L['flow'])] = self._edge_rows(L, ctx, plan)
        self._prefetched = ((data.x.data_ptr(), data.edge_index.data_ptr()), rows)

    def _take_prefetched(self, data):
        pre = getattr(self, '_prefetched', None)
        self._prefetched = None
        if pre is not None and pre[0] == (data.x.data_ptr(), data.edge_index.data_ptr()):
            return pre[1]
        return {}

    def _node_rows(self, L, ctx):
        ids, vptr, id_off, vcat = ctx['ids'], ctx['vptr'], ctx['id_off'], ctx['vcat']
        m = self.model
        node_cols = []
        if L['x_cat']:
            node_cols.append((ctx['xi'][:, 0], None, 0, int(m.input_node_encoder.encoder.d_in[0])))
        if L['uses_ids'] and not L['local']:
            id_dims = m.id_encoder[0].encoder.d_in
            node_cols += [(ids[:, c], vptr[c], id_off[c] + L['Tn_off_ids'], 0 if vptr[c] is not None else int(id_dims[c]))
                          for c in range(ids.shape[1])]
        return ops.encode_rows(node_cols, vcat, ctx['N'], ctx['dev'], status=ctx['status']) if node_cols else None

    def _edge_rows(self, L, ctx, plan):
        ids, vptr, id_off, vcat, dev = ctx['ids'], ctx['vptr'], ctx['id_off'], ctx['vcat'], ctx['dev']
        m = self.model
        edge_cols = []
        if L['uses_ids'] and L['local']:
            id_dims = m.id_encoder[0].encoder.d_in
            edge_cols += [(ids[:, c], vptr[c], id_off[c], 0 if vptr[c] is not None else int(id_dims[c]))
                          for c in range(ids.shape[1])]
        if L['uses_ef'] and L['ef_cat']:
            edge_cols.append((ctx['efi'][:, 0], None, L['Te_off_ef'], int(L['ef_rows'])))
        eg = L['edge_groups']
        if eg is not None:
            edge_cols = [(s, v, eg['off'][c], r) for c, (s, v, _, r) in enumerate(edge_cols)]
        st = ctx['status']
        # edge rows are produced directly in CSR order (perm = plan.eid): the message kernel then reads them
        # sequentially instead of chasing eid -> row
        only_ef = len(edge_cols) == 1 and L['uses_ef'] and L['ef_cat']
        key = (L['Te_off_ef'], L['flow'])
        cache = ctx['ef_rows_cache']
        if only_ef and key in cache:
            return cache[key]
        if eg is not None:
            edge_rows = ops.encode_rows_grouped(edge_cols, eg['group'], eg['mult'], eg['n_groups'], vcat, ctx['E'], dev,
                                                perm=plan.eid, status=st)
        else:
            edge_rows = ops.encode_rows(edge_cols, vcat, ctx['E'], dev, perm=plan.eid, status=st) if edge_cols else None
        if only_ef:
            cache[key] = edge_rows
        return edge_rows

    def _layer_rows(self, L, ctx, plan):
        """(node_rows int32 [N, n_node_cols] | None, edge_rows int32 [E, n_groups] in CSR order | None) of one layer"""
        pre = ctx['pre']
        li = next(i for i, l_ in enumerate(self.layers) if l_ is L)
        node_rows = pre[('n', li)] if ('n', li) in pre else self._node_rows(L, ctx)
        ekey = ('e', L['Te_off_ef'], L['flow'])
        ids_edges = L['uses_ids'] and L['local']
        edge_rows = pre[ekey] if (ekey in pre and not ids_edges) else self._edge_rows(L, ctx, plan)
        return node_rows, edge_rows

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, data, raw_identifiers: Optional[torch.Tensor] = None, vocab: Optional[List[torch.Tensor]] = None):
        """data: same attribute bag as GNNSubstructures.forward plus `node_ptr` (int64 [G+1]).
        raw_identifiers/vocab: un-encoded COUNT output and the one_hot_unique vocabulary; when omitted,
        data.identifiers must hold the encoded ranks (as in the reference)."""
        m = self.model
        ctx = self._context(data, raw_identifiers, vocab)
        edge_index, N, node_ptr = data.edge_index, ctx['N'], data.node_ptr
        x = ctx['x']
        x_interm = [x]
        for i, (conv, L) in enumerate(zip(m.conv, self.layers)):
            plan = ops.edge_plan(edge_index, N, L['flow'])
            dh = L['dh']
            node_rows, edge_rows = self._layer_rows(L, ctx, plan)
            # ---- dense parts
            P = ops.linear(x, L['Wp'], bias=L['bp']) if L['Wp'] is not None else None
            Q = None
            if L['uses_ef'] and not L['ef_cat']:
                ee = m.edge_encoder[i if m.inject_edge_features else 0]
                Q = ops.linear(ee(data.edge_features).contiguous(), L['Wq_ef'])
            S = ops.general_edge_idx(plan, dh, P=P, Q=Q, node_rows=node_rows, Tn=L['Tn'], edge_rows=edge_rows,
                                     Te=L['Te'], activation=L['act_mlp'],
                                     edge_rows_csr=True)
            # ---- update_fn (first Linear carries the folded second message Linear) + BN + act
            if L['x_cat']:
                H = ops.linear(S, L['Wu'], bias=L['c1'], row_scale=plan.degree(), row_vec=L['vf'],
                               tab_idx=node_rows[:, 0].contiguous(), tab=L['Tu'], scale=L['su'], shift=L['tu'],
                               activation=L['act_mlp'])
            else:
                H = ops.linear(x, L['Wu'], A2=S, bias=L['c1'], row_scale=plan.degree(), row_vec=L['vf'],
                               scale=L['su'], shift=L['tu'], activation=L['act_mlp'])
            # ---- second Linear + model-level BatchNorm + activation (models_graph_classification.py:231-233)
            x = ops.linear(H, L['U2'], bias=L['c2'], scale=L['sm'], shift=L['tm'], activation=self._act_model)
            x_interm.append(x)

        self.last_x_interm = x_interm          # layer outputs of the last call (parity checks of the one-kernel path)
        mean = m.readout == 'mean'
        return self._project([None if pr is None else ops.pool_ptr(x_interm[i], node_ptr, mean)
                              for i, pr in enumerate(self.proj)])

    def _project(self, pooled_list):
        """sum of the JK projections of the pooled layer outputs (models_graph_classification.py:236-240)"""
        out = None
        for pr, pooled in zip(self.proj, pooled_list):
            if pr is None:
                continue
            if pr[0] == 'linear':
                out = ops.linear(pooled, pr[1], bias=pr[2], out=out, accumulate=out is not None)
            else:
                _, W0, b0, s, t, W1, b1, act = pr
                h = ops.linear(pooled, W0, bias=b0, scale=s, shift=t, activation=act)
                out = ops.linear(h, W1, bias=b1, out=out, accumulate=out is not None)
        return out
