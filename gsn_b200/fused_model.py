"""Whole-model fused inference forward: GNNSubstructures.forward in ONE kernel launch.

Host side of `gsn_fused_model_fwd` (csrc/fused_model.cu).  A PyG batch is block-diagonal,
so a row tile that holds whole graphs never reads outside itself; the kernel keeps a
tile's activations on chip across all layers of
/root/reference/models_graph_classification.py:204-247 (eval mode, 'general' message
kind, GSN_edge_sparse.py:111-166 + models_misc.py:52-59) and writes only the per-graph
readout (utils_graph_learning.py:23-41).  The arithmetic re-association (split first
Linear, folded second message Linear, BatchNorm as scale / shift, categorical inputs as
table rows) is the one of gsn_b200/fused.py; this module pads every matrix to D x D
(D = 64 | 128), scales weight rows by exact powers of two and splits them into fp16
(hi, lo) pairs for the 3 x fp16 tensor-core products.

Parity: tests/test_fused_model_gpu.py (reference goldens, per-layer fused path, oracle).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch

from . import _lib, ops
from .fused import FusedForward, supported as _fused_supported

MAX_LAYERS = 8
MAX_TILE_ROWS = 128
_ACT = ops.ACTIVATIONS
JK_IN_KERNEL = True     # evaluate the JK head inside the model kernel when its shape allows (False: two more GEMM launches)
FILL_ROWS = 100.0      # rows' worth of graphs per unit of work for small batches (see FusedModel.__call__)
(V_CJ, V_CI, V_SHIFT, V_CU, V_CF, V_CV, V_CB, V_C2S, V_C2B) = range(9)


class GsnFusedLayer(ctypes.Structure):
    """ctypes image of `struct GsnFusedLayer` (include/gsn_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ('node_rows', 'Tn', 'tu_rows', 'Tu', 'edge_rows', 'Te', 'vec', 'pooled',
                                               'x_out')] + \
               [(n, ctypes.c_int32) for n in ('n_node_cols', 'tu_stride', 'n_edge_cols', 'te_rows', 'has_dense', 'mat0',
                                              'act_msg', 'act_upd', 'act_out', 'pool')] + \
               [(n, ctypes.c_void_p) for n in ('jk_W0T', 'jk_vec', 'jk_W1', 'jk_b1')] + \
               [('jk_kind', ctypes.c_int32), ('jk_act', ctypes.c_int32)]


class GsnFusedModel(ctypes.Structure):
    """ctypes image of `struct GsnFusedModel`."""
    _fields_ = [('layers', GsnFusedLayer * MAX_LAYERS), ('n_layers', ctypes.c_int32), ('D', ctypes.c_int32),
                ('n_mats', ctypes.c_int32), ('graphs_per_unit', ctypes.c_int32), ('Whi', ctypes.c_void_p),
                ('Wlo', ctypes.c_void_p), ('rowptr', ctypes.c_void_p), ('nbr', ctypes.c_void_p),
                ('node_ptr', ctypes.c_void_p), ('x0', ctypes.c_void_p), ('x0_ld', ctypes.c_int32), ('x0_d', ctypes.c_int32),
                ('N', ctypes.c_int64), ('E', ctypes.c_int64), ('G', ctypes.c_int64), ('status', ctypes.c_void_p),
                ('tile_plan', ctypes.c_void_p), ('max_tiles', ctypes.c_int32), ('n_out', ctypes.c_int32),
                ('out', ctypes.c_void_p)]


def split_fp16_rows(W: torch.Tensor):
    """W fp32 [n, k] -> (hi fp16, lo fp16, inverse row scale fp32 [n]): rows scaled by the power of two that puts
    their largest magnitude into [2^14, 2^15); hi = rn_fp16(W s), lo = rn_fp16(W s - hi).  Exact scaling: the
    product a*w is recovered by multiplying the accumulator with the two inverse scales."""
    m = W.abs().amax(1)
    e = torch.floor(torch.log2(torch.where(m > 0, m, torch.ones_like(m))))
    e = torch.where(m > 0, e, torch.full_like(e, 14.0))
    scale = torch.exp2(14.0 - e)
    Ws = W * scale[:, None]
    hi = Ws.to(torch.float16)
    lo = (Ws - hi.float()).to(torch.float16)
    return hi, lo, (1.0 / scale).float()


def supported(model, max_nodes_per_graph: Optional[int] = None) -> bool:
    """the configurations the one-kernel forward covers (everything else: FusedForward / the per-layer path)"""
    if not _fused_supported(model):
        return False
    if max_nodes_per_graph is not None and max_nodes_per_graph > MAX_TILE_ROWS:
        return False
    if len(model.conv) > MAX_LAYERS:
        return False
    from .fused import _is_onehot
    for i, conv in enumerate(model.conv):
        f, u = conv.msg_fn, conv.update_fn
        dims = [f.fc[0].weight.shape[0], u.fc[0].weight.shape[0], u.fc[1].weight.shape[0]]
        if i > 0 or not _is_onehot(model.input_node_encoder):
            dims.append(conv._dims[0])
        if max(dims) > 128:
            return False
        if conv.uses_ef:
            ee = model.edge_encoder[i if model.inject_edge_features else 0]
            if not _is_onehot(ee):
                return False          # dense edge features would need a per-edge operand (Q)
    if model.final_projection[0]:
        return False                  # JK term of the raw input features: not pooled by the kernel
    return True


class FusedModel(FusedForward):
    """Drop-in for FusedForward.__call__ that runs all layers + readout in one launch."""

    def __init__(self, model, graphs_per_unit: Optional[int] = None):
        super().__init__(model)
        if not supported(model):
            raise NotImplementedError('fused model kernel: unsupported model configuration')
        self.graphs_per_unit = graphs_per_unit
        self._fm_stamp = None
        self.debug_x_out = False
        self.last_x_out: List[torch.Tensor] = []

    # ------------------------------------------------------------------ weight preparation
    @torch.no_grad()
    def prepare(self):
        super().prepare()
        m = self.model
        dev = next(m.parameters()).device
        widths = []
        for L in self.layers:
            widths += [L['dh'], L['U2'].shape[0], L['U2'].shape[1]]
            if not L['x_cat']:
                widths.append(L['d_in'])
        D = 64 if max(widths) <= 64 else 128
        self.D = D

        def pad_mat(W):                       # [n, k] -> [D, D]
            out = torch.zeros((D, D), dtype=torch.float32, device=dev)
            out[:W.shape[0], :W.shape[1]] = W
            return out

        def pad_vec(v, fill=0.0):
            out = torch.full((D,), fill, dtype=torch.float32, device=dev)
            if v is not None:
                out[:v.numel()] = v
            return out

        def pad_cols(T, width):               # [rows, w] -> [rows, width]
            out = torch.zeros((T.shape[0], width), dtype=torch.float32, device=dev)
            out[:, :T.shape[1]] = T
            return out

        mats, self.fm_layers = [], []
        for L in self.layers:
            dh, d_in = L['dh'], L['d_in']
            vec = torch.zeros((9, D), dtype=torch.float32, device=dev)
            F = {'has_dense': not L['x_cat'], 'mat0': len(mats)}
            su = L['su'] if L['su'] is not None else torch.ones_like(L['c1'])
            tu = L['tu'] if L['tu'] is not None else torch.zeros_like(L['c1'])
            ones_up = torch.ones_like(L['c2'])
            sm = L['sm'] if L['sm'] is not None else ones_up
            tm = L['tm'] if L['tm'] is not None else torch.zeros_like(L['c2'])
            if F['has_dense']:
                Wxi, Wxj = L['Wp'][:dh], L['Wp'][dh:]
                U1x, Wf = L['Wu'][:, :d_in], L['Wu'][:, d_in:]
                vec[V_SHIFT] = pad_vec(L['bp'][:dh])
                order = [(Wxj, V_CJ, None), (Wxi, V_CI, None), (U1x, V_CU, su), (Wf, V_CF, su), (L['U2'], V_C2S, sm)]
            else:
                Wf = L['Wu']
                order = [(Wf, V_CF, su), (L['U2'], V_C2S, sm)]
            for W, slot, post in order:
                hi, lo, inv = split_fp16_rows(pad_mat(W))
                if post is not None:
                    inv = inv * pad_vec(post, 1.0)
                vec[slot] = inv
                mats.append((hi, lo))
            vec[V_CV] = pad_vec(L['vf'] * su)
            vec[V_CB] = pad_vec(L['c1'] * su + tu)
            vec[V_C2B] = pad_vec(L['c2'] * sm + tm)
            F['vec'] = vec.contiguous()
            # tables: node side [rows, 2D] (P_i half | P_j half), edge side [rows, D], update-side rows of a one-hot x
            if L['Tn'] is not None:
                Tn = torch.zeros((L['Tn'].shape[0], 2 * D), dtype=torch.float32, device=dev)
                Tn[:, :dh] = L['Tn'][:, :dh]
                Tn[:, D:D + dh] = L['Tn'][:, dh:]
                F['Tn'] = Tn
            else:
                F['Tn'] = None
            F['Te'] = pad_cols(L['Te'], D).contiguous() if L['Te'] is not None else None
            F['Tu'] = pad_cols(L['Tu'] * su[None, :], D).contiguous() if L['Tu'] is not None else None
            F['act_msg'] = F['act_upd'] = _ACT[L['act_mlp']]
            F['act_out'] = _ACT[self._act_model]
            self.fm_layers.append(F)
        # JK head in the kernel (models_graph_classification.py:236-240) when every projection reads a layer output of
        # width <= D and the head is narrow; otherwise the kernel writes the pooled rows and the caller projects them
        self.jk = None
        prs = [pr for pr in self.proj if pr is not None]
        if prs and self.proj[0] is None and JK_IN_KERNEL:
            n_out = prs[0][5].shape[0] if prs[0][0] == 'mlp' else prs[0][1].shape[0]
            ok = n_out <= 32
            jk = [None]
            for pr in self.proj[1:]:
                if pr is None:
                    jk.append(None)
                    continue
                if pr[0] == 'linear':
                    _, W1, b1 = pr
                    ok = ok and W1.shape[1] <= D and W1.shape[0] == n_out
                    jk.append({'kind': 1, 'act': 0, 'W0T': None, 'vec': None, 'W1': pad_cols(W1, D).contiguous(),
                               'b1': b1.float().contiguous()} if ok else None)
                else:
                    _, W0, b0, sc, sh, W1, b1, act = pr
                    ok = ok and max(W0.shape) <= D and W1.shape[1] <= D and W1.shape[0] == n_out
                    if ok:
                        vec = torch.stack([pad_vec(b0), pad_vec(sc, 1.0), pad_vec(sh)])
                        jk.append({'kind': 2, 'act': _ACT[act], 'W0T': pad_mat(W0).t().contiguous(), 'vec': vec.contiguous(),
                                   'W1': pad_cols(W1, D).contiguous(), 'b1': b1.float().contiguous()})
            if ok:
                self.jk, self.n_out = jk, int(n_out)
        self.n_mats = len(mats)
        self.Whi = torch.cat([h for h, _ in mats], 0).contiguous()
        self.Wlo = torch.cat([l for _, l in mats], 0).contiguous()
        self._fm_stamp = self._stamp

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, data, raw_identifiers: Optional[torch.Tensor] = None, vocab: Optional[List[torch.Tensor]] = None,
                 tile_plan: Optional[torch.Tensor] = None):
        """tile_plan: optional result of `tile_plan()` for this batch's node_ptr (small batches: one CTA per greedily
        packed tile); without it a CTA takes `graphs_per_unit` consecutive graphs"""
        m = self.model
        ctx = self._context(data, raw_identifiers, vocab)       # (re)prepares when the weights changed
        dev, N, E = ctx['dev'], ctx['N'], ctx['E']
        node_ptr = data.node_ptr
        G = int(node_ptr.numel() - 1)
        D = self.D
        flows = {L['flow'] for L in self.layers}
        if len(flows) != 1:
            raise NotImplementedError('fused model kernel: layers with different flow directions')
        plan = ops.edge_plan(data.edge_index, N, flows.pop())
        mean = m.readout == 'mean'
        fm = GsnFusedModel()
        keep = []
        pooled_list: List[Optional[torch.Tensor]] = [None]
        self.last_x_out = []
        for i, (L, F) in enumerate(zip(self.layers, self.fm_layers)):
            node_rows, edge_rows = self._layer_rows(L, ctx, plan)
            keep += [node_rows, edge_rows]
            fl = fm.layers[i]
            fl.vec = F['vec'].data_ptr()
            fl.has_dense, fl.mat0 = int(F['has_dense']), F['mat0']
            fl.act_msg, fl.act_upd, fl.act_out = F['act_msg'], F['act_upd'], F['act_out']
            if node_rows is not None:
                fl.node_rows, fl.n_node_cols, fl.Tn = node_rows.data_ptr(), node_rows.shape[1], F['Tn'].data_ptr()
            if edge_rows is not None:
                fl.edge_rows, fl.n_edge_cols = edge_rows.data_ptr(), edge_rows.shape[1]
                fl.Te, fl.te_rows = F['Te'].data_ptr(), F['Te'].shape[0]
            if F['Tu'] is not None:
                fl.Tu, fl.tu_rows, fl.tu_stride = F['Tu'].data_ptr(), node_rows.data_ptr(), node_rows.shape[1]
            if self.proj[i + 1] is not None and self.jk is not None:
                J = self.jk[i + 1]
                fl.pool, fl.jk_kind, fl.jk_act = (2 if mean else 1), J['kind'], J['act']
                fl.jk_W1, fl.jk_b1 = J['W1'].data_ptr(), J['b1'].data_ptr()
                if J['kind'] == 2:
                    fl.jk_W0T, fl.jk_vec = J['W0T'].data_ptr(), J['vec'].data_ptr()
                pooled_list.append(None)
            elif self.proj[i + 1] is not None:
                pooled = torch.empty((G, D), dtype=torch.float32, device=dev)
                fl.pool, fl.pooled = (2 if mean else 1), pooled.data_ptr()
                d_up = L['U2'].shape[0]
                pooled_list.append(pooled[:, :d_up])
            else:
                pooled_list.append(None)
            if self.debug_x_out:
                xo = torch.zeros((N, D), dtype=torch.float32, device=dev)
                fl.x_out = xo.data_ptr()
                self.last_x_out.append(xo)
        fm.n_layers, fm.D, fm.n_mats = len(self.layers), D, self.n_mats
        gpu = self.graphs_per_unit
        if gpu is None:
            # a tile costs about the same whether it holds 20 rows or 128 (its phases are latency-bound), so tiles are
            # filled: ~100 rows' worth of graphs per unit (one tile, rarely two) for small batches -- a step then occupies
            # few SMs and steps of other streams run beside it -- and >= 2 units per SM for large ones
            avg = max(1.0, N / max(G, 1))
            gpu = max(1, int(FILL_ROWS / avg), -(-G // (2 * 148)))
        fm.graphs_per_unit = int(gpu)
        if tile_plan is not None:
            fm.tile_plan, fm.max_tiles = tile_plan.data_ptr(), int(tile_plan.numel() - 2)
            keep.append(tile_plan)
        fm.Whi, fm.Wlo = self.Whi.data_ptr(), self.Wlo.data_ptr()
        fm.rowptr, fm.nbr, fm.node_ptr = plan.rowptr.data_ptr(), plan.nbr.data_ptr(), node_ptr.data_ptr()
        x0 = ctx['x']
        if x0 is not None:
            x0 = x0.float().contiguous()
            keep.append(x0)
            fm.x0, fm.x0_ld, fm.x0_d = x0.data_ptr(), x0.stride(0), x0.shape[1]
        fm.N, fm.E, fm.G = N, E, G
        fm.status = self.status.data_ptr()
        out = None
        if self.jk is not None:
            out = torch.empty((G, self.n_out), dtype=torch.float32, device=dev)
            fm.out, fm.n_out = out.data_ptr(), self.n_out
        with torch.cuda.device(dev):
            _lib.call('fused_model', 'gsn_fused_model_fwd', ctypes.byref(fm), _lib.stream_ptr())
        self._keep = keep
        return out if out is not None else self._project(pooled_list)

    def raise_on_status(self):
        bits = (int(self.status.item()) if getattr(self, 'status', None) is not None else 0) & _lib.S_FATAL
        if bits:
            raise RuntimeError('fused model kernel: ' + _lib.status_message(bits))


TILE_PLAN_MAX_GRAPHS = 1024      # beyond this a CTA takes a run of graphs_per_unit graphs and cuts it into tiles itself (the plan is a
#                                  one-warp scan: 10 us at 1,024 graphs, 90 us at 4,096 -- measured -- and no longer pays for itself)


def tile_plan(node_ptr: torch.Tensor, N: int, status: torch.Tensor) -> Optional[torch.Tensor]:
    """Greedy packing of the batch's graphs into tiles of <= 128 rows (`gsn_tile_plan`): int32 [max_tiles + 2] on the
    device, or None for batches too large for the one-CTA scan (they do not need it: with hundreds of graphs per CTA only
    the last tile of a run is underfilled)."""
    G = int(node_ptr.numel() - 1)
    if G < 1 or G > TILE_PLAN_MAX_GRAPHS:
        return None
    cap = min(G, 2 * int(N) // 128 + G // 32 + 2)
    plan = torch.empty(cap + 2, dtype=torch.int32, device=node_ptr.device)
    with torch.cuda.device(node_ptr.device):
        _lib.call('tile_plan', 'gsn_tile_plan', node_ptr.data_ptr(), G, plan.data_ptr(), cap, status.data_ptr(),
                  _lib.stream_ptr())
    return plan
