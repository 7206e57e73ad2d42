"""preprocess + forward as one device-side pipeline (the BASELINE.json metric):

    edge_index, node_ptr ──COUNT──▶ identifiers int64 ──encode──▶ ranks
                                   (utils_ids.py:7-29)   (utils_encoding.py:37-59, one_hot_unique)
    ranks, x, edge_features ──GNNSubstructures.forward──▶ prediction [G, out]
                                   (models_graph_classification.py:204-247)

The reference runs COUNT + encode offline on the CPU for the whole data set and
only the forward per batch; here all three stages run per batch on the GPU, and
the whole step can be captured in one CUDA graph (static shapes: one graph per
batch shape) so that a 128-graph batch is not bound by Python launch overhead.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import counting
from .patterns import total_columns


class UniqueEncoder:
    """one_hot_unique (utils_encoding.py:37-59): every identifier column is
    replaced by the rank of its value among the data-set-wide sorted distinct
    values.  `fit` learns the vocabulary from identifiers of a calibration set;
    values never seen map to the rank of the next larger known value (the
    reference cannot encode unseen values at all: it encodes the data set it
    was fitted on)."""

    def __init__(self, vocab: Sequence[torch.Tensor]):
        self.vocab = [v.contiguous() for v in vocab]
        self.d = [int(v.numel()) for v in self.vocab]

    @staticmethod
    def fit(identifiers: torch.Tensor) -> 'UniqueEncoder':
        return UniqueEncoder([torch.unique(identifiers[:, c]) for c in range(identifiers.shape[1])])

    def __call__(self, identifiers: torch.Tensor) -> torch.Tensor:
        cols = [torch.bucketize(identifiers[:, c].contiguous(), self.vocab[c]).clamp_(max=self.d[c] - 1)
                for c in range(identifiers.shape[1])]
        return torch.stack(cols, 1)


class Batch:
    """attribute bag with the fields GNNSubstructures.forward reads"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class GSNPipeline:

    def __init__(self, model, subgraph_dicts, induced: bool, id_scope: str, encoder: UniqueEncoder,
                 max_nodes_per_graph: int, fused=True):
        """fused: True = best available (one-kernel forward when the model allows it, else the per-layer fused path),
        'model' / 'layers' force one of the two, False = the drop-in per-layer modules"""
        from . import fused as _fused, fused_model as _fm
        self.model, self.subgraph_dicts, self.induced, self.id_scope = model, subgraph_dicts, induced, id_scope
        # inference fast path (indices instead of one-hot tensors, BN folded, own GEMM kernels) when the
        # model configuration allows it; otherwise the per-layer path
        self.fused = None
        if fused and not model.training:
            if fused in (True, 'model') and _fm.supported(model, max_nodes_per_graph):
                self.fused = _fm.FusedModel(model)
            elif fused == 'model':
                raise NotImplementedError('one-kernel forward: unsupported model / graph size')
            elif _fused.supported(model):
                self.fused = _fused.FusedForward(model)
        self.encoder, self.max_nodes = encoder, int(max_nodes_per_graph)
        self.n_cols = total_columns(subgraph_dicts)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._exec = None               # cudaGraphExec_t of the captured step (gsn_submit_step)
        self._static: Dict[str, torch.Tensor] = {}
        self._out: Optional[torch.Tensor] = None
        self.last_status: Optional[torch.Tensor] = None
        self._side: Optional[torch.cuda.Stream] = None
        self.use_tile_plan = True       # small batches: greedy packing of graphs into the forward's tiles (gsn_tile_plan)

    # -- one eager step on device-resident inputs --------------------------------
    def step(self, t: Dict[str, torch.Tensor]) -> torch.Tensor:
        from . import encoders, fused_model as _fm, ops
        ops.clear_plan_cache()            # a step is a new batch: never reuse groupings across steps
        encoders._pool_plans.clear()
        N, G = int(t['x'].shape[0]), int(t['node_ptr'].numel() - 1)
        # the grouping of edge_index for the message kernels does not depend on COUNT: build it on a side stream
        # while the counting kernels run (a fork / join that CUDA-graph capture records as parallel branches)
        # one device-side status word per step: COUNT, the CSR build and the encoders OR their GSN_S_* bits into it
        status = torch.zeros(1, dtype=torch.int32, device=t['edge_index'].device)
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        one_kernel = isinstance(self.fused, _fm.FusedModel)
        data = Batch(x=t['x'], edge_index=t['edge_index'], edge_features=t['edge_features'], batch=t['batch'],
                     degrees=t['degrees'], node_ptr=t['node_ptr'], num_graphs=G)
        with torch.cuda.stream(self._side):
            flow = self.model.conv[0].flow
            plan = ops.edge_plan(t['edge_index'], N, flow, status=status)
            tiles = None
            if one_kernel:
                if self.use_tile_plan:
                    tiles = _fm.tile_plan(t['node_ptr'], N, status)    # which graphs share a tile (small batches)
            else:
                plan.degree()
            if self.fused is not None:
                self.fused.prefetch_rows(data)        # index rows of the inputs that do not involve the identifiers
        ids = counting.count_batch(t['edge_index'], t['node_ptr'], self.subgraph_dicts, self.induced, self.id_scope,
                                   num_nodes=N, max_nodes_per_graph=self.max_nodes, check=False, status=status)
        self.last_status = status
        cur.wait_stream(self._side)
        if self.fused is not None:
            if tiles is not None:
                return self.fused(data, raw_identifiers=ids, vocab=self.encoder.vocab, tile_plan=tiles)
            return self.fused(data, raw_identifiers=ids, vocab=self.encoder.vocab)
        data = Batch(x=t['x'], edge_index=t['edge_index'], edge_features=t['edge_features'], batch=t['batch'],
                     degrees=t['degrees'], identifiers=self.encoder(ids), num_graphs=G)
        return self.model(data)

    # -- CUDA-graph capture of the whole step -------------------------------------
    def capture(self, example: Dict[str, torch.Tensor], warmup: int = 3, static: Optional[Dict[str, torch.Tensor]] = None):
        """static: optional pre-allocated input buffers to capture on (e.g. views into ONE packed buffer, so that a
        step's inputs arrive with a single copy); they are filled from `example`"""
        if static is not None:
            for k, v in example.items():
                static[k].copy_(v)
            self._static = static
        else:
            self._static = {k: v.clone() for k, v in example.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.step(self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph, self._exec = torch.cuda.CUDAGraph(), None
        with torch.no_grad(), torch.cuda.graph(self._graph):
            self._out = self.step(self._static)
        return self

    def load(self, t: Dict[str, torch.Tensor]):
        """copy a same-shaped batch (device or pinned host tensors) into the captured inputs"""
        for k, v in t.items():
            self._static[k].copy_(v, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self._graph.replay()
        return self._out

    def submit(self, stream: torch.cuda.Stream, d_in: Optional[torch.Tensor] = None, src: Optional[torch.Tensor] = None,
               out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The captured step on `stream` through ONE native call (`gsn_submit_step`): copy `src` (pinned host or device)
        into the static input buffer `d_in`, replay, copy the predictions into `out_host` (pinned).  Same effect as
        `d_in.copy_(src); replay(); out_host.copy_(out)` under `torch.cuda.stream(stream)` at a fraction of the host time."""
        from . import _lib
        out = self._out
        nin = 0 if src is None else src.numel() * src.element_size()
        nout = 0 if out_host is None else out_host.numel() * out_host.element_size()
        if out_host is not None and nout > out.numel() * out.element_size():
            raise ValueError('out_host is larger than the step\'s output')
        if self._exec is None:
            self._exec = self._graph.raw_cuda_graph_exec()
        _lib.check(_lib.lib().gsn_submit_step(self._exec, None if d_in is None else d_in.data_ptr(),
                                              None if src is None else src.data_ptr(), nin,
                                              None if out_host is None else out_host.data_ptr(), out.data_ptr(), nout,
                                              stream.cuda_stream), 'gsn_submit_step')
        return out


# ---------------------------------------------------------------------------------------------------------------------
# variable-shape batches through captured CUDA graphs
# ---------------------------------------------------------------------------------------------------------------------
FIELDS = ('edge_index', 'node_ptr', 'x', 'edge_features', 'batch', 'degrees')


class Packing:
    """byte layout of one (padded) batch inside a single buffer, 16-byte aligned fields: a step's inputs move with ONE
    copy (device to device, or pinned host to device)"""

    def __init__(self, shapes: Dict[str, tuple], dtypes: Dict[str, torch.dtype]):
        self.fields, off = [], 0
        for k in FIELDS:
            n = 1
            for d in shapes[k]:
                n *= int(d)
            nbytes = n * torch.empty((), dtype=dtypes[k]).element_size()
            self.fields.append((k, off, nbytes, dtypes[k], tuple(int(d) for d in shapes[k])))
            off += (nbytes + 15) // 16 * 16
        self.nbytes = off

    def views(self, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {k: buf[off:off + n].view(dt).view(shape) for k, off, n, dt, shape in self.fields}

    def pack(self, t: Dict[str, torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tensors (CPU) -> one uint8 buffer"""
        if out is None:
            out = torch.zeros(self.nbytes, dtype=torch.uint8)
        for k, off, n, dt, shape in self.fields:
            out[off:off + n] = t[k].contiguous().view(-1).view(torch.uint8)
        return out


class BucketedPipeline:
    """The reference's DataLoader (main.py:243-258) yields a different (N, E) every step, a captured CUDA graph needs
    static shapes.  Batches are therefore padded to BUCKET shapes and every bucket owns one captured GSNPipeline step
    (an LRU of `max_buckets` of them):

      * nodes  N -> N_cap: the padding rows form extra `sentinel` graphs (<= pad_graph_nodes nodes each, G_cap - G of
        them, possibly empty) after the real ones -- graphs are independent in COUNT, in message passing and in the
        readout, so the real graphs' predictions are unchanged and the sentinels' rows are dropped;
      * edges  E -> E_cap: the padding columns are self loops on sentinel nodes (at most 4 per node).  COUNT drops self
        loops like remove_self_loops (utils_ids.py:11-15), message passing only sends them to sentinel rows;
      * graphs G -> G_cap = G + a fixed number of sentinel graphs (so that the shape depends on the bucket only).

    pad() works on host arrays (it is part of collating a batch); run() copies one packed buffer into the bucket's
    static inputs and replays its graph."""

    def __init__(self, model, subgraph_dicts, induced: bool, id_scope: str, encoder: UniqueEncoder, max_nodes_per_graph: int,
                 node_step: int = 128, edge_step: int = 256, max_buckets: int = 16, fused=True):
        self.args = (model, subgraph_dicts, induced, id_scope, encoder, int(max_nodes_per_graph))
        self.fused = fused
        self.node_step, self.edge_step, self.max_buckets = int(node_step), int(edge_step), int(max_buckets)
        self.pad_graph_nodes = min(64, int(max_nodes_per_graph))
        self._lru: 'OrderedDict[tuple, tuple]' = __import__('collections').OrderedDict()
        self.captures = 0

    # ---- shapes
    def bucket(self, N: int, E: int, G: int):
        """(N_cap, E_cap, G_cap) of a batch with N nodes, E edge columns, G graphs"""
        E_cap = -(-max(E, 1) // self.edge_step) * self.edge_step
        need_nodes = max(1, -(-(E_cap - E) // 4))                       # <= 4 padding self loops per sentinel node
        N_cap = -(-(N + need_nodes) // self.node_step) * self.node_step
        # sentinel graphs: enough for the largest padding a bucket can see
        max_pad_nodes = self.node_step + -(-self.edge_step // 4)
        G_cap = G + -(-max_pad_nodes // self.pad_graph_nodes) + 1
        return N_cap, E_cap, G_cap

    def pad(self, b: Dict, bucket=None) -> Dict[str, torch.Tensor]:
        """numpy / tensor batch (fields FIELDS) -> padded CPU tensors of the bucket's shapes"""
        import numpy as np
        t = {k: (torch.from_numpy(np.ascontiguousarray(b[k])) if not torch.is_tensor(b[k]) else b[k].cpu()) for k in FIELDS}
        N, E, G = int(t['x'].shape[0]), int(t['edge_index'].shape[1]), int(t['node_ptr'].numel() - 1)
        N_cap, E_cap, G_cap = bucket or self.bucket(N, E, G)
        pn, pe = N_cap - N, E_cap - E
        if pn < max(1, -(-pe // 4)) or pe < 0 or G_cap <= G:
            raise ValueError('batch does not fit the bucket')
        out = {}
        loops = N + (torch.arange(pe, dtype=torch.int64) // 4)
        out['edge_index'] = torch.cat((t['edge_index'], torch.stack((loops, loops))), 1)
        ptr = torch.full((G_cap + 1,), N_cap, dtype=torch.int64)
        ptr[:G + 1] = t['node_ptr']
        k = torch.arange(1, G_cap - G + 1, dtype=torch.int64)
        ptr[G + 1:] = torch.clamp(N + k * self.pad_graph_nodes, max=N_cap)
        if int(ptr[-1]) != N_cap:
            raise ValueError('not enough sentinel graphs for this padding')
        out['node_ptr'] = ptr
        xs = t['x']
        out['x'] = torch.cat((xs, torch.zeros((pn,) + tuple(xs.shape[1:]), dtype=xs.dtype)), 0)
        ef = t['edge_features']
        out['edge_features'] = torch.cat((ef, torch.ones((pe,) + tuple(ef.shape[1:]), dtype=ef.dtype)), 0)
        sizes = ptr[1:] - ptr[:-1]
        out['batch'] = torch.repeat_interleave(torch.arange(G_cap, dtype=torch.int64), sizes)
        dg = t['degrees']
        out['degrees'] = torch.cat((dg, torch.zeros((pn,) + tuple(dg.shape[1:]), dtype=dg.dtype)), 0)
        return out

    def packing(self, padded: Dict[str, torch.Tensor]) -> Packing:
        return Packing({k: tuple(v.shape) for k, v in padded.items()}, {k: v.dtype for k, v in padded.items()})

    # ---- captured step per bucket
    def _entry(self, key, padded_example: Dict[str, torch.Tensor], device):
        ent = self._lru.get(key)
        if ent is not None:
            self._lru.move_to_end(key)
            return ent
        model, sds, induced, scope, enc, mx = self.args
        pk = self.packing(padded_example)
        buf = torch.zeros(pk.nbytes, dtype=torch.uint8, device=device)
        pipe = GSNPipeline(model, sds, induced, scope, enc, mx, fused=self.fused)
        pipe.capture({k: v.to(device) for k, v in padded_example.items()}, warmup=2, static=pk.views(buf))
        ent = (pipe, pk, buf)
        self._lru[key] = ent
        self.captures += 1
        while len(self._lru) > self.max_buckets:
            self._lru.popitem(last=False)
        return ent

    def prepare(self, b: Dict, device, pin: bool = False):
        """collate-time half of a step: pad + pack one batch.  Returns (bucket key, packed uint8 buffer on the CPU
        [pinned], number of real graphs) and makes sure the bucket's graph is captured."""
        padded = self.pad(b)
        key = (padded['x'].shape[0], padded['edge_index'].shape[1], padded['node_ptr'].numel() - 1)
        _, pk, _ = self._entry(key, padded, device)
        packed = pk.pack(padded)
        return key, (packed.pin_memory() if pin else packed), int(b['node_ptr'].shape[0] - 1)

    def run(self, key, packed: torch.Tensor, n_graphs: int) -> torch.Tensor:
        """one step: copy the packed inputs (host or device) into the bucket's static buffer, replay; the returned
        tensor is a view of the bucket's static output (valid until the bucket runs again)"""
        pipe, pk, buf = self._lru[key]
        buf.copy_(packed, non_blocking=True)
        return pipe.replay()[:n_graphs]

    def submit(self, key, packed: torch.Tensor, n_graphs: int, stream: torch.cuda.Stream,
               out_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """run() on an explicit stream through one native call (copy in, replay, optional copy of the first
        `out_host.numel()` predictions to pinned host memory); see GSNPipeline.submit"""
        pipe, pk, buf = self._lru[key]
        return pipe.submit(stream, buf, packed, out_host)[:n_graphs]
