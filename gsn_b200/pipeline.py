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
        self._static: Dict[str, torch.Tensor] = {}
        self._out: Optional[torch.Tensor] = None
        self.last_status: Optional[torch.Tensor] = None
        self._side: Optional[torch.cuda.Stream] = None

    # -- one eager step on device-resident inputs --------------------------------
    def step(self, t: Dict[str, torch.Tensor]) -> torch.Tensor:
        from . import encoders, ops
        ops.clear_plan_cache()            # a step is a new batch: never reuse groupings across steps
        encoders._pool_plans.clear()
        N, G = int(t['x'].shape[0]), int(t['node_ptr'].numel() - 1)
        # the grouping of edge_index for the message kernels does not depend on COUNT: build it on a side stream
        # while the counting kernels run (a fork / join that CUDA-graph capture records as parallel branches)
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            flow = self.model.conv[0].flow
            ops.edge_plan(t['edge_index'], N, flow).degree()
        graph = counting.BatchedGraph(t['edge_index'], t['node_ptr'], num_nodes=N, max_nodes_per_graph=self.max_nodes)
        ids = counting.count_batch(t['edge_index'], t['node_ptr'], self.subgraph_dicts, self.induced, self.id_scope,
                                   num_nodes=N, max_nodes_per_graph=self.max_nodes, check=False, graph=graph)
        self.last_status = graph.status
        cur.wait_stream(self._side)
        if self.fused is not None:
            data = Batch(x=t['x'], edge_index=t['edge_index'], edge_features=t['edge_features'], batch=t['batch'],
                         degrees=t['degrees'], node_ptr=t['node_ptr'], num_graphs=G)
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
        self._graph = torch.cuda.CUDAGraph()
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
