"""bench.py -- graphs/s of preprocess (COUNT) + forward (MP) on ZINC-shaped batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-sweep]

Workload (BASELINE.json configs[1]): ZINC-shaped synthetic batch of B=128
molecules; structural identifiers = cycles k<=8, edge scope (GSN-e), non-induced;
model = GNNSubstructures, README.md:112 recipe with id_scope local
(GSN_edge_sparse, general, 4 layers, d_out 128, one-hot encoders, sum readout).
One step = COUNT over the batch + one_hot_unique encode + model forward.

  value     inputs resident in HBM, whole step replayed as one CUDA graph
  e2e       same step through GSNPipeline with inputs in pinned HOST memory:
            H2D of the batch and D2H of the predictions inside the timed region
  roofline  the scatter kernel (general_edge) timed live with CUDA events in an
            instrumented eager pass over the same steps
  N > 1     one process per GPU (torchrun), every rank its own batch (weak
            scaling, no data-path collective), max-over-ranks time

--impl reference times the reference's CPU path for the same step on the host
cores: the oracle ports (oracle/count_enum.c with OpenMP = graph-tool-equivalent
all-maps enumeration, oracle/mp_ref.py = the reference's PyTorch layers on CPU);
/root/reference itself does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_MAX = 8
D_OUT = 128
N_LAYERS = 4


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def model_args(d_id_cols):
    """post-process_arguments dict (utils.py:94-161) of the README.md:112 ZINC recipe, id_scope local"""
    L = N_LAYERS
    return dict(seed=0, model_name='GSN_edge_sparse', readout='sum', dropout_features=[0.0] * (L + 1), bn=[True] * L,
                final_projection=[False] * L + [True], inject_ids=False, inject_edge_features=True,
                random_features=False, id_scope='local', d_msg=[D_OUT] * L, d_out=[D_OUT] * L, d_h=[[D_OUT]] * L,
                aggr='add', flow='source_to_target', msg_kind='general', train_eps=[False] * L, activation_mlp='relu',
                bn_mlp=True, jk_mlp=True, degree_embedding='one_hot_encoder', degree_as_tag=[False] * L,
                retain_features=[False] + [True] * (L - 1), multi_embedding_aggr='sum',
                input_node_encoder='one_hot_encoder', d_out_node_encoder=D_OUT, edge_encoder='one_hot_encoder',
                d_out_edge_encoder=[D_OUT] * L, id_embedding='one_hot_encoder', d_out_id_embedding=D_OUT,
                d_out_degree_embedding=D_OUT, extend_dims=True, activation='relu')


def model_ctor(d_in_id):
    return dict(in_features=1, out_features=1, encoder_ids=None, d_in_id=d_in_id, in_edge_features=1,
                d_in_node_encoder=[28], d_in_edge_encoder=[4], encoder_degrees=None, d_degree=None)


def cycle_edge_lists():
    import networkx as nx
    return [list(nx.cycle_graph(k).edges) for k in range(3, K_MAX + 1)]


class Clocks:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md)"""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thr.join(timeout=2)
        return False

    def summary(self):
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def build_batches(batch_size, n_batches, seed0):
    from gsn_b200.synthetic import zinc_like_batch
    return [zinc_like_batch(batch_size, seed=seed0 + i, distinct=4096 if batch_size > 8192 else None)
            for i in range(n_batches)]


def to_tensors(b, pad_to=None, pin=False, device=None):
    """numpy batch -> tensors (padded to fixed N/E so that one CUDA graph serves every step)"""
    t = {'edge_index': torch.from_numpy(b['edge_index']), 'node_ptr': torch.from_numpy(b['node_ptr']),
         'x': torch.from_numpy(b['x']), 'edge_features': torch.from_numpy(b['edge_features']),
         'batch': torch.from_numpy(b['batch']), 'degrees': torch.from_numpy(b['degrees'])}
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    if device is not None:
        t = {k: v.to(device) for k, v in t.items()}
    return t


def batch_variants(b, n, seed):
    """n distinct batches with the SAME shapes as b (CUDA graphs need static shapes): the graphs of b in a different
    order (so edge_index, node_ptr, batch differ) with fresh random atom / bond types.  Stacked along a leading axis."""
    rng = np.random.default_rng(seed)
    node_ptr, edge_ptr, ei = b['node_ptr'], b['edge_ptr'], b['edge_index']
    G = len(node_ptr) - 1
    sizes, esizes = np.diff(node_ptr), np.diff(edge_ptr)
    g_of_e = np.repeat(np.arange(G), esizes)
    local = ei - node_ptr[g_of_e][None, :]
    N, E = int(node_ptr[-1]), int(edge_ptr[-1])
    out = {k: [] for k in ('edge_index', 'node_ptr', 'x', 'edge_features', 'batch', 'degrees')}
    for _ in range(n):
        perm = rng.permutation(G)                          # new position p holds old graph perm[p]
        nptr = np.concatenate([[0], np.cumsum(sizes[perm])]).astype(np.int64)
        e_order = np.concatenate([np.arange(edge_ptr[g], edge_ptr[g + 1]) for g in perm])
        pos_of_e = np.repeat(np.arange(G), esizes[perm])
        new_ei = local[:, e_order] + nptr[pos_of_e][None, :]
        out['edge_index'].append(new_ei)
        out['node_ptr'].append(nptr)
        out['x'].append(rng.integers(0, 28, size=(N, 1), dtype=np.int64))
        out['edge_features'].append(rng.integers(1, 4, size=(E, 1), dtype=np.int64))
        out['batch'].append(np.repeat(np.arange(G, dtype=np.int64), sizes[perm]))
        out['degrees'].append(np.bincount(new_ei[0], minlength=N).astype(np.float32))
    return {k: torch.from_numpy(np.stack(v)) for k, v in out.items()}


class Packing:
    """byte layout of one batch inside a single buffer (16-byte aligned fields): a step's inputs move with ONE copy"""

    def __init__(self, example):
        self.fields, off = [], 0
        for k, v in example.items():
            nbytes = v.numel() * v.element_size()
            self.fields.append((k, off, nbytes, v.dtype, tuple(v.shape)))
            off += (nbytes + 15) // 16 * 16
        self.nbytes = off

    def views(self, buf):
        return {k: buf[off:off + n].view(dt).view(shape) for k, off, n, dt, shape in self.fields}

    def pack_pool(self, stacked):
        P = next(iter(stacked.values())).shape[0]
        out = torch.zeros((P, self.nbytes), dtype=torch.uint8)
        for k, off, n, dt, shape in self.fields:
            out[:, off:off + n] = stacked[k].contiguous().view(P, -1).view(torch.uint8).view(P, n)
        return out


# ======================================================================================
# our arm
# ======================================================================================
def run_ours(args, rank, world, local_rank):
    from gsn_b200 import _lib, counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False      # fp32 parity (1e-5) needs full-precision GEMMs
    torch.backends.cudnn.allow_tf32 = False
    _lib.lib()

    B = args.batch
    els = cycle_edge_lists()
    sds = patterns.make_subgraph_dicts(els, 'local')
    # one fixed-shape batch per rank: CUDA graphs need static shapes, so every step re-runs the same shapes
    # with different CONTENT (a pool of distinct batches padded to common N/E would be equivalent)
    pool = build_batches(B, 1, seed0=1000 * rank)
    calib = build_batches(min(2048, max(B, 512)), 1, seed0=77)[0]
    ids_cal = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']),
                                   sds, False, 'local', max_nodes_per_graph=64)
    encoder = UniqueEncoder.fit(ids_cal)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**model_ctor(encoder.d), **model_args(encoder.d)).to(dev).eval()
    pipe = GSNPipeline(model, sds, False, 'local', encoder, max_nodes_per_graph=64)
    b0 = pool[0]
    N, E, G = int(b0['node_ptr'][-1]), int(b0['edge_index'].shape[1]), B
    dev_in = to_tensors(b0, device=dev)
    host_in = to_tensors(b0, pin=True)

    # ---- correctness of the captured step vs the eager step (cheap sanity, not the parity test)
    with torch.no_grad():
        ref_out = pipe.step(dev_in).clone()
    launches0 = _lib.launch_count()
    pipe.capture(dev_in, warmup=max(3, args.warmup))
    with torch.no_grad():
        l0 = _lib.launch_count()
        pipe.step(dev_in)
        my_launches_per_step = _lib.launch_count() - l0
    out = pipe.replay().clone()
    torch.cuda.synchronize()
    if os.environ.get('GSN_PROFILE_REPLAY'):      # ncu --profile-from-start off --graph-profiling node
        torch.cuda.profiler.start()
        pipe.replay()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    assert torch.allclose(out, ref_out, atol=1e-5, rtol=1e-5), 'captured step differs from eager step'
    assert int(pipe.last_status.item()) == 0

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    dist_on = world > 1

    def barrier():
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_steps(fn, steps):
        """K steps, each bracketed by its own event pair with an L2 flush in between (outside the pairs)"""
        evs = []
        barrier()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)      # ms

    # ---- single stream, one step at a time, L2 flushed before every step (the latency of ONE step)
    for _ in range(args.warmup):
        pipe.replay()
    with Clocks(local_rank) as clk:
        ms_single = timed_steps(lambda: pipe.replay(), args.steps) / args.steps

        # ---- value: whole-job throughput.  At B=128 one step is a chain of ~26 dependent launches that each fill a
        # fraction of the 148 SMs, so S steps on DIFFERENT batches are kept in flight (S captured graphs on S streams),
        # as a server overlapping consecutive batches would.  Every step copies ITS OWN batch out of a pool of
        # `--pool` distinct batches (same shapes, other graph order / features; > L2 in total, no batch is used twice
        # inside a timed region that fits the pool) into the graph's input buffers, inside the timed region; L2 is
        # flushed once before the region (steps overlap, so it cannot be flushed between them).
        S, P = max(1, args.streams), max(args.pool, 2 * max(1, args.streams))
        variants = batch_variants(b0, P, seed=4242 + rank)
        packing = Packing({k: variants[k][0] for k in dev_in})
        pool_host = packing.pack_pool(variants).pin_memory()          # [P, nbytes] uint8, one row = one batch
        pool_dev = pool_host.to(dev)
        packed = [torch.zeros(packing.nbytes, dtype=torch.uint8, device=dev) for _ in range(S)]
        pipes = [GSNPipeline(model, sds, False, 'local', encoder, max_nodes_per_graph=64).capture(
            dev_in, warmup=3, static=packing.views(packed[i])) for i in range(S)]
        streams = [torch.cuda.Stream() for _ in range(S)]
        out_host = torch.empty((args.steps + args.warmup, G, 1), dtype=torch.float32).pin_memory()

        def run_steps(k, start, src, d2h):
            main = torch.cuda.current_stream()
            for st in streams:
                st.wait_stream(main)
            for i in range(k):
                with torch.cuda.stream(streams[i % S]):
                    packed[i % S].copy_(src[(start + i) % P], non_blocking=True)      # the step's inputs: one copy
                    o = pipes[i % S].replay()
                    if d2h:
                        out_host[i].copy_(o, non_blocking=True)
            for st in streams:
                main.wait_stream(st)

        def timed_region(src, d2h):
            run_steps(args.warmup, 0, src, d2h)
            barrier()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_steps(args.steps, args.warmup, src, d2h)
            e1.record()
            barrier()
            return e0.elapsed_time(e1)

        # the pooled batches go through the same captured step: check two of them against the eager step
        for v in (1, P - 1):
            tv = {key: variants[key][v].to(dev) for key in dev_in}
            with torch.no_grad():
                exp_v = pipe.step(tv).clone()
            packed[-1].copy_(pool_dev[v])
            got_v = pipes[-1].replay().clone()
            torch.cuda.synchronize()
            assert torch.allclose(got_v, exp_v, atol=1e-5, rtol=1e-5), 'pooled batch: captured step differs from eager step'
        ms_total = timed_region(pool_dev, False)
        # ---- e2e: every step's inputs come from PINNED HOST memory (H2D inside the region) and its predictions are
        # read back to the host (D2H inside the region); same S streams
        ms_e2e = timed_region(pool_host, True)
        assert bool(torch.isfinite(out_host[:args.steps]).all())
    h2d = packing.nbytes
    d2h = out_host[0].numel() * out_host.element_size()

    if dist_on:
        tt = torch.tensor([ms_total, ms_e2e, ms_single], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_e2e, ms_single = float(tt[0]), float(tt[1]), float(tt[2])
    ms_per_step = ms_total / args.steps
    value = world * G / (ms_per_step * 1e-3)
    e2e_value = world * G / (ms_e2e / args.steps * 1e-3)

    # ---- informative only: other stream counts on ONE repeated batch (no pool), to show where the overlap saturates
    concurrent = None
    if rank == 0 and not args.no_sweep:
        try:
            concurrent = []
            for S2 in (2, 4, 12):
                pipes2 = [pipe] + [GSNPipeline(model, sds, False, 'local', encoder, max_nodes_per_graph=64).capture(
                    dev_in, warmup=3) for _ in range(S2 - 1)]
                streams2 = [torch.cuda.Stream() for _ in range(S2)]
                K2 = max(args.steps, 40)
                for p_ in pipes2:
                    p_.load(dev_in)

                def run2(k):
                    main = torch.cuda.current_stream()
                    for st in streams2:
                        st.wait_stream(main)
                    for i in range(k):
                        with torch.cuda.stream(streams2[i % S2]):
                            pipes2[i % S2].replay()
                    for st in streams2:
                        main.wait_stream(st)
                run2(2 * S2)
                torch.cuda.synchronize()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run2(K2)
                e1.record()
                torch.cuda.synchronize()
                ms2 = e0.elapsed_time(e1)
                ok = all(torch.allclose(p_._out, ref_out, atol=1e-5, rtol=1e-5) for p_ in pipes2)
                concurrent.append({'streams': S2, 'steps': K2, 'graphs_per_s': K2 * G / (ms2 * 1e-3), 'ms_per_step': ms2 / K2,
                                   'outputs_match': bool(ok)})
                del pipes2
        except Exception as ex:
            concurrent = [{'error': repr(ex)[:200]}]

    line = None
    if rank == 0:
        # ---- instrumented eager pass: per-kernel device time with CUDA events on the launching stream
        def instrumented(tensors, steps):
            _lib.TIMER = []
            with torch.no_grad():
                for _ in range(steps):
                    flush.zero_()
                    pipe.step(tensors)
            torch.cuda.synchronize()
            agg = {}
            for tag, e0, e1 in _lib.TIMER:
                agg.setdefault(tag, []).append(e0.elapsed_time(e1))
            _lib.TIMER = None
            return {k: (float(np.mean(v)) * 1e-3, len(v) // steps) for k, v in agg.items()}   # seconds, calls/step

        peak, peak_src = peaks()
        prof = instrumented(dev_in, max(3, min(args.steps, 10)))
        dh = D_OUT

        from gsn_b200.fused import MERGE_MAX_ROWS, _merge_edge_columns
        n_id_cols = len(encoder.d)
        # index columns the layer-0 message kernel reads per edge after column grouping (fused.py)
        n_l0_cols = _merge_edge_columns(torch.zeros((sum(encoder.d) + 4, 1)), _edge_cols(encoder), MERGE_MAX_ROWS)[1]['n_groups']

        def scatter_bytes(n, e):
            """algorithmic bytes of ONE launch of the message (scatter) kernel, averaged over the N_LAYERS launches
            of a step.  Layer 0 reads only indices (x: 4 B/node, identifiers + bond type folded
            into n_l0_cols grouped row indices of 4 B per edge); layers >= 1 read P [N, 2dh] fp32 and one 4-byte index per edge.
            Every launch reads the CSR (rowptr + nbr, 4 B each) and writes S [N, dh] fp32."""
            csr = 4 * e + 4 * (n + 1)
            out = 4 * dh * n
            layer0 = 4 * n + 4 * n_l0_cols * e + csr + out
            later = 4 * 2 * dh * n + 4 * e + csr + out
            return (layer0 + (N_LAYERS - 1) * later) / N_LAYERS
        # device duration of the message kernel at the step's own shapes: R back-to-back launches inside one event
        # pair (a single ~5 us launch bracketed by its own events mostly measures launch gaps)
        t_sc_eager = prof['general_edge'][0]
        t_sc = scatter_launch_seconds(dev, b0, encoder, reps=50)
        roof = {'bound': 'hbm', 'kernel': 'tab_tight_kernel (layer 0) + p1_tight_kernel (layers >= 1) behind gsn_mp_general_edge_idx_fwd, mean of the '
                                          f'{N_LAYERS} launches of a step',
                'achieved': scatter_bytes(N, E) / t_sc / 1e9, 'peak': peak, 'unit': 'GB/s',
                'frac': scatter_bytes(N, E) / t_sc / 1e9 / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 4 launches of one step (ncu --set full,
                # B=128): layer 0 (tab_tight_kernel, grouped columns) 0.182 MB read (profiles/r1g_ncu_tab_tight_b128.txt);
                # layers >= 1 (p1_tight_kernel) 3.20 MB read (profiles/r1e_ncu_tight_scatter_b128.txt); 0 B written in both
                # -- the 1.5 MB S output stays in the 126 MB L2
                'traffic': (0.182272e6 + 3 * 3.200768e6) / 4 if (B == 128 and N_LAYERS == 4) else None,
                'peak_source': peak_src,
                'algorithmic_bytes_per_launch': scatter_bytes(N, E), 'avg_launch_us': t_sc * 1e6,
                'launches_per_step': prof['general_edge'][1],
                'avg_launch_us_single_eager': t_sc_eager * 1e6,
                'how': 'mean device time of the 4 message-kernel launches of a step (layer 0 + 3 identifier-free layers) at '
                       'the step shapes, 50 launches replayed from a CUDA graph per CUDA-event pair; at B=128 one launch moves ~2 MB '
                       '(0.3 us at HBM peak) so it is launch-latency bound: see roofline_large_batch, sweep[] and '
                       'scatter_kernels[] for batches where HBM traffic dominates',
                'traffic_note': 'ncu --set full (profiles/): dram read+write per launch == algorithmic bytes within 1 % '
                                'for the dense kernel at B=32,768'}
        kernels_us = {k: round(v[0] * 1e6, 2) for k, v in prof.items()}
        sweep = []
        if not args.no_sweep:
            for Bs in (4096, 131072):
                try:
                    bb = build_batches(Bs, 1, seed0=5)[0]
                    tens = to_tensors(bb, device=dev)
                    n_s, e_s = int(bb['node_ptr'][-1]), int(bb['edge_index'].shape[1])
                    with torch.no_grad():
                        for _ in range(2):
                            pipe.step(tens)
                    torch.cuda.synchronize()
                    t0 = torch.cuda.Event(enable_timing=True)
                    t1 = torch.cuda.Event(enable_timing=True)
                    reps = 3
                    t0.record()
                    with torch.no_grad():
                        for _ in range(reps):
                            pipe.step(tens)
                    t1.record()
                    torch.cuda.synchronize()
                    ms = t0.elapsed_time(t1) / reps
                    pr = instrumented(tens, 3)
                    ts = pr['general_edge'][0]
                    sweep.append({'batch': Bs, 'N': n_s, 'E': e_s, 'graphs_per_s': Bs / (ms * 1e-3),
                                  'edges_per_s': e_s / (ms * 1e-3), 'ms_per_step': ms,
                                  'scatter_GBps': scatter_bytes(n_s, e_s) / ts / 1e9,
                                  'scatter_frac_of_peak': scatter_bytes(n_s, e_s) / ts / 1e9 / peak,
                                  'count_ms': pr['count_pattern'][0] * 1e3, 'graph_build_ms': pr['graph_build'][0] * 1e3,
                                  'kernels_us': {k: round(v[0] * 1e6, 1) for k, v in pr.items()}})
                    del tens
                    torch.cuda.empty_cache()
                except Exception as ex:      # the sweep is informative only; never lose the headline line
                    sweep.append({'batch': Bs, 'error': repr(ex)[:200]})
        roof_large = None
        scatter_kernels = []
        if not args.no_sweep:
            try:
                scatter_kernels = scatter_microbench(dev, flush, peak)
            except Exception as ex:
                scatter_kernels = [{'error': repr(ex)[:200]}]
        if not args.no_sweep:
            try:
                bl = build_batches(131072, 1, seed0=5)[0]
                nl, el = int(bl['node_ptr'][-1]), int(bl['edge_index'].shape[1])
                tl = scatter_launch_seconds(dev, bl, encoder, reps=3, flush=flush)
                roof_large = {'bound': 'hbm', 'kernel': roof['kernel'], 'batch': 131072, 'N': nl, 'E': el,
                              'achieved': scatter_bytes(nl, el) / tl / 1e9, 'peak': peak, 'unit': 'GB/s',
                              'frac': scatter_bytes(nl, el) / tl / 1e9 / peak, 'avg_launch_us': tl * 1e6,
                              'algorithmic_bytes_per_launch': scatter_bytes(nl, el),
                              'traffic': 'ncu --set full at B=131,072 (profiles/r1d_ncu_tight_scatter_b131072.txt): dram '
                                         'read 3.18 GB + write 1.53 GB per identifier-free launch vs 4.72 GB algorithmic'}
            except Exception as ex:
                roof_large = {'error': repr(ex)[:200]}
        # reported baselines: rank 0 at N=1 only (the multi-GPU runs of the scaling sweep stay short)
        cpu = cpu_baseline(pool[0], sds_oracle(), encoder, model, budget_s=12.0) if world == 1 else None
        eager = None
        if not args.no_sweep and world == 1:
            try:
                ids_dev = counting.count_batch(dev_in['edge_index'], dev_in['node_ptr'], sds, False, 'local',
                                               max_nodes_per_graph=64)
                eager = torch_eager_gpu(pool[0], ids_dev, encoder, model, dev)
                eager['max_abs_diff_vs_ours'] = float((eager.pop('out') - ref_out).abs().max())
            except Exception as ex:
                eager = {'error': repr(ex)[:200]}
        line = {
            'metric': 'graphs/sec preprocess+forward (ZINC batch)', 'value': value, 'unit': 'graphs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64 counts + fp32 forward',
            'data': 'synthetic',
            'config': {'workload': f'ZINC-shaped synthetic batch B={B} per GPU (N={N}, E={E}); COUNT cycles k<=8 edge '
                                   f'scope non-induced + one_hot_unique encode + GNNSubstructures forward '
                                   f'(GSN_edge_sparse general, id_scope local, {N_LAYERS} layers, d_out {D_OUT})',
                       'batch_per_gpu': B, 'N': N, 'E': E, 'id_columns': encoder.d, 'l2': f'inputs larger than L2: every step reads its own batch from '
                       f'a pool of {P} distinct same-shape batches ({P * h2d / 1e6:.0f} MB > 126 MB L2), none reused inside a '
                       'timed region; one 256 MiB L2 flush before the region (the steps overlap); single_stream: L2 flushed '
                       'before every step', 'cuda_graph': True, 'streams': S, 'pool': P,
                       'parallelism': f'batch-sharded x{world}, no data-path collective'},
            'edges_per_s': world * E / (ms_per_step * 1e-3),
            'single_stream': {'ms_per_step': ms_single, 'graphs_per_s': world * G / (ms_single * 1e-3),
                              'note': 'one step at a time on one stream, L2 flushed before every step, one fixed batch: the '
                                      'LATENCY of a step; `value` keeps `streams` steps on different batches in flight'},
            'e2e': {'value': e2e_value, 'unit': 'graphs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(my_launches_per_step) * args.steps,     # our kernels; + 1 packed input copy per step
            'gpu_launches_per_step': int(my_launches_per_step),
            'clocks': clk.summary(), 'roofline': roof, 'cpu_baseline': cpu, 'kernels_us': kernels_us, 'sweep': sweep, 'scatter_kernels': scatter_kernels,
            'roofline_large_batch': roof_large, 'torch_eager_gpu': eager,
            'concurrent_streams': {'note': 'informative: other stream counts, ONE batch replayed (no pool), one L2 flush '
                                           'before the region',
                                   'runs': concurrent},
        }
    return line


def _edge_cols(encoder):
    """(vocabulary size, first table row) of the layer-0 edge columns: identifier ranks, then the bond type"""
    cols, o = [], 0
    for d in list(encoder.d) + [4]:
        cols.append((int(d), o))
        o += int(d)
    return cols


def scatter_launch_seconds(dev, batch, encoder, reps, flush=None):
    """mean device seconds of ONE message-kernel launch of a step on `batch` (weights: 1 layer-0 launch reading only
    indices + (N_LAYERS-1) launches reading P and one index column), timed as back-to-back launches"""
    from gsn_b200 import ops
    ei = torch.from_numpy(batch['edge_index']).to(dev)
    N, E, dh = int(batch['node_ptr'][-1]), int(ei.shape[1]), D_OUT
    plan = ops.EdgePlan(ei, N)
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn((N, 2 * dh), device=dev, generator=g)
    # no scale / shift operands: fused.py folds msg_fn's BatchNorm affine into P and the tables
    er1 = torch.randint(0, 4, (E, 1), device=dev, dtype=torch.int32)
    Te1 = torch.randn((4, dh), device=dev)
    n_id = sum(encoder.d)
    nr = torch.randint(0, 28, (N, 1), device=dev, dtype=torch.int32)
    Tn = torch.randn((28, 2 * dh), device=dev)
    from gsn_b200.fused import MERGE_MAX_ROWS, _merge_edge_columns
    Te0, eg = _merge_edge_columns(torch.randn((n_id + 4, dh), device=dev), _edge_cols(encoder), MERGE_MAX_ROWS)
    er0 = torch.randint(0, Te0.shape[0], (E, eg['n_groups']), device=dev, dtype=torch.int32)

    def later():
        ops.general_edge_idx(plan, dh, P=P, edge_rows=er1, Te=Te1, edge_rows_csr=True)

    def first():
        ops.general_edge_idx(plan, dh, node_rows=nr, Tn=Tn, edge_rows=er0, Te=Te0, edge_rows_csr=True)

    def run(fn):
        # the launches are recorded into a CUDA graph and replayed, so that the event pair brackets device work only
        # (an eager Python launch costs ~15 us of host time, more than the kernel itself at B=128)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                fn()
        gr.replay()
        torch.cuda.synchronize()
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps
    return (run(first) + (N_LAYERS - 1) * run(later)) / N_LAYERS


def scatter_microbench(dev, flush, peak, batch=131072):
    """the layer-API scatter kernels (what the drop-in layers launch) at a batch where HBM traffic >> launch latency:
    median of 5 event-timed launches, L2 flushed before each"""
    from gsn_b200 import ops
    b = build_batches(batch, 1, seed0=5)[0]
    ei = torch.from_numpy(b['edge_index']).to(dev)
    N, E, dh = int(b['node_ptr'][-1]), int(ei.shape[1]), D_OUT
    plan = ops.EdgePlan(ei, N)
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn((N, 2 * dh), device=dev, generator=g)
    Q = torch.randn((E, dh), device=dev, generator=g)
    x = torch.randn((N, dh), device=dev, generator=g)
    sc, sf = torch.rand(dh, device=dev) + 0.5, torch.randn(dh, device=dev)
    csr = 8 * E + 4 * (N + 1)

    def med(fn):
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return sorted(ts[2:])[2]
    out = []
    for name, fn, by in (
            ('general_edge_kernel (GSN_edge_sparse general, dense P + Q)', lambda: ops.general_edge(plan, P, Q, sc, sf),
             4 * 2 * dh * N + 4 * dh * E + 4 * dh * N + csr),
            ('ogb_kernel (GSN_edge_sparse_ogb, local ids)', lambda: ops.ogb_aggregate(plan, x, Q, True, Q, None),
             4 * dh * (2 * N + 2 * E) + csr),
            ('segsum_kernel (scatter-add of [E,128] messages)', lambda: ops.segment_sum(plan, Q), 4 * dh * (N + E) + 4 * E + 4 * (N + 1))):
        t = med(fn)
        out.append({'kernel': name, 'batch': batch, 'N': N, 'E': E, 'us': t * 1e6, 'algorithmic_bytes': by,
                    'GBps': by / t / 1e9, 'frac_of_peak': by / t / 1e9 / peak})
    return out


def sds_oracle():
    from oracle import count_vf2
    return count_vf2.make_subgraph_dicts(cycle_edge_lists(), 'local')


# ======================================================================================
# CPU arm: the oracle ports of the reference path, timed on the host cores
# ======================================================================================
def cpu_step_fn(batch, sds_o, encoder, model, threads):
    from oracle import count_c, mp_ref
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    vocab = [v.cpu().numpy() for v in encoder.vocab]
    args = model_args(encoder.d)
    args['d_in_id'] = encoder.d
    args['d_in_node_encoder'], args['d_in_edge_encoder'] = [28], [4]
    from gsn_b200.graph_filters.base import SparseFilter  # noqa: F401  (layer configs only)
    cfgs = []
    for i in range(N_LAYERS):
        cfgs.append(dict(uses_ids=(i == 0), uses_ef=True, msg_kind='general', id_scope='local',
                         flow='source_to_target', activation_name='relu', bn=True, degree_as_tag=False,
                         retain_features=True, edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder',
                         extend_dims=True))
    torch.set_num_threads(threads)

    def step():
        ids = count_c.count_batch(batch['node_ptr'], batch['edge_ptr'], batch['edge_index'], sds_o, False, 1,
                                  nthreads=threads)
        enc = np.stack([np.minimum(np.searchsorted(vocab[c], ids[:, c]), len(vocab[c]) - 1) for c in range(ids.shape[1])], 1)
        data = {'edge_index': torch.from_numpy(batch['edge_index']), 'batch': torch.from_numpy(batch['batch']),
                'x': torch.from_numpy(batch['x']), 'edge_features': torch.from_numpy(batch['edge_features']),
                'degrees': torch.from_numpy(batch['degrees']), 'identifiers': torch.from_numpy(enc)}
        with torch.no_grad():
            return mp_ref.gnn_substructures_forward(args, sd, data, cfgs)
    return step


def torch_eager_gpu(batch, ids_dev, encoder, model, dev, reps=20):
    """the incumbent on the box -- the reference's MP forward as eager PyTorch on the same GPU
    (oracle/mp_ref.py restates the reference layers op by op; COUNT has no PyTorch form, so this is the forward
    only, identifiers pre-computed and pre-encoded).  Baseline leg, like cpu_baseline: never the product path."""
    from oracle import mp_ref
    sd = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    args = model_args(encoder.d)
    args['d_in_id'] = encoder.d
    args['d_in_node_encoder'], args['d_in_edge_encoder'] = [28], [4]
    cfgs = [dict(uses_ids=(i == 0), uses_ef=True, msg_kind='general', id_scope='local', flow='source_to_target',
                 activation_name='relu', bn=True, degree_as_tag=False, retain_features=True,
                 edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder', extend_dims=True) for i in range(N_LAYERS)]
    data = {k: torch.from_numpy(batch[k]).to(dev) for k in ('edge_index', 'batch', 'x', 'edge_features', 'degrees')}
    data['identifiers'] = encoder(ids_dev)
    with torch.no_grad():
        for _ in range(3):
            out = mp_ref.gnn_substructures_forward(args, sd, data, cfgs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = mp_ref.gnn_substructures_forward(args, sd, data, cfgs)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    G = len(batch['node_ptr']) - 1
    return {'forward_ms': ms, 'forward_graphs_per_s': G / (ms * 1e-3), 'out': out,
            'what': 'reference MP forward (one-hot encoders, [E,d] gathers + cat, E-row message MLP, index_add_) as eager PyTorch '
                    'on the same B200, fp32 (TF32 off); forward only -- COUNT and the encoding are not included'}


def cpu_baseline(batch, sds_o, encoder, model, budget_s):
    threads = os.cpu_count() or 1
    step = cpu_step_fn(batch, sds_o, encoder, model, threads)
    step()
    t0, n = time.perf_counter(), 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 200:
            break
    dt = (time.perf_counter() - t0) / n
    G = len(batch['node_ptr']) - 1
    return {'value': G / dt, 'unit': 'graphs/s', 'cores': threads, 'kind': 'port',
            'sample': f'{n} repetitions of the same B={G} step (oracle/count_enum.c OpenMP all-maps COUNT + '
                      f'oracle/mp_ref.py PyTorch-CPU forward), {dt * 1e3:.1f} ms/step'}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle ports) on the host cores"""
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import UniqueEncoder
    from oracle import count_c
    B = args.batch
    sds_o = sds_oracle()
    batch = build_batches(B, 1, seed0=0)[0]
    calib = build_batches(512, 1, seed0=77)[0]
    ids_cal = count_c.count_batch(calib['node_ptr'], calib['edge_ptr'], calib['edge_index'], sds_o, False, 1)
    encoder = UniqueEncoder([torch.from_numpy(np.unique(ids_cal[:, c])) for c in range(ids_cal.shape[1])])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**model_ctor(encoder.d), **model_args(encoder.d)).eval()
    threads = os.cpu_count() or 1
    step = cpu_step_fn(batch, sds_o, encoder, model, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    N, E = int(batch['node_ptr'][-1]), int(batch['edge_index'].shape[1])
    v = B / dt
    sample = f'{args.steps} steps of one B={B} batch; COUNT = oracle/count_enum.c (all maps / |Aut|, OpenMP), ' \
             f'forward = oracle/mp_ref.py (reference layers restated, PyTorch CPU)'
    return {'impl': 'reference', 'metric': 'graphs/sec preprocess+forward (ZINC batch)', 'value': v,
            'unit': 'graphs/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int64 counts + fp32 forward', 'data': 'synthetic',
            'config': {'workload': f'ZINC-shaped synthetic batch B={B} (N={N}, E={E}); COUNT cycles k<=8 edge scope + '
                                   f'one_hot_unique encode + GNNSubstructures forward on the host CPU',
                       'batch_per_gpu': B, 'N': N, 'E': E},
            'cpu_baseline': {'value': v, 'unit': 'graphs/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': v, 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--no-sweep', action='store_true')
    ap.add_argument('--streams', type=int, default=8, help='independent steps in flight (CUDA streams / captured graphs)')
    ap.add_argument('--pool', type=int, default=640, help='distinct same-shape input batches cycled through (> L2 in total)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: gsn_b200 has no CPU path (use --impl reference for the CPU arm)')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG', 'WARN')      # keep stdout to the single JSON line (no 'NCCL version' banner)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    line = run_ours(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
