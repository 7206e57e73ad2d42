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
REGION_REPEATS = 3      # timed regions per reported number (median)
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
    """numpy batch -> tensors"""
    t = {'edge_index': torch.from_numpy(b['edge_index']), 'node_ptr': torch.from_numpy(b['node_ptr']),
         'x': torch.from_numpy(b['x']), 'edge_features': torch.from_numpy(b['edge_features']),
         'batch': torch.from_numpy(b['batch']), 'degrees': torch.from_numpy(b['degrees'])}
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    if device is not None:
        t = {k: v.to(device) for k, v in t.items()}
    return t


def workload_string(B):
    """identical in both arms (the driver compares config.workload)"""
    return (f'ZINC-shaped synthetic batches of B={B} molecules per GPU (mean 23.2 nodes / 49.8 directed edges per graph); '
            f'step = COUNT cycles k<={K_MAX} edge scope non-induced + one_hot_unique encode + GNNSubstructures forward '
            f'(GSN_edge_sparse general, id_scope local, {N_LAYERS} layers, d_out {D_OUT})')


FLOPS_PER_NODE = 2 * (2 * D_OUT * D_OUT) + (N_LAYERS - 1) * 2 * (5 * D_OUT * D_OUT)
"""fp32-equivalent GEMM flops of one node through the re-associated model (DESIGN.md sec. 4): layer 0 (categorical
input) S.Wf^T and H.U2^T; layers >= 1 additionally x.Wxi^T, x.Wxj^T, x.U1x^T -- each a [D,D] matrix, 2 flops per MAC.
On the tensor cores every product costs 3 fp16 MMAs."""


def bytes_per_step(N, E, G):
    """HBM bytes one step has to move at the API boundary (every boundary tensor once):
    edge_index int64 [2,E], node_ptr int64 [G+1], x int64 [N,1], edge_features int64 [E,1] in; predictions fp32 [G,1] out"""
    return 16 * E + 8 * (G + 1) + 8 * N + 8 * E + 4 * G


# ======================================================================================
# our arm
# ======================================================================================
def other_configs(local_rank):
    """BASELINE configs 3 / 4 / 5 through their own scripts (scripts/bench_ogb.py, bench_imdb.py, bench_count.py) as child
    processes on the same GPU; each prints one JSON line.  A failure or a time-out is recorded, never raised."""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', str(local_rank)))
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    jobs = {'config3_molhiv_train_step_b512': ['scripts/bench_ogb.py', '--steps', '10', '--warmup', '3'],
            'config4_imdb_cliques_k5': ['scripts/bench_imdb.py', '--reps', '10'],
            'config5_count_cycles_k12_131072_graphs': ['scripts/bench_count.py', '--graphs', '131072', '--k', '12', '--check', '100']}
    out = {}
    for name, cmd in jobs.items():
        try:
            r = subprocess.run([sys.executable] + cmd, cwd=here, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                               text=True, timeout=240)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
            out[name] = json.loads(lines[-1]) if lines else {'error': f'no JSON line (exit {r.returncode})'}
        except Exception as ex:
            out[name] = {'error': repr(ex)[:200]}
    return out


def run_ours(args, rank, world, local_rank):
    from gsn_b200 import _lib, counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import BucketedPipeline, GSNPipeline, UniqueEncoder
    from gsn_b200.synthetic import zinc_like_batch
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cuda.matmul.allow_tf32 = False      # fp32 parity (1e-5) needs full-precision GEMMs
    torch.backends.cudnn.allow_tf32 = False
    _lib.lib()

    B = args.batch
    G = B
    els = cycle_edge_lists()
    sds = patterns.make_subgraph_dicts(els, 'local')
    calib = build_batches(min(2048, max(B, 512)), 1, seed0=77)[0]
    ids_cal = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']),
                                   sds, False, 'local', max_nodes_per_graph=64)
    encoder = UniqueEncoder.fit(ids_cal)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**model_ctor(encoder.d), **model_args(encoder.d)).to(dev).eval()

    # ---- the pool: P DISTINCT batches (own generator seed each: different molecules, different N and E), as a
    # DataLoader would deliver them (main.py:243-258).  Every batch is padded to its shape bucket (sentinel graphs /
    # self loops, gsn_b200/pipeline.py) and packed into one buffer at collate time; every (stream, bucket) pair owns
    # one captured CUDA graph of the whole step.
    S, P = max(1, args.streams), max(args.pool, 2 * max(1, args.streams))
    raw = [zinc_like_batch(B, seed=100_003 * (rank + 1) + i) for i in range(P)]
    bps = [BucketedPipeline(model, sds, False, 'local', encoder, max_nodes_per_graph=64) for _ in range(S)]
    keys, pool_host, shapes = [], [], []
    for b in raw:
        padded = bps[0].pad(b)
        key = (padded['x'].shape[0], padded['edge_index'].shape[1], padded['node_ptr'].numel() - 1)
        for bp in bps:
            bp._entry(key, padded, dev)
        keys.append(key)
        pool_host.append(bps[0].packing(padded).pack(padded).pin_memory())
        shapes.append((int(b['node_ptr'][-1]), int(b['edge_index'].shape[1])))
    pool_dev = [t.to(dev) for t in pool_host]
    pool_bytes = sum(t.numel() for t in pool_host)
    n_buckets = len(set(keys))
    N_mean, E_mean = float(np.mean([s_[0] for s_ in shapes])), float(np.mean([s_[1] for s_ in shapes]))

    # ---- correctness of the captured + padded step vs the eager step on the UNPADDED batch (cheap sanity, not the parity test)
    pipe = GSNPipeline(model, sds, False, 'local', encoder, max_nodes_per_graph=64)
    for v in (0, 1, P // 2, P - 1):
        with torch.no_grad():
            exp_v = pipe.step(to_tensors(raw[v], device=dev)).clone()
        got_v = bps[-1].run(keys[v], pool_dev[v], G).clone()
        torch.cuda.synchronize()
        assert torch.allclose(got_v, exp_v, atol=1e-5, rtol=1e-5), 'padded + captured step differs from the eager step'
        assert int(pipe.last_status.item()) & _lib.S_FATAL == 0 and int(pipe.fused.status.item()) & _lib.S_FATAL == 0
    dev_in = to_tensors(raw[0], device=dev)
    with torch.no_grad():
        ref_out = pipe.step(dev_in).clone()
        l0 = _lib.launch_count()
        pipe.step(dev_in)
        my_launches_per_step = _lib.launch_count() - l0
    if os.environ.get('GSN_PROFILE_REPLAY'):      # ncu --profile-from-start off
        torch.cuda.profiler.start()
        bps[0].run(keys[0], pool_dev[0], G)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    dist_on = world > 1

    def barrier():
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    streams = [torch.cuda.Stream() for _ in range(S)]
    out_host = torch.empty((args.steps + args.warmup, G, 1), dtype=torch.float32).pin_memory()

    def run_steps(k, start, src, d2h):
        main = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(main)
        for i in range(k):
            s_ = i % S
            idx = (start + i) % P
            # the step's inputs (one packed copy), the replay and the read-back of the predictions: one native call
            bps[s_].submit(keys[idx], src[idx], G, streams[s_], out_host[i] if d2h else None)
        for st in streams:
            main.wait_stream(st)

    def timed_region(src, d2h):
        run_steps(args.warmup, 0, src, d2h)
        barrier()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        run_steps(args.steps, args.warmup, src, d2h)
        host_enqueue_us.append((time.perf_counter() - t_host) * 1e6 / args.steps)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    host_enqueue_us = []       # host time to enqueue one step (copy + graph launch), per timed region

    with Clocks(local_rank) as clk:
        # ---- single stream, one step at a time, L2 flushed before every step, a different batch every step: the latency
        # of ONE step (device time between an event pair around copy + replay)
        for i in range(args.warmup):
            bps[0].run(keys[i % P], pool_dev[i % P], G)
        evs = []
        cur_stream = torch.cuda.current_stream()
        barrier()
        for i in range(args.steps):
            idx = (args.warmup + i) % P
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bps[0].submit(keys[idx], pool_dev[idx], G, cur_stream)
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms_single = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        # ---- value: whole-job throughput with S steps on different batches in flight (S streams), inputs resident in HBM
        # each region = W warm-up steps + exactly K timed steps; a K = 50 region lasts 1.5 ms, so it is repeated and the MEDIAN
        # region is reported (max over ranks of one 1.5 ms region at 8 processes measured one rank's scheduling hiccup)
        ms_total = float(np.median([timed_region(pool_dev, False) for _ in range(REGION_REPEATS)]))
        # ---- e2e: every step's inputs come from PINNED HOST memory (H2D inside the region) and its predictions are read
        # back to the host (D2H inside the region); same S streams
        ms_e2e = float(np.median([timed_region(pool_host, True) for _ in range(REGION_REPEATS)]))
        assert bool(torch.isfinite(out_host[:args.steps]).all())
    h2d = int(np.mean([t.numel() for t in pool_host]))
    d2h = out_host[0].numel() * out_host.element_size()

    if dist_on:
        tt = torch.tensor([ms_total, ms_e2e, ms_single], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_e2e, ms_single = float(tt[0]), float(tt[1]), float(tt[2])
    ms_per_step = ms_total / args.steps
    value = world * G / (ms_per_step * 1e-3)
    e2e_value = world * G / (ms_e2e / args.steps * 1e-3)

    line = None
    if rank == 0:
        flush_ = flush

        # ---- instrumented eager pass: per-call device time with CUDA events on the launching stream
        def instrumented(p_, tensors, steps):
            _lib.TIMER = []
            with torch.no_grad():
                for _ in range(steps):
                    flush_.zero_()
                    p_.step(tensors)
            torch.cuda.synchronize()
            agg = {}
            for tag, e0, e1 in _lib.TIMER:
                agg.setdefault(tag, []).append(e0.elapsed_time(e1))
            _lib.TIMER = None
            return {k: (float(np.median(v)) * 1e-3, len(v) // steps) for k, v in agg.items()}   # seconds, calls/step

        hbm_peak, hbm_src = peaks()
        tc_peak, tc_src = tensor_peak()
        N0, E0 = shapes[0]
        prof = instrumented(pipe, dev_in, max(5, min(args.steps, 10)))
        kernels_us = {k: round(v[0] * 1e6, 2) for k, v in prof.items()}

        fmod = pipe.fused
        w_bytes = int(fmod.Whi.numel() * 2 * 2 + sum(F['vec'].numel() * 4 + sum(F[k].numel() * 4 for k in ('Tn', 'Te', 'Tu')
                                                                               if F[k] is not None) for F in fmod.fm_layers))
        d_id_local = int(sum(encoder.d))

        def forward_roofline(n, e, g, t_fm):
            fl = FLOPS_PER_NODE * n
            # the kernel's own operands: index inputs + CSR in, weights / tables once, predictions out
            by = 4 * n + 4 * e * 2 + 8 * (g + 1) + 4 * (n + 1) + 4 * e + 4 * g + w_bytes
            # every tensor at the reference's layer boundary once (one-hot x / identifiers / bond types as the
            # fp32 tensors the reference materialises, x in and out of every layer, edge_index 16 B per edge)
            by_8d = (4 * n * 28 + 4 * e * d_id_local + 4 * e * 4 + 16 * e + 4 * n * D_OUT) + \
                    (N_LAYERS - 1) * (4 * n * D_OUT + 4 * e * 4 + 16 * e + 4 * n * D_OUT) + w_bytes // 2
            return {'bound': 'tensor',
                    'kernel': 'fused_model_kernel<128> (gsn_fused_model_fwd): all layers of the model for a tile of whole graphs; '
                              'the dominant launch of a step',
                    'achieved': fl / t_fm / 1e12, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': fl / t_fm / 1e12 / tc_peak,
                    'peak_source': tc_src, 'algorithmic_flops_per_launch': fl, 'flops_per_node': FLOPS_PER_NODE,
                    'mma_flops_per_launch': 3 * fl, 'mma_frac_of_peak': 3 * fl / t_fm / 1e12 / tc_peak,
                    'avg_launch_us': t_fm * 1e6,
                    'hbm_view': {'algorithmic_bytes_per_launch': by, 'GBps': by / t_fm / 1e9, 'frac_of_hbm_peak': by / t_fm / 1e9 / hbm_peak,
                                 'peak': hbm_peak, 'peak_source': hbm_src,
                                 'note': 'the kernel reads index columns + CSR + weights and writes the predictions only: '
                                         'activations never leave the SM, so HBM is not what bounds it'},
                    'survey_8d_view': {'algorithmic_bytes_per_launch': by_8d, 'GBps': by_8d / t_fm / 1e9,
                                       'frac_of_hbm_peak': by_8d / t_fm / 1e9 / hbm_peak,
                                       'note': 'SURVEY 8(d) numerator: the tensors of the reference layer API read / written once '
                                               'per layer; the one-kernel forward does not move them at all'},
                    'how': 'median device time of the launch (CUDA events on the launching stream, L2 flushed before every step) '
                           'in an eager pass over the step; algorithmic flops = FLOPS_PER_NODE x nodes (fp32-equivalent; the '
                           'tensor cores execute 3 fp16 MMAs per product, mma_* counts those)'}
        roof = forward_roofline(N0, E0, G, prof['fused_model'][0])
        traffic = {}
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r2_ncu_traffic.json')) as f:
                traffic = json.load(f)['bytes_per_launch']
        except Exception:
            pass
        roof['traffic'] = traffic.get(str(B))            # ncu dram bytes of this kernel at this batch size (committed capture)
        roof['launches_per_step'] = prof['fused_model'][1]
        sweep = []
        roof_large = None
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
                    pr = instrumented(pipe, tens, 3)
                    sweep.append({'batch': Bs, 'N': n_s, 'E': e_s, 'graphs_per_s': Bs / (ms * 1e-3),
                                  'edges_per_s': e_s / (ms * 1e-3), 'ms_per_step': ms,
                                  'step_bytes_GBps': bytes_per_step(n_s, e_s, Bs) / (ms * 1e-3) / 1e9,
                                  'count_ms': pr['count_pattern'][0] * 1e3,
                                  'kernels_us': {k: round(v[0] * 1e6, 1) for k, v in pr.items()}})
                    if Bs == 131072:
                        roof_large = forward_roofline(n_s, e_s, Bs, pr['fused_model'][0])
                        roof_large['batch'] = Bs
                        roof_large['traffic'] = traffic.get(str(Bs))
                    del tens
                    torch.cuda.empty_cache()
                except Exception as ex:      # the sweep is informative only; never lose the headline line
                    sweep.append({'batch': Bs, 'error': repr(ex)[:200]})
        scatter_kernels = []
        if not args.no_sweep:
            try:
                scatter_kernels = scatter_microbench(dev, flush, hbm_peak)
            except Exception as ex:
                scatter_kernels = [{'error': repr(ex)[:200]}]
        # ---- GSN-v (id_scope global: the README.md:112 recipe as written), same pipeline, informative second line
        gsn_v = None
        if not args.no_sweep and world == 1:
            try:
                gsn_v = gsn_v_line(dev, raw[:64], flush, S)
            except Exception as ex:
                gsn_v = {'error': repr(ex)[:200]}
        # ---- the other BASELINE configs on this GPU (their own scripts, one JSON line each; informative, never fatal):
        # 3 = molhiv-recipe training step, 4 = IMDB-BINARY fixture (cliques k <= 5), 5 = counting throughput (cycles k <= 12,
        # a 131,072-graph sample of the 1 M-graph workload)
        other = None
        if not args.no_sweep and world == 1:
            other = other_configs(local_rank)
        # reported baselines: rank 0 at N=1 only (the multi-GPU runs of the scaling sweep stay short)
        cpu = cpu_baseline(raw[0], sds_oracle(), encoder, model, budget_s=12.0) if world == 1 else None
        eager = None
        if not args.no_sweep and world == 1:
            try:
                ids_dev = counting.count_batch(dev_in['edge_index'], dev_in['node_ptr'], sds, False, 'local',
                                               max_nodes_per_graph=64)
                eager = torch_eager_gpu(raw[0], ids_dev, encoder, model, dev)
                eager['max_abs_diff_vs_ours'] = float((eager.pop('out') - ref_out).abs().max())
            except Exception as ex:
                eager = {'error': repr(ex)[:200]}
        line = {
            'metric': 'graphs/sec preprocess+forward (ZINC batch)', 'value': value, 'unit': 'graphs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64 counts + fp32 forward',
            'data': 'synthetic',
            'config': {'workload': workload_string(B), 'batch_per_gpu': B, 'N_mean': N_mean, 'E_mean': E_mean,
                       'id_columns': encoder.d,
                       'l2': f'inputs larger than L2: every step reads its own batch from a pool of {P} DISTINCT batches (own '
                             f'generator seed each, {pool_bytes / 1e6:.0f} MB packed > 126 MB L2), none reused inside a timed '
                             f'region that fits the pool; one 256 MiB L2 flush before the region (the steps overlap); '
                             f'single_stream: L2 flushed before every step',
                       'shapes': f'variable (N, E) per batch, padded to {n_buckets} shape buckets (node step 128, edge step 256); '
                                 f'one captured CUDA graph per (stream, bucket)',
                       'cuda_graph': True, 'streams': S, 'pool': P, 'buckets': n_buckets,
                       'timed_regions': f'{REGION_REPEATS} regions of W warm-up + K timed steps each, median region reported',
                       'parallelism': f'batch-sharded x{world}, no data-path collective'},
            'edges_per_s': world * E_mean / (ms_per_step * 1e-3),
            'single_stream': {'ms_per_step': ms_single, 'graphs_per_s': world * G / (ms_single * 1e-3),
                              'note': 'one step at a time on one stream, L2 flushed before every step, a different batch every '
                                      'step: the LATENCY of a step; `value` keeps `streams` steps on different batches in flight'},
            'e2e': {'value': e2e_value, 'unit': 'graphs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(my_launches_per_step) * args.steps,     # our kernels; + 1 packed input copy per step
            'gpu_launches_per_step': int(my_launches_per_step),
            'host_enqueue_us_per_step': {'value': float(np.median(host_enqueue_us[:REGION_REPEATS])),
                                         'e2e': float(np.median(host_enqueue_us[REGION_REPEATS:])),
                                         'note': 'host time to enqueue a step (BucketedPipeline.submit = one native call: packed copy + graph launch [+ read-back]) in the two timed regions'},
            'clocks': clk.summary(), 'roofline': roof, 'cpu_baseline': cpu, 'kernels_us': kernels_us, 'sweep': sweep,
            'scatter_kernels': scatter_kernels, 'roofline_large_batch': roof_large, 'torch_eager_gpu': eager, 'gsn_v': gsn_v,
            'other_configs': other,
        }
    return line


def tensor_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['bf16_tflops']), 'measured (MEASURED_PEAKS.json bf16_tflops, burst; kind::f16 runs at the bf16 rate)'
    except Exception:
        return 2250.0, 'fallback (nominal dense bf16/fp16 2.25 PFLOP/s)'


def gsn_v_line(dev, raw, flush, S):
    """the same step with vertex-scope identifiers (GSN-v, id_scope global, README.md:112): latency of one captured
    step and throughput with S steps in flight, over 64 distinct batches"""
    from gsn_b200 import counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import BucketedPipeline, UniqueEncoder
    sds = patterns.make_subgraph_dicts(cycle_edge_lists(), 'global')
    calib = build_batches(512, 1, seed0=77)[0]
    ids_cal = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']),
                                   sds, False, 'global', max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids_cal)
    margs = model_args(enc.d)
    margs['id_scope'] = 'global'
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**model_ctor(enc.d), **margs).to(dev).eval()
    G = len(raw[0]['node_ptr']) - 1
    bps = [BucketedPipeline(model, sds, False, 'global', enc, max_nodes_per_graph=64) for _ in range(S)]
    keys, packed = [], []
    for b in raw:
        padded = bps[0].pad(b)
        key = (padded['x'].shape[0], padded['edge_index'].shape[1], padded['node_ptr'].numel() - 1)
        for bp in bps:
            bp._entry(key, padded, dev)
        keys.append(key)
        packed.append(bps[0].packing(padded).pack(padded).to(dev))
    P = len(raw)
    streams = [torch.cuda.Stream() for _ in range(S)]
    for i in range(8):
        bps[0].run(keys[i], packed[i], G)
    torch.cuda.synchronize()
    ts = []
    for i in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bps[0].run(keys[i % P], packed[i % P], G)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))

    def run(k):
        main = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(main)
        for i in range(k):
            with torch.cuda.stream(streams[i % S]):
                bps[i % S].run(keys[i % P], packed[i % P], G)
        for st in streams:
            main.wait_stream(st)
    run(2 * S)
    torch.cuda.synchronize()
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / P
    return {'config': 'GSN-v: id_scope global (vertex-scope cycle counts, README.md:112), otherwise the headline workload',
            'single_stream_ms_per_step': float(np.median(ts)), 'single_stream_graphs_per_s': G / (float(np.median(ts)) * 1e-3),
            'value_graphs_per_s': G / (ms * 1e-3), 'ms_per_step': ms, 'streams': S, 'distinct_batches': P,
            'id_columns': enc.d}


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


def settle_cpu(step, seconds=2.0, min_steps=3):
    """un-timed warm-up of the CPU arm: thread pools, page faults and the clock ramp of the host cores made the first
    dozens of steps up to 1.7x slower than the steady state (round 1: two different CPU numbers for one step)"""
    t0, n = time.perf_counter(), 0
    while n < min_steps or time.perf_counter() - t0 < seconds:
        step()
        n += 1


def cpu_baseline(batch, sds_o, encoder, model, budget_s):
    threads = os.cpu_count() or 1
    step = cpu_step_fn(batch, sds_o, encoder, model, threads)
    settle_cpu(step)
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
    settle_cpu(step)                     # same un-timed settling as the cpu_baseline leg of our arm: one CPU number
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
            'config': {'workload': workload_string(B), 'batch_per_gpu': B, 'N': N, 'E': E,
                       'where': 'host CPU: oracle ports of the reference path'},
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
    ap.add_argument('--fill-rows', type=float, default=None, help='experiment: gsn_b200.fused_model.FILL_ROWS')
    ap.add_argument('--streams', type=int, default=8, help='independent steps in flight (CUDA streams / captured graphs)')
    ap.add_argument('--pool', type=int, default=640, help='distinct same-shape input batches cycled through (> L2 in total)')
    args = ap.parse_args()
    if args.fill_rows is not None and args.impl == 'ours':
        from gsn_b200 import fused_model as _fm
        _fm.FILL_ROWS = float(args.fill_rows)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    # the driver reads ONE JSON line from stdout: everything else this process (or a library: NCCL prints its version
    # banner to stdout) writes goes to stderr; the line itself is written to the saved descriptor at the end
    real_stdout = os.fdopen(os.dup(1), 'w')
    sys.stdout.flush()
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if args.impl == 'reference':
        if rank == 0:
            emit(run_reference(args))
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: gsn_b200 has no CPU path (use --impl reference for the CPU arm)')
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # NCCL's banner / debug lines never reach stdout (one JSON line)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    line = run_ours(args, rank, world, local_rank)
    if rank == 0:
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
