"""BASELINE config 3: ogbg-molhiv recipe (README.md:121) -- GSN-v, graphlets k<=5 (induced, vertex scope: 29 patterns,
72 orbit columns), GNN_OGB with GSN_edge_sparse_ogb layers (5 layers, width 300, hidden 600, virtual node, dropout 0.5),
B = 512 synthetic molhiv-shaped graphs: COUNT time and one TRAINING step (forward + backward + Adam).

    python scripts/bench_ogb.py                    # this package on cuda:0
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_ogb.py
                                                   # N GPUs: every rank trains on its own B graphs (weak scaling), weights
                                                   # broadcast from rank 0, ONE flat-buffer NCCL all-reduce of the gradients
                                                   # per step (the only collective of the system; BatchNorm stays per shard)
    python scripts/bench_ogb.py --impl eager_torch # the reference's formulation op by op in eager PyTorch on the same GPU
                                                   # (index_select gathers, relu, index_add_ scatter, one nn.Embedding per
                                                   # column, cuBLAS MLPs, no CUDA graph): the incumbent on the box
    python scripts/bench_ogb.py --impl reference   # the unmodified reference model on the host CPU (needs /root/reference)

ours: fused ogb message forward + backward kernels, embedding-bag kernels for the 72 identifier columns / atom / bond
encoders, and the whole step (forward + BCE + backward [+ all-reduce] + Adam) replayed from CUDA graphs (--no-graph: eager).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
ATOM, BOND = [119, 4, 12, 12, 10, 6, 6, 2, 2], [5, 6, 2]


def graphlets():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'graphlets.npz'))
    out = []
    for k in (3, 4, 5):
        ptr, ed = z[f'k{k}_ptr'], z[f'k{k}_edges']
        out += [ed[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]
    return out


def model_args(L, d, dh, dropout):
    return dict(seed=0, model_name='GSN_edge_sparse_ogb', readout='mean', dropout_features=[dropout] * (L + 1),
                bn=[True] * L, final_projection=[False] * L + [True], residual=False, inject_ids=False,
                inject_edge_features=True, vn=True, vn_pooling='sum', input_vn_encoder='embedding',
                d_out_vn_encoder=d, d_out_vn=[d] * (L - 1), id_scope='global', d_msg=[d] * L, d_out=[d] * L,
                d_h=[[dh]] * L, aggr='add', flow='source_to_target', msg_kind='ogb', train_eps=[False] * L,
                activation_mlp='relu', bn_mlp=True, jk_mlp=False, degree_embedding='one_hot_encoder',
                degree_as_tag=[False] * L, retain_features=[False] + [True] * (L - 1), multi_embedding_aggr='sum',
                features_scope='full', input_node_encoder='atom_encoder', d_out_node_encoder=d,
                edge_encoder='bond_encoder', d_out_edge_encoder=[d] * L, id_embedding='embedding',
                d_out_id_embedding=d, d_out_degree_embedding=d, extend_dims=True, activation='relu')


def _clocks(index):
    """one nvidia-smi sample right after the timed region (B200_PROFILING.md clocks line)"""
    import subprocess
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(',')
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        return {'sm_mhz': float(out[0]), 'sm_max_mhz': float(out[1]),
                'reasons': [n for n, v in zip(names, out[2:]) if 'Active' in v and 'Not' not in v]}
    except Exception as ex:
        return {'error': repr(ex)[:100]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--batch', type=int, default=512)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--id-cap', type=int, default=64, help='identifier embedding rows per column (counts are clamped)')
    ap.add_argument('--no-graph', action='store_true', help='eager training step (no CUDA graph)')
    ap.add_argument('--profile', action='store_true', help='torch.profiler table of the eager step (top kernels by device time)')
    a = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), 'w')      # stdout carries the JSON line only (NCCL prints a banner to stdout)
    sys.stdout.flush()
    os.dup2(2, 1)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    gpu_impl = a.impl in ('ours', 'eager_torch')
    if world > 1 and gpu_impl:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')       # stdout carries the JSON line only
        torch.cuda.set_device(local)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    from gsn_b200.synthetic import zinc_like_batch
    b = zinc_like_batch(a.batch, seed=21 + rank, mean_nodes=25.5, sd_nodes=12.0, min_nodes=2, max_nodes=222)
    N, E = int(b['node_ptr'][-1]), int(b['edge_index'].shape[1])
    rng = np.random.default_rng(5)
    x = torch.from_numpy(np.stack([rng.integers(0, d, N) for d in ATOM], 1))
    ef = torch.from_numpy(np.stack([rng.integers(0, d, E) for d in BOND], 1))
    y = torch.from_numpy(rng.integers(0, 2, (a.batch, 1))).float()
    ei, node_ptr = torch.from_numpy(b['edge_index']), torch.from_numpy(b['node_ptr'])
    els = graphlets()
    L, d, dh = 5, 300, 600
    args = model_args(L, d, dh, 0.5)
    res = {'config': f'molhiv-shaped synthetic batch B={a.batch} (N={N}, E={E}); all_simple_graphs k<=5 induced, id_scope '
                     f'global ({len(els)} patterns); GNN_OGB + GSN_edge_sparse_ogb, {L} layers, d_out {d}, d_h {dh}, vn, '
                     'dropout 0.5; train step = forward + BCE + backward + Adam', 'impl': a.impl, 'steps': a.steps}

    class Obj:
        pass

    if gpu_impl:
        from gsn_b200 import counting, ops, patterns
        from gsn_b200.graph_filters import autograd as gf_autograd
        if a.impl == 'eager_torch':
            gf_autograd.PURE_TORCH, ops.EMBEDDING_BAG, ops.TC_TRAINING = True, False, False
        from gsn_b200.network import GNN_OGB
        dev = torch.device('cuda', local)
        torch.cuda.set_device(dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        sds = patterns.make_subgraph_dicts(els, 'global')
        ei_d = ei.to(dev)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            ids = counting.count_batch(ei_d, node_ptr, sds, True, 'global')
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            ids = counting.count_batch(ei_d, node_ptr, sds, True, 'global')
        e1.record()
        torch.cuda.synchronize()
        res['count_ms'] = e0.elapsed_time(e1) / 5
        res['id_columns'] = int(ids.shape[1])
        ids = ids.clamp_max(a.id_cap - 1)
        ctor = dict(in_features=9, out_features=1, encoder_ids=None, d_in_id=[a.id_cap] * ids.shape[1], in_edge_features=3,
                    d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = GNN_OGB(**ctor, **args).to(dev).train()
        data = Obj()
        data.edge_index, data.batch, data.x, data.edge_features = ei_d, torch.from_numpy(b['batch']).to(dev), x.to(dev), ef.to(dev)
        data.identifiers, data.degrees, data.node_ptr = ids, torch.from_numpy(b['degrees']).to(dev), node_ptr.to(dev)
        data.num_graphs = a.batch
        yd = y.to(dev)
        use_graph = a.impl == 'ours' and not a.no_graph
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=use_graph)
        if world > 1:
            from gsn_b200 import distributed as gd
            gd.broadcast_parameters(model, src=0)
        params = [p_ for p_ in model.parameters()]
        from gsn_b200 import _lib
        from gsn_b200.distributed import FlatGradients
        fg = FlatGradients(params) if a.impl == 'ours' else None      # gradients live in one flat buffer: 1 NCCL call

        def fwd_bwd():
            if fg is not None:
                fg.zero()
            else:
                opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(model(data), yd)
            loss.backward()
            return loss

        def step():
            loss = fwd_bwd()
            if world > 1:
                fg.allreduce() if fg is not None else gd.allreduce_gradients(params)     # one flat fp32 buffer, averaged
            opt.step()
            return loss
        ar_ms = None
        if use_graph:
            # whole step from CUDA graphs: forward + loss + backward in one graph, Adam (capturable) in a second one;
            # between them (N > 1 only) the one collective of the system, an NCCL all-reduce of the flat gradient buffer
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(3, a.warmup)):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            g_fb, g_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fb):
                static_loss = fwd_bwd()
            own_launches = _lib.launch_count() - l0
            with torch.cuda.graph(g_opt):
                opt.step()

            def step():                                  # noqa: F811
                g_fb.replay()
                if world > 1:
                    fg.allreduce()
                g_opt.replay()
                return static_loss
            for _ in range(a.warmup):
                step()
            torch.cuda.synchronize()
            if world > 1:                                # the all-reduce alone, device time
                e0, e1 = ev(), ev()
                e0.record()
                for _ in range(a.steps):
                    fg.allreduce()
                e1.record()
                torch.cuda.synchronize()
                ar_ms = e0.elapsed_time(e1) / a.steps
        else:
            for _ in range(a.warmup):
                step()
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            step()
            own_launches = _lib.launch_count() - l0
            torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            ms = float(tt[0])
        if a.profile and rank == 0:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    loss = step()
                torch.cuda.synchronize()
            sys.stderr.write(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90) + '\n')
        n_par = sum(p_.numel() for p_ in params)
        res.update(n_gpus=world, train_step_ms=ms, graphs_per_s=world * a.batch / (ms * 1e-3), loss=float(loss),
                   cuda_graph=bool(use_graph), gsn_kernel_launches_per_step=own_launches,
                   parameters=n_par, allreduce_bytes=4 * n_par if world > 1 else 0, allreduce_ms=ar_ms,
                   count_plus_step_graphs_per_s=world * a.batch / ((ms + res['count_ms']) * 1e-3),
                   clocks=_clocks(local))
        model.eval()
        with torch.no_grad():
            for _ in range(2):
                model(data)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(a.steps):
                model(data)
            e1.record()
            torch.cuda.synchronize()
        res['eval_forward_ms'] = e0.elapsed_time(e1) / a.steps
    else:
        from oracle import ref_import
        M = ref_import.models()
        ids = torch.from_numpy(rng.integers(0, a.id_cap, (N, 72)))      # counting is not timed here (graph-tool absent)
        ctor = dict(in_features=9, out_features=1, encoder_ids=None, d_in_id=[a.id_cap] * 72, in_edge_features=3,
                    d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = M['GNN_OGB'](**ctor, **args).train()
        data = Obj()
        data.edge_index, data.batch, data.x, data.edge_features = ei, torch.from_numpy(b['batch']), x, ef
        data.identifiers, data.degrees = ids, torch.from_numpy(b['degrees'])
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        ts = []
        for i in range(1 + min(a.steps, 3)):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(model(data), y)
            loss.backward()
            opt.step()
            ts.append(time.perf_counter() - t0)
        ms = float(np.median(ts[1:])) * 1e3
        res.update(train_step_ms=ms, graphs_per_s=a.batch / (ms * 1e-3), cores=torch.get_num_threads(),
                   note='unmodified reference model (models_graph_classification_ogb_original.py) on the host CPU, PyTorch '
                        'threads = cores; identifiers random (graph-tool absent, COUNT not timed)')
    if rank == 0:
        real_stdout.write(json.dumps(res) + '\n')
        real_stdout.flush()
    if world > 1 and gpu_impl:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
