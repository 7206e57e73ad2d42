"""One-kernel forward (csrc/fused_model.cu) vs the reference's outputs, the per-layer fused path and the CPU oracle."""
import contextlib
import io
import os

import pytest
import torch

from oracle import mp_ref
from tests.conftest import GOLDEN
from tests.test_oracle_mp import model_layer_cfgs

pytestmark = pytest.mark.gpu
MODEL_GOLDEN = torch.load(os.path.join(GOLDEN, 'mp_models.pt'))
TOL = 1e-5          # BASELINE.json: forward activations within 1e-5 fp32 (relative to the output scale)


def _golden_model(name):
    from gsn_b200.network import GNNSubstructures
    c = MODEL_GOLDEN[name]
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**c['ctor'], **c['args'])
    model.load_state_dict(c['state_dict'], strict=True)
    model = model.cuda().eval()

    class B:
        pass
    b = B()
    for k, v in c['data'].items():
        setattr(b, k, v.cuda())
    G = int(c['data']['batch'].max()) + 1
    b.node_ptr = torch.searchsorted(c['data']['batch'], torch.arange(G + 1)).cuda()
    return model, b, c['out']


@pytest.mark.parametrize('unit', [None, 1, 3])
@pytest.mark.parametrize('name', ['zinc_gsnv_general', 'zinc_gsne_general', 'sr_general_local_nobn', 'mpnn_general'])
def test_one_kernel_forward_matches_reference_output(name, unit):
    """outputs of the UNMODIFIED reference model (tests/golden/mp_models.pt); hidden width 16 runs zero-padded to D=64"""
    from gsn_b200 import fused_model
    model, b, ref = _golden_model(name)
    assert fused_model.supported(model)
    fm = fused_model.FusedModel(model, graphs_per_unit=unit)
    out = fm(b)
    fm.raise_on_status()
    scale = max(float(ref.abs().max()), 1.0)
    torch.testing.assert_close(out.cpu(), ref, atol=TOL * scale, rtol=TOL)


def _zinc_setup(B, seed, d_out=128, id_scope='local'):
    import bench
    from gsn_b200 import counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import UniqueEncoder
    dev = torch.device('cuda')
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), id_scope)
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids_cal = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']),
                                   sds, False, id_scope, max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids_cal)
    torch.manual_seed(0)
    args = bench.model_args(enc.d)
    args['id_scope'] = id_scope
    if d_out != 128:
        n = len(args['d_out'])
        args.update(d_out=[d_out] * n, d_msg=[d_out] * n, d_h=[[d_out]] * n)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **args)
    g = torch.Generator().manual_seed(5)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    model = model.to(dev).eval()
    b = bench.build_batches(B, 1, seed0=seed)[0]
    t = bench.to_tensors(b, device=dev)
    return model, sds, enc, b, t, args


@pytest.mark.parametrize('id_scope', ['local', 'global'])
@pytest.mark.parametrize('B,unit,d_out', [(128, None, 128), (128, 7, 128), (700, None, 128), (128, None, 64), (300, 2, 96)])
def test_one_kernel_pipeline_vs_per_layer_fused_path(B, unit, d_out, id_scope):
    """BASELINE config 2 (ZINC-shaped batch, cycles k<=8): COUNT + encode + forward; the one-kernel forward against the
    per-layer tcgen05 path on the same inputs, layer by layer and at the output"""
    from gsn_b200 import fused, fused_model
    from gsn_b200.pipeline import GSNPipeline
    model, sds, enc, b, t, _ = _zinc_setup(B, 3, d_out, id_scope)
    with torch.no_grad():
        p_ref = GSNPipeline(model, sds, False, id_scope, enc, 64, fused='layers')
        p_one = GSNPipeline(model, sds, False, id_scope, enc, 64, fused='model')
        assert isinstance(p_one.fused, fused_model.FusedModel) and not isinstance(p_ref.fused, fused_model.FusedModel)
        p_one.fused.graphs_per_unit = unit
        p_one.use_tile_plan = unit is None          # a forced unit exercises the in-kernel cutting of runs into tiles
        p_one.fused.debug_x_out = True
        out_ref = p_ref.step(t)
        out_one = p_one.step(t)
        p_one.fused.raise_on_status()
    for i, (xr, xo) in enumerate(zip(p_ref.fused.last_x_interm[1:], p_one.fused.last_x_out)):
        scale = max(float(xr.abs().max()), 1.0)
        torch.testing.assert_close(xo[:, :xr.shape[1]], xr, atol=TOL * scale, rtol=TOL, msg=lambda m: f'layer {i}: {m}')
        assert float(xo[:, xr.shape[1]:].abs().max() if xo.shape[1] > xr.shape[1] else 0) == 0
    scale = max(float(out_ref.abs().max()), 1.0)
    torch.testing.assert_close(out_one, out_ref, atol=TOL * scale, rtol=TOL)


def test_one_kernel_pipeline_vs_oracle():
    """config 2 end to end against the CPU oracle stack (C COUNT oracle + fp32 restatement of the reference layers)"""
    import numpy as np
    import bench
    from gsn_b200.pipeline import GSNPipeline
    from oracle import count_c
    model, sds, enc, b, t, args = _zinc_setup(128, 3)
    with torch.no_grad():
        pipe = GSNPipeline(model, sds, False, 'local', enc, 64, fused='model')
        out = pipe.step(t)
        pipe.fused.raise_on_status()
    ids = count_c.count_batch(b['node_ptr'], b['edge_ptr'], b['edge_index'], bench.sds_oracle(), False, 1)
    vocab = [v.cpu().numpy() for v in enc.vocab]
    ranks = np.stack([np.minimum(np.searchsorted(vocab[c], ids[:, c]), len(vocab[c]) - 1) for c in range(ids.shape[1])], 1)
    args = dict(args)
    args.update(d_in_id=enc.d, d_in_node_encoder=[28], d_in_edge_encoder=[4])
    data = {'edge_index': torch.from_numpy(b['edge_index']), 'batch': torch.from_numpy(b['batch']),
            'x': torch.from_numpy(b['x']), 'edge_features': torch.from_numpy(b['edge_features']),
            'degrees': torch.from_numpy(b['degrees']), 'identifiers': torch.from_numpy(ranks)}
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref = mp_ref.gnn_substructures_forward(args, sd, data, model_layer_cfgs(args))
    scale = max(float(ref.abs().max()), 1.0)
    torch.testing.assert_close(out.cpu(), ref, atol=TOL * scale, rtol=TOL)


def test_one_kernel_wide_dynamic_range_rows():
    """the power-of-two row scaling keeps fp32 accuracy when the rows of a batch differ by many orders of magnitude"""
    from gsn_b200 import fused, fused_model
    model, b, _ = _golden_model('sr_general_local_nobn')
    g = torch.Generator().manual_seed(1)
    b.x = (b.x.float().cpu() * torch.exp(torch.randn(b.x.shape[0], generator=g) * 6).reshape(b.x.shape[0], *([1] * (b.x.dim() - 1)))).cuda()
    ref = fused.FusedForward(model)
    ref_out = ref(b)
    fm = fused_model.FusedModel(model)
    fm.debug_x_out = True
    out = fm(b)
    for xr, xo in zip(ref.last_x_interm[1:], fm.last_x_out):
        rows = xr.abs().amax(1, keepdim=True).clamp_min(1.0)          # per-row scale: rows are independent until the readout
        assert float(((xo[:, :xr.shape[1]] - xr).abs() / rows).max()) < 2 * TOL
    scale = max(float(ref_out.abs().max()), 1.0)
    torch.testing.assert_close(out, ref_out, atol=2 * TOL * scale, rtol=2 * TOL)


def test_tile_plan_greedy_packing():
    """gsn_tile_plan: whole graphs, <= 128 rows and <= 32 graphs per tile, greedy; oversized graphs alone; the forward
    with the plan equals the forward without it bit for bit (a row's result does not depend on its tile mates)"""
    from gsn_b200 import fused_model
    from gsn_b200.pipeline import GSNPipeline
    g = torch.Generator().manual_seed(3)
    for sizes in (torch.randint(9, 38, (128,), generator=g), torch.randint(0, 4, (300,), generator=g),
                  torch.tensor([128, 1, 127, 129, 5, 200, 64, 64, 1]), torch.tensor([7])):
        ptr = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)])
        status = torch.zeros(1, dtype=torch.int32, device='cuda')
        plan = fused_model.tile_plan(ptr.cuda(), int(ptr[-1]), status).cpu()
        exp, g0, G = [], 0, len(sizes)
        while g0 < G:
            g1 = g0
            while g1 < G and g1 - g0 < 32 and int(ptr[g1 + 1] - ptr[g0]) <= 128:
                g1 += 1
            g1 = max(g1, g0 + 1)
            exp.append(g0)
            g0 = g1
        assert int(plan[0]) == len(exp) <= plan.numel() - 2 and int(status.item()) == 0
        assert plan[1:1 + len(exp)].tolist() == exp and int(plan[1 + len(exp)]) == G
    model, sds, enc, b, t, _ = _zinc_setup(128, 3)
    with torch.no_grad():
        p = GSNPipeline(model, sds, False, 'local', enc, 64, fused='model')
        out_plan = p.step(t).clone()
        p.use_tile_plan = False
        out_unit = p.step(t).clone()
    assert torch.equal(out_plan, out_unit)


@pytest.mark.parametrize('name', ['zinc_gsnv_general', 'zinc_gsne_general', 'sr_general_local_nobn', 'mpnn_general'])
def test_jk_head_in_kernel_vs_projection_launches(name, monkeypatch):
    """the JK head (models_graph_classification.py:236-240) evaluated inside the model kernel equals the pooled rows
    projected by separate GEMM launches"""
    from gsn_b200 import fused_model
    model, b, ref = _golden_model(name)
    outs = []
    for flag in (True, False):
        monkeypatch.setattr(fused_model, 'JK_IN_KERNEL', flag)
        fm = fused_model.FusedModel(model)
        outs.append(fm(b).clone())
        fm.raise_on_status()
        assert (fm.jk is not None) == (flag and fm.proj[0] is None)
    scale = max(float(ref.abs().max()), 1.0)
    torch.testing.assert_close(outs[0], outs[1], atol=TOL * scale, rtol=TOL)
    torch.testing.assert_close(outs[0].cpu(), ref, atol=TOL * scale, rtol=TOL)


def test_one_kernel_rejects_oversized_graphs():
    from gsn_b200 import fused_model
    model, b, _ = _golden_model('zinc_gsne_general')
    assert not fused_model.supported(model, max_nodes_per_graph=129)
    assert fused_model.supported(model, max_nodes_per_graph=128)


def test_bucketed_pipeline_variable_shapes():
    """the reference's DataLoader yields a different (N, E) every step (main.py:243-258): batches padded to shape
    buckets (sentinel graphs, self loops) and replayed through captured CUDA graphs give the eager results on the
    unpadded batch, and revisiting a bucket does not capture again"""
    import bench
    from gsn_b200.pipeline import BucketedPipeline, GSNPipeline
    from gsn_b200.synthetic import zinc_like_batch
    model, sds, enc, _, _, _ = _zinc_setup(64, 3)
    dev = torch.device('cuda')
    bp = BucketedPipeline(model, sds, False, 'local', enc, 64, node_step=64, edge_step=128)
    eager = GSNPipeline(model, sds, False, 'local', enc, 64)
    raws = [zinc_like_batch(48, seed=100 + i) for i in range(6)] + [zinc_like_batch(17, seed=7)]
    shapes = {(int(b['node_ptr'][-1]), b['edge_index'].shape[1]) for b in raws}
    assert len(shapes) == len(raws)                                  # genuinely different batches
    outs = []
    for rnd in range(2):
        for i, b in enumerate(raws):
            key, packed, G = bp.prepare(b, dev, pin=(i % 2 == 0))
            out = bp.run(key, packed.to(dev) if i % 3 == 0 else packed, G).clone()
            torch.cuda.synchronize()
            with torch.no_grad():
                exp = eager.step(bench.to_tensors(b, device=dev))
            assert out.shape == exp.shape
            scale = max(float(exp.abs().max()), 1.0)
            torch.testing.assert_close(out, exp, atol=TOL * scale, rtol=TOL)
            assert int(bp._lru[key][0].last_status.item()) == 0
            outs.append(out)
        if rnd == 0:
            first = bp.captures
    assert bp.captures == first <= len(raws)
    # padding never leaks: a batch alone and the same batch after another one of the bucket give identical bits
    for a, b_ in zip(outs[:len(raws)], outs[len(raws):]):
        assert torch.equal(a, b_)
    # submit(): the same step through one native call on an explicit stream, inputs from pinned host memory,
    # predictions read back to pinned host memory
    st = torch.cuda.Stream()
    for i, b in enumerate(raws[:3]):
        key, packed, G = bp.prepare(b, dev, pin=True)
        host = torch.empty((G, outs[i].shape[1]), dtype=torch.float32).pin_memory()
        st.wait_stream(torch.cuda.current_stream())
        got = bp.submit(key, packed, G, st, host)
        st.synchronize()
        assert torch.equal(host, outs[i].cpu()) and torch.equal(got.cpu(), host)
