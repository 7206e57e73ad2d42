"""Dense-tail kernels and the fused inference forward vs the reference outputs / the oracle."""
import contextlib
import io
import os

import pytest
import torch

from tests.tolerance import FWD, GRAD, close

from oracle import mp_ref
from tests.conftest import GOLDEN
from tests.test_oracle_mp import model_layer_cfgs

pytestmark = pytest.mark.gpu
MODEL_GOLDEN = torch.load(os.path.join(GOLDEN, 'mp_models.pt'))


@pytest.mark.parametrize('mode', ['auto', 'off', 'presplit', 'one_tile'])
@pytest.mark.parametrize('M,K1,K2,Nout', [(1, 7, 0, 5), (37, 28, 0, 256), (130, 128, 128, 128), (2924, 156, 0, 128),
                                          (5000, 64, 33, 70), (70000, 128, 128, 128), (128, 128, 0, 1), (300, 8, 0, 24),
                                          (1000, 300, 0, 600), (513, 600, 0, 300), (260, 128, 4, 200),
                                          (40000, 128, 0, 256), (38000, 64, 0, 70), (60000, 32, 0, 48), (45000, 128, 128, 128)])
def test_linear_kernel_vs_torch(M, K1, K2, Nout, mode, monkeypatch):
    """fp32 FFMA kernel (mode off) and tcgen05 3xTF32 kernel (mode auto, where shapes allow)"""
    from gsn_b200 import ops
    from gsn_b200 import _lib
    monkeypatch.setattr(ops, 'TENSOR_CORES', 'off' if mode == 'off' else 'auto')
    monkeypatch.setattr(ops, 'TC_PATH', {'presplit': 1, 'one_tile': 2}.get(mode, 0))     # GsnLinear.tc_path, per call
    g = torch.Generator().manual_seed(M + K1)
    A1 = torch.randn((M, K1), generator=g)
    A2 = torch.randn((M, K2), generator=g) if K2 else None
    W = torch.randn((Nout, K1 + K2), generator=g) / (K1 + K2) ** 0.5
    bias, rs, rv = torch.randn(Nout, generator=g), torch.rand(M, generator=g) * 3, torch.randn(Nout, generator=g)
    tab, tidx = torch.randn((11, Nout), generator=g), torch.randint(0, 11, (M,), generator=g, dtype=torch.int32)
    scale, shift = torch.rand(Nout, generator=g) + 0.5, torch.randn(Nout, generator=g)
    A = A1 if A2 is None else torch.cat((A1, A2), 1)
    ref = torch.relu((A.double() @ W.double().t() + rs.double()[:, None] * rv.double()[None] + tab.double()[tidx.long()]
                      + bias.double()) * scale.double() + shift.double()).float()
    c = lambda t: None if t is None else t.cuda()
    out = ops.linear(c(A1), c(W), bias=c(bias), A2=c(A2), row_scale=c(rs), row_vec=c(rv), tab_idx=c(tidx), tab=c(tab),
                     scale=c(scale), shift=c(shift), activation='relu')
    close(out, ref)
    out2 = ops.linear(c(A1), c(W), A2=c(A2), activation='identity')
    close(out2, (A.double() @ W.double().t()).float())
    acc = out2.clone()
    ops.linear(c(A1), c(W), A2=c(A2), out=acc, accumulate=True)
    close(acc, 2 * out2)


def test_pool_ptr_and_encode_rows():
    from gsn_b200 import ops
    g = torch.Generator().manual_seed(0)
    sizes = torch.randint(0, 9, (50,), generator=g)
    ptr = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)])
    x = torch.randn((int(ptr[-1]), 24), generator=g)
    batch = torch.repeat_interleave(torch.arange(50), sizes)
    for mean in (False, True):
        torch.testing.assert_close(ops.pool_ptr(x.cuda(), ptr.cuda(), mean).cpu(),
                                   mp_ref.pool(x, batch, 'mean' if mean else 'sum', 50), atol=1e-5, rtol=1e-5)
    ids = torch.randint(0, 40, (1000, 3), generator=g)
    vocab = [torch.unique(ids[:, c]) for c in range(3)]
    vcat = torch.cat(vocab).cuda()
    ptrs, o = [], 0
    for v in vocab:
        ptrs.append((o, o + v.numel()))
        o += v.numel()
    ids_c = ids.cuda()
    extra = torch.randint(0, 4, (1000,), generator=g).cuda()
    rows = ops.encode_rows([(ids_c[:, 0], ptrs[0], 0), (ids_c[:, 1], ptrs[1], 100), (ids_c[:, 2], ptrs[2], 200),
                            (extra, None, 300)], vcat, 1000, ids_c.device).cpu()
    perm = torch.randperm(1000, generator=g).to(torch.int32)
    rows_p = ops.encode_rows([(ids_c[:, 0], ptrs[0], 0), (extra, None, 300)], vcat, 1000, ids_c.device, perm=perm.cuda()).cpu()
    assert torch.equal(rows_p, rows[perm.long()][:, [0, 3]])
    for c in range(3):
        exp = torch.bucketize(ids[:, c], vocab[c]) + 100 * c
        assert torch.equal(rows[:, c].long(), exp)
    assert torch.equal(rows[:, 3].long(), extra.cpu() + 300)


@pytest.mark.parametrize('name', ['zinc_gsnv_general', 'zinc_gsne_general', 'sr_general_local_nobn', 'mpnn_general'])
def test_fused_forward_matches_reference_output(name):
    from gsn_b200 import fused
    from gsn_b200.network import GNNSubstructures
    c = MODEL_GOLDEN[name]
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**c['ctor'], **c['args'])
    model.load_state_dict(c['state_dict'], strict=True)
    model = model.cuda().eval()
    assert fused.supported(model)

    class B:
        pass
    b = B()
    for k, v in c['data'].items():
        setattr(b, k, v.cuda())
    G = int(c['data']['batch'].max()) + 1
    b.node_ptr = torch.searchsorted(c['data']['batch'], torch.arange(G + 1)).cuda()
    out = fused.FusedForward(model)(b)
    close(out, c['out'])


def test_fused_pipeline_vs_generic_pipeline_and_oracle():
    """BASELINE config 2 end to end (COUNT + one_hot_unique + forward) at B=128: fused path vs per-layer path vs
    the CPU oracle stack (C COUNT oracle + fp32 reference-layer restatement)"""
    import numpy as np
    import bench
    from gsn_b200 import counting, patterns
    from gsn_b200.network import GNNSubstructures
    from gsn_b200.pipeline import GSNPipeline, UniqueEncoder
    from oracle import count_c
    dev = torch.device('cuda')
    sds = patterns.make_subgraph_dicts(bench.cycle_edge_lists(), 'local')
    calib = bench.build_batches(512, 1, seed0=77)[0]
    ids_cal = counting.count_batch(torch.from_numpy(calib['edge_index']).to(dev), torch.from_numpy(calib['node_ptr']),
                                   sds, False, 'local', max_nodes_per_graph=64)
    enc = UniqueEncoder.fit(ids_cal)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = GNNSubstructures(**bench.model_ctor(enc.d), **bench.model_args(enc.d))
    g = torch.Generator().manual_seed(5)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
    model = model.to(dev).eval()
    b = bench.build_batches(128, 1, seed0=3)[0]
    t = bench.to_tensors(b, device=dev)
    with torch.no_grad():
        out_f = GSNPipeline(model, sds, False, 'local', enc, 64, fused=True).step(t)
        out_g = GSNPipeline(model, sds, False, 'local', enc, 64, fused=False).step(t)
    # oracle stack on the CPU
    ids = count_c.count_batch(b['node_ptr'], b['edge_ptr'], b['edge_index'], bench.sds_oracle(), False, 1)
    vocab = [v.cpu().numpy() for v in enc.vocab]
    ranks = np.stack([np.minimum(np.searchsorted(vocab[c], ids[:, c]), len(vocab[c]) - 1) for c in range(ids.shape[1])], 1)
    args = bench.model_args(enc.d)
    args.update(d_in_id=enc.d, d_in_node_encoder=[28], d_in_edge_encoder=[4])
    data = {'edge_index': torch.from_numpy(b['edge_index']), 'batch': torch.from_numpy(b['batch']),
            'x': torch.from_numpy(b['x']), 'edge_features': torch.from_numpy(b['edge_features']),
            'degrees': torch.from_numpy(b['degrees']), 'identifiers': torch.from_numpy(ranks)}
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref = mp_ref.gnn_substructures_forward(args, sd, data, model_layer_cfgs(args))
    scale = float(ref.abs().max())
    close(out_g, ref)
    close(out_f, ref)
