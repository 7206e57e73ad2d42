"""The drop-in claim, checked in the build container: the UNMODIFIED reference model files
(/root/reference/models_graph_classification.py, models_graph_classification_ogb_original.py) construct on top of
gsn_b200.graph_filters (aliased as `graph_filters`) and get the SAME state_dict keys and the same
seed-0 weights as on the reference's own layers.  Skipped where /root/reference is absent (the GPU box)."""
import contextlib
import importlib
import io
import os
import sys

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason='/root/reference is not present')
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _fresh_import(name):
    for k in [k for k in sys.modules if k == name or k.startswith(name + '.')]:
        del sys.modules[k]
    return importlib.import_module(name)


def _build(model_module, cls_name, ctor, args, alias):
    """import the reference model file with `graph_filters` bound to the reference's package or to ours"""
    ref_import.install()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'graph_filters' or k.startswith('graph_filters.')}
    try:
        if alias:
            import gsn_b200.graph_filters as ours
            sys.modules['graph_filters'] = ours
            for n in ('GSN_sparse', 'GSN_edge_sparse', 'GSN_edge_sparse_ogb', 'MPNN_sparse', 'MPNN_edge_sparse',
                      'MPNN_edge_sparse_ogb'):
                sys.modules[f'graph_filters.{n}'] = importlib.import_module(f'gsn_b200.graph_filters.{n}')
        mod = _fresh_import(model_module)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = getattr(mod, cls_name)(**ctor, **args)
        return model, mod
    finally:
        for k in [k for k in sys.modules if k == 'graph_filters' or k.startswith('graph_filters.')]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.modules.pop(model_module, None)


@pytest.mark.parametrize('fixture,module,cls', [('mp_models.pt', 'models_graph_classification', 'GNNSubstructures'),
                                                ('mp_ogb.pt', 'models_graph_classification_ogb_original', 'GNN_OGB')])
def test_unmodified_reference_models_construct_on_our_layers(fixture, module, cls):
    cases = torch.load(os.path.join(GOLDEN, fixture))
    for name, c in cases.items():
        ref_model, _ = _build(module, cls, c['ctor'], c['args'], alias=False)
        our_model, mod = _build(module, cls, c['ctor'], c['args'], alias=True)
        layer_mods = {type(l).__module__ for l in our_model.conv}
        assert all(m.startswith('gsn_b200.graph_filters') for m in layer_mods), (name, layer_mods)
        sd_ref, sd_our = ref_model.state_dict(), our_model.state_dict()
        assert list(sd_ref) == list(sd_our), name                       # same keys, same order
        for k in sd_ref:
            assert torch.equal(sd_ref[k], sd_our[k]), (name, k)          # same seed-0 initialisation
        our_model.load_state_dict(c['state_dict'], strict=True)          # and the golden checkpoint loads


def test_reference_model_forward_on_our_layers_matches_golden_on_cpu_inputs_is_refused():
    """our layers have no CPU path: the reference model built on them refuses CPU tensors instead of falling back"""
    c = torch.load(os.path.join(GOLDEN, 'mp_models.pt'))['zinc_gsne_general']
    model, _ = _build('models_graph_classification', 'GNNSubstructures', c['ctor'], c['args'], alias=True)
    model.load_state_dict(c['state_dict'], strict=True)
    model.eval()

    class B:
        pass
    b = B()
    for k, v in c['data'].items():
        setattr(b, k, v)
    with pytest.raises(RuntimeError, match='CUDA'):
        with torch.no_grad():
            model(b)
