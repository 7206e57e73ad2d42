"""MP golden vectors: runs the UNMODIFIED reference layers / models
(/root/reference/graph_filters/*.py, models_graph_classification.py) on CPU with
seeded random weights and inputs and stores config + inputs + state_dict +
outputs in tests/golden/mp_layers.pt and mp_models.pt.

    python scripts/make_golden.py mp
"""
import contextlib
import io
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def rand_graph(g, n_graphs=5, lo=4, hi=9, p=0.45, sort=True):
    """block-diagonal symmetric batch without self loops"""
    src, dst, batch, off = [], [], [], 0
    for gi in range(n_graphs):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        a = torch.rand((n, n), generator=g) < p
        a = torch.triu(a, 1)
        r, c = a.nonzero(as_tuple=True)
        src += (r + off).tolist() + (c + off).tolist()
        dst += (c + off).tolist() + (r + off).tolist()
        batch += [gi] * n
        off += n
    ei = torch.tensor([src, dst], dtype=torch.int64)
    if not sort:
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    return ei, torch.tensor(batch, dtype=torch.int64), off


def randomize(module, g):
    """non-trivial BatchNorm running stats / affine so eval-mode BN is exercised"""
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
    if hasattr(module, 'eps') and isinstance(getattr(module, 'eps'), torch.Tensor):
        module.eps.data.fill_(0.25)


LAYER_CASES = []


def case(name, cls, **kw):
    LAYER_CASES.append((name, cls, kw))


COMMON = dict(d_degree=3, degree_as_tag=False, retain_features=True, d_msg=12, d_up=16, d_h=[20], seed=0,
              activation_name='relu', bn=True, edge_embedding='one_hot_encoder', id_embedding='one_hot_encoder',
              extend_dims=True)
case('gsn_gin_local_onehot', 'GSN_sparse', d_in=1, d_id=7, id_scope='local', msg_kind='gin', **COMMON)
case('gsn_gin_global', 'GSN_sparse', d_in=8, d_id=8, id_scope='global', msg_kind='gin', train_eps=True, **COMMON)
case('gsn_general_local_nobn', 'GSN_sparse', d_in=8, d_id=5, id_scope='local', msg_kind='general',
     **{**COMMON, 'bn': False})
case('gsn_general_global', 'GSN_sparse', d_in=8, d_id=5, id_scope='global', msg_kind='general', **COMMON)
case('gsn_general_local_t2s_elu', 'GSN_sparse', d_in=6, d_id=4, id_scope='local', msg_kind='general',
     flow='target_to_source', **{**COMMON, 'activation_name': 'elu'})
case('gsne_general_global', 'GSN_edge_sparse', d_in=28, d_ef=4, d_id=9, id_scope='global', msg_kind='general', **COMMON)
case('gsne_general_local', 'GSN_edge_sparse', d_in=28, d_ef=4, d_id=9, id_scope='local', msg_kind='general', **COMMON)
case('gsne_general_local_tag', 'GSN_edge_sparse', d_in=5, d_ef=4, d_id=6, id_scope='local', msg_kind='general',
     **{**COMMON, 'degree_as_tag': True})
case('gsne_gin_local_onehot', 'GSN_edge_sparse', d_in=8, d_ef=4, d_id=8, id_scope='local', msg_kind='gin', **COMMON)
case('gsne_gin_local_embedding', 'GSN_edge_sparse', d_in=8, d_ef=8, d_id=8, id_scope='local', msg_kind='gin',
     **{**COMMON, 'edge_embedding': 'embedding', 'id_embedding': 'embedding'})
case('gsne_gin_global_noextend', 'GSN_edge_sparse', d_in=8, d_ef=4, d_id=8, id_scope='global', msg_kind='gin',
     **{**COMMON, 'extend_dims': False})
case('gsn_ogb_local', 'GSN_edge_sparse_ogb', d_in=12, d_ef=12, d_id=12, id_scope='local', msg_kind='ogb',
     **{**COMMON, 'd_up': 12})
case('gsn_ogb_global', 'GSN_edge_sparse_ogb', d_in=10, d_ef=10, d_id=10, id_scope='global', msg_kind='ogb',
     train_eps=True, **{**COMMON, 'd_up': 10})
case('mpnn_gin', 'MPNN_sparse', d_in=8, msg_kind='gin', **COMMON)
case('mpnn_general', 'MPNN_sparse', d_in=16, msg_kind='general', **COMMON)
case('mpnne_general', 'MPNN_edge_sparse', d_in=16, d_ef=4, msg_kind='general', **COMMON)
case('mpnne_gin', 'MPNN_edge_sparse', d_in=8, d_ef=4, msg_kind='gin', **COMMON)
case('mpnn_ogb', 'MPNN_edge_sparse_ogb', d_in=12, d_ef=12, msg_kind='ogb', **{**COMMON, 'd_up': 12})


def layer_inputs(kw, g, sort):
    ei, batch, n = rand_graph(g, sort=sort)
    E = ei.shape[1]
    inp = {'edge_index': ei, 'x': torch.randn((n, kw['d_in']), generator=g),
           'degrees': torch.randn((n, kw['d_degree']), generator=g)}
    if 'd_id' in kw:
        rows = E if kw['id_scope'] == 'local' else n
        if kw['msg_kind'] == 'ogb' or kw['id_embedding'] == 'embedding':
            inp['identifiers'] = torch.randn((rows, kw['d_id']), generator=g)
        else:   # one-hot rows like the reference's id encoder produces
            idx = torch.randint(0, kw['d_id'], (rows,), generator=g)
            inp['identifiers'] = torch.nn.functional.one_hot(idx, kw['d_id']).float()
    else:
        inp['identifiers'] = None
    if 'd_ef' in kw:
        inp['edge_features'] = torch.randn((E, kw['d_ef']), generator=g)
    return inp


def model_args(**over):
    """post-process_arguments dict (utils.py:94-161) for GNNSubstructures"""
    L = over.pop('num_layers', 3)
    d_out = over.pop('d_out', 16)
    a = dict(seed=0, model_name='GSN_edge_sparse', readout='sum', dropout_features=[0.0] * (L + 1), bn=[True] * L,
             final_projection=[False] * L + [True], inject_ids=False, inject_edge_features=True,
             random_features=False, id_scope='global', d_msg=[d_out] * L, d_out=[d_out] * L, d_h=[[d_out]] * L,
             aggr='add', flow='source_to_target', msg_kind='general', train_eps=[False] * L, activation_mlp='relu',
             bn_mlp=True, jk_mlp=True, degree_embedding='one_hot_encoder', degree_as_tag=[False] * L,
             retain_features=[False] + [True] * (L - 1), multi_embedding_aggr='sum',
             input_node_encoder='one_hot_encoder', d_out_node_encoder=d_out, edge_encoder='one_hot_encoder',
             d_out_edge_encoder=[d_out] * L, id_embedding='one_hot_encoder', d_out_id_embedding=d_out,
             d_out_degree_embedding=d_out, extend_dims=True, activation='relu')
    a.update(over)
    return a


def make_mp(out_dir):
    from oracle import ref_import
    warnings.filterwarnings('ignore')
    L = ref_import.layers()
    g = torch.Generator().manual_seed(1234)
    golden = {}
    for i, (name, cls, kw) in enumerate(LAYER_CASES):
        torch.manual_seed(100 + i)
        with contextlib.redirect_stdout(io.StringIO()):
            layer = L[cls](**kw)
        randomize(layer, g)
        layer.eval()
        inp = layer_inputs(kw, g, sort=(i % 2 == 0))
        with torch.no_grad():
            out = layer(inp['x'], inp['edge_index'], identifiers=inp['identifiers'], degrees=inp['degrees'],
                        edge_features=inp.get('edge_features'))
        golden[name] = {'cls': cls, 'ctor': kw, 'inputs': inp, 'state_dict': layer.state_dict(), 'out': out}
        print(name, tuple(out.shape), float(out.abs().mean()))
    torch.save(golden, os.path.join(out_dir, 'mp_layers.pt'))

    M = ref_import.models()
    mg = {}
    specs = {
        'zinc_gsnv_general': dict(args=model_args(), d_in_id=[3, 4, 2], n_x=[28], n_ef=[4], scope='global'),
        'zinc_gsne_general': dict(args=model_args(id_scope='local'), d_in_id=[3, 4, 2], n_x=[28], n_ef=[4], scope='local'),
        'imdb_gin_local': dict(args=model_args(model_name='GSN_sparse', msg_kind='gin', id_scope='local',
                                               readout='mean', jk_mlp=False, final_projection=[True] * 4,
                                               input_node_encoder='None', edge_encoder='None'),
                               d_in_id=[5, 6, 4], n_x=None, n_ef=None, scope='local'),
        'sr_general_local_nobn': dict(args=model_args(model_name='GSN_sparse', id_scope='local', bn=[False] * 3,
                                                      input_node_encoder='None', edge_encoder='None',
                                                      num_layers=2, d_out=16)
                                      if False else model_args(model_name='GSN_sparse', id_scope='local',
                                                               bn=[False] * 3, input_node_encoder='None',
                                                               edge_encoder='None'),
                                      d_in_id=[1, 3, 5], n_x=None, n_ef=None, scope='local'),
        'mpnn_general': dict(args=model_args(model_name='MPNN_edge_sparse'), d_in_id=[3], n_x=[28], n_ef=[4], scope='local'),
    }
    for j, (name, sp) in enumerate(specs.items()):
        torch.manual_seed(500 + j)
        args = sp['args']
        ei, batch, n = rand_graph(g, n_graphs=6, sort=(j % 2 == 0))
        E = ei.shape[1]
        rows = E if sp['scope'] == 'local' else n
        data = {'edge_index': ei, 'batch': batch,
                'identifiers': torch.stack([torch.randint(0, d, (rows,), generator=g) for d in sp['d_in_id']], 1),
                'degrees': torch.randint(0, 5, (n,), generator=g)}
        if sp['n_x'] is not None:
            data['x'] = torch.randint(0, sp['n_x'][0], (n, 1), generator=g)
            in_features = 1
        else:
            data['x'] = torch.ones((n, 1))
            in_features = 1
        if sp['n_ef'] is not None:
            data['edge_features'] = torch.randint(0, sp['n_ef'][0], (E, 1), generator=g)
        ctor = dict(in_features=in_features, out_features=1 if 'zinc' in name else 2, encoder_ids=None,
                    d_in_id=sp['d_in_id'], in_edge_features=1 if sp['n_ef'] else None, d_in_node_encoder=sp['n_x'],
                    d_in_edge_encoder=sp['n_ef'], encoder_degrees=None, d_degree=[5])
        with contextlib.redirect_stdout(io.StringIO()):
            model = M['GNNSubstructures'](**ctor, **args)
        randomize(model, g)
        model.eval()

        class Obj:
            pass
        d = Obj()
        for k, v in data.items():
            setattr(d, k, v)
        with torch.no_grad():
            out = model(d)
        mg[name] = {'ctor': ctor, 'args': args, 'data': data, 'state_dict': model.state_dict(), 'out': out}
        print(name, tuple(out.shape), out.flatten()[:3].tolist())
    torch.save(mg, os.path.join(out_dir, 'mp_models.pt'))


if __name__ == '__main__' and (len(sys.argv) < 2 or sys.argv[1] not in ('sr', 'ogb')):
    make_mp(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden'))


def make_sr_isomorphism(out_dir):
    """README.md:84 recipe on SR(25,12,5,6) with the reference's own model class and seed-0 initialisation:
    stores the weights and the 15 graph embeddings (test_isomorphism, train_test_funcs.py:262-277)."""
    import numpy as np
    from oracle import count_c, count_vf2, ref_import
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
    z = np.load(os.path.join(out_dir, 'sr251256.npz'))
    eis = z['edge_index'].astype(np.int64)
    node_ptr = np.arange(16, dtype=np.int64) * 25
    edge_ptr = np.arange(16, dtype=np.int64) * 300
    ei = np.concatenate([eis[i] + 25 * i for i in range(15)], 1)
    sds = count_vf2.make_subgraph_dicts(count_vf2.pattern_edge_lists('cycle_graph', 6), 'local')
    ids = count_c.count_batch(node_ptr, edge_ptr, ei, sds, True, 1)
    ranks = np.stack([np.unique(ids[:, c], return_inverse=True)[1] for c in range(ids.shape[1])], 1)
    d_id = [int(len(np.unique(ids[:, c]))) for c in range(ids.shape[1])]
    M = ref_import.models()
    L, d = 2, 64
    args = dict(seed=0, model_name='GSN_sparse', readout='sum', dropout_features=[0.0] * (L + 1), bn=[False] * L,
                final_projection=[False] * L + [True], inject_ids=False, inject_edge_features=True, random_features=False,
                id_scope='local', d_msg=[d] * L, d_out=[d] * L, d_h=[[d]] * L, aggr='add', flow='source_to_target',
                msg_kind='general', train_eps=[False] * L, activation_mlp='relu', bn_mlp=True, jk_mlp=True,
                degree_embedding='one_hot_encoder', degree_as_tag=[False] * L, retain_features=[False, True],
                multi_embedding_aggr='sum', input_node_encoder='None', d_out_node_encoder=d, edge_encoder='None',
                d_out_edge_encoder=[d] * L, id_embedding='one_hot_encoder', d_out_id_embedding=d,
                d_out_degree_embedding=d, extend_dims=True, activation='relu')
    ctor = dict(in_features=1, out_features=2, encoder_ids=None, d_in_id=d_id, in_edge_features=None,
                d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None)
    torch.manual_seed(0)                                   # main.py:43-50
    with contextlib.redirect_stdout(io.StringIO()):
        model = M['GNNSubstructures'](**ctor, **args)
    model.eval()

    class Obj:
        pass
    dobj = Obj()
    dobj.x = torch.ones((375, 1))
    dobj.edge_index = torch.from_numpy(ei)
    dobj.identifiers = torch.from_numpy(ranks)
    dobj.degrees = torch.full((375,), 12.0)
    dobj.batch = torch.repeat_interleave(torch.arange(15), 25)
    warnings.filterwarnings('ignore')
    with torch.no_grad():
        y = model(dobj)
    mm = torch.pdist(y, p=2)
    print('SR isomorphism: d_id', d_id, 'failures', int((mm < 1e-2).sum()), 'of', mm.numel(), 'min dist', float(mm.min()))
    torch.save({'ctor': ctor, 'args': args, 'state_dict': model.state_dict(), 'y': y, 'identifiers': torch.from_numpy(ids)},
               os.path.join(out_dir, 'sr_isomorphism.pt'))


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'sr':
    make_sr_isomorphism(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden'))


def make_ogb(out_dir):
    """GNN_OGB (models_graph_classification_ogb_original.py) on a molhiv-shaped batch (README.md:121 recipe,
    reduced width/depth): eval forward, and a training step's loss + gradients (dropout 0)."""
    from oracle import ref_import
    warnings.filterwarnings('ignore')
    M = ref_import.models()
    g = torch.Generator().manual_seed(77)
    L, d, dh = 3, 32, 64
    out = {}
    # residual=True cannot be differentiated in the reference under torch 2.x (`x += x_interm[-1]` is an in-place
    # update of a ReLU output, models_graph_classification_ogb_original.py:247) -> eval-only golden for it
    for name, scope, vn, residual, do_train in (('ogb_local_vn', 'local', True, False, True),
                                               ('ogb_global', 'global', False, False, True),
                                               ('ogb_global_vn_res_eval', 'global', True, True, False)):
        args = dict(seed=0, model_name='GSN_edge_sparse_ogb', readout='mean', dropout_features=[0.0] * (L + 1),
                    bn=[True] * L, final_projection=[False] * L + [True], residual=residual, inject_ids=False,
                    inject_edge_features=True, vn=vn, vn_pooling='sum', input_vn_encoder='embedding',
                    d_out_vn_encoder=d, d_out_vn=[d] * (L - 1), id_scope=scope, d_msg=[d] * L, d_out=[d] * L,
                    d_h=[[dh]] * L, aggr='add', flow='source_to_target', msg_kind='ogb', train_eps=[False] * L,
                    activation_mlp='relu', bn_mlp=True, jk_mlp=False, degree_embedding='one_hot_encoder',
                    degree_as_tag=[False] * L, retain_features=[False] + [True] * (L - 1), multi_embedding_aggr='sum',
                    features_scope='full', input_node_encoder='atom_encoder', d_out_node_encoder=d,
                    edge_encoder='bond_encoder', d_out_edge_encoder=[d] * L, id_embedding='embedding',
                    d_out_id_embedding=d, d_out_degree_embedding=d, extend_dims=True, activation='relu')
        d_in_id = [4, 6, 3]
        ctor = dict(in_features=9, out_features=1, encoder_ids=None, d_in_id=d_in_id, in_edge_features=3,
                    d_in_node_encoder=None, d_in_edge_encoder=None, encoder_degrees=None, d_degree=None)
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            model = M['GNN_OGB'](**ctor, **args)
        randomize(model, g)
        if vn:      # the vn embedding is zero-initialised; make it non-trivial
            for p in model.vn_encoder.parameters():
                p.data.normal_(0, 0.3, generator=g)
        ei, batch, n = rand_graph(g, n_graphs=7, lo=3, hi=9, sort=False)
        E = ei.shape[1]
        rows = E if scope == 'local' else n
        atom, bond = [119, 4, 12, 12, 10, 6, 6, 2, 2], [5, 6, 2]
        data = {'edge_index': ei, 'batch': batch,
                'x': torch.stack([torch.randint(0, a, (n,), generator=g) for a in atom], 1),
                'edge_features': torch.stack([torch.randint(0, b, (E,), generator=g) for b in bond], 1),
                'identifiers': torch.stack([torch.randint(0, k, (rows,), generator=g) for k in d_in_id], 1),
                'degrees': torch.randint(0, 5, (n,), generator=g)}

        class Obj:
            pass
        dobj = Obj()
        for k, v in data.items():
            setattr(dobj, k, v)
        model.eval()
        with torch.no_grad():
            y_eval = model(dobj)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        out[name] = {'ctor': ctor, 'args': args, 'data': data, 'state_dict': sd, 'y_eval': y_eval}
        if do_train:
            model.train()
            target = torch.randint(0, 2, (7, 1), generator=g).float()
            y_train = model(dobj)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(y_train, target)
            loss.backward()
            grads = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
            out[name].update(y_train=y_train.detach(), target=target, loss=loss.detach(), grads=grads)
            print(name, 'eval', y_eval.flatten()[:3].tolist(), 'loss', float(loss), 'n grads', len(grads))
        else:
            print(name, 'eval', y_eval.flatten()[:3].tolist())
    torch.save(out, os.path.join(out_dir, 'mp_ogb.pt'))


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'ogb':
    make_ogb(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden'))
