"""Regenerates tests/golden/*.npz from the reference checkout (run in the build
container only: /root/reference does not exist on the GPU box).

  python scripts/make_golden.py count      # COUNT fixtures (graph-tool output shipped by the reference)
  python scripts/make_golden.py mp         # MP fixtures (the reference's own layers run on CPU)

COUNT fixtures
  imdb_k5_edge_counts.npz  re-pack of datasets/social/IMDBBINARY/processed/local/complete_graph_5.pt
                           (graph-tool output: #K3/#K4/#K5 per directed edge, 1000 graphs)
  sr251256.npz             the 15 SR(25,12,5,6) graphs of datasets/SR_graphs/sr251256/sr251256.g6 as
                           to_undirected edge lists (utils_data_prep.py:197-212)
  graphlets.npz            datasets/all_simple_graphs/graph{3..6}c.g6 as edge lists in file order
                           (utils.py:16-33 reads them with nx.read_graph6(...).edges)
"""
import hashlib
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _stub_pyg_data():
    import torch  # noqa: F401
    tg = types.ModuleType('torch_geometric')
    tgd = types.ModuleType('torch_geometric.data')
    tgdd = types.ModuleType('torch_geometric.data.data')

    class Data:
        def __setstate__(self, s):
            self.__dict__.update(s)
    tgdd.Data = Data
    tgd.Data = Data
    tgd.data = tgdd
    tg.data = tgd
    sys.modules.update({'torch_geometric': tg, 'torch_geometric.data': tgd, 'torch_geometric.data.data': tgdd})


def make_count():
    import torch
    import networkx as nx
    _stub_pyg_data()
    path = os.path.join(REF, 'datasets/social/IMDBBINARY/processed/local/complete_graph_5.pt')
    sha = hashlib.sha256(open(path, 'rb').read()).hexdigest()
    graphs, num_classes, orbit_sizes = torch.load(path, weights_only=False)
    node_ptr, edge_ptr, ei, ids = [0], [0], [], []
    for g in graphs:
        node_ptr.append(node_ptr[-1] + int(g.x.shape[0]))
        edge_ptr.append(edge_ptr[-1] + int(g.edge_index.shape[1]))
        ei.append(g.edge_index.numpy())
        ids.append(g.identifiers.numpy())
    ei = np.concatenate(ei, 1)
    ids = np.concatenate(ids, 0)
    assert ei.max() < 256 and ids.max() < 65536
    np.savez_compressed(os.path.join(OUT, 'imdb_k5_edge_counts.npz'),
                        node_ptr=np.array(node_ptr, np.int32), edge_ptr=np.array(edge_ptr, np.int32),
                        edge_index=ei.astype(np.uint8),          # per-graph LOCAL vertex ids
                        identifiers=ids.astype(np.uint16),
                        orbit_partition_sizes=np.array(orbit_sizes, np.int32),
                        source_sha256=np.array(sha))
    print('imdb', len(graphs), 'graphs', ei.shape, ids.sum(0), 'sha', sha[:16])

    # SR(25,12,5,6): utils_data_prep.py:203-207
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), '..'))
    from oracle.count_vf2 import to_undirected
    dataset = nx.read_graph6(os.path.join(REF, 'datasets/SR_graphs/sr251256/sr251256.g6'))
    els = [to_undirected(np.array(list(d.edges())).T) for d in dataset]
    np.savez_compressed(os.path.join(OUT, 'sr251256.npz'),
                        edge_index=np.stack(els).astype(np.uint8),
                        num_nodes=np.array([d.number_of_nodes() for d in dataset], np.int32))
    print('sr', len(els), els[0].shape)

    # graphlets: utils.py:27-31
    out = {}
    for k in range(2, 7):
        gs = nx.read_graph6(os.path.join(REF, f'datasets/all_simple_graphs/graph{k}c.g6'))
        gs = gs if isinstance(gs, list) else [gs]
        ptr, flat = [0], []
        for g in gs:
            e = list(g.edges)
            flat += e
            ptr.append(ptr[-1] + len(e))
        out[f'k{k}_ptr'] = np.array(ptr, np.int32)
        out[f'k{k}_edges'] = np.array(flat, np.uint8).reshape(-1, 2)
        print('graphlets k', k, len(gs))
    np.savez_compressed(os.path.join(OUT, 'graphlets.npz'), **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else 'count'
    if what == 'count':
        make_count()
    elif what == 'mp':
        from make_golden_mp import make_mp
        make_mp(OUT)
