"""repro of a count_small failure under compute-sanitizer: python scripts/cs_repro.py family scope induced"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import count_c, count_vf2  # noqa: E402
from tests.util import batch_graphs, random_graph  # noqa: E402


def main():
    family, scope_name, induced = sys.argv[1], sys.argv[2], sys.argv[3] == '1'
    from gsn_b200 import counting, patterns
    rng = np.random.default_rng(1)
    els = count_vf2.pattern_edge_lists(family, int(sys.argv[4]))
    graphs = []
    for _ in range(20):
        n = int(rng.integers(1, 30))
        graphs.append((random_graph(rng, n, float(rng.uniform(0.05, 0.5))), n))
    graphs.append((random_graph(rng, 64, 0.08), 64))
    node_ptr, edge_ptr, ei = batch_graphs(graphs)
    scope = 1 if scope_name == 'local' else 0
    sds = patterns.make_subgraph_dicts(els, scope_name)
    exp = count_c.count_batch(node_ptr, edge_ptr, ei, count_vf2.make_subgraph_dicts(els, scope_name), induced, scope)
    got = counting.count_batch(torch.from_numpy(ei).cuda(), torch.from_numpy(node_ptr), sds, induced, scope_name).cpu().numpy()
    print('equal', np.array_equal(got, exp))


if __name__ == '__main__':
    main()
