"""kernel-level breakdown of COUNT (cliques k<=5, edge scope) on the IMDB-BINARY fixture: python scripts/prof_imdb_count.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import networkx as nx  # noqa: E402
from gsn_b200 import counting, patterns  # noqa: E402

z = np.load(os.path.join(ROOT, 'tests', 'golden', 'imdb_k5_edge_counts.npz'))
node_ptr, edge_ptr = z['node_ptr'].astype(np.int64), z['edge_ptr'].astype(np.int64)
ei = z['edge_index'].astype(np.int64)
for g in range(len(node_ptr) - 1):
    ei[:, edge_ptr[g]:edge_ptr[g + 1]] += node_ptr[g]
dev = torch.device('cuda', 0)
sds = patterns.make_subgraph_dicts([list(nx.complete_graph(k).edges) for k in (3, 4, 5)], 'local')
ei_t, ptr_t = torch.from_numpy(ei).to(dev), torch.from_numpy(node_ptr)
mx = int(np.diff(node_ptr).max())
for _ in range(2):
    counting.count_batch(ei_t, ptr_t, sds, False, 'local', max_nodes_per_graph=mx)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        counting.count_batch(ei_t, ptr_t, sds, False, 'local', max_nodes_per_graph=mx, check=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=70))
