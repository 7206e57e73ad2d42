"""Shared helpers of the test-suite (synthetic graph generators, batching)."""
import numpy as np


def random_graph(rng, n, p, shuffle=True):
    a = np.triu(rng.random((n, n)) < p, 1)
    r, c = np.nonzero(a)
    ei = np.stack([np.concatenate([r, c]), np.concatenate([c, r])]).astype(np.int64)
    if shuffle and ei.shape[1]:
        ei = ei[:, rng.permutation(ei.shape[1])]
    return ei


def batch_graphs(graphs):
    """[(edge_index local, n)] -> node_ptr, edge_ptr, edge_index global (PyG collate)"""
    node_ptr, edge_ptr, eis = [0], [0], []
    for ei, n in graphs:
        eis.append(ei + node_ptr[-1])
        node_ptr.append(node_ptr[-1] + n)
        edge_ptr.append(edge_ptr[-1] + ei.shape[1])
    ei = np.concatenate(eis, 1) if eis else np.zeros((2, 0), np.int64)
    return np.array(node_ptr, np.int64), np.array(edge_ptr, np.int64), ei
