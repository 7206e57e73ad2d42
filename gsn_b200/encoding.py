"""Data-set level encoding of identifiers / degrees (mirror of /root/reference/utils_encoding.py:8-69).

`encode(graphs, id_encoding, degree_encoding, **kwargs)` keeps the reference's call signature and return tuple
(main.py:147).  `one_hot_unique` ranks every column's values among the sorted distinct values of the WHOLE data set
(np.unique(return_inverse) per column in the reference, :37-59); here one torch.unique per column on whatever device
the identifiers live on (the GPU when they come straight from count_batch)."""
from __future__ import annotations

import sys

import torch


class one_hot_unique:
    """utils_encoding.py:37-59"""

    def __init__(self, tensor_list, **kwargs):
        cat = torch.cat(tensor_list, 0)
        self.d, self.corrs, self.uniques = [], {}, []
        for col in range(cat.shape[1]):
            uniques, inverse = torch.unique(cat[:, col], return_inverse=True)
            self.d.append(int(uniques.numel()))
            self.corrs[col] = inverse
            self.uniques.append(uniques)

    def fit(self, tensor_list):
        pointer, out = 0, []
        for t in tensor_list:
            n = t.shape[0]
            out.append(torch.stack([self.corrs[col][pointer:pointer + n] for col in range(t.shape[1])], 1).long())
            pointer += n
        return out


class one_hot_max:
    """utils_encoding.py:62-69"""

    def __init__(self, tensor_list, **kwargs):
        cat = torch.cat(tensor_list, 0)
        self.d = [int(cat[:, i].max() + 1) for i in range(cat.shape[1])]

    def fit(self, tensor_list):
        return tensor_list


def encode(graphs, id_encoding, degree_encoding=None, **kwargs):
    """utils_encoding.py:8-34"""
    encoder_ids, d_id = None, [1] * graphs[0].identifiers.shape[1]
    encoded_ids = encoded_degrees = None
    if id_encoding is not None:
        fn = getattr(sys.modules[__name__], id_encoding)
        ids = [g.identifiers for g in graphs]
        encoder_ids = fn(ids, **(kwargs.get('ids') or {}))
        encoded_ids = encoder_ids.fit(ids)
        d_id = encoder_ids.d
    encoder_degrees, d_degree = None, []
    if degree_encoding is not None:
        fn = getattr(sys.modules[__name__], degree_encoding)
        degrees = [g.degrees.unsqueeze(1) for g in graphs]
        encoder_degrees = fn(degrees, **(kwargs.get('degree') or {}))
        encoded_degrees = encoder_degrees.fit(degrees)
        d_degree = encoder_degrees.d
    for i, g in enumerate(graphs):
        if encoded_ids is not None:
            setattr(g, 'identifiers', encoded_ids[i])
        if encoded_degrees is not None:
            setattr(g, 'degrees', encoded_degrees[i])
    return graphs, encoder_ids, d_id, encoder_degrees, d_degree
