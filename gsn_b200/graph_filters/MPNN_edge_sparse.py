"""MPNN_edge_sparse: constructor-compatible with /root/reference/graph_filters/MPNN_edge_sparse.py
(same positional order, keyword names and state_dict keys); the arithmetic is in
graph_filters/base.py + libgsn_b200.so."""
from .base import SparseFilter


class MPNN_edge_sparse(SparseFilter):

    def __init__(self, d_in, d_ef, d_degree, degree_as_tag, retain_features,
                 d_msg, d_up, d_h, seed, activation_name, bn, aggr='add', msg_kind='general', eps=0,
                 train_eps=False, flow='source_to_target', **kwargs):
        super().__init__()
        self._configure(d_in=d_in, d_degree=d_degree, degree_as_tag=degree_as_tag,
                        retain_features=retain_features, d_msg=d_msg, d_up=d_up, d_h=d_h, seed=seed,
                        activation_name=activation_name, bn=bn, aggr=aggr, msg_kind=msg_kind, eps=eps,
                        train_eps=train_eps, flow=flow, d_ef=d_ef, extras=kwargs)
