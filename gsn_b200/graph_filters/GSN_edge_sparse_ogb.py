"""GSN_edge_sparse_ogb: constructor-compatible with /root/reference/graph_filters/GSN_edge_sparse_ogb.py
(same positional order, keyword names and state_dict keys); the arithmetic is in
graph_filters/base.py + libgsn_b200.so."""
from .base import SparseFilter


class GSN_edge_sparse_ogb(SparseFilter):

    def __init__(self, d_in, d_ef, d_id, d_degree, degree_as_tag, retain_features, id_scope,
                 d_msg, d_up, d_h, seed, activation_name, bn, aggr='add', msg_kind='ogb', eps=0,
                 train_eps=False, flow='source_to_target', **kwargs):
        super().__init__()
        self._configure(d_in=d_in, d_degree=d_degree, degree_as_tag=degree_as_tag,
                        retain_features=retain_features, d_msg=d_msg, d_up=d_up, d_h=d_h, seed=seed,
                        activation_name=activation_name, bn=bn, aggr=aggr, msg_kind=msg_kind, eps=eps,
                        train_eps=train_eps, flow=flow, d_ef=d_ef, d_id=d_id, id_scope=id_scope, extras=kwargs)
