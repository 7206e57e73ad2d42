from .GSN_sparse import GSN_sparse
from .GSN_edge_sparse import GSN_edge_sparse
from .GSN_edge_sparse_ogb import GSN_edge_sparse_ogb
from .MPNN_sparse import MPNN_sparse
from .MPNN_edge_sparse import MPNN_edge_sparse
from .MPNN_edge_sparse_ogb import MPNN_edge_sparse_ogb

__all__ = ['GSN_sparse', 'GSN_edge_sparse', 'GSN_edge_sparse_ogb', 'MPNN_sparse', 'MPNN_edge_sparse',
           'MPNN_edge_sparse_ogb']
