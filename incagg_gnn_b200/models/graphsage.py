"""GraphSAGE on the GAS / IncAgg runtime (reference: torch_geometric_autoscale/models/graphsage.py).

Same skeleton as GCN with SAGEConv (mean aggregation, edge values stripped, root weight on the
in-batch rows).  As written in the reference (graphsage.py:634 vs :898), the IncAgg delta term is a
mean over the IN-BATCH degree while M_ag is a mean over the FULL degree."""
import torch
from torch import Tensor

from ..nn import SAGEConv
from ..sparse import SparseTensor, spmm
from .gcn import GCN


class GraphSAGE(GCN):
    _delta_reduce = 'mean'

    @property
    def _share_refresh_aggregate(self):
        # SAGEConv aggregates its input directly; with linear=True layer 0 aggregates lins[0](x)
        # while M_ag0 aggregates the raw x (graphsage.py:886-898), so the two differ
        return not self.linear

    def _make_conv(self, in_dim, out_dim):
        return SAGEConv(in_dim, out_dim, normalize=False)

    def _grad_rows(self, i, batch_size):
        # SAGEConv aggregates x itself; after the first push_and_pull rows >= B are constants
        return batch_size if i > 0 else None

    def _conv_no_agg(self, conv, x):
        raise NotImplementedError('use_aggregation=False calls conv.lin, which SAGEConv does not '
                                  'have in the reference either (graphsage.py:216)')

    def _conv_after_agg(self, conv, h, x_root):
        out = conv.lin_l(h)
        if conv.lin_r is not None:
            out = out + conv.lin_r(x_root[:h.size(0)])
        return out

    def _refresh_aggregate(self, adj_t: SparseTensor, x: Tensor) -> Tensor:
        adj = adj_t.set_value(None) if adj_t.value is not None else adj_t
        return spmm(adj, x, reduce='mean')  # graphsage.py:897-898
