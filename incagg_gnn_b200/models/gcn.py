"""GCN on the GAS / IncAgg runtime (reference: torch_geometric_autoscale/models/gcn.py)."""
from typing import Optional

import torch
from torch import Tensor
import torch.nn.functional as F
from torch.nn import ModuleList, BatchNorm1d

from ..nn import GCNConv, Linear
from ..sparse import SparseTensor, spmm_delta
from .base import ScalableGNN
from ._masking import select_edges


class GCN(ScalableGNN):
    def __init__(self, num_nodes: int, in_channels, hidden_channels: int, out_channels: int,
                 num_layers: int, dropout: float = 0.0, drop_input: bool = True,
                 batch_norm: bool = False, residual: bool = False, linear: bool = False,
                 pool_size: Optional[int] = None, buffer_size: Optional[int] = None, device=None):
        super().__init__(num_nodes, hidden_channels, num_layers, pool_size, buffer_size, device,
                         in_channels=in_channels)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.dropout = dropout
        self.drop_input = drop_input
        self.batch_norm = batch_norm
        self.residual = residual
        self.linear = linear

        self.lins = ModuleList()
        if linear:
            self.lins.append(Linear(in_channels, hidden_channels))
            self.lins.append(Linear(hidden_channels, out_channels))
        self.convs = ModuleList()
        for i in range(num_layers):
            in_dim = out_dim = hidden_channels
            if i == 0 and not linear:
                in_dim = in_channels
            if i == num_layers - 1 and not linear:
                out_dim = out_channels
            self.convs.append(self._make_conv(in_dim, out_dim))
        self.bns = ModuleList()
        for i in range(num_layers):
            self.bns.append(BatchNorm1d(hidden_channels))

    def _make_conv(self, in_dim, out_dim):
        return GCNConv(in_dim, out_dim, normalize=False)

    @property
    def reg_modules(self):
        if self.linear:
            return ModuleList(list(self.convs) + list(self.bns))
        return ModuleList(list(self.convs[:-1]) + list(self.bns))

    @property
    def nonreg_modules(self):
        return self.lins if self.linear else self.convs[-1:]

    def reset_parameters(self):
        super().reset_parameters()
        for lin in self.lins:
            lin.reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    def _gas_pull_histories(self):
        # layer-l output is pushed to / pulled from histories[l+1] (gcn.py:154)
        return list(self.histories)[1:self.num_layers]

    def _post(self, i, h, x):
        if self.batch_norm:
            h = self.bns[i](h)
        if self.residual and h.size(-1) == x.size(-1):
            h = h + x[:h.size(0)]
        return h.relu_()

    def _head(self, h, x):
        if not self.linear:
            return h
        h = self._post(self.num_layers - 1, h, x)
        h = F.dropout(h, p=self.dropout, training=self.training)
        return self.lins[1](h)

    # transform without aggregation ("degraded to MLP", gcn.py:168-189)
    def _conv_no_agg(self, conv, x):
        h = conv.lin(x)
        if conv.bias is not None:
            h = h + conv.bias
        return h

    # the dense part applied to an already aggregated input (IncAgg, gcn.py:241-244)
    def _conv_after_agg(self, conv, h, x_root):
        return self._conv_no_agg(conv, h)

    _delta_reduce = 'sum'

    def _grad_rows(self, i, batch_size):
        return None  # GCNConv transforms before it aggregates: halo rows feed dW at every layer

    # GAS step (gcn.py:97-205)
    def forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int = 2,
                aggregate_combined: bool = True, use_aggregation=True, *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        if self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        if self.linear:
            x = self.lins[0](x).relu_()
            x = F.dropout(x, p=self.dropout, training=self.training)
        t_all = 0
        if use_aggregation:
            adj_t = select_edges(adj_t, batch_size, aggregate_combined)
            ahead = self.pull_ahead(self._gas_pull_histories(), x, batch_size, n_id)
            for i, conv in enumerate(self.convs[:-1]):
                h = conv(x, adj_t, grad_rows=self._grad_rows(i, batch_size))
                x = self._post(i, h, x)
                x, t = self.push_and_pull(self.histories[i + 1], x, batch_size, n_id, offset, count,
                                          ahead=ahead[i] if ahead else None)
                t_all += t
                x = F.dropout(x, p=self.dropout, training=self.training)
            h = self.convs[-1](x, adj_t, grad_rows=self._grad_rows(self.num_layers - 1, batch_size))
        else:
            x = x[:batch_size]
            for i, (conv, hist) in enumerate(zip(self.convs[:-1], self.histories)):
                h = self._conv_no_agg(conv, x)
                x = self._post(i, h, x)
                x, t = self.push_and_pull(hist, x, batch_size, n_id[:batch_size], offset, count)
                t_all += t
                x = F.dropout(x, p=self.dropout, training=self.training)
            h = self._conv_no_agg(self.convs[-1], x)
        return self._head(h, x), t_all

    # IncAgg step (gcn.py:209-279)
    def VR_forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int, epoch: int, batch_idx: int,
                   *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        if self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        if self.linear:
            x = self.lins[0](x).relu_()
            x = F.dropout(x, p=self.dropout, training=self.training)
        adj = adj_t.set_value(None) if self._delta_reduce == 'mean' and adj_t.value is not None else adj_t
        h = None
        for i, conv in enumerate(self.convs):
            x = x[:batch_size]
            m_in, m_ag, gid = self._incagg_tables(i, batch_size, x.shape[1], n_id, offset, count)
            h = spmm_delta(adj, x, m_in, m_ag, gid, reduce=self._delta_reduce)
            h = self._conv_after_agg(conv, h, x)
            self._incagg_release()
            if i < self.num_layers - 1:
                x = self._post(i, h, x)
                x = F.dropout(x, p=self.dropout, training=self.training)
        if not self.linear:
            return h, 0, 0, 0
        return self._head(h, x), 0, 0, 0

    # layer-wise sweep (gcn.py:282-332)
    @torch.no_grad()
    def forward_layer(self, layer, x, adj_t, state, use_aggregation=True, agg=None):
        if layer == 0:
            if self.drop_input:
                x = F.dropout(x, p=self.dropout, training=self.training)
            if self.linear:
                x = self.lins[0](x).relu_()
                x = F.dropout(x, p=self.dropout, training=self.training)
        else:
            x = F.dropout(x, p=self.dropout, training=self.training)
        if not use_aggregation:
            h = self.convs[layer].lin(x)  # as written in the reference (no bias here, gcn.py:316)
        elif agg is not None:
            h = self._conv_after_agg(self.convs[layer], agg, x[:adj_t.size(0)])
        else:
            h = self.convs[layer](x, adj_t)
        if layer < self.num_layers - 1 or self.linear:
            if self.batch_norm:
                h = self.bns[layer](h)
            if self.residual and h.size(-1) == x.size(-1):
                h = h + x[:h.size(0)]
            h = h.relu_()
        if self.linear:
            h = F.dropout(h, p=self.dropout, training=self.training)
            h = self.lins[1](h)
        return h

    def _refresh_layer0_input(self, x: Tensor) -> Tensor:
        # gcn.py:353-361: M_in0 is the raw input of layer 0 (zero-padded to the table width on push)
        return x
