"""APPNP on the GAS / IncAgg runtime (reference: torch_geometric_autoscale/models/appnp.py)."""
from typing import Optional

import torch
from torch import Tensor
import torch.nn.functional as F
from torch.nn import ModuleList

from ..nn import Linear
from ..sparse import SparseTensor, spmm, spmm_delta
from .base import ScalableGNN
from ._masking import select_edges


class APPNP(ScalableGNN):
    _share_refresh_aggregate = True

    def __init__(self, num_nodes: int, in_channels, hidden_channels: int, out_channels: int,
                 num_layers: int, alpha: float, dropout: float = 0.0,
                 pool_size: Optional[int] = None, buffer_size: Optional[int] = None, device=None):
        # histories are out_channels wide (appnp.py:24)
        super().__init__(num_nodes, out_channels, num_layers, pool_size, buffer_size, device,
                         in_channels=in_channels)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.alpha = alpha
        self.dropout = dropout
        self.lins = ModuleList()
        self.lins.append(Linear(in_channels, hidden_channels))
        self.lins.append(Linear(hidden_channels, out_channels))
        self.reg_modules = self.lins[:1]
        self.nonreg_modules = self.lins[1:]

    def reset_parameters(self):
        super().reset_parameters()
        for lin in self.lins:
            lin.reset_parameters()

    def _gas_pull_histories(self):
        return list(self.histories)  # forward() consumes one pull per history (appnp.py:84-88)

    def _mlp(self, x):
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.lins[0](x, relu=True)
        x = F.dropout(x, p=self.dropout, training=self.training)
        return self.lins[1](x)

    # GAS step (appnp.py:44-106): L+1 propagations over L histories, as written in the fork
    def forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int = 2,
                aggregate_combined: bool = True, use_aggregation=True, *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        t_all = 0
        if use_aggregation:
            adj_t = select_edges(adj_t, batch_size, aggregate_combined)
            x = self._mlp(x)
            x_0 = x[:adj_t.size(0)]
            ahead = self.pull_ahead(self.histories, x, batch_size, n_id, width=x.size(1))
            for i, history in enumerate(self.histories):
                x = (1 - self.alpha) * spmm(adj_t, x, grad_rows=batch_size if i > 0 else None) \
                    + self.alpha * x_0
                x, t = self.push_and_pull(history, x, batch_size, n_id, offset, count,
                                          ahead=ahead[i] if ahead else None)
                t_all += t
            x = (1 - self.alpha) * spmm(adj_t, x, grad_rows=batch_size) + self.alpha * x_0
        else:
            x = self._mlp(x[:batch_size])
            x_0 = x[:adj_t.size(0)]
            for history in self.histories:
                x = (1 - self.alpha) * x + self.alpha * x_0
                x, t = self.push_and_pull(history, x, batch_size, n_id[:batch_size], offset, count)
                t_all += t
            x = (1 - self.alpha) * x + self.alpha * x_0
        return x, t_all

    # IncAgg step (appnp.py:108-137): L propagations
    def VR_forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int, epoch: int, batch_idx: int,
                   *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        x = self._mlp(x[:batch_size])
        x_0 = x[:adj_t.size(0)]
        for i in range(self.num_layers):
            m_in, m_ag, gid = self._incagg_tables(i, batch_size, x.shape[1], n_id, offset, count)
            x_vr = spmm_delta(adj_t, x, m_in, m_ag, gid)
            x = (1 - self.alpha) * x_vr + self.alpha * x_0
            self._incagg_release()
        return x, 0, 0, 0

    # layer-wise sweep (appnp.py:140-166)
    @torch.no_grad()
    def forward_layer(self, layer, x, adj_t, state, use_aggregation=True, agg=None):
        if not use_aggregation:
            x = x[:adj_t.size(0)]
        if layer == 0:
            x = x_0 = state['m_in0'] if 'm_in0' in state else self._mlp(x)
            state['x_0'] = x_0[:adj_t.size(0)]
        if not use_aggregation:
            return (1 - self.alpha) * x + self.alpha * state['x_0']
        ax = agg if agg is not None else adj_t @ x
        return (1 - self.alpha) * ax + self.alpha * state['x_0']

    def _refresh_layer0_input(self, x: Tensor) -> Tensor:
        return self.lins[1](self.lins[0](x, relu=True))  # appnp.py:249-251
