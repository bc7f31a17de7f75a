"""GCNII on the GAS / IncAgg runtime (reference: torch_geometric_autoscale/models/gcn2.py)."""
import os
from typing import Optional

import torch
from torch import Tensor
import torch.nn.functional as F
from torch.nn import ModuleList, BatchNorm1d

from ..nn import GCN2Conv, Linear, X0GradSink
from ..sparse import SparseTensor, spmm_delta
from .base import ScalableGNN, _PULL_PRIORITY
from ._masking import select_edges


class GCN2(ScalableGNN):
    _share_refresh_aggregate = True  # the layer aggregates its input directly: A @ x

    def __init__(self, num_nodes: int, in_channels, hidden_channels: int, out_channels: int,
                 num_layers: int, alpha: float, theta: float, shared_weights: bool = True,
                 dropout: float = 0.0, drop_input: bool = True, batch_norm: bool = False,
                 residual: bool = False, pool_size: Optional[int] = None,
                 buffer_size: Optional[int] = None, device=None):
        super().__init__(num_nodes, hidden_channels, num_layers, pool_size, buffer_size, device,
                         in_channels=in_channels)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.dropout = dropout
        self.drop_input = drop_input
        self.batch_norm = batch_norm
        self.residual = residual

        self.lins = ModuleList()
        self.lins.append(Linear(in_channels, hidden_channels))
        self.lins.append(Linear(hidden_channels, out_channels))
        self.convs = ModuleList()
        for i in range(num_layers):
            self.convs.append(GCN2Conv(hidden_channels, alpha=alpha, theta=theta, layer=i + 1,
                                       shared_weights=shared_weights, normalize=False))
        self.bns = ModuleList()
        for i in range(num_layers):
            self.bns.append(BatchNorm1d(hidden_channels))

    @property
    def reg_modules(self):
        return ModuleList(list(self.convs) + list(self.bns))

    @property
    def nonreg_modules(self):
        return self.lins

    def reset_parameters(self):
        super().reset_parameters()
        for lin in self.lins:
            lin.reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    @property
    def _fuse_relu(self) -> bool:
        return not self.batch_norm and not self.residual

    def _first_linear(self, x: Tensor):
        """x_0 = ReLU(lins[0] x) (gcn2.py:87) and, in training, the sink that collects the x_0 gradients
        of the layers in their GEMM epilogues (nn.X0GradSink; unshared weights only)."""
        sink = None
        if (torch.is_grad_enabled() and self.training and self.convs[0].weight2 is not None
                and self.lins[0].weight.requires_grad
                and os.environ.get('INCAGG_X0_SINK', '1') != '0'):   # (A/B switch)
            sink = X0GradSink()
        return self.lins[0](x, relu=True, x0_sink=sink), sink

    def _post(self, i: int, h: Tensor, x: Tensor) -> Tensor:
        if self.batch_norm:
            h = self.bns[i](h)
        if self.residual:
            h = h + x[:h.size(0)]
        return h.relu_()

    # GAS step (gcn2.py:78-185)
    def forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int = 2,
                aggregate_combined: bool = True, use_aggregation=True, *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        if self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        fuse = self._fuse_relu  # no batch norm / residual: ReLU rides in the GEMM epilogue
        # The halo pulls depend on nothing this step computes: issued before the first Linear, they run
        # beside it (87 K x 100 -> 128, alone on the GPU for 50 us) instead of beside the first two layers,
        # whose SpMM / GEMM they slowed by a fifth.
        ahead = None
        if use_aggregation and fuse:
            ahead = self.pull_ahead(self.histories[:self.num_layers - 1], x, batch_size, n_id,
                                    width=self.hidden_channels)
        x_0, sink = self._first_linear(x)
        x = F.dropout(x_0, p=self.dropout, training=self.training)
        t_all = 0
        x0b = x_0[:adj_t.size(0)]
        if use_aggregation:
            adj_t = select_edges(adj_t, batch_size, aggregate_combined)
            for i, (conv, hist) in enumerate(zip(self.convs[:-1], self.histories)):
                # rows >= B of x are constants (pulled history) after the first push_and_pull
                if ahead is not None:
                    # the layer GEMM writes rows [0, B) of the buffer whose tail the early pull fills
                    buf, pulled = ahead[i]
                    # the ReLU of layer i rides in its GEMM epilogue; its backward mask rides in the
                    # epilogue of the transposed SpMM of layer i + 1 (x passes only through dropout in between)
                    x = conv(x, x0b, adj_t, grad_rows=batch_size if i > 0 else None, relu=True, out_full=buf,
                             relu_input=i > 0, defer_relu_bwd=True, x0_sink=sink)
                    # the push only reads the rows the GEMM just wrote and nothing in this step reads
                    # the table rows it writes: it rides on the pull stream, joined after the loop
                    main, side = torch.cuda.current_stream(), self._pull_stream
                    if side is None:
                        side = self._pull_stream = torch.cuda.Stream(x.device, priority=_PULL_PRIORITY)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        hist.push(x[:batch_size].detach(), n_id[:batch_size], offset, count)
                    if pulled is not None:
                        main.wait_event(pulled)
                else:
                    h = conv(x, x0b, adj_t, grad_rows=batch_size if i > 0 else None, relu=fuse, x0_sink=sink)
                    x = h if fuse else self._post(i, h, x)
                    x, t = self.push_and_pull(hist, x, batch_size, n_id, offset, count)
                    t_all += t
                x = F.dropout(x, p=self.dropout, training=self.training)
            if ahead is not None:
                torch.cuda.current_stream().wait_stream(self._pull_stream)   # pushes done before the step ends
            h = self.convs[-1](x, x0b, adj_t,
                               grad_rows=batch_size if self.num_layers > 1 else None, relu=fuse,
                               relu_input=ahead is not None and self.num_layers > 1, x0_sink=sink,
                               defer_relu_bwd=fuse)   # lins[1] gates its input gradient (relu_input)
        else:  # no neighbour information (gcn2.py:151-181)
            x, x_0 = x[:batch_size], x_0[:batch_size]
            for i, conv in enumerate(self.convs[:-1]):
                h = conv.forward_no_neighbor(x, x_0, relu=fuse, x0_sink=sink)
                x = h if fuse else self._post(i, h, x)
                x = F.dropout(x, p=self.dropout, training=self.training)
            h = self.convs[-1].forward_no_neighbor(x, x_0, relu=fuse, x0_sink=sink, defer_relu_bwd=fuse)
        x = h if fuse else self._post(self.num_layers - 1, h, x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        return self.lins[1](x, relu_input=fuse), t_all

    # IncAgg step (gcn2.py:187-323)
    def VR_forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int, epoch: int, batch_idx: int,
                   *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        if self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        x_0, sink = self._first_linear(x)
        x = F.dropout(x_0, p=self.dropout, training=self.training)
        fuse = self._fuse_relu
        x0b = x_0[:adj_t.size(0)]
        for i, conv in enumerate(self.convs):
            if i == self.num_layers - 1:
                x = x[:batch_size]
            m_in, m_ag, gid = self._incagg_tables(i, batch_size, x.shape[1], n_id, offset, count)
            # the ReLU of layer i rides in its GEMM epilogue; its backward mask rides in the epilogue of the
            # next consumer (the transposed SpMM of layer i + 1, the classifier's input-gradient GEMM)
            h = spmm_delta(adj_t, x, m_in, m_ag, gid, relu_input=fuse and i > 0)  # A_BB (x - M_in) + M_ag, one kernel
            h = conv.forward_after_propagate(h, x0b, relu=fuse, x0_sink=sink, defer_relu_bwd=fuse)
            self._incagg_release()
            x = h if fuse else self._post(i, h, x)
            x = F.dropout(x, p=self.dropout, training=self.training)
        return self.lins[1](x, relu_input=fuse), 0, 0, 0

    # layer-wise sweep (gcn2.py:325-374); `agg` = precomputed A @ (layer input) in eval mode
    @torch.no_grad()
    def forward_layer(self, layer, x, adj_t, state, use_aggregation=True, agg=None):
        if not use_aggregation:
            x = x[:adj_t.size(0)]
        if layer == 0:
            if 'm_in0' in state:
                x = x_0 = state['m_in0']
            else:
                if self.drop_input:
                    x = F.dropout(x, p=self.dropout, training=self.training)
                x = x_0 = self.lins[0](x, relu=True)
            state['x_0'] = x_0[:adj_t.size(0)]
        x = F.dropout(x, p=self.dropout, training=self.training)
        conv = self.convs[layer]
        if not use_aggregation:
            h = conv.forward_no_neighbor(x, state['x_0'])
        elif agg is not None:
            h = conv.forward_after_propagate(agg, state['x_0'])
        else:
            h = conv(x, state['x_0'], adj_t)
        if self.batch_norm:
            h = self.bns[layer](h)
        if self.residual and h.size(-1) == x.size(-1):
            h = h + x[:h.size(0)]
        x = h.relu_()
        if layer == self.num_layers - 1:
            x = F.dropout(x, p=self.dropout, training=self.training)
            x = self.lins[1](x)
        return x

    def _refresh_layer0_input(self, x: Tensor) -> Tensor:
        return self.lins[0](x, relu=True)  # gcn2.py:452
