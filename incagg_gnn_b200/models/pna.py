"""PNA on the GAS runtime (reference: torch_geometric_autoscale/models/pna.py:24-158,281-295).

PNA has no working IncAgg path in the reference; its GAS ``forward`` and
``forward_layer`` are what is built here.  ``PNAConv.message_and_aggregate`` runs K =
|aggregators| x |scalers| separate  pre_lin -> relu -> matmul(reduce) -> post_lin -> scaler  passes in
the reference; here the K pre_lin outputs are written side by side into one ``[n, K*F]`` operand and
the K reductions run in ONE launch of the multi-aggregator SpMM kernel - in the inference sweeps and
in training (``sparse.spmm_multi``: the backward pass is one transposed SpMM over the sum / mean slabs
plus one scatter over the min / max slabs).  In training the K pre_lin GEMMs are one GEMM over the
row-concatenated weights, and the K post_lin GEMMs of one scaler one GEMM over the column-concatenated
weights (out = sum_k agg_k W_k^T = [agg_0 | ... | agg_K-1] [W_0 | ... | W_K-1]^T).
"""
from itertools import product
from typing import Optional, List

import torch
from torch import Tensor
import torch.nn.functional as F
from torch.nn import ModuleList, BatchNorm1d

from .. import ops
from ..nn import Linear, linear
from ..sparse import SparseTensor, spmm, spmm_multi
from .base import ScalableGNN

EPS = 1e-5


class PNAConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, aggregators: List[str],
                 scalers: List[str], deg: Tensor, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.aggregators = list(aggregators)
        self.scalers = list(scalers)
        deg = deg.to(torch.float)
        self.avg_deg = {'lin': deg.mean().item(), 'log': (deg + 1).log().mean().item()}
        K = len(self.aggregators) * len(self.scalers)
        self.pre_lins = ModuleList([Linear(in_channels, out_channels) for _ in range(K)])
        self.post_lins = ModuleList([Linear(out_channels, out_channels) for _ in range(K)])
        self.lin = Linear(in_channels, out_channels)
        self.reset_parameters()

    def reset_parameters(self):
        for lin in self.pre_lins:
            lin.reset_parameters()
        for lin in self.post_lins:
            lin.reset_parameters()
        self.lin.reset_parameters()

    def forward(self, x: Tensor, adj_t: SparseTensor) -> Tensor:
        out = self.message_and_aggregate(adj_t, x)
        out = out + self.lin(x)[:out.size(0)]
        return out

    def _scale(self, h: Tensor, scaler: str, deg: Tensor) -> Tensor:
        if scaler == 'amplification':
            h = h * ((deg + 1).log() / self.avg_deg['log'])
        elif scaler == 'attenuation':
            h = h * (self.avg_deg['log'] / ((deg + 1).log() + EPS))
        return h

    STD_EPS = 1e-5

    def _aggregate_one(self, adj_t: SparseTensor, h: Tensor, aggr: str) -> Tensor:
        """One aggregator with autograd.  'std' / 'var' are composed from two mean aggregations,
        sqrt(relu(mean(h^2) - mean(h)^2) + eps) (PyG's definition); torch_sparse.matmul itself has
        sum / mean / min / max only, so these two go beyond the reference."""
        if aggr in ('std', 'var'):
            m1 = spmm(adj_t, h, reduce='mean')
            m2 = spmm(adj_t, h * h, reduce='mean')
            var = (m2 - m1 * m1).relu()
            return var if aggr == 'var' else (var + self.STD_EPS).sqrt()
        return spmm(adj_t, h, reduce=aggr)

    def message_and_aggregate(self, adj_t: SparseTensor, x: Tensor) -> Tensor:
        deg = adj_t.storage.rowcount().to(x.dtype).view(-1, 1)
        combos = list(product(self.aggregators, self.scalers))
        # slabs of the fused multi-aggregator launch: one per combo, two for std / var
        slabs = []
        for k, (aggr, _) in enumerate(combos):
            slabs += [(k, 'mean', False), (k, 'mean', True)] if aggr in ('std', 'var') else [(k, aggr, False)]
        fused = not torch.is_grad_enabled() and len(slabs) <= 16 and self.out_channels % 4 == 0
        out = 0
        if fused:
            # inference sweeps: the K pre_lin outputs are written side by side by the GEMM epilogues
            # (bias + ReLU fused, strided output), all K reductions run in ONE SpMM launch
            Fo = self.out_channels
            hs = torch.empty((x.size(0), len(slabs) * Fo), dtype=x.dtype, device=x.device)
            for j, (k, _, squared) in enumerate(slabs):
                dst = hs[:, j * Fo:(j + 1) * Fo]
                if squared:
                    torch.mul(hs[:, (j - 1) * Fo:j * Fo], hs[:, (j - 1) * Fo:j * Fo], out=dst)
                else:
                    pl = self.pre_lins[k]
                    ops.gemm(x, pl.weight, trans_b=True, bias=pl.bias, relu=True, out=dst)
            agg = ops.spmm_multi_raw(adj_t.rowptr, adj_t.col, adj_t.value, hs, Fo,
                                     [r for _, r, _ in slabs], rows=adj_t.size(0), plan=adj_t.plan())
            j = 0
            for k, ((aggr, scaler), post_lin) in enumerate(zip(combos, self.post_lins)):
                a = agg[:, j * Fo:(j + 1) * Fo]
                if aggr in ('std', 'var'):
                    m2 = agg[:, (j + 1) * Fo:(j + 2) * Fo]
                    var = (m2 - a * a).relu()
                    a = var if aggr == 'var' else (var + self.STD_EPS).sqrt()
                    j += 1
                j += 1
                h = post_lin(a)
                out = out + self._scale(h, scaler, deg)
            return out
        if (all(a in ('sum', 'add', 'mean', 'min', 'max') for a in self.aggregators) and len(combos) <= 16
                and self.out_channels % 4 == 0 and x.is_cuda and x.dtype == torch.float32):
            return self._fused_training_path(adj_t, x, combos, deg)
        for (aggr, scaler), pre_lin, post_lin in zip(combos, self.pre_lins, self.post_lins):
            h = pre_lin(x, relu=True)
            h = self._aggregate_one(adj_t, h, aggr)
            h = post_lin(h)
            out = out + self._scale(h, scaler, deg)
        return out

    def _fused_training_path(self, adj_t: SparseTensor, x: Tensor, combos, deg: Tensor) -> Tensor:
        """sum / mean / min / max aggregators with autograd: 1 GEMM (all pre_lins, bias + ReLU in the
        epilogue) -> 1 multi-aggregator SpMM -> 1 GEMM per scaler (its post_lins), instead of K of each."""
        Fo, K = self.out_channels, len(combos)
        # slabs grouped by scaler, so that the slabs one post GEMM consumes are adjacent columns
        groups = [[k for k, (_, sc) in enumerate(combos) if sc == scaler] for scaler in self.scalers]
        order = [k for ks in groups for k in ks]
        w_pre = torch.cat([self.pre_lins[k].weight for k in order], 0)     # [K*Fo, in]
        b_pre = torch.cat([self.pre_lins[k].bias for k in order], 0)
        hs = linear(x, w_pre, b_pre, relu=True)                            # [n, K*Fo]
        agg = spmm_multi(adj_t, hs, Fo, [combos[k][0] for k in order])     # [rows, K*Fo]
        out = 0
        c0 = 0
        for scaler, ks in zip(self.scalers, groups):
            a_s = agg if len(ks) == K else agg[:, c0 * Fo:(c0 + len(ks)) * Fo]
            c0 += len(ks)
            w_post = torch.cat([self.post_lins[k].weight for k in ks], 1)  # [Fo, len(ks)*Fo]
            b_post = torch.stack([self.post_lins[k].bias for k in ks], 0).sum(0)
            h = linear(a_s, w_post, b_post)
            out = out + self._scale(h, scaler, deg)
        return out


class PNA(ScalableGNN):
    def __init__(self, num_nodes: int, in_channels: int, hidden_channels: int, out_channels: int,
                 num_layers: int, aggregators: List[str], scalers: List[str], deg: Tensor,
                 dropout: float = 0.0, drop_input: bool = True, batch_norm: bool = False,
                 residual: bool = False, pool_size: Optional[int] = None,
                 buffer_size: Optional[int] = None, device=None):
        super().__init__(num_nodes, hidden_channels, num_layers, pool_size, buffer_size, device)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.dropout = dropout
        self.drop_input = drop_input
        self.batch_norm = batch_norm
        self.residual = residual
        self.convs = ModuleList()
        for i in range(num_layers):
            in_dim = in_channels if i == 0 else hidden_channels
            out_dim = out_channels if i == num_layers - 1 else hidden_channels
            self.convs.append(PNAConv(in_dim, out_dim, aggregators=aggregators, scalers=scalers, deg=deg))
        self.bns = ModuleList()
        for i in range(num_layers - 1):
            self.bns.append(BatchNorm1d(hidden_channels))

    @property
    def reg_modules(self):
        return ModuleList(list(self.convs[:-1]) + list(self.bns))

    @property
    def nonreg_modules(self):
        return self.convs[-1:]

    def reset_parameters(self):
        super().reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()
        for bn in self.bns:
            bn.reset_parameters()

    # GAS step (pna.py:138-158); the leading arguments follow ScalableGNN.__call__ of this repo
    def forward(self, x: Tensor, adj_t: SparseTensor, drift_norm: int = 2,
                aggregate_combined: bool = True, use_aggregation=True, *args):
        batch_size, n_id, offset, count = (list(args) + [None] * 4)[:4]
        t_all = 0
        if self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        hists = list(self.histories)[:len(self.convs) - 1]
        ahead = self.pull_ahead(hists, x, batch_size, n_id)
        for i, (conv, bn, hist) in enumerate(zip(self.convs[:-1], self.bns, self.histories)):
            h = conv(x, adj_t)
            if self.batch_norm:
                h = bn(h)
            if self.residual and h.size(-1) == x.size(-1):
                h = h + x[:h.size(0)]
            x = h.relu_()
            x, t = self.push_and_pull(hist, x, batch_size, n_id, offset, count,
                                      ahead=ahead[i] if ahead else None)
            t_all += t
            x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.convs[-1](x, adj_t)
        return x, t_all

    def VR_forward(self, *args, **kwargs):
        raise NotImplementedError('PNA has no working incremental-aggregation path in the reference '
                                  '(pna.py:162-278 is a debugging mock, SURVEY F10)')

    # layer-wise sweep (pna.py:281-295); `use_aggregation` accepted for mini_inference's call shape
    @torch.no_grad()
    def forward_layer(self, layer, x, adj_t, state, use_aggregation=True, agg=None):
        if layer == 0 and self.drop_input:
            x = F.dropout(x, p=self.dropout, training=self.training)
        h = self.convs[layer](x, adj_t)
        if layer < self.num_layers - 1:
            if self.batch_norm:
                h = self.bns[layer](h)
            if self.residual and h.size(-1) == x.size(-1):
                h = h + x[:h.size(0)]
            h = h.relu_()
            h = F.dropout(h, p=self.dropout, training=self.training)
        return h
