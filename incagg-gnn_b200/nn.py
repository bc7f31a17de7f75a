"""Conv layers the reference takes from torch_geometric (not a dependency here), restated on top of
the hand-written SpMM kernels.  Parameter names follow PyG so state_dicts line up:
``GCNConv.lin/.bias``, ``GCN2Conv.weight1/.weight2``, ``SAGEConv.lin_l/.lin_r``.

Semantics follow SURVEY.md §8a "Conv dense parts" (upstream PyG definitions) plus the two methods of
the reference's locally patched GCN2Conv:
``forward_after_propagate(h, x_0)`` = everything in ``GCN2Conv.forward`` after ``propagate``;
``forward_no_neighbor(x, x_0)`` = the same with ``h = x``.
"""
import math
from typing import Optional

import torch
from torch import Tensor
from torch.nn import Linear, Parameter

from .sparse import SparseTensor, spmm


def glorot_(w: Tensor) -> Tensor:
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        return w.uniform_(-a, a)


class GCNConv(torch.nn.Module):
    """out = A (x W) + b   (PyG GCNConv with normalize=False; reference gcn.py:63)."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False, bias: bool = True):
        super().__init__()
        if normalize:
            raise NotImplementedError('normalisation is applied once to the graph (main.py:151)')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = Linear(in_channels, out_channels, bias=False)
        self.bias = Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.lin.weight)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, adj_t: SparseTensor, grad_rows: Optional[int] = None) -> Tensor:
        x = self.lin(x)
        out = spmm(adj_t, x, reduce='sum', grad_rows=grad_rows)
        if self.bias is not None:
            out = out + self.bias
        return out


class GCN2Conv(torch.nn.Module):
    """GCNII layer (PyG GCN2Conv, normalize=False) with the reference's two extra entry points."""

    def __init__(self, channels: int, alpha: float, theta: float = None, layer: int = None,
                 shared_weights: bool = True, normalize: bool = False):
        super().__init__()
        if normalize:
            raise NotImplementedError('normalisation is applied once to the graph (main.py:151)')
        self.channels = channels
        self.alpha = alpha
        self.beta = 1.
        if theta is not None or layer is not None:
            assert theta is not None and layer is not None
            self.beta = math.log(theta / layer + 1)
        self.weight1 = Parameter(torch.empty(channels, channels))
        if shared_weights:
            self.register_parameter('weight2', None)
        else:
            self.weight2 = Parameter(torch.empty(channels, channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight1)
        if self.weight2 is not None:
            glorot_(self.weight2)

    def forward_after_propagate(self, h: Tensor, x_0: Tensor) -> Tensor:
        x = h * (1 - self.alpha)
        x_0 = self.alpha * x_0[:x.size(0)]
        if self.weight2 is None:
            out = x + x_0
            out = torch.addmm(out, out, self.weight1, beta=1. - self.beta, alpha=self.beta)
        else:
            out = torch.addmm(x, x, self.weight1, beta=1. - self.beta, alpha=self.beta)
            out = out + torch.addmm(x_0, x_0, self.weight2, beta=1. - self.beta, alpha=self.beta)
        return out

    def forward_no_neighbor(self, x: Tensor, x_0: Tensor) -> Tensor:
        return self.forward_after_propagate(x, x_0)

    def forward(self, x: Tensor, x_0: Tensor, adj_t: SparseTensor,
                grad_rows: Optional[int] = None) -> Tensor:
        h = spmm(adj_t, x, reduce='sum', grad_rows=grad_rows)
        return self.forward_after_propagate(h, x_0)


class SAGEConv(torch.nn.Module):
    """out = lin_l(mean_j x_j) + lin_r(x_root)   (PyG SAGEConv, aggr='mean', normalize=False)."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False,
                 root_weight: bool = True, bias: bool = True, aggr: str = 'mean'):
        super().__init__()
        if normalize:
            raise NotImplementedError
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggr = aggr
        self.normalize = normalize
        self.root_weight = root_weight
        self.project = False
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        self.lin_r = Linear(in_channels, out_channels, bias=False) if root_weight else None

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.lin_r is not None:
            self.lin_r.reset_parameters()

    def forward(self, x: Tensor, adj_t: SparseTensor, grad_rows: Optional[int] = None) -> Tensor:
        adj = adj_t.set_value(None) if adj_t.value is not None else adj_t
        out = spmm(adj, x, reduce=self.aggr, grad_rows=grad_rows)
        out = self.lin_l(out)
        if self.lin_r is not None:
            out = out + self.lin_r(x[:adj_t.size(0)])
        return out
