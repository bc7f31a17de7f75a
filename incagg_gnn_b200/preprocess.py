"""One-off graph preprocessing the reference applies before the hot path (main.py:147-151,
data.py:59): ``to_symmetric``, ``set_diag`` and ``gcn_norm(add_self_loops=False)``.

These produce the *inputs* of the propagation path (they run once per graph, not per batch) and are
plain tensor programs that run wherever the graph lives (the GPU for the BASELINE shapes)."""
from typing import Optional

import torch
from torch import Tensor

from .sparse import SparseTensor


def _coo(adj_t: SparseTensor):
    return adj_t.storage.row(), adj_t.col.to(torch.int64), adj_t.value


def _from_sorted_coo(row, col, value, sizes) -> SparseTensor:
    counts = torch.bincount(row, minlength=sizes[0])
    rowptr = torch.zeros(sizes[0] + 1, dtype=torch.int64, device=row.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return SparseTensor(rowptr=rowptr, col=col, value=value, sparse_sizes=sizes, is_sorted=True)


def coalesce(row: Tensor, col: Tensor, value: Optional[Tensor], sizes, reduce: str = 'sum') -> SparseTensor:
    """Sort by (row, col) and merge duplicates (values summed), as torch_sparse.coalesce does."""
    key = row * sizes[1] + col
    key, perm = torch.sort(key, stable=True)
    uniq, inv = torch.unique_consecutive(key, return_inverse=True)
    if value is not None:
        v = torch.zeros(uniq.numel(), dtype=value.dtype, device=value.device)
        v.index_add_(0, inv, value[perm])
        value = v
    return _from_sorted_coo(uniq // sizes[1], uniq % sizes[1], value, sizes)


def to_symmetric(adj_t: SparseTensor) -> SparseTensor:
    """A + A^T with duplicate entries merged (torch_sparse ``to_symmetric``; data.py:59)."""
    row, col, value = _coo(adj_t)
    N = max(adj_t.size(0), adj_t.size(1))
    r = torch.cat([row, col])
    c = torch.cat([col, row])
    v = torch.cat([value, value]) if value is not None else None
    return coalesce(r, c, v, (N, N))


def set_diag(adj_t: SparseTensor, values: Optional[Tensor] = None) -> SparseTensor:
    """Replace the diagonal by ones (torch_sparse ``set_diag``; main.py:148)."""
    row, col, value = _coo(adj_t)
    N = min(adj_t.size(0), adj_t.size(1))
    keep = row != col
    d = torch.arange(N, device=row.device)
    r = torch.cat([row[keep], d])
    c = torch.cat([col[keep], d])
    v = None
    if value is not None:
        dv = values if values is not None else torch.ones(N, dtype=value.dtype, device=value.device)
        v = torch.cat([value[keep], dv])
    key = r * adj_t.size(1) + c
    key, perm = torch.sort(key, stable=True)
    return _from_sorted_coo(key // adj_t.size(1), key % adj_t.size(1),
                            v[perm] if v is not None else None, adj_t.sparse_sizes())


def gcn_norm(adj_t: SparseTensor, add_self_loops: bool = False) -> SparseTensor:
    """value <- d_row^-1/2 * value * d_col^-1/2 with d = row sums (PyG ``gcn_norm`` on a
    SparseTensor; main.py:151 passes add_self_loops=False after set_diag)."""
    if add_self_loops:
        adj_t = set_diag(adj_t)
    row, col, value = _coo(adj_t)
    if value is None:
        value = torch.ones(col.numel(), dtype=torch.float32, device=col.device)
    deg = torch.zeros(adj_t.size(0), dtype=value.dtype, device=value.device)
    deg.index_add_(0, row, value)
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float('inf'), 0.)
    value = dis[row] * value * dis[col]
    return adj_t.set_value(value)
