"""Per-step edge selection of the GAS forward (reference gcn.py:117-141, same in gcn2/appnp/sage).

The reference materialises row()/col(), three boolean masks and two SparseTensors every step; with
``aggregate_combined=True`` the mask is all-true (``m | ~m``) so the adjacency is unchanged, which is
what this returns without touching memory.  ``aggregate_combined=False`` keeps only edges whose
column is in the batch (rows are always < B for a bipartite batch adjacency) but keeps the
``[B, B+H]`` shape."""
import torch

from ..sparse import SparseTensor


def select_edges(adj_t: SparseTensor, batch_size, aggregate_combined: bool) -> SparseTensor:
    if aggregate_combined or batch_size is None:
        return adj_t
    row = adj_t.storage.row()
    col = adj_t.storage.col()
    mask = (row < batch_size) & (col < batch_size)
    val = adj_t.value[mask] if adj_t.value is not None else None
    return SparseTensor(row=row[mask], col=col[mask], value=val,
                        sparse_sizes=adj_t.sparse_sizes(), is_sorted=True)
