"""``metis`` / ``permute`` (reference: torch_geometric_autoscale/metis.py:14-63).

The reference calls ``torch.ops.torch_sparse.partition`` (METIS k-way through torch_sparse, which is
not available here).  This module calls METIS itself: ``csrc/metis_shim.c`` links the
``libmetis_static.a`` that ships with the CUDA toolkit (``METIS_PartGraphKway`` /
``METIS_PartGraphRecursive``, default options, exactly the call torch_sparse makes) and the result is
turned into ``(perm, ptr)`` the way the reference does (``cluster.sort()``, ``ind2ptr``).
Partitioning is the step immediately *before* the hot path.  The synthetic graphs of ``synthetic.py``
are generated already clustered (contiguous equal blocks); for them ``metis`` returns the identity
permutation and the block boundaries unless ``force=True``.  If the METIS library is missing, a
deterministic locality ordering (reverse Cuthill-McKee, equal cuts) stands in."""
import copy
import ctypes
import os
import time
from typing import Tuple

import torch
from torch import Tensor

from .sparse import SparseTensor


def block_ptr(num_nodes: int, num_parts: int) -> Tensor:
    """Boundaries of ``num_parts`` near-equal contiguous node ranges."""
    base, rem = divmod(num_nodes, num_parts)
    sizes = torch.full((num_parts,), base, dtype=torch.int64)
    sizes[:rem] += 1
    ptr = torch.zeros(num_parts + 1, dtype=torch.int64)
    torch.cumsum(sizes, 0, out=ptr[1:])
    return ptr


def _bfs_order(rowptr: Tensor, col: Tensor) -> Tensor:
    import numpy as np
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    n = rowptr.numel() - 1
    m = csr_matrix((np.ones(col.numel(), dtype=np.int8), col.cpu().numpy(), rowptr.cpu().numpy()),
                   shape=(n, n))
    return torch.from_numpy(np.ascontiguousarray(reverse_cuthill_mckee(m, symmetric_mode=False))).long()


_METIS = None


def _metis_lib():
    global _METIS
    if _METIS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'libincagg_metis.so')
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.incagg_metis_partition.restype = ctypes.c_int
            lib.incagg_metis_partition.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
            _METIS = lib
        else:
            _METIS = False
    return _METIS


def metis_available() -> bool:
    return bool(_metis_lib())


def partition(rowptr: Tensor, col: Tensor, num_parts: int, recursive: bool = False):
    """torch.ops.torch_sparse.partition(rowptr, col, None, num_parts, recursive): cluster id per node
    (METIS k-way).  Returns (cluster, edgecut)."""
    lib = _metis_lib()
    if not lib:
        raise RuntimeError('libincagg_metis.so is not built (make -C incagg_gnn_b200/csrc metis)')
    rowptr = rowptr.cpu().to(torch.int64).contiguous()
    col = col.cpu().to(torch.int64).contiguous()
    n = rowptr.numel() - 1
    part = torch.empty(n, dtype=torch.int64)
    cut = ctypes.c_int64(-1)
    rc = lib.incagg_metis_partition(n, rowptr.data_ptr(), col.data_ptr(), int(num_parts), int(recursive),
                                    part.data_ptr(), ctypes.byref(cut))
    if rc != 1:
        raise RuntimeError(f'METIS failed with status {rc}')
    return part, int(cut.value)


def metis(adj_t: SparseTensor, num_parts: int, recursive: bool = False,
          log: bool = True, force: bool = False) -> Tuple[Tensor, Tensor]:
    r"""Returns the "clustered" permutation :obj:`perm` and the cluster slices :obj:`ptr`."""
    if log:
        t = time.perf_counter()
        print(f'Computing METIS partitioning with {num_parts} parts...', end=' ', flush=True)
    num_nodes = adj_t.size(0)
    if num_parts <= 1:
        perm, ptr = torch.arange(num_nodes), torch.tensor([0, num_nodes])
    elif getattr(adj_t, 'clustered_parts', None) == num_parts and not force:
        perm, ptr = torch.arange(num_nodes), block_ptr(num_nodes, num_parts)
    elif metis_available():
        rowptr, col, _ = adj_t.csr()
        cluster, _ = partition(rowptr, col, num_parts, recursive)
        cluster, perm = cluster.sort()                                    # metis.py:32
        ptr = torch.zeros(num_parts + 1, dtype=torch.int64)               # ind2ptr, metis.py:33
        torch.cumsum(torch.bincount(cluster, minlength=num_parts), 0, out=ptr[1:])
    else:
        rowptr, col, _ = adj_t.csr()
        perm = _bfs_order(rowptr, col)
        ptr = block_ptr(num_nodes, num_parts)
    if log:
        print(f'Done! [{time.perf_counter() - t:.2f}s]')
    return perm, ptr


def permute_adj(adj_t: SparseTensor, perm: Tensor) -> SparseTensor:
    """adj_t[perm][:, perm] (torch_sparse ``SparseTensor.permute``)."""
    dev = adj_t.device
    perm = perm.to(dev)
    n = adj_t.size(0)
    inv = torch.empty(n, dtype=torch.int64, device=dev)
    inv[perm] = torch.arange(n, device=dev)
    row = inv[adj_t.storage.row()]
    col = inv[adj_t.col.to(torch.int64)]
    key = row * adj_t.size(1) + col
    key, order = torch.sort(key, stable=True)
    value = adj_t.value[order] if adj_t.value is not None else None
    return SparseTensor(row=key // adj_t.size(1), col=key % adj_t.size(1), value=value,
                        sparse_sizes=adj_t.sparse_sizes(), is_sorted=True)


def permute(data, perm: Tensor, log: bool = True):
    r"""Permutes a :obj:`data` object according to a given permutation :obj:`perm`."""
    if log:
        t = time.perf_counter()
        print('Permuting data...', end=' ', flush=True)
    identity = bool((perm == torch.arange(perm.numel(), device=perm.device)).all())
    data = copy.copy(data)
    if not identity:
        for key, value in data:
            if isinstance(value, Tensor) and value.size(0) == data.num_nodes:
                data[key] = value[perm.to(value.device)]
            elif isinstance(value, Tensor) and value.size(0) == data.num_edges:
                raise NotImplementedError
            elif isinstance(value, SparseTensor):
                data[key] = permute_adj(value, perm)
    if log:
        print(f'Done! [{time.perf_counter() - t:.2f}s]')
    return data
