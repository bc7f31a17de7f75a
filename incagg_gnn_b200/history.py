"""``History`` — one table of historical embeddings, ``emb[num_embeddings, embedding_dim]`` fp32
(reference: torch_geometric_autoscale/history.py:9-74; same constructor, ``emb`` / ``_device``
attributes, ``pull`` / ``push`` / ``reset_parameters``).

Placement and data movement are what changed:
  * ``device='cuda'``: the table lives in HBM (a B200 holds all 2*L tables of every BASELINE config,
    12.5 GB for GCNII / products) - the default of this repo;
    ``device=None`` / ``'cpu'``: page-locked host memory, the reference's layout;
  * ``pull(n_id)`` is the indexed-row gather kernel; a host table is read by that kernel through UVA
    (no CPU index_select, no pageable staging copy);
  * ``push`` is the slice-copy kernel (``offset`` / ``count`` given: partitions are contiguous row
    ranges), the indexed scatter kernel (``n_id`` only) or a plain full-table copy;
  * ``row_offset`` > 0 marks a rank's shard of a table that is partitioned over GPUs: ids and offsets
    are global, the tensor holds rows ``[row_offset, row_offset + num_embeddings)``.
There is no CPU execution path: the owning module must have been moved to a CUDA device.
"""
from typing import Optional

import torch
from torch import Tensor

from . import ops


class History(torch.nn.Module):
    r"""A historical embedding storage module."""

    def __init__(self, num_embeddings: int, embedding_dim: int, device=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        on_host = device is None or torch.device(device).type == 'cpu'
        # page-locked when on the host (needs a CUDA runtime; plain memory keeps CPU-only unit tests alive)
        self.emb = torch.zeros(num_embeddings, embedding_dim, device=None if on_host else device,
                               pin_memory=on_host and torch.cuda.is_available())
        self._device = torch.device('cpu')   # device of the owning module, set by .to() / .cuda()
        self.row_offset = 0

    def reset_parameters(self):
        self.emb.zero_()

    def _apply(self, fn):
        # .to(device) moves the consumer side only; the table stays where it was created (history.py:28-31)
        self._device = fn(torch.zeros(1)).device
        return self

    def _kernel_device(self) -> torch.device:
        if self.emb.is_cuda:
            return self.emb.device
        if self._device.type == 'cuda':
            return self._device
        raise RuntimeError('History.pull/push need a CUDA module device (no CPU fallback)')

    def _local_ids(self, n_id: Tensor, dev: torch.device) -> Tensor:
        ids = n_id.to(device=dev, dtype=torch.int64, non_blocking=True)
        return (ids - self.row_offset if self.row_offset else ids).contiguous()

    @torch.no_grad()
    def pull(self, n_id: Optional[Tensor] = None) -> Tensor:
        """Rows ``n_id`` of the table (all rows if None) on the module's device."""
        if n_id is None:
            return self.emb.to(device=self._device)
        dev = self._kernel_device()
        with torch.cuda.device(dev):
            rows = ops.gather_rows(self.emb, self._local_ids(n_id, dev))
        return rows.to(device=self._device)

    @torch.no_grad()
    def push(self, x, n_id: Optional[Tensor] = None, offset: Optional[Tensor] = None,
             count: Optional[Tensor] = None):
        """Three forms, as in the reference: whole table (``n_id is None``), rows ``n_id``, or
        contiguous chunks ``emb[offset_i : offset_i + count_i] = x[...]`` (``n_id`` is then ignored)."""
        if n_id is None:
            if x.size(0) != self.num_embeddings:
                raise ValueError
            self.emb.copy_(x)
            return
        dev = self._kernel_device()
        src = x.to(dev).contiguous()
        with torch.cuda.device(dev):
            if offset is not None and count is not None:
                ops.copy_slices(src, self.emb, offset - self.row_offset if self.row_offset else offset, count, 1)
            else:
                ops.scatter_rows(src, self._local_ids(n_id, dev), self.emb)

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def extra_repr(self) -> str:
        return (f'{self.num_embeddings}, {self.embedding_dim}, emb_device={self.emb.device}, '
                f'device={self._device}')
