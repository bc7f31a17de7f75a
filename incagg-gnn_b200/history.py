"""``History`` — the historical-embedding table (reference: torch_geometric_autoscale/history.py:9-74).

Same constructor, attributes and methods.  What differs is where the bytes move:
  * the table lives either in HBM (``device='cuda'``; 180 GB per B200 holds all 2*L tables of every
    BASELINE config, e.g. 12.5 GB for GCNII/products) or in pinned host memory (``device=None``),
  * ``pull`` is the indexed-gather kernel (reading pinned memory through UVA when the table is on
    the host: no CPU index_select, no pageable copy),
  * ``push`` is the slice-copy / indexed-scatter kernel (or DMA copies towards pinned memory).
There is no CPU execution path: the module device must be CUDA for pull/push.
"""
from typing import Optional

import torch
from torch import Tensor

from . import ops


class History(torch.nn.Module):
    r"""A historical embedding storage module."""

    def __init__(self, num_embeddings: int, embedding_dim: int, device=None):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        pin_memory = device is None or str(device) == 'cpu'
        if pin_memory and not torch.cuda.is_available():
            pin_memory = False  # host-logic tests without a GPU; pull/push will refuse to run
        self.emb = torch.empty(num_embeddings, embedding_dim, device=device, pin_memory=pin_memory)
        self._device = torch.device('cpu')
        # first global row held by this table (> 0 for a rank's shard of a partitioned table)
        self.row_offset = 0
        self.reset_parameters()

    def reset_parameters(self):
        self.emb.fill_(0)

    def _apply(self, fn):
        # Set the `_device` of the module without transfering `self.emb` (history.py:28-31).
        self._device = fn(torch.zeros(1)).device
        return self

    def _compute_device(self) -> torch.device:
        if self.emb.is_cuda:
            return self.emb.device
        if self._device.type != 'cuda':
            raise RuntimeError('History.pull/push need a CUDA module device (no CPU fallback)')
        return self._device

    @torch.no_grad()
    def pull(self, n_id: Optional[Tensor] = None) -> Tensor:
        if n_id is None:
            return self.emb.to(device=self._device)
        dev = self._compute_device()
        idx = n_id.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        if self.row_offset:
            idx = idx - self.row_offset
        with torch.cuda.device(dev):
            out = ops.gather_rows(self.emb, idx)
        return out.to(device=self._device)

    @torch.no_grad()
    def push(self, x, n_id: Optional[Tensor] = None, offset: Optional[Tensor] = None,
             count: Optional[Tensor] = None):
        if n_id is None and x.size(0) != self.num_embeddings:
            raise ValueError
        elif n_id is None and x.size(0) == self.num_embeddings:
            self.emb.copy_(x)
        elif offset is None or count is None:
            dev = self._compute_device()
            idx = n_id.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
            if self.row_offset:
                idx = idx - self.row_offset
            with torch.cuda.device(dev):
                ops.scatter_rows(x.to(dev).contiguous(), idx, self.emb)
        else:  # push in chunks (history.py:60-65); n_id is ignored here as in the reference
            dev = self._compute_device()
            if self.row_offset:
                offset = offset - self.row_offset
            with torch.cuda.device(dev):
                ops.copy_slices(x.to(dev).contiguous(), self.emb, offset, count, 1)

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.num_embeddings}, '
                f'{self.embedding_dim}, emb_device={self.emb.device}, '
                f'device={self._device})')
