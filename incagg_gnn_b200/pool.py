"""``AsyncIOPool`` — pinned-host <-> device history staging (reference: pool.py:15-134, over
csrc/async.cpp).  Same constructor, same five methods, same FIFO semantics:

    async_pull(src, offset, count, index)   enqueue a pull; at most ``pool_size`` are in flight, the rest
                                            wait in the queue until ``free_pull`` releases a slot
    synchronize_pull() -> Tensor            device buffer of the OLDEST pull: rows [0, sum(count)) are the
                                            slices, the following |index| rows are src[index]
    free_pull()                             release the oldest pull's slot, start the next queued pull
    async_push(src, offset, count, dst)     dst[offset_i : +count_i] <- src rows, on a side stream
    synchronize_push(idx=None)              wait for one / all outstanding pushes

What differs from the reference is how ordering is enforced: every slot owns a pull stream, a push
stream and three CUDA events (filled / released / pushed), and all dependencies between the side
streams and the compute stream are event waits.  The reference's ``torch.cuda.synchronize(<Stream>)``
(a whole-device synchronisation in torch 2.x) and the trailing ``cudaStreamSynchronize`` of
``read_async`` / ``write_async`` do not exist here, and the indexed part of a pull is gathered by a
kernel straight out of the pinned table (UVA), so no pinned bounce buffer per slot is needed.
"""
from collections import deque
from typing import Callable, Deque, List, NamedTuple, Optional

import torch
from torch import Tensor

from . import ops


class _PullRequest(NamedTuple):
    slot: int
    src: Tensor
    offset: Optional[Tensor]
    count: Optional[Tensor]
    index: Tensor


class _Slot:
    """Resources of one pool slot, created lazily on the pool's device."""

    def __init__(self, rows: int, width: int):
        self.rows, self.width = rows, width
        self.buffer: Optional[Tensor] = None            # [rows, width] on the device
        self.pull_stream: Optional[torch.cuda.Stream] = None
        self.push_stream: Optional[torch.cuda.Stream] = None
        self.filled: Optional[torch.cuda.Event] = None    # the pull into `buffer` has completed
        self.released: Optional[torch.cuda.Event] = None  # the consumer has finished reading `buffer`
        self.pushed: Optional[torch.cuda.Event] = None    # the push issued from this slot has completed
        self.push_src: Optional[Tensor] = None            # keeps the pushed tensor alive (pool.py:107)

    def materialise(self, device: torch.device) -> None:
        if self.buffer is None:
            if device.type != 'cuda':
                raise RuntimeError('AsyncIOPool needs a CUDA device (move the model with .to(device))')
            self.buffer = torch.empty(self.rows, self.width, device=device)
            self.pull_stream = torch.cuda.Stream(device)
            self.push_stream = torch.cuda.Stream(device)


class AsyncIOPool(torch.nn.Module):
    def __init__(self, pool_size: int, buffer_size: int, embedding_dim: int):
        super().__init__()
        self.pool_size, self.buffer_size, self.embedding_dim = pool_size, buffer_size, embedding_dim
        self._device = torch.device('cpu')
        self._slots: List[_Slot] = [_Slot(buffer_size, embedding_dim) for _ in range(pool_size)]
        self._pull_queue: Deque[_PullRequest] = deque()
        self._next_pull = 0   # slot assigned to the next async_pull (round robin)
        self._next_push = 0

    def _apply(self, fn: Callable) -> 'AsyncIOPool':
        self._device = fn(torch.zeros(1)).device
        return self

    def _slot(self, i: int) -> _Slot:
        s = self._slots[i]
        s.materialise(self._device)
        return s

    def _compute_stream(self) -> torch.cuda.Stream:
        return torch.cuda.current_stream(self._device)

    # ---- pulls --------------------------------------------------------------------------------
    @torch.no_grad()
    def async_pull(self, src: Tensor, offset: Optional[Tensor], count: Optional[Tensor],
                   index: Tensor) -> None:
        req = _PullRequest(self._next_pull, src, offset, count, index)
        self._next_pull = (self._next_pull + 1) % self.pool_size
        self._pull_queue.append(req)
        if len(self._pull_queue) <= self.pool_size:  # a slot is free: start right away
            self._start_pull(req)

    def _start_pull(self, req: _PullRequest) -> None:
        slot = self._slot(req.slot)
        stream = slot.pull_stream
        if slot.released is not None:          # previous content of the buffer still being read?
            stream.wait_event(slot.released)
        for other in self._slots:              # the table may still be written by an outstanding push
            if other.pushed is not None:
                stream.wait_event(other.pushed)
        if req.index.is_cuda:                  # device-resident index produced on the compute stream
            stream.wait_stream(self._compute_stream())
        with torch.cuda.stream(stream):
            ops.read_async(req.src, req.offset, req.count, req.index, slot.buffer, None)
            ops._PENDING_READS.pop()           # this pool keeps its own completion events
            slot.filled = torch.cuda.Event()
            slot.filled.record(stream)

    @torch.no_grad()
    def synchronize_pull(self) -> Tensor:
        slot = self._slot(self._pull_queue[0].slot)
        # the COMPUTE STREAM waits for the copy; the host does not block
        self._compute_stream().wait_event(slot.filled)
        return slot.buffer

    @torch.no_grad()
    def free_pull(self) -> None:
        done = self._pull_queue.popleft()
        slot = self._slot(done.slot)
        slot.released = torch.cuda.Event()
        slot.released.record(self._compute_stream())
        if len(self._pull_queue) >= self.pool_size:
            # the request that has just moved into the in-flight window reuses the released slot
            self._start_pull(self._pull_queue[self.pool_size - 1])
        elif not self._pull_queue:
            self._next_pull = 0

    # ---- pushes -------------------------------------------------------------------------------
    @torch.no_grad()
    def async_push(self, src: Tensor, offset: Tensor, count: Tensor, dst: Tensor) -> None:
        i = self._next_push
        self._next_push = (self._next_push + 1) % self.pool_size
        self.synchronize_push(i)               # at most one outstanding push per slot
        slot = self._slot(i)
        slot.push_src = src.contiguous()
        slot.push_stream.wait_stream(self._compute_stream())  # src must have been produced
        with torch.cuda.stream(slot.push_stream):
            ops.write_async(slot.push_src, offset, count, dst)
            slot.pushed = torch.cuda.Event()
            slot.pushed.record(slot.push_stream)

    @torch.no_grad()
    def synchronize_push(self, idx: Optional[int] = None) -> None:
        if idx is None:
            for i in range(self.pool_size):
                self.synchronize_push(i)
            self._next_push = 0
            return
        slot = self._slots[idx]
        if slot.pushed is not None:
            slot.pushed.synchronize()          # this slot's push only, never the whole device
            slot.pushed = None
        slot.push_src = None

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def extra_repr(self) -> str:
        return (f'pool_size={self.pool_size}, buffer_size={self.buffer_size}, '
                f'embedding_dim={self.embedding_dim}, device={self._device}')
