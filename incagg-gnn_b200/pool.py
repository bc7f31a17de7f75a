"""``AsyncIOPool`` — pinned-host <-> device history staging (reference: pool.py:15-134).

Same slots / FIFO state machine / method names (§3.5 of SURVEY.md).  Differences:
  * dependencies are CUDA events between the pull / push side streams and the compute stream; the
    reference's ``torch.cuda.synchronize(<Stream>)`` (a whole-device sync in torch 2.x, F8) and the
    trailing ``cudaStreamSynchronize`` of read/write_async are gone,
  * the indexed part of a pull is gathered by a kernel straight out of the pinned table (UVA), so no
    pinned bounce buffer per slot is allocated (``_cpu_buffer`` is kept for API compatibility and
    allocates lazily only if somebody asks for it).
"""
from typing import Optional, Callable

import torch
from torch import Tensor
from torch.cuda import Stream

from . import ops


class AsyncIOPool(torch.nn.Module):
    def __init__(self, pool_size: int, buffer_size: int, embedding_dim: int):
        super().__init__()
        self.pool_size = pool_size
        self.buffer_size = buffer_size
        self.embedding_dim = embedding_dim

        self._device = torch.device('cpu')
        self._pull_queue = []
        self._push_cache = [None] * pool_size
        self._push_streams = [None] * pool_size
        self._pull_streams = [None] * pool_size
        self._cpu_buffers = [None] * pool_size
        self._cuda_buffers = [None] * pool_size
        self._pull_events = [None] * pool_size   # pull into slot finished
        self._free_events = [None] * pool_size   # consumer finished reading slot
        self._push_events = [None] * pool_size   # push from slot finished
        self._pull_index = -1
        self._push_index = -1

    def _apply(self, fn: Callable) -> None:
        self._device = fn(torch.zeros(1)).device
        return self

    def _pull_stream(self, idx: int) -> Stream:
        if self._pull_streams[idx] is None:
            assert str(self._device)[:4] == 'cuda'
            self._pull_streams[idx] = torch.cuda.Stream(self._device)
        return self._pull_streams[idx]

    def _push_stream(self, idx: int) -> Stream:
        if self._push_streams[idx] is None:
            assert str(self._device)[:4] == 'cuda'
            self._push_streams[idx] = torch.cuda.Stream(self._device)
        return self._push_streams[idx]

    def _cpu_buffer(self, idx: int) -> Tensor:
        if self._cpu_buffers[idx] is None:
            self._cpu_buffers[idx] = torch.empty(self.buffer_size, self.embedding_dim, pin_memory=True)
        return self._cpu_buffers[idx]

    def _cuda_buffer(self, idx: int) -> Tensor:
        if self._cuda_buffers[idx] is None:
            assert str(self._device)[:4] == 'cuda'
            self._cuda_buffers[idx] = torch.empty(self.buffer_size, self.embedding_dim,
                                                  device=self._device)
        return self._cuda_buffers[idx]

    @torch.no_grad()
    def async_pull(self, src: Tensor, offset: Optional[Tensor], count: Optional[Tensor],
                   index: Tensor) -> None:
        # Start pulling `src` at ([offset, count] and index positions (pool.py:64-74):
        self._pull_index = (self._pull_index + 1) % self.pool_size
        data = (self._pull_index, src, offset, count, index)
        self._pull_queue.append(data)
        if len(self._pull_queue) <= self.pool_size:
            self._async_pull(self._pull_index, src, offset, count, index)

    @torch.no_grad()
    def _async_pull(self, idx: int, src: Tensor, offset: Optional[Tensor], count: Optional[Tensor],
                    index: Tensor) -> None:
        stream = self._pull_stream(idx)
        # the slot may still be read by the consumer of its previous content, and the table may
        # still be written by an outstanding push:
        if self._free_events[idx] is not None:
            stream.wait_event(self._free_events[idx])
        for ev in self._push_events:
            if ev is not None:
                stream.wait_event(ev)
        if index.is_cuda:
            stream.wait_stream(torch.cuda.current_stream(self._device))
        with torch.cuda.stream(stream):
            ops.read_async(src, offset, count, index, self._cuda_buffer(idx), None)
            ops._PENDING_READS.pop()  # this pool tracks its own events
            ev = torch.cuda.Event()
            ev.record(stream)
            self._pull_events[idx] = ev

    @torch.no_grad()
    def synchronize_pull(self) -> Tensor:
        idx = self._pull_queue[0][0]
        # the compute stream waits for the copy; the host does not (reference: device-wide sync)
        torch.cuda.current_stream(self._device).wait_event(self._pull_events[idx])
        return self._cuda_buffer(idx)

    @torch.no_grad()
    def free_pull(self) -> None:
        # Free the buffer space and start pulling from remaining queue (pool.py:90-99):
        idx = self._pull_queue[0][0]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self._device))
        self._free_events[idx] = ev
        self._pull_queue.pop(0)
        if len(self._pull_queue) >= self.pool_size:
            data = self._pull_queue[self.pool_size - 1]
            idx, src, offset, count, index = data
            self._async_pull(idx, src, offset, count, index)
        elif len(self._pull_queue) == 0:
            self._pull_index = -1

    @torch.no_grad()
    def async_push(self, src: Tensor, offset: Tensor, count: Tensor, dst: Tensor) -> None:
        # Start pushing `src` to ([offset, count] and index positions to `dst` (pool.py:101-109):
        self._push_index = (self._push_index + 1) % self.pool_size
        self.synchronize_push(self._push_index)
        src = src.contiguous()
        self._push_cache[self._push_index] = src
        stream = self._push_stream(self._push_index)
        stream.wait_stream(torch.cuda.current_stream(self._device))  # src must be produced first
        with torch.cuda.stream(stream):
            ops.write_async(src, offset, count, dst)
            ev = torch.cuda.Event()
            ev.record(stream)
            self._push_events[self._push_index] = ev

    @torch.no_grad()
    def synchronize_push(self, idx: Optional[int] = None) -> None:
        # Synchronize the push command of stream `idx` or all commands (pool.py:111-123):
        if idx is None:
            for idx in range(self.pool_size):
                self.synchronize_push(idx)
            self._push_index = -1
        else:
            ev = self._push_events[idx]
            if ev is not None:
                ev.synchronize()  # this stream's work only, not the device
                self._push_events[idx] = None
            self._push_cache[idx] = None

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def __repr__(self):
        return (f'{self.__class__.__name__}(pool_size={self.pool_size}, '
                f'buffer_size={self.buffer_size}, '
                f'embedding_dim={self.embedding_dim}, '
                f'device={self._device})')
