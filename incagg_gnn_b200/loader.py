"""``SubgraphLoader`` / ``EvalSubgraphLoader`` (reference: torch_geometric_autoscale/loader.py:95-285).

Same constructor arguments, same ``SubData`` tuples ``(data, batch_size, n_id, offset, count)``, same
batch order (the reference's DataLoader sampler classes are reused for the index stream, so a seeded
shuffle yields the same partition order).  The collate itself runs on the GPU:

  * relabel_one_hop / relabel_one_hop_within_batch are the bit-exact CUDA kernels (the reference has
    no CUDA relabel, csrc/relabel.cpp:15-20, and runs a single-threaded unordered_map in DataLoader
    worker processes),
  * every node-level tensor is gathered with the indexed-row kernel; the source may sit in HBM or in
    pinned host memory (read through UVA), so there is no CPU gather + pickling + H2D copy,
  * batch CSR structures use int32 indices.

``num_workers`` / ``persistent_workers`` are accepted and ignored (no worker processes are needed).
"""
import time
from typing import List, NamedTuple, Tuple

import torch
from torch import Tensor
from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler

from . import ops
from .data import Data
from .sparse import SparseTensor


class SubData(NamedTuple):
    data: Data
    batch_size: int
    n_id: Tensor  # The indices of mini-batched nodes
    offset: Tensor  # The offset of contiguous mini-batched nodes
    count: Tensor  # The number of contiguous mini-batched nodes

    def to(self, *args, **kwargs):
        return SubData(self.data.to(*args, **kwargs), self.batch_size,
                       self.n_id, self.offset, self.count)


def pack_node_records(fields):
    """Pack narrow per-node attributes ``[(name, tensor[N, ...])]`` (row size <= 8 bytes each: labels,
    masks) into one ``uint8 [N, R]`` record table, R a multiple of 16 bytes, every field at its
    natural alignment.  Returns ``(table, [(name, dtype, trailing shape, byte offset, bytes)])``."""
    n = fields[0][1].size(0)
    layout, off = [], 0
    for k, v in fields:
        w = v[0].numel() * v.element_size()
        off = (off + w - 1) // w * w
        layout.append((k, v.dtype, tuple(v.shape[1:]), off, w))
        off += w
    rec = (off + 15) // 16 * 16
    table = torch.zeros((n, rec), dtype=torch.uint8)
    for (k, dt, shp, o, w), (_, v) in zip(layout, fields):
        src = v.view(torch.uint8) if v.dtype == torch.bool else v
        table[:, o:o + w] = src.contiguous().view(n, -1).view(torch.uint8)
    return table, layout


def unpack_node_records(rec: Tensor, layout):
    """Inverse of :func:`pack_node_records` on a (gathered) record table: ``{name: tensor}``."""
    out, n = {}, rec.size(0)
    for k, dt, shp, o, w in layout:
        col = rec[:, o:o + w].contiguous()
        col = col.view(torch.uint8 if dt == torch.bool else dt).view((n,) + shp)
        out[k] = col.view(torch.bool) if dt == torch.bool else col
    return out


class SubgraphLoader:
    r"""A simple subgraph loader that, given a pre-partioned :obj:`data` object,
    generates subgraphs from mini-batches in :obj:`ptr` (including their 1-hop
    neighbors)."""

    def __init__(self, data: Data, ptr: Tensor, batch_size: int = 1, bipartite: bool = True,
                 log: bool = True, num_neighbors=-1, type='eval', IB=False, shuffle: bool = False,
                 num_workers: int = 0, persistent_workers: bool = False, device=None,
                 prefetch: bool = True, shard=None, halo_plans: bool = False, **kwargs):
        self.data = data
        self.ptr = ptr.cpu()
        self.bipartite = bipartite
        self.log = log
        self.num_neighbors = num_neighbors
        if num_neighbors is not None and num_neighbors >= 0:
            raise NotImplementedError('neighbour sampling is out of scope (the reference call site is '
                                      'broken, loader.py:235; num_neighbors=-1 is the identity)')
        self.shuffle = shuffle
        self.prefetch = prefetch
        self._collate_stream = None
        self.batch_size = batch_size
        self.shuffled_batch_id = []
        self.device = torch.device(device) if device is not None else data.adj_t.device
        if self.device.type != 'cuda':
            raise RuntimeError('SubgraphLoader collates on the GPU: pass device="cuda" (with data in '
                               'pinned host memory) or a CUDA data.adj_t; there is no CPU collate')

        self.num_parts = self.ptr.numel() - 1
        # multi-GPU: this rank iterates over the partitions it owns only (parallel.Shard); every rank
        # yields the same number of batches per epoch (wrapping around) so collectives stay matched
        self.shard = shard
        self.halo_plans = halo_plans  # NCCL transport: exchange the halo-row plan of every batch
        if shard is not None:
            self._parts = [p for p in range(self.num_parts) if shard.lo <= int(self.ptr[p]) < shard.hi]
            assert all(int(self.ptr[p + 1]) <= shard.hi for p in self._parts), \
                'partition blocks of the loader must nest in the rank shards'
        else:
            self._parts = list(range(self.num_parts))
        # global CSR for relabel: int64 rowptr, int32 col, fp32 values (device resident)
        # (device resident, or pinned host memory that the relabel kernels read through UVA)
        adj = data.adj_t
        self._rowptr64 = adj.rowptr.to(torch.int64)
        self._rowptr_host = self._rowptr64.cpu()
        self._col = adj.col
        self._val = adj.value
        if not adj.col.is_cuda:
            self._rowptr64 = self._rowptr64.pin_memory()
            self._col = self._col if self._col.is_pinned() else self._col.pin_memory()
            if self._val is not None and not self._val.is_pinned():
                self._val = self._val.pin_memory()
        self._ws = ops.RelabelWorkspace(adj.size(0), self.device)
        self._known_sizes = {}
        self._host_graph = not adj.col.is_cuda
        # host-resident node attributes narrower than 16 bytes per row (labels, masks) are packed
        # into one pinned record table: one PCIe gather per step instead of one per attribute (each
        # of them is bound by the latency of ~10^5 tiny reads, not by bytes)
        self._packed, self._packed_fields = None, []
        narrow = [(k, v) for k, v in data
                  if isinstance(v, Tensor) and v.dim() >= 1 and v.size(0) == data.num_nodes and not v.is_cuda
                  and v.is_pinned() and v[0].numel() * v.element_size() <= 8]
        if len(narrow) > 1:
            packed, self._packed_fields = pack_node_records(narrow)
            self._packed = packed.pin_memory()

        n_local = len(self._parts)
        sampler = RandomSampler(range(n_local)) if shuffle else SequentialSampler(range(n_local))
        self._batch_sampler = BatchSampler(sampler, batch_size, drop_last=False)
        self._steps = len(self._batch_sampler)
        if shard is not None and shard.world_size > 1:
            import torch.distributed as dist
            t = torch.tensor([self._steps], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=shard.group)
            self._steps = int(t)

        if type == 'train':
            self._collate = self.compute_subgraph_IB if IB else self.compute_subgraph
            self._cached = None
        else:
            self._collate = self.compute_subgraph
            self._cached = None
            if batch_size == 1:  # pre-process the subgraph generation (loader.py:153-170)
                if log:
                    t = time.perf_counter()
                    print('Pre-processing subgraphs...', end=' ', flush=True)
                # one entry per step; every rank runs the same number of steps (halo plans are
                # exchanged pairwise per step), a rank with fewer partitions wraps around
                self._cached = [self.compute_subgraph([self._parts[j % len(self._parts)]])
                                for j in range(self._steps)]
                if log:
                    torch.cuda.synchronize(self.device)
                    print(f'Done! [{time.perf_counter() - t:.2f}s]')

    # -- helpers -----------------------------------------------------------------------------
    def _batch_nodes(self, batch_ids: List[int]):
        ptr = self.ptr
        ranges = [(int(ptr[b]), int(ptr[b + 1])) for b in batch_ids]
        n_id = torch.cat([torch.arange(lo, hi, device=self.device) for lo, hi in ranges]) \
            if len(ranges) > 1 else torch.arange(ranges[0][0], ranges[0][1], device=self.device)
        batch_id = torch.tensor(batch_ids)
        offset = ptr[batch_id]
        count = ptr[batch_id + 1] - offset
        rp = self._rowptr_host
        nnz_b = sum(int(rp[hi]) - int(rp[lo]) for lo, hi in ranges)  # partitions are contiguous rows
        return n_id, offset, count, nnz_b

    @property
    def fixed_batches(self) -> bool:
        """True when every epoch draws its batches from one fixed set (single partitions, or groups
        formed in sequential order): the precondition for replaying a captured step per batch."""
        return self.batch_size == 1 or not self.shuffle

    def _graph_window(self, batch_ids):
        """Host-resident graph, one contiguous partition: bulk-copy its CSR rows (three DMA transfers)
        and let the relabel kernels run on the device copy instead of chasing rowptr -> col through
        PCIe.  Returns (rowptr, col, value, window) for ops.relabel_*."""
        if not self._host_graph or len(batch_ids) != 1:
            return self._rowptr64, self._col, self._val, None
        lo, hi = int(self.ptr[batch_ids[0]]), int(self.ptr[batch_ids[0] + 1])
        e0, e1 = int(self._rowptr_host[lo]), int(self._rowptr_host[hi])
        dev = self.device
        rp = torch.empty(hi - lo + 1, dtype=torch.int64, device=dev)
        rp.copy_(self._rowptr64[lo:hi + 1], non_blocking=True)
        col = torch.empty(e1 - e0, dtype=self._col.dtype, device=dev)
        col.copy_(self._col[e0:e1], non_blocking=True)
        val = None
        if self._val is not None:
            val = torch.empty(e1 - e0, dtype=self._val.dtype, device=dev)
            val.copy_(self._val[e0:e1], non_blocking=True)
        return rp, col, val, (lo, e0, self._rowptr64.numel() - 1)

    def _finish(self, rowptr, col, value, n_id, batch_size, offset, count) -> SubData:
        adj_t = SparseTensor(rowptr=rowptr, col=col, value=value,
                             sparse_sizes=(rowptr.numel() - 1, n_id.numel()), is_sorted=True)
        if self.shard is not None and self.shard.world_size > 1 and self.halo_plans:
            from .parallel import HaloPlan
            n_id.halo_plan = HaloPlan(n_id[batch_size:], self.shard)
        data = self.data.__class__(adj_t=adj_t)
        packed_keys = ()
        if self._packed is not None:
            rec = ops.gather_rows(self._packed, n_id)             # [n, record bytes] uint8
            packed_keys = {f[0] for f in self._packed_fields}
            for k, v in unpack_node_records(rec, self._packed_fields).items():
                data[k] = v
        for k, v in self.data:
            if k in packed_keys:
                continue
            if isinstance(v, Tensor) and v.size(0) == self.data.num_nodes:
                if not (v.is_cuda or v.is_pinned()):
                    raise RuntimeError(f'data.{k} must be a CUDA or pinned host tensor')
                if v.dtype == torch.bool:
                    data[k] = ops.gather_rows(v.view(torch.uint8), n_id).view(torch.bool)
                else:
                    data[k] = ops.gather_rows(v, n_id)
        return SubData(data, batch_size, n_id, offset, count)

    # -- collates (same names as the reference) ---------------------------------------------
    # The output sizes of a relabel are data dependent (halo count / kept edges) and are read back
    # once per distinct batch; the loader remembers them, so a batch that comes round again (every
    # epoch with batch_size = 1, every evaluation sweep) is collated without any host synchronisation.
    def compute_subgraph(self, batches) -> SubData:
        batch_ids = [b[0] if isinstance(b, tuple) else int(b) for b in batches]
        n_id, offset, count, nnz_b = self._batch_nodes(batch_ids)
        batch_size = n_id.numel()
        key = ('gas',) + tuple(batch_ids)
        with torch.cuda.device(self.device):
            g_rowptr, g_col, g_val, window = self._graph_window(batch_ids)
            rowptr, col, value, n_id = ops.relabel_one_hop(
                g_rowptr, g_col, g_val, n_id, self.bipartite, ws=self._ws,
                out_int32=True, nnz_b=nnz_b, known=self._known_sizes.get(key), window=window)
            self._remember(key, self._ws.last_count)
            return self._finish(rowptr, col, value, n_id, batch_size, offset, count)

    def compute_subgraph_IB(self, batches) -> SubData:
        batch_ids = [b[0] if isinstance(b, tuple) else int(b) for b in batches]
        n_id, offset, count, nnz_b = self._batch_nodes(batch_ids)
        batch_size = n_id.numel()
        key = ('ib',) + tuple(batch_ids)
        with torch.cuda.device(self.device):
            g_rowptr, g_col, g_val, window = self._graph_window(batch_ids)
            rowptr, col, value, n_id = ops.relabel_one_hop_within_batch(
                g_rowptr, g_col, g_val, n_id, self.bipartite, ws=self._ws,
                out_int32=True, nnz_b=nnz_b, known=self._known_sizes.get(key), window=window)
            self._remember(key, self._ws.last_count)
            return self._finish(rowptr, col, value, n_id, batch_size, offset, count)

    def _remember(self, key, size):
        if len(self._known_sizes) < 65536:
            self._known_sizes[key] = size

    # the reference's train collate without IncAgg is compute_subgraph_NS, which equals
    # compute_subgraph for num_neighbors=-1
    compute_subgraph_NS = compute_subgraph

    def __len__(self):
        return self._steps

    def _batches_of_epoch(self):
        """Lists of partition ids, one per step; a rank with fewer batches than the slowest rank
        wraps around."""
        groups = [[self._parts[i] for i in ids] for ids in self._batch_sampler]
        j = 0
        while len(groups) < self._steps:
            groups.append(groups[j])
            j += 1
        return groups

    # -- prefetching iterator ------------------------------------------------------------------
    # The collate of batch i+1 is issued on a side stream while batch i trains on the caller's
    # stream (the role the reference gives to DataLoader worker processes, main.py:158-160).  The one
    # host synchronisation of a collate (reading the halo count) then waits for the side stream only,
    # so the host keeps running ahead of the training kernels.
    def _collate_async(self, batch_ids):
        if self._collate_stream is None:
            self._collate_stream = torch.cuda.Stream(self.device)
        st = self._collate_stream
        with torch.cuda.stream(st):
            sub = self._collate(batch_ids)
            ev = torch.cuda.Event()
            ev.record(st)
        return sub, ev

    def _hand_over(self, sub: SubData, ev) -> SubData:
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        # tensors were allocated on the side stream: tell the caching allocator about their consumer
        for _, v in sub.data:
            if isinstance(v, Tensor) and v.is_cuda:
                v.record_stream(cur)
            elif isinstance(v, SparseTensor):
                for t in (v.rowptr, v.col, v.value):
                    if t is not None and t.is_cuda:
                        t.record_stream(cur)
        if sub.n_id.is_cuda:
            sub.n_id.record_stream(cur)
        plan = getattr(sub.n_id, 'halo_plan', None)
        if plan is not None:
            for t in plan.tensors():
                if t.is_cuda:
                    t.record_stream(cur)
        return sub

    def __iter__(self):
        self.shuffled_batch_id = []
        if self._cached is not None:  # pre-materialised (evaluation): fixed step order
            for sub in self._cached:
                yield sub
            return
        if not self.prefetch:
            for batch_ids in self._batches_of_epoch():
                if self.shuffle:
                    self.shuffled_batch_id.append(batch_ids)
                yield self._collate(batch_ids)
            return
        pending = None
        for batch_ids in self._batches_of_epoch():
            if self.shuffle:
                self.shuffled_batch_id.append(batch_ids)
            nxt = self._collate_async(batch_ids)
            if pending is not None:
                yield self._hand_over(*pending)
            pending = nxt
        if pending is not None:
            yield self._hand_over(*pending)

    def __repr__(self):
        return f'{self.__class__.__name__}()'


class EvalSubgraphLoader(SubgraphLoader):
    r"""Like :class:`SubgraphLoader`, but merges ``batch_size`` consecutive partitions into one
    evaluation batch, never shuffles and pre-materialises every subgraph (loader.py:266-284)."""

    def __init__(self, data: Data, ptr: Tensor, batch_size: int = 1, bipartite: bool = True,
                 log: bool = True, **kwargs):
        ptr = ptr.cpu()[::batch_size]
        if int(ptr[-1]) != data.num_nodes:
            ptr = torch.cat([ptr, torch.tensor([data.num_nodes])], dim=0)
        kwargs.pop('shuffle', None)
        kwargs.pop('num_workers', None)
        super().__init__(data=data, ptr=ptr, batch_size=1, bipartite=bipartite, log=log,
                         shuffle=False, num_workers=0, **kwargs)
