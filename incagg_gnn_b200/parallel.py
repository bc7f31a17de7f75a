"""Multi-GPU sharding of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink /
NVSwitch), METIS partitions sharded over ranks (the reference itself is single-GPU).

* rank ``r`` owns a contiguous block of partitions = a contiguous node range ``[lo, hi)``, and with it
  those rows of every history table (``histories[l]``, ``histories_ag[l]``) as an HBM-resident shard;
* a rank trains only on batches of its own partitions, so every push is local;
* halo rows owned by other ranks are fetched by an all-to-all-v: the row ids a batch needs from each
  owner are exchanged once per batch (``HaloPlan``, known at collate time), then each layer's pull is
  one exchange of ``[H_peer, D]`` fp32 rows — the owner gathers them out of its shard with the
  indexed-row kernel, ``ncclSend/ncclRecv`` (``batch_isend_irecv``: one NCCL group = an all-to-all-v),
  and the requester scatters them into the tail of its layer input;
* gradients are averaged with one ``all_reduce`` per step over a flat buffer.

The IncAgg training step needs no halo traffic (``A_BB`` and the rank's own ``M_in``/``M_ag`` slices);
only the per-epoch refresh sweep exchanges halos.

The row gather / scatter used on each side are injectable so that the protocol itself is testable
with the ``gloo`` backend on CPU tensors (tests/test_parallel.py); the product path uses the CUDA
kernels and refuses CPU tensors.
"""
from typing import Callable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


class Shard:
    """Ownership map: partitions -> ranks -> node ranges."""

    def __init__(self, ptr: Tensor, rank: int, world_size: int, group=None):
        self.rank, self.world_size, self.group = int(rank), int(world_size), group
        ptr = ptr.cpu().to(torch.int64)
        P = ptr.numel() - 1
        if world_size > P:
            raise ValueError(f'{world_size} ranks for {P} partitions')
        # contiguous blocks of partitions, sizes differing by at most one
        self.part_bounds = [(r * P) // world_size for r in range(world_size + 1)]
        self.node_bounds = ptr[torch.tensor(self.part_bounds)].contiguous()  # [W+1] on the host
        self.part_lo, self.part_hi = self.part_bounds[rank], self.part_bounds[rank + 1]
        self.lo, self.hi = int(self.node_bounds[rank]), int(self.node_bounds[rank + 1])
        self.num_local = self.hi - self.lo
        self._bounds_dev = {}

    @property
    def parts(self) -> range:
        return range(self.part_lo, self.part_hi)

    def bounds_on(self, device) -> Tensor:
        key = str(device)
        if key not in self._bounds_dev:
            self._bounds_dev[key] = self.node_bounds.to(device)
        return self._bounds_dev[key]

    def owner_of(self, ids: Tensor) -> Tensor:
        """Rank that owns each global node id."""
        b = self.bounds_on(ids.device)
        return torch.bucketize(ids, b[1:], right=True)

    def steps_per_epoch(self, batch_size: int) -> int:
        """Steps every rank runs per epoch (ranks with fewer batches wrap around): collectives must
        be entered the same number of times on all ranks."""
        most = max(self.part_bounds[r + 1] - self.part_bounds[r] for r in range(self.world_size))
        return -(-most // batch_size)


class HaloPlan:
    """Who serves which halo rows of one batch.  Built once per batch (one small id exchange),
    reused by every layer's pull."""

    def __init__(self, halo_ids: Tensor, shard: Shard):
        self.shard = shard
        W, dev = shard.world_size, halo_ids.device
        self.n_halo = halo_ids.numel()
        owner = shard.owner_of(halo_ids)
        order = torch.argsort(owner, stable=True)              # halo positions grouped by owner
        self.order = order
        counts = torch.bincount(owner, minlength=W)
        self.req_counts = counts.tolist()                      # rows I request from each rank
        ids_sorted = halo_ids[order].contiguous()
        # counts of the requests the others make to me
        if W > 1:
            theirs = _exchange_counts(counts, shard)
        else:
            theirs = counts.clone()
        self.serve_counts = theirs.tolist()
        # ids the others request from me (global ids; I own all of them)
        req_splits = list(torch.split(ids_sorted, self.req_counts))
        self.local_ids = req_splits[shard.rank]                # my own halo rows: no exchange
        self.local_pos = torch.split(order, self.req_counts)[shard.rank]
        self.serve_ids: List[Tensor] = [torch.empty(0, dtype=torch.int64, device=dev) for _ in range(W)]
        ops_list, keep = [], []
        for r in range(W):
            if r == shard.rank:
                continue
            if self.serve_counts[r] > 0:
                buf = torch.empty(self.serve_counts[r], dtype=torch.int64, device=dev)
                self.serve_ids[r] = buf
                ops_list.append(dist.P2POp(dist.irecv, buf, _global_rank(r, shard), group=shard.group))
            if self.req_counts[r] > 0:
                keep.append(req_splits[r].contiguous())
                ops_list.append(dist.P2POp(dist.isend, keep[-1], _global_rank(r, shard), group=shard.group))
        if ops_list:
            for w in dist.batch_isend_irecv(ops_list):
                w.wait()
        self.recv_pos = torch.split(order, self.req_counts)    # where each rank's rows go in the halo block

    def tensors(self):
        """Device tensors of the plan (for allocator stream bookkeeping when it was built on a side
        stream)."""
        return [self.order, self.local_ids] + [t for t in self.serve_ids]


def _global_rank(r: int, shard: Shard) -> int:
    return r if shard.group is None else dist.get_global_rank(shard.group, r)


def _exchange_counts(counts: Tensor, shard: Shard) -> Tensor:
    """all-to-all of one int64 per peer (all_gather of the count vectors: W*W integers)."""
    W = shard.world_size
    gathered = [torch.empty_like(counts) for _ in range(W)]
    dist.all_gather(gathered, counts.contiguous(), group=shard.group)
    return torch.stack([g[shard.rank] for g in gathered])


def _default_gather(table: Tensor, idx: Tensor, out: Optional[Tensor] = None) -> Tensor:
    from . import ops
    return ops.gather_rows(table, idx, out=out)


def _default_scatter(src: Tensor, idx: Tensor, dst: Tensor) -> None:
    from . import ops
    ops.scatter_rows(src, idx, dst)


def pull_halo_rows(table_local: Tensor, plan: HaloPlan, out: Tensor, width: Optional[int] = None,
                   gather: Callable = _default_gather, scatter: Callable = _default_scatter) -> Tensor:
    """out[j] = table[halo_ids[j]] for the halo block of one batch, where `table_local` holds only the
    rows ``[shard.lo, shard.hi)`` of the global table.  One all-to-all-v of rows."""
    shard = plan.shard
    W, lo = shard.world_size, shard.lo
    D = table_local.size(1)
    # 1. the rows I own myself
    if plan.local_ids.numel() > 0:
        mine = gather(table_local, plan.local_ids - lo)
        scatter(mine, plan.local_pos, out)
    if W == 1:
        return out
    # 2. serve the others, receive mine
    ops_list, keep, recv = [], [], {}
    for r in range(W):
        if r == shard.rank:
            continue
        if plan.req_counts[r] > 0:
            recv[r] = torch.empty((plan.req_counts[r], D), dtype=table_local.dtype, device=out.device)
            ops_list.append(dist.P2POp(dist.irecv, recv[r], _global_rank(r, shard), group=shard.group))
        if plan.serve_counts[r] > 0:
            keep.append(gather(table_local, plan.serve_ids[r] - lo))
            ops_list.append(dist.P2POp(dist.isend, keep[-1], _global_rank(r, shard), group=shard.group))
    if ops_list:
        for w in dist.batch_isend_irecv(ops_list):
            w.wait()
    for r, buf in recv.items():
        scatter(buf, plan.recv_pos[r], out)
    return out


def open_peer_views(t: Tensor, shard: Shard):
    """Map every rank's copy of a (differently sized) table into this process: returns one tensor per
    rank (this rank's own tensor at its index).  The peers' storages are opened through CUDA IPC *in
    this rank's device context*, so kernels running on this GPU can load the peers' HBM directly over
    NVLink / NVSwitch.  Collective: every rank of the group must call it with its own table."""
    W = shard.world_size
    if W == 1:
        return [t]
    info = t.untyped_storage()._share_cuda_()
    meta = (tuple(info), tuple(t.shape), tuple(t.stride()), t.storage_offset(), t.dtype)
    gathered = [None] * W
    dist.all_gather_object(gathered, meta, group=shard.group)
    dev = t.device.index
    views = []
    for r in range(W):
        if r == shard.rank:
            views.append(t)
            continue
        inf, shape, stride, off, dtype = gathered[r]
        st = torch.UntypedStorage._new_shared_cuda(dev, *inf[1:])
        views.append(torch.empty(0, dtype=dtype, device=t.device).set_(st, off, shape, stride))
    return views


class GradAverager:
    """Gradient averaging over the ranks with one all_reduce per step.  All gradients live in ONE flat
    buffer (every ``p.grad`` is a view into it, the DDP bucket idea), so the step is
    ``zero() -> backward (accumulates in place) -> all_reduce(flat) -> flat /= W`` without any packing
    copies, and the buffer addresses are static (CUDA-graph friendly)."""

    def __init__(self, params, shard: Shard, flat: Optional[Tensor] = None):
        """`flat`: an existing flat gradient buffer that every ``p.grad`` already views (e.g.
        ``train.FlatAdam.flat_g``); otherwise one is created and the gradients re-pointed into it."""
        self.params = [p for p in params if p.requires_grad]
        self.shard = shard
        p0 = self.params[0]
        self.numel = sum(p.numel() for p in self.params)
        if flat is not None:
            self.flat = flat
            return
        self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        o = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[o:o + n].view_as(p)
            o += n

    def zero(self):
        """Instead of optimizer.zero_grad(): keeps the views in place."""
        self.flat.zero_()

    def all_reduce(self):
        if self.shard.world_size > 1:
            dist.all_reduce(self.flat, group=self.shard.group)

    def scale(self):
        if self.shard.world_size > 1:
            self.flat.div_(self.shard.world_size)

    def __call__(self):
        self.all_reduce()
        self.scale()


class FusedGradSync:
    """Gradient exchange of the data-parallel step WITHOUT a collective call: one kernel per rank
    (``incagg_allreduce_adam_step``) stages its gradient slice, signals the peers through NVLink peer
    memory, reads the peers' slices out of their HBM, adds them in rank order, and applies the Adam
    update - all-reduce, scaling and optimizer in a single launch that a CUDA graph can contain, so a
    training step at N GPUs is ONE graph replay (the NCCL path splits it into two graphs around a
    host-issued ``all_reduce``).  Needs ``train.FlatAdam`` (flat parameter / gradient / moment buffers)
    and no gradient clipping; every rank computes bit-identical parameters.

    Same surface as :class:`GradAverager` where the training loop touches it (``zero``), plus ``step``
    which replaces ``all_reduce`` + ``scale`` + ``optimizer.step``."""

    fused = True

    def __init__(self, optimizer, shard: Shard):
        from . import _lib
        if not hasattr(optimizer, 'flat_g'):
            raise RuntimeError('FusedGradSync needs train.FlatAdam (flat gradient / moment buffers)')
        if shard.world_size > _lib.lib.incagg_allreduce_adam_max_ranks():
            raise RuntimeError('too many ranks for the fused gradient exchange')
        self.opt, self.shard = optimizer, shard
        self.flat = optimizer.flat_g
        n = self.flat.numel()
        dev = self.flat.device
        self.stage = torch.zeros(2 * n, dtype=torch.float32, device=dev)
        self.signal = torch.zeros(_lib.lib.incagg_allreduce_adam_blocks() * _lib.lib.incagg_allreduce_adam_max_ranks(),
                                  dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        self.stage_views = open_peer_views(self.stage, shard)
        self.signal_views = open_peer_views(self.signal, shard)
        # the step counters of all ranks must agree (they are the epoch numbers of the signals)
        if shard.world_size > 1:
            t = optimizer.step_t.clone()
            dist.broadcast(t, _global_rank(0, shard), group=shard.group)
            if float(t) != float(optimizer.step_t):
                raise RuntimeError('FusedGradSync: the ranks have taken different numbers of optimizer steps')
            dist.barrier(group=shard.group)   # every signal array is zeroed and mapped before the first step

    def zero(self):
        self.flat.zero_()

    def step(self):
        from . import ops
        o = self.opt
        g0 = o.param_groups[0]
        wd_rest = o.param_groups[1]['weight_decay'] if len(o.param_groups) > 1 else g0['weight_decay']
        ops.allreduce_adam_step(self.stage_views, self.signal_views, self.shard.rank, o.flat_g, o.flat_p,
                                o.exp_avg, o.exp_avg_sq, o.n_first, g0['lr'], g0['betas'][0], g0['betas'][1],
                                g0['eps'], g0['weight_decay'], wd_rest, o.step_t, o._arrivals)

    # GradAverager surface, for loops written against it
    def all_reduce(self):
        raise RuntimeError('FusedGradSync exchanges gradients inside step()')

    def scale(self):
        pass

    def __call__(self):
        raise RuntimeError('FusedGradSync: call step() instead of averager() + optimizer.step()')


def make_grad_sync(model, optimizer, shard: Shard, grad_norm=None, transport: str = 'p2p'):
    """The gradient exchange of a sharded run: the fused peer-memory kernel when it applies (p2p
    transport, FlatAdam, no clipping), else one NCCL all_reduce per step."""
    import os
    if (shard.world_size > 1 and transport == 'p2p' and grad_norm is None and hasattr(optimizer, 'flat_g')
            and os.environ.get('INCAGG_FUSED_ALLREDUCE', '1') != '0'):
        return FusedGradSync(optimizer, shard)
    return GradAverager(model.parameters(), shard, flat=getattr(optimizer, 'flat_g', None))
