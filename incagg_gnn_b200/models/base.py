"""``ScalableGNN`` — the GAS / IncAgg runtime (reference: torch_geometric_autoscale/models/base.py).

Same public surface: ``__call__`` (GAS step, base.py:126-240), ``VR_call`` (IncAgg step, :242-378),
``push_and_pull`` (:380-456), ``push_only`` (:458-499), ``mini_inference`` (:509-603),
``mini_inference_vr`` (per model in the reference, e.g. gcn2.py:432-507; shared here through two small
per-model hooks), ``histories`` / ``histories_ag`` / ``pool`` / ``pool_ag`` / ``_out``.

Two placements of the history tables:
  * ``device='cuda'`` (HBM-resident, the B200-native default): no pools; push / pull are kernels on
    the compute stream and the IncAgg step reads ``M_in`` / ``M_ag`` in place from the tables,
  * ``device=None`` (pinned host memory, the reference's layout): ``AsyncIOPool`` pairs exactly as in
    the reference, with event dependencies instead of device-wide synchronisation.

History-slot map follows the fork as written: GCN pushes/pulls layer-l output at
``histories[l+1]``, GCN2/APPNP/PNA at ``histories[l]``; the sweeps write layer-l output to
``histories[l+1]``; IncAgg reads ``histories[l]`` / ``histories_ag[l]``.
"""
import os
import time
import warnings
from typing import Optional, Callable, Dict, Any

import torch
from torch import Tensor

from .. import ops
from ..history import History
from ..pool import AsyncIOPool
from ..sparse import SparseTensor


_NO_AHEAD = os.environ.get('INCAGG_NO_AHEAD') == '1'   # A/B switch: issue the pulls in line


class _PushPull(torch.autograd.Function):
    """cat([x[:B], pulled]) without the intermediate: the pulled rows are gathered straight into
    the tail of the output buffer.  grad flows to x[:B] only (base.py:426,451)."""

    @staticmethod
    def forward(ctx, x: Tensor, fill_tail, batch_size: int, n_tail: int, out: Optional[Tensor] = None):
        if out is None:
            out = torch.empty((batch_size + n_tail, x.size(1)), dtype=x.dtype, device=x.device)
        else:  # a buffer whose tail an early pull (pull_ahead) is filling
            ctx.mark_dirty(out)
        out[:batch_size].copy_(x[:batch_size])
        if n_tail > 0:
            fill_tail(out[batch_size:])
        ctx.batch_size, ctx.n_in = batch_size, x.size(0)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B = ctx.batch_size
        if ctx.n_in == B:
            return grad_out[:B], None, None, None, None
        g = grad_out.new_zeros((ctx.n_in, grad_out.size(1)))
        g[:B] = grad_out[:B]
        return g, None, None, None, None


# Priority of the stream the halo pulls (and pushes) are issued on.  The pulls are issued before the first
# Linear, whose 680 CTAs the block scheduler dispatches to the end before it places a CTA of a younger
# kernel of the same priority.  With a high-priority pull stream (INCAGG_PULL_PRIORITY=-1; the priority is
# kept by the captured graph's kernel nodes) the gathers do run beside that GEMM, but it then takes 74
# instead of 52 us and the step 0.802 instead of 0.792 ms: default priority.
_PULL_PRIORITY = int(os.environ.get('INCAGG_PULL_PRIORITY', '0'))


class ScalableGNN(torch.nn.Module):
    r"""An abstract class for implementing scalable GNNs via historical embeddings."""

    def __init__(self, num_nodes: int, hidden_channels: int, num_layers: int,
                 pool_size: Optional[int] = None, buffer_size: Optional[int] = None, device=None,
                 in_channels=None):
        super().__init__()
        self.num_nodes = num_nodes
        self.hidden_channels = hidden_channels
        self.num_layers = num_layers
        self.pool_size = num_layers - 1 if pool_size is None else pool_size
        self.buffer_size = buffer_size

        # L histories (the fork; upstream PyGAS had L-1), all hidden-wide (base.py:67-72)
        self.histories = torch.nn.ModuleList(
            [History(num_nodes, hidden_channels, device) for _ in range(num_layers)])
        self.pool: Optional[AsyncIOPool] = None
        # M_ag tables (base.py:76-81)
        self.histories_ag = torch.nn.ModuleList(
            [History(num_nodes, hidden_channels, device) for _ in range(num_layers)])
        self.pool_ag: Optional[AsyncIOPool] = None

        self._async = False
        self._pull_stream = None
        self.__out: Optional[Tensor] = None
        self.shard = None      # parallel.Shard when the history tables are sharded over ranks
        self._row_lo = 0       # first global row of this rank's shard
        # Pinned-host tables WITHOUT the pool's slot / queue protocol: the gather kernel reads the halo
        # rows straight out of host memory (UVA) on the pull stream, pushes are DMA slice copies, all of
        # it enqueued like any other kernel - which is what lets a CUDA graph contain the whole step
        # (train.GraphedTrainer sets this for host-resident tables).  Same bytes, same order per step.
        self._direct_host = False

    def shard_histories(self, shard, transport: str = 'p2p') -> 'ScalableGNN':
        """Keep only this rank's rows ``[shard.lo, shard.hi)`` of every history table (HBM-resident
        shards; multi-GPU design of DESIGN.md §6).  Halo rows owned by other ranks are fetched
          * ``transport='p2p'``  by the gather kernel itself, loading the peers' shards over NVLink
            (tables mapped through CUDA IPC; no collective, no lockstep, capturable in CUDA graphs),
          * ``transport='nccl'`` by an all-to-all-v of row ids and rows (``parallel.pull_halo_rows``).
        Collective: every rank must call it."""
        if self.pool is not None or self.histories[0].emb.device.type != 'cuda':
            raise RuntimeError('sharded histories are HBM-resident (device="cuda"); the pinned-host '
                               'AsyncIOPool layout is single-GPU')
        assert transport in ('p2p', 'nccl')
        self.shard, self._row_lo, self.transport = shard, shard.lo, transport
        self._peer_views = {}
        for h in list(self.histories) + list(self.histories_ag):
            h.emb = torch.zeros(shard.num_local, h.embedding_dim, device=h.emb.device)
            h.num_embeddings = shard.num_local
            h.row_offset = shard.lo
            if transport == 'p2p' and shard.world_size > 1:
                from ..parallel import open_peer_views
                self._peer_views[h.emb.data_ptr()] = open_peer_views(h.emb, shard)
        self._shard_bounds = shard.node_bounds.tolist()
        self.__out = None
        return self

    def _pull_rows(self, table: Tensor, idx: Tensor, n_id: Tensor, dst: Tensor) -> None:
        """dst[j] = table_global[idx[j]] for halo ids `idx` (= n_id[B:]); `table` is this rank's
        shard (or the whole table on one GPU)."""
        if dst.size(0) == 0:
            return
        views = getattr(self, '_peer_views', {}).get(table.data_ptr()) if self.shard is not None else None
        if views is not None:      # NVLink peer loads inside the gather kernel
            ops.gather_rows_sharded(views, self._shard_bounds, idx.contiguous(), dst)
            return
        plan = getattr(n_id, 'halo_plan', None)
        if plan is not None:       # NCCL all-to-all-v
            from ..parallel import pull_halo_rows
            pull_halo_rows(table, plan, dst)
            return
        if self._row_lo:
            idx = idx - self._row_lo
        ops.gather_rows(table, idx.contiguous(), out=dst)

    @property
    def emb_device(self):
        return self.histories[0].emb.device

    @property
    def device(self):
        return self.histories[0]._device

    def _apply(self, fn: Callable) -> None:
        super()._apply(fn)
        # We only initialize the AsyncIOPool in case histories are on CPU (base.py:96-120):
        if (str(self.emb_device) == 'cpu' and str(self.device)[:4] == 'cuda'
                and self.pool_size is not None and self.buffer_size is not None):
            self.pool = AsyncIOPool(self.pool_size, self.buffer_size, self.histories[0].embedding_dim)
            self.pool.to(self.device)
            self.pool_ag = AsyncIOPool(self.pool_size, self.buffer_size,
                                       self.histories_ag[0].embedding_dim)
            self.pool_ag.to(self.device)
        return self

    def reset_parameters(self):
        for history in self.histories:
            history.reset_parameters()

    # ------------------------------------------------------------------------------------------
    # GAS step
    # ------------------------------------------------------------------------------------------
    def __call__(self, x: Optional[Tensor] = None, adj_t: Optional[SparseTensor] = None,
                 batch_size: Optional[int] = None, n_id: Optional[Tensor] = None,
                 offset: Optional[Tensor] = None, count: Optional[Tensor] = None, loader=None,
                 drift_norm: int = 2, aggregate_combined: bool = True, use_aggregation: bool = True,
                 **kwargs) -> Dict[str, Any]:
        if loader is not None:
            return self.mini_inference(loader, use_aggregation)

        self._async = (self.pool is not None and batch_size is not None and n_id is not None
                       and offset is not None and count is not None and not self._direct_host)
        if (batch_size is not None and not self._async and str(self.emb_device) == 'cpu'
                and str(self.device)[:4] == 'cuda' and not self._direct_host):
            warnings.warn('Asynchronous I/O disabled, although history and model sit on different devices.')

        if self._async:
            # one pull per history consumed by forward(); the fork issues L and leaks the extra
            # queue entries -- here exactly as many as push_and_pull will consume
            for hist in self._gas_pull_histories():
                self.pool.async_pull(hist.emb, None, None, n_id[batch_size:])

        out, t_mov = self.forward(x, adj_t, drift_norm, aggregate_combined, use_aggregation,
                                  batch_size, n_id, offset, count, **kwargs)

        if self._async:
            self.pool.synchronize_push()
        self._async = False
        return {'out': out, 'time_pool_pull': 0, 'time_pool_push': 0,
                'time_forward_histories_movement': t_mov, 'time_forward_total': 0}

    def _gas_pull_histories(self):
        """Histories pulled by one GAS forward, in consumption order (model specific)."""
        return list(self.histories)[:self.num_layers - 1]

    # ------------------------------------------------------------------------------------------
    # IncAgg step
    # ------------------------------------------------------------------------------------------
    def VR_call(self, x: Optional[Tensor] = None, adj_t: Optional[SparseTensor] = None,
                batch_size: Optional[int] = None, n_id: Optional[Tensor] = None,
                offset: Optional[Tensor] = None, count: Optional[Tensor] = None, loader=None,
                debug_flag: bool = False, drift_norm: int = 2, epoch: int = 0, batch_idx: int = 0,
                **kwargs) -> Dict[str, Any]:
        if loader is not None:
            return self.mini_inference(loader)
        self._async = (self.pool is not None and batch_size is not None and n_id is not None
                       and offset is not None and count is not None and not self._direct_host)
        if self._async:
            empty = torch.empty(0, dtype=torch.int64)
            for i in range(len(self.histories)):  # M_in / M_ag slices of the batch (base.py:318-323)
                self.pool.async_pull(self.histories[i].emb, offset, count, empty)
                self.pool_ag.async_pull(self.histories_ag[i].emb, offset, count, empty)
        out, t_mov, n_ib, n_ob = self.VR_forward(x, adj_t, drift_norm, epoch, batch_idx, batch_size,
                                                 n_id, offset, count, **kwargs)
        self._async = False
        return {'out': out, 'time_pool_pull': 0, 'time_pool_push': 0,
                'time_forward_histories_movement': t_mov, 'time_forward_total': 0,
                'num_in_batch_neighbors': n_ib, 'num_out_batch_neighbors': n_ob}

    def _incagg_tables(self, layer: int, batch_size: int, width: int, n_id: Tensor, offset, count):
        """(M_in, M_ag, n_id_or_None) for ``spmm_delta`` at `layer`.

        Pool mode (pinned-host tables): the pulled ``[:B, :F]`` slices as in gcn2.py:246-247 (no
        clone: the slot is released only after the kernel that reads it has been enqueued).
        HBM mode: the tables themselves; a single contiguous partition is a plain slice, several
        partitions are addressed through n_id inside the kernel."""
        if self._async:
            m_in = self.pool.synchronize_pull()[:batch_size, :width]
            m_ag = self.pool_ag.synchronize_pull()[:batch_size, :width]
            return m_in, m_ag, None
        hist, hist_ag = self.histories[layer].emb, self.histories_ag[layer].emb
        if not hist.is_cuda:
            if not self._direct_host:
                raise RuntimeError('IncAgg step with host-resident histories needs the AsyncIOPool '
                                   '(pool_size and buffer_size must be set)')
            # the batch's own rows of M_in / M_ag: partition slices, staged host -> device by DMA
            m_in = torch.empty((batch_size, hist.size(1)), dtype=hist.dtype, device=self.device)
            m_ag = torch.empty((batch_size, hist_ag.size(1)), dtype=hist_ag.dtype, device=self.device)
            ops.copy_slices(hist, m_in, offset, count, 0)
            ops.copy_slices(hist_ag, m_ag, offset, count, 0)
            return m_in[:, :width], m_ag[:, :width], None
        if offset is not None and offset.numel() == 1:
            o = int(offset[0]) - self._row_lo
            return hist[o:o + batch_size, :width], hist_ag[o:o + batch_size, :width], None
        gid = n_id[:batch_size]
        return hist, hist_ag, (gid - self._row_lo if self._row_lo else gid)

    def _incagg_release(self):
        if self._async:
            self.pool.free_pull()
            self.pool_ag.free_pull()

    # ------------------------------------------------------------------------------------------
    # push / pull
    # ------------------------------------------------------------------------------------------
    def push_and_pull(self, history, x: Tensor, batch_size: Optional[int] = None,
                      n_id: Optional[Tensor] = None, offset: Optional[Tensor] = None,
                      count: Optional[Tensor] = None, ahead=None):
        r"""Pushes and pulls information from :obj:`x` to :obj:`history` and vice versa.
        ``ahead`` = this history's (buffer, event) of :meth:`pull_ahead`: the pull is already under
        way on the side stream."""
        if n_id is None and x.size(0) != self.num_nodes:
            return x  # Do nothing...
        if n_id is None and x.size(0) == self.num_nodes:
            history.push(x)
            return x, 0.
        assert n_id is not None
        if batch_size is None:
            history.push(x, n_id)
            return x
        n_tail = n_id.numel() - batch_size
        if ahead is not None:
            buf, pulled = ahead
            history.push(x[:batch_size].detach(), n_id[:batch_size], offset, count)

            def fill(dst):
                if pulled is not None:   # (None: pulled by an earlier graph, already complete)
                    torch.cuda.current_stream().wait_event(pulled)
            return _PushPull.apply(x, fill, batch_size, n_tail, buf), 0.
        if not self._async:  # synchronous branch = the semantic definition (base.py:411-426)
            history.push(x[:batch_size].detach(), n_id[:batch_size], offset, count)
            idx = n_id[batch_size:]

            def fill(dst):
                self._pull_rows(history.emb, idx, n_id, dst)
            return _PushPull.apply(x, fill, batch_size, n_tail), 0.
        pulled = self.pool.synchronize_pull()

        def fill(dst):
            dst.copy_(pulled[:n_tail, :dst.size(1)])
        self.pool.async_push(x[:batch_size].detach(), offset, count, history.emb)
        out = _PushPull.apply(x, fill, batch_size, n_tail)
        self.pool.free_pull()
        return out, 0.

    def pull_ahead(self, histories, x: Tensor, batch_size: Optional[int], n_id: Optional[Tensor],
                   width: Optional[int] = None):
        """GAS steps pull the halo rows n_id[B:] and push the batch rows n_id[:B]: disjoint rows, so no
        pull depends on anything the step computes.  All of a step's pulls are therefore issued up
        front on a side stream (a parallel branch of the captured graph), each into the tail of a
        [B + H, F] buffer whose head the layer's GEMM later writes in place.  Returns one
        (buffer, event) per history - the consumer waits for the event - or None when the step does
        not take this path (host histories / NCCL transport / full-batch)."""
        pre = getattr(n_id, 'prefetched_pulls', None)
        if pre is not None:  # pulled one step ahead (train.GraphedTrainer host_prefetch): nothing to issue
            return pre
        if (_NO_AHEAD or n_id is None or batch_size is None or self._async or not x.is_cuda
                or not (self.emb_device.type == 'cuda' or self._direct_host)
                or getattr(n_id, 'halo_plan', None) is not None):
            return None
        n_tail = n_id.numel() - batch_size
        main = torch.cuda.current_stream(x.device)
        side = self._pull_stream
        if side is None:
            side = self._pull_stream = torch.cuda.Stream(x.device, priority=_PULL_PRIORITY)
        bufs = [torch.empty((batch_size + n_tail, width or h.emb.size(1)), dtype=x.dtype, device=x.device)
                for h in histories]
        side.wait_stream(main)
        idx = n_id[batch_size:]
        out = []
        with torch.cuda.stream(side):
            for h, buf in zip(histories, bufs):
                if n_tail > 0:
                    self._pull_rows(h.emb, idx, n_id, buf[batch_size:])
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((buf, ev))
        return out

    def prefetch_pulls(self, batch_size: int, n_id: Tensor, dtype=torch.float32):
        """The halo pulls of a GAS step, issued outside the step (train.GraphedTrainer replays them with
        the collate of the NEXT batch on a side stream, so with pinned-host tables the PCIe transfer of
        step i + 1 overlaps the compute of step i).  The buffers are attached to ``n_id`` and picked up
        by :meth:`pull_ahead`.  Rows pushed by the step in flight are read one step staler than in the
        sequential loop (or mid-copy: every 16-byte piece is one of the two valid versions) - the
        staleness GAS is built on, but not the reference's exact schedule: an opt-in mode."""
        hists = self._gas_pull_histories()
        n_tail = n_id.numel() - batch_size
        idx = n_id[batch_size:]
        out = []
        for h in hists:
            buf = torch.empty((batch_size + n_tail, h.emb.size(1)), dtype=dtype, device=self.device)
            if n_tail > 0:
                self._pull_rows(h.emb, idx, n_id, buf[batch_size:])
            out.append((buf, None))
        n_id.prefetched_pulls = out
        return out

    def push_only(self, history, x: Tensor, batch_size: Optional[int] = None,
                  n_id: Optional[Tensor] = None, offset: Optional[Tensor] = None,
                  count: Optional[Tensor] = None):
        """Pushes updated embeddings to `history` without pulling (base.py:458-499)."""
        if n_id is None and x.size(0) != self.num_nodes:
            return x
        if n_id is None and x.size(0) == self.num_nodes:
            history.push(x)
            return x, 0.
        assert n_id is not None
        if batch_size is None:
            history.push(x, n_id)
            return x, 0.
        if not self._async:
            history.push(x[:batch_size].detach(), n_id[:batch_size], offset, count)
        else:
            self.pool.async_push(x[:batch_size].detach(), offset, count, history.emb)
        return x[:batch_size], 0.

    @property
    def _out(self):
        if self.__out is None:
            if self.shard is not None:  # logits of this rank's rows only
                self.__out = torch.empty(self.shard.num_local, self.out_channels, device=self.emb_device)
            elif self.emb_device.type == 'cuda':
                self.__out = torch.empty(self.num_nodes, self.out_channels, device=self.emb_device)
            else:
                self.__out = torch.empty(self.num_nodes, self.out_channels,
                                         pin_memory=torch.cuda.is_available())
        return self.__out

    # ------------------------------------------------------------------------------------------
    # layer-wise sweeps
    # ------------------------------------------------------------------------------------------
    def _sweep_push(self, pool, x: Tensor, offset, count, table: Tensor):
        """x rows -> table[offset_i : +count_i] (pool.async_push in the reference)."""
        if table.size(1) > x.size(1):  # zero-pad to the table width (gcn.py:355-359)
            xp = x.new_zeros((x.size(0), table.size(1)))
            xp[:, :x.size(1)] = x
            x = xp
        if pool is not None:
            pool.async_push(x, offset, count, table)
        else:
            if self._row_lo:
                offset = offset - self._row_lo
            ops.copy_slices(x.contiguous(), table, offset, count, 1)

    def _sweep_pull_all(self, loader, table: Tensor):
        """Pool mode: enqueue the pulls of every batch (base.py:555-557)."""
        if self.pool is not None:
            for _, batch_size, n_id, offset, count, _ in loader:
                self.pool.async_pull(table, offset, count, n_id[batch_size:])

    def _sweep_pull(self, table: Tensor, batch_size, n_id, offset, count) -> Tensor:
        """[x_B ; x_halo] of one batch from `table` (pool.synchronize_pull()[:|n_id|])."""
        if self.pool is not None:
            return self.pool.synchronize_pull()[:n_id.numel()]
        out = torch.empty((n_id.numel(), table.size(1)), dtype=table.dtype, device=self.device)
        ops.copy_slices(table, out, offset - self._row_lo if self._row_lo else offset, count, 0)
        self._pull_rows(table, n_id[batch_size:], n_id, out[batch_size:])
        return out

    def _sweep_sync(self):
        if self.pool is not None:
            self.pool.synchronize_push()
        if self.pool_ag is not None:
            self.pool_ag.synchronize_push()
        if self.shard is not None and self.shard.world_size > 1 and getattr(self, 'transport', '') == 'p2p':
            # peers read this rank's rows directly: a layer phase must be complete on every rank before
            # the next one starts (the NCCL transport gets this ordering from its collectives)
            hook = getattr(self, '_sweep_phase_hook', None)
            if hook is not None:   # train.GraphedSweep: one CUDA graph per layer phase, it barriers itself
                hook()
                return
            import torch.distributed as dist
            torch.cuda.current_stream(self.device).synchronize()
            dist.barrier(group=self.shard.group)

    @torch.no_grad()
    def mini_inference(self, loader, use_aggregation=True) -> Tensor:
        """Layer-wise evaluation sweep (base.py:509-603): layer-l outputs -> histories[l+1], logits
        -> ``_out``."""
        loader = [sub_data + ({}, ) for sub_data in loader]
        for data, batch_size, n_id, offset, count, state in loader:
            x = data.x.to(self.device)
            adj_t = data.adj_t.to(self.device)
            out = self.forward_layer(0, x, adj_t, state, use_aggregation)[:batch_size]
            self._sweep_push(self.pool, out, offset, count, self.histories[1].emb)
        self._sweep_sync()
        for i in range(1, len(self.histories) - 1):
            self._sweep_pull_all(loader, self.histories[i].emb)
            for batch, batch_size, n_id, offset, count, state in loader:
                adj_t = batch.adj_t.to(self.device)
                x = self._sweep_pull(self.histories[i].emb, batch_size, n_id, offset, count)
                out = self.forward_layer(i, x, adj_t, state, use_aggregation)[:batch_size]
                self._sweep_push(self.pool, out, offset, count, self.histories[i + 1].emb)
                if self.pool is not None:
                    self.pool.free_pull()
            self._sweep_sync()
        self._sweep_pull_all(loader, self.histories[-1].emb)
        for batch, batch_size, n_id, offset, count, state in loader:
            adj_t = batch.adj_t.to(self.device)
            x = self._sweep_pull(self.histories[-1].emb, batch_size, n_id, offset, count)
            out = self.forward_layer(self.num_layers - 1, x, adj_t, state, use_aggregation)[:batch_size]
            self._sweep_push(self.pool, out, offset, count, self._out)
            if self.pool is not None:
                self.pool.free_pull()
        self._sweep_sync()
        return self._out

    # per-model hooks of the IncAgg refresh --------------------------------------------------
    def _refresh_layer0_input(self, x: Tensor) -> Tensor:
        """M_in of layer 0 for all B+H rows (what the first propagation aggregates)."""
        raise NotImplementedError

    def _refresh_aggregate(self, adj_t: SparseTensor, x: Tensor) -> Tensor:
        """M_ag = aggregate of stale layer inputs (sum; GraphSAGE overrides with mean)."""
        return adj_t @ x

    @torch.no_grad()
    def mini_inference_vr(self, loader, use_aggregation=True) -> Tensor:
        """Per-epoch refresh of M_in (``histories``) and M_ag (``histories_ag``) by a layer-wise sweep
        (gcn.py:335-410, gcn2.py:432-507, appnp.py:228-314, graphsage.py:862-960).  Where the layer
        aggregates its input directly (GCN2 / APPNP / GraphSAGE) the aggregate is computed once and
        handed to ``forward_layer`` instead of being recomputed (the reference runs the SpMM twice)."""
        loader = [sub_data + ({}, ) for sub_data in loader]
        share = getattr(self, '_share_refresh_aggregate', False) and not self.training
        for data, batch_size, n_id, offset, count, state in loader:
            x = data.x.to(self.device)
            adj_t = data.adj_t.to(self.device)
            m_in0 = self._refresh_layer0_input(x)
            m_ag0 = self._refresh_aggregate(adj_t, m_in0)
            if share:
                state['m_in0'] = m_in0
                out = self.forward_layer(0, x, adj_t, state, use_aggregation, agg=m_ag0)[:batch_size]
                state.pop('m_in0', None)
            else:
                out = self.forward_layer(0, x, adj_t, state, use_aggregation)[:batch_size]
            self._sweep_push(self.pool_ag, m_ag0, offset, count, self.histories_ag[0].emb)
            self._sweep_push(self.pool, m_in0[:batch_size], offset, count, self.histories[0].emb)
            self._sweep_push(self.pool, out, offset, count, self.histories[1].emb)
        self._sweep_sync()
        for i in range(1, len(self.histories) - 1):
            self._sweep_pull_all(loader, self.histories[i].emb)
            for batch, batch_size, n_id, offset, count, state in loader:
                adj_t = batch.adj_t.to(self.device)
                x = self._sweep_pull(self.histories[i].emb, batch_size, n_id, offset, count)
                m_ag = self._refresh_aggregate(adj_t, x)
                self._sweep_push(self.pool_ag, m_ag, offset, count, self.histories_ag[i].emb)
                if share:
                    out = self.forward_layer(i, x, adj_t, state, use_aggregation, agg=m_ag)[:batch_size]
                else:
                    out = self.forward_layer(i, x, adj_t, state, use_aggregation)[:batch_size]
                self._sweep_push(self.pool, out, offset, count, self.histories[i + 1].emb)
                if self.pool is not None:
                    self.pool.free_pull()
            self._sweep_sync()
        self._sweep_pull_all(loader, self.histories[-1].emb)
        for batch, batch_size, n_id, offset, count, state in loader:
            adj_t = batch.adj_t.to(self.device)
            x = self._sweep_pull(self.histories[-1].emb, batch_size, n_id, offset, count)
            m_ag = self._refresh_aggregate(adj_t, x)
            if share:
                out = self.forward_layer(self.num_layers - 1, x, adj_t, state, use_aggregation,
                                         agg=m_ag)[:batch_size]
            else:
                out = self.forward_layer(self.num_layers - 1, x, adj_t, state, use_aggregation)[:batch_size]
            self._sweep_push(self.pool_ag, m_ag, offset, count, self.histories_ag[-1].emb)
            self._sweep_push(self.pool, out, offset, count, self._out)
            if self.pool is not None:
                self.pool.free_pull()
        self._sweep_sync()
        return self._out

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def forward_layer(self, *args, **kwargs):
        raise NotImplementedError
