"""The driver loop that calls the hot path (reference: main.py:47-110 ``mini_train`` / ``mini_test``,
main.py:140-236 setup) and the hyper-parameters of the five BASELINE configs.

Not a product by itself: it is the caller the benchmark and the parity tests need.
Differences from the reference loop, none of which change results:
  * ``total_loss`` / ``total_examples`` are initialised (the fork uses them unassigned),
  * no ``torch.cuda.synchronize()`` after every forward (main.py:72,76) and no autograd anomaly mode
    (main.py:41); the loss is accumulated on the device and read back once per epoch.
"""
from typing import Any, Dict, Optional

import contextlib
import os
import torch

from . import models
from .loader import EvalSubgraphLoader, SubgraphLoader
from .metis import metis, permute
from .preprocess import gcn_norm, set_diag
from .synthetic import get_data
from .utils import dropout

CONFIG_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'conf', 'configs.yaml')


def _parse_scalar(text: str):
    import yaml
    return yaml.safe_load(text)


def apply_overrides(conf: Dict[str, Any], overrides) -> Dict[str, Any]:
    """hydra-style command-line overrides (reference main.py:112-122 runs under @hydra.main):
    ``++key=value`` / ``key=value`` set a top-level entry, ``++architecture.hidden_channels=64`` a
    nested one; values are parsed as YAML scalars / lists.  Returns a new dict."""
    import copy
    out = copy.deepcopy(conf)
    for item in overrides or ():
        key, sep, value = item.lstrip('+').partition('=')
        if not sep or not key:
            raise ValueError(f'override {item!r} is not of the form [++]key=value')
        node = out
        parts = key.split('.')
        for k in parts[:-1]:
            node = node.setdefault(k, {})
            if not isinstance(node, dict):
                raise ValueError(f'override {item!r}: {k!r} is not a section')
        node[parts[-1]] = _parse_scalar(value)
    return out


def load_configs(path: Optional[str] = None, overrides=None) -> Dict[str, Dict[str, Any]]:
    """The per-config hyper-parameters (conf/configs.yaml, values from the reference's
    conf/model/*.yaml where the fork has them), with optional hydra-style overrides applied to every
    config."""
    import yaml
    with open(path or CONFIG_FILE) as f:
        table = yaml.safe_load(f)
    return {k: apply_overrides(v, overrides) for k, v in table.items()}


# Hyper-parameters per BASELINE config
CONFIGS: Dict[str, Dict[str, Any]] = load_configs()


class FlatAdam:
    """``torch.optim.Adam`` (the reference's optimizer, main.py:196-201: a weight-decayed and a
    non-decayed parameter group) with every parameter, gradient and moment in ONE flat buffer each and
    the update as a single kernel launch (``incagg_adam_step``) instead of ~12 multi-tensor launches.
    Parameters and ``.grad`` become views into the flat buffers (static addresses: CUDA-graph friendly,
    and the gradient all-reduce of ``parallel.GradAverager`` runs on the same buffer without copies).
    Same update arithmetic as torch's Adam (no amsgrad; L2 weight decay added to the gradient)."""

    def __init__(self, param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        groups = [dict(g) for g in param_groups]
        assert 1 <= len(groups) <= 2, 'FlatAdam handles the one or two parameter groups of main.py'
        seen, flat_list = set(), []
        for g in groups:
            ps = [p for p in g['params'] if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in ps)
            g['params'] = ps
            g.setdefault('weight_decay', 0.)
            g['lr'], g['betas'], g['eps'], g['capturable'] = lr, betas, eps, True
            flat_list += ps
        assert flat_list and all(p.is_cuda and p.dtype == torch.float32 for p in flat_list)
        self.param_groups = groups
        self.params = flat_list
        dev = flat_list[0].device
        n = sum(p.numel() for p in flat_list)
        self.n_first = sum(p.numel() for p in groups[0]['params'])
        self.flat_p = torch.empty(n, device=dev)
        self.flat_g = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.step_t = torch.zeros(1, device=dev)
        self._arrivals = torch.zeros(1, dtype=torch.int32, device=dev)
        o = 0
        for p in flat_list:
            k = p.numel()
            self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + k].view_as(p)
            p.grad = self.flat_g[o:o + k].view_as(p)
            p._flat_grad = True   # nn._grad_buffer: weight-gradient GEMMs accumulate into the view
            o += k
        self.state = {'step': self.step_t, 'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq}

    def zero_grad(self, set_to_none: bool = False):
        self.flat_g.zero_()  # the views stay in place

    @torch.no_grad()
    def step(self):
        from . import ops
        g0 = self.param_groups[0]
        wd_rest = self.param_groups[1]['weight_decay'] if len(self.param_groups) > 1 else g0['weight_decay']
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.n_first, g0['lr'],
                      g0['betas'][0], g0['betas'][1], g0['eps'], g0['weight_decay'], wd_rest, self.step_t,
                      self._arrivals)


def mini_train(model, loader, criterion, optimizer, max_steps, grad_norm=None, edge_dropout=0.0,
               epoch=0, VR_update=False, drift_norm=2, aggregate_combined=True,
               use_aggregation=True) -> Dict[str, float]:
    """One training epoch over the partition mini-batches (main.py:47-96)."""
    model.train()
    total_loss = torch.zeros((), dtype=torch.float64, device=model.device)
    total_examples = torch.zeros((), dtype=torch.float64, device=model.device)
    steps = 0
    for i, sub in enumerate(loader):
        batch, batch_size, n_id, offset, count = sub
        if edge_dropout > 0.:
            batch.adj_t = dropout(batch.adj_t.to(model.device), p=edge_dropout)
        # forward (model.VR_call / model.__call__) -> masked CE -> backward -> clip -> optimizer step
        ln, n = train_step(model, sub, optimizer, VR_update=VR_update, grad_norm=grad_norm, epoch=epoch,
                           batch_idx=i)
        total_loss += ln.double()
        total_examples += n.double()
        steps += 1
        if (i + 1) >= max_steps and (i + 1) < len(loader):
            break
    tl, te = float(total_loss), float(total_examples)
    from . import ops
    ops.check_device_errors()  # an index outside a table anywhere in the epoch raises here
    return {'loss': tl / max(te, 1.), 'steps': steps}


_SIDE_STREAMS = {}


def _backward_prep(adj_t, batch_size, VR_update):
    """Fork: build adj_t's transposed CSR + SpMM plans on a side stream.  Returns the stream to join."""
    if not adj_t.col.is_cuda or adj_t.nnz() == 0 or os.environ.get('INCAGG_NO_AHEAD') == '1':
        return None
    dev = adj_t.col.device
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        adj_t.t_csr()
        adj_t.t_plan()
        if not VR_update and batch_size < adj_t.size(1):
            adj_t.t_plan_prefix(batch_size)
    return side


def forward_backward(model, sub, optimizer, VR_update=False, averager=None, epoch=0, batch_idx=0,
                     loss_acc=None):
    """Forward + loss + backward of one mini_train iteration on an already collated batch.  Returns
    (loss * n_train, n_train) as device scalars (no host synchronisation); ``loss_acc`` (float64[2] on the
    device, optional) additionally receives ``+= [loss * n_train, n_train]`` (main.py:82) from the loss
    kernel itself."""
    batch, batch_size, n_id, offset, count = sub
    x, adj_t = batch.x, batch.adj_t
    y, train_mask = batch.y[:batch_size], batch.train_mask[:batch_size]
    # The transposed CSR and the plans that only the backward pass needs are built on a side stream
    # while the forward pass runs (a fork / join that CUDA-graph capture records as parallel branches).
    side = _backward_prep(adj_t, batch_size, VR_update)
    fused_ce = y.dim() == 1 and x.is_cuda
    n_train = None
    zeroed = False
    if side is not None:
        from . import ops
        with torch.cuda.stream(side):
            if fused_ce:   # the row count of the loss depends on the batch alone
                n_train = ops.mask_count(train_mask)
            # flat gradient buffers are cleared here, beside the forward pass (a fill on the main stream
            # would sit between the last forward kernel and the loss)
            if averager is not None:
                averager.zero()  # gradients are views into the averager's flat buffer
                zeroed = True
            elif hasattr(optimizer, 'flat_g'):
                optimizer.zero_grad(set_to_none=True)
                zeroed = True
    if VR_update:
        out = model.VR_call(x, adj_t, batch_size, n_id, offset, count, epoch=epoch, batch_idx=batch_idx)['out']
    else:
        out = model(x, adj_t, batch_size, n_id, offset, count)['out']
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)
    if zeroed:
        pass
    elif averager is not None:
        averager.zero()  # gradients are views into the averager's flat buffer
    else:
        optimizer.zero_grad(set_to_none=True)
    if fused_ce and out.dtype == torch.float32:
        # criterion(out[mask], y[mask]) as one fused masked cross-entropy whose gradient feeds
        # out.backward directly (no autograd node: loss.backward() would materialise a ones tensor and
        # multiply the gradient by it).  Only the gradient kernel sits on the step's critical path: the
        # row count was computed beside the forward pass, the loss value (and the running epoch loss)
        # is finished beside the backward pass.
        from . import ops
        from .nn import weight_grads_on_side_stream, side_stream_of_weight_grads
        if n_train is None:
            n_train = ops.mask_count(train_mask)
        dlogits, ce_ws = ops.masked_ce_rows(out.detach(), y, train_mask, n_train)
        # (single-GPU path; with a gradient averager the step keeps the one-stream backward that the
        # multi-GPU runs of this round were measured and checked with)
        one_graph = averager is None or getattr(averager, 'fused', False)
        with (weight_grads_on_side_stream(out.device) if one_graph else contextlib.nullcontext()):
            wside = side_stream_of_weight_grads()
            if wside is not None:
                wside.wait_stream(torch.cuda.current_stream(out.device))
                with torch.cuda.stream(wside):
                    out3 = ops.masked_ce_finish(ce_ws, out.size(0), n_train, loss_acc)
            else:
                out3 = ops.masked_ce_finish(ce_ws, out.size(0), n_train, loss_acc)
            out.backward(dlogits)
        return out3[0], out3[2]
    w = train_mask.to(out.dtype)
    n = w.sum()
    per_row = torch.nn.functional.binary_cross_entropy_with_logits(
        out, y.to(out.dtype), reduction='none').mean(dim=-1)
    loss = (per_row * w).sum() / n.clamp(min=1.)
    loss.backward()
    if loss_acc is not None:
        loss_acc += torch.stack([(loss.detach() * n).double(), n.double()])
    return loss.detach() * n, n


def apply_update(model, optimizer, grad_norm=None, averager=None):
    """Second half of the iteration: (averaged) gradients -> clip -> optimizer step."""
    if getattr(averager, 'fused', False):   # all-reduce + scaling + Adam in one peer-memory kernel
        assert grad_norm is None
        averager.step()
        return
    if averager is not None:
        averager.scale()
    if grad_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), grad_norm)
    optimizer.step()


def train_step(model, sub, optimizer, VR_update=False, grad_norm=None, averager=None, epoch=0,
               batch_idx=0):
    """One iteration of the mini_train loop body on an already collated batch."""
    ln, n = forward_backward(model, sub, optimizer, VR_update, averager, epoch, batch_idx)
    if averager is not None and not getattr(averager, 'fused', False):
        averager.all_reduce()
    apply_update(model, optimizer, grad_norm, averager)
    return ln, n


class GraphedTrainer:
    """Replays CUDA graphs of the training step, one per distinct batch (group of partitions).

    A training step is ~200 kernel launches of a few microseconds each, so issuing it from Python is
    launch-bound.  With a fixed partition -> batch assignment (``batch_size`` partitions per step drawn
    from a fixed set of groups; C3 uses batch_size = 1, i.e. 150 distinct batches) every step of a given
    batch has the same kernel sequence and the same sizes: the whole step - GPU collate (relabel +
    gathers), forward, history push / pull, loss, backward, Adam - is captured once per batch and
    replayed every epoch.  Nothing is cached between replays: each replay runs every kernel again on
    the current weights and history tables.  All graphs share one memory pool (they are replayed one at
    a time).  Data-dependent sizes (halo counts) are read back once, before the capture.

    Multi-GPU: the step is captured as two graphs around the NCCL gradient all-reduce, which is issued
    eagerly between them (collate + forward + backward | all_reduce | scale + clip + Adam).  Halo rows
    of other ranks are read by the captured gather kernels straight out of the peers' HBM (p2p
    transport), so the graphs contain no collective.

    ``pipeline_collate=True``: the collate of a batch (relabel + gathers; with host-resident inputs
    the kernels and DMA transfers that read the graph, features, labels and masks out of pinned host
    memory over PCIe) is captured as its own graph with a private memory pool and persistent outputs,
    and :meth:`run` replays the collate of step i+1 on a side stream while step i computes - what the
    reference's DataLoader workers + ``non_blocking`` copies do for its loop.  Every step still
    collates its batch (and moves its inputs host -> device); nothing is reused between steps.
    """

    def __init__(self, model, loader, optimizer, VR_update=False, grad_norm=None, averager=None,
                 pipeline_collate=False, host_prefetch=False):
        """``host_prefetch`` (with ``pipeline_collate``, GAS mode): the halo pulls of a step are part
        of its collate graph, i.e. they run one step ahead on the side stream - with pinned-host history
        tables the PCIe reads of step i + 1 overlap the compute of step i (see
        ``ScalableGNN.prefetch_pulls`` for what that changes)."""
        self.model, self.loader, self.optimizer = model, loader, optimizer
        self.host_prefetch = bool(host_prefetch) and bool(pipeline_collate) and not VR_update
        if model.pool is not None:
            # pinned-host tables: the step is captured with the direct UVA / DMA path instead of the
            # pool's slot protocol (whose host-side queue cannot be part of a graph)
            model._direct_host = True
        self.vr, self.grad_norm, self.averager = VR_update, grad_norm, averager
        self.graphs = {}
        self.pool = None
        self.pipeline = bool(pipeline_collate)
        self.in_graphs = {}          # batch -> (collate graph, its persistent SubData)
        self._in_stream = None
        self._in_done, self._step_done = {}, {}
        self.acc = torch.zeros(2, dtype=torch.float64, device=model.device)  # sum(loss * n), sum(n)
        for g in optimizer.param_groups:
            if not g.get('capturable', False):
                raise RuntimeError('GraphedTrainer needs an optimizer built with capturable=True')

    def _body_a(self, ids):
        sub = self.loader._collate(list(ids))
        forward_backward(self.model, sub, self.optimizer, self.vr, self.averager, loss_acc=self.acc)

    def _body_b(self):
        apply_update(self.model, self.optimizer, self.grad_norm, self.averager)

    @property
    def _two_graphs(self) -> bool:
        """NCCL gradient all-reduce: issued eagerly between two graphs of a step.  Single GPU, or the
        fused peer-memory exchange (parallel.FusedGradSync): the whole step is one graph."""
        return self.averager is not None and not getattr(self.averager, 'fused', False)

    def _step_on(self, sub):
        forward_backward(self.model, sub, self.optimizer, self.vr, self.averager, loss_acc=self.acc)
        if not self._two_graphs:
            apply_update(self.model, self.optimizer, self.grad_norm, self.averager)

    def _launch_collate(self, key):
        """Replay the collate graph of batch `key` on the side stream (after the last step that read
        its output buffers)."""
        side = self._in_stream
        with torch.cuda.stream(side):
            prev = self._step_done.get(key)
            if prev is not None:
                side.wait_event(prev)
            self.in_graphs[key][0].replay()
            ev = torch.cuda.Event()
            ev.record(side)
        self._in_done[key] = ev

    def run(self, seq, after_step=None):
        """Steps over the batches `seq` in order; with ``pipeline_collate`` the collate of step i+1
        overlaps step i.  ``after_step(i)`` runs on the host after step i has been issued."""
        if not self.pipeline:
            for i, ids in enumerate(seq):
                self.step(ids)
                if after_step is not None:
                    after_step(i)
            return
        keys = [tuple(ids) for ids in seq]
        if not keys:
            return
        for k in keys:
            self.capture(k)
        main = torch.cuda.current_stream(self.model.device)
        self._in_stream.wait_stream(main)
        self._launch_collate(keys[0])
        for i, k in enumerate(keys):
            main.wait_event(self._in_done[k])
            g = self.graphs[k]
            g[0].replay()
            ev = torch.cuda.Event()
            ev.record(main)
            self._step_done[k] = ev
            if i + 1 < len(keys):
                self._launch_collate(keys[i + 1])
            if g[1] is not None:
                self.averager.all_reduce()  # NCCL, eager, between the two graphs
                g[1].replay()
            if after_step is not None:
                after_step(i)

    def _body(self, ids):
        self._body_a(ids)
        if self._two_graphs:
            self.averager.all_reduce()
        self._body_b()

    def warmup(self, ids, steps: int = 3):
        """Eager steps before the first capture (lazy optimizer state, cuBLAS workspaces, scratch)."""
        self.model.train()
        s = torch.cuda.Stream(self.model.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(steps):
                self._body(ids)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()

    def capture(self, ids):
        key = tuple(ids)
        if key in self.graphs:
            return
        self.model.train()
        if len(self.optimizer.state) == 0:
            raise RuntimeError('run GraphedTrainer.warmup() first: the optimizer state must exist before '
                               'a step is captured (its lazy initialisation cannot be part of a graph)')
        # make sure the loader knows the data-dependent sizes of this batch (one eager collate)
        lk = (('ib',) if self.vr else ('gas',)) + key
        if lk not in self.loader._known_sizes:
            self.loader._collate(list(ids))
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        if self.pipeline:
            if self._in_stream is None:
                self._in_stream = torch.cuda.Stream(self.model.device)
            # Every collate graph gets a memory pool of its OWN: its outputs are read by a step graph
            # that runs concurrently with the NEXT collate, so they must not share addresses with
            # another collate graph's temporaries (which a common pool would hand out again as soon
            # as they are freed at capture time).
            gi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gi):
                sub = self.loader._collate(list(ids))
                # the forward SpMM plan depends on the batch structure alone: built here, one step ahead,
                # instead of between the first Linear and the first SpMM of the step
                plan = sub.data.adj_t.plan() if sub.data.adj_t.col.is_cuda and sub.data.adj_t.nnz() else None
                if self.host_prefetch:
                    self.model.prefetch_pulls(sub.batch_size, sub.n_id)
            self.in_graphs[key] = (gi, sub, plan)   # the outputs stay allocated: the step graph reads them
            g, gb = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(g, pool=self.pool):
                self._step_on(sub)
            sub.data.adj_t.drop_caches()
            if self._two_graphs:
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, pool=self.pool):
                    self._body_b()
            self.graphs[key] = (g, gb)
        elif not self._two_graphs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool):
                self._body(ids)
            self.graphs[key] = (g, None)
        else:
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, pool=self.pool):
                self._body_a(ids)
            with torch.cuda.graph(gb, pool=self.pool):
                self._body_b()
            self.graphs[key] = (ga, gb)

    MAX_GRAPHS = 4096

    def step(self, ids):
        key = tuple(ids)
        if self.pipeline:
            return self.run([ids])
        g = self.graphs.get(key)
        if g is None:
            # shuffled groups of several partitions never repeat: nothing to replay, issue eagerly
            if not self.loader.fixed_batches or len(self.graphs) >= self.MAX_GRAPHS:
                self.model.train()
                return self._body(ids)
            self.capture(ids)
            g = self.graphs[key]
        g[0].replay()
        if g[1] is not None:
            self.averager.all_reduce()  # NCCL, eager, between the two graphs
            g[1].replay()

    def epoch(self, max_steps=None):
        """One epoch in the loader's (shuffled) batch order.  Returns the mean training loss."""
        self.acc.zero_()
        steps = 0
        if self.pipeline and self.loader.fixed_batches:
            seq = self.loader._batches_of_epoch()
            seq = seq if max_steps is None else seq[:max_steps]
            self.run(seq)
            a = self.acc.tolist()
            from . import ops
            ops.check_device_errors()
            return {'loss': a[0] / max(a[1], 1.), 'steps': len(seq)}
        for ids in self.loader._batches_of_epoch():
            self.step(ids)
            steps += 1
            if max_steps is not None and steps >= max_steps:
                break
        a = self.acc.tolist()
        return {'loss': a[0] / max(a[1], 1.), 'steps': steps}


class GraphedSweep:
    """The per-epoch layer-wise sweep (``mini_inference`` / ``mini_inference_vr``: all partitions, all
    layers, ~10^4 launches of small kernels) captured once and replayed every epoch.  The evaluation
    loader pre-materialises its subgraphs (as the reference's does, loader.py:153-170), so the kernel
    sequence is fixed; every replay recomputes all tables from the current weights.

    Single GPU: ONE graph.  Sharded tables (p2p transport): peers read this rank's rows directly, so a
    layer phase must be complete on every rank before the next one starts; the sweep is captured as one
    graph PER LAYER PHASE (cut where the eager sweep synchronises, ``ScalableGNN._sweep_sync``) and the
    replays are separated by a stream synchronisation + ``dist.barrier`` (L + 1 barriers per sweep)."""

    def __init__(self, model, loader, VR_update=False, use_aggregation=True):
        if model.pool is not None:
            raise RuntimeError('GraphedSweep needs HBM-resident histories')
        if model.shard is not None and model.shard.world_size > 1 and getattr(model, 'transport', '') != 'p2p':
            raise RuntimeError('GraphedSweep on sharded tables needs the p2p transport (NCCL collectives '
                               'of the all-to-all-v transport are issued eagerly)')
        self.model, self.loader, self.vr, self.use_aggregation = model, loader, VR_update, use_aggregation
        self.graph = None
        self.phases = None
        self.sharded = model.shard is not None and model.shard.world_size > 1

    @torch.no_grad()
    def _body(self):
        if self.vr:
            return self.model.mini_inference_vr(loader=self.loader, use_aggregation=self.use_aggregation)
        return self.model.mini_inference(self.loader, self.use_aggregation)

    def _phase_barrier(self):
        import torch.distributed as dist
        torch.cuda.current_stream(self.model.device).synchronize()
        dist.barrier(group=self.model.shard.group)

    @torch.no_grad()
    def _capture_phases(self):
        """One graph per layer phase: the capture is cut at every ``_sweep_sync`` of the sweep."""
        dev = self.model.device
        pool = torch.cuda.graph_pool_handle()
        phases, cur, dirty = [], [None], [False]
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))

        def begin():
            cur[0] = torch.cuda.CUDAGraph()
            cur[0].capture_begin(pool=pool)

        def cut():
            cur[0].capture_end()
            phases.append(cur[0])
            begin()

        with torch.cuda.stream(s):
            begin()
            self.model._sweep_phase_hook = cut
            try:
                self._body()
            finally:
                self.model._sweep_phase_hook = None
                cur[0].capture_end()   # the sweep ends with a cut: this last graph is empty, not kept
        torch.cuda.current_stream(dev).wait_stream(s)
        return phases

    @torch.no_grad()
    def __call__(self):
        self.model.eval()
        if self.graph is None and self.phases is None:
            s = torch.cuda.Stream(self.model.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._body()  # eager warm-up (plans, transposes, scratch, output buffer)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if self.sharded:
                self.phases = self._capture_phases()
            else:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._body()
        if self.sharded:
            for g in self.phases:
                g.replay()
                self._phase_barrier()
        else:
            self.graph.replay()
        return self.model._out


@torch.no_grad()
def mini_test(model, loader, use_aggregation=True, VR_update=False):
    model.eval()
    if VR_update:
        return model.mini_inference_vr(loader=loader, use_aggregation=use_aggregation)
    return model(loader=loader, use_aggregation=use_aggregation)


def build(config: str, device='cuda', seed: int = 0, scale: int = 1, history_device='cuda',
          overrides: Optional[Dict[str, Any]] = None, log: bool = False, data_device=None,
          shuffle: bool = True, host_resident: bool = False, data=None, rank: int = 0,
          eval_batch_size: Optional[int] = None,
          world_size: int = 1, transport: str = 'p2p', force_metis: bool = False, fused_adam: bool = True):
    """Everything main.py:140-201 sets up for one named config on synthetic data: returns a dict
    with data, ptr, loaders, model, optimizer, criterion and the config."""
    conf = dict(CONFIGS[config])
    if overrides:
        conf.update(overrides)
    torch.manual_seed(seed)
    # data_device='cpu' generates with the CPU generator (identical inputs for the CPU oracle)
    if data is None:
        data, in_channels, out_channels = get_data('', conf['dataset'], seed=seed,
                                                   device=data_device or device, scale=scale,
                                                   num_parts=conf['num_parts'])
        raw = data
        data = data.to(device)
        data.adj_t.clustered_parts = getattr(raw.adj_t, 'clustered_parts', None)
        # force_metis: run the real METIS k-way partitioner even on a pre-clustered synthetic graph
        perm, ptr = metis(data.adj_t, num_parts=conf['num_parts'], log=log, force=force_metis)
        data = permute(data, perm, log=log)
        if force_metis:
            raw = data.to('cpu')  # the partition-ordered graph is what the oracle is given
        if conf['loop']:
            data.adj_t = set_diag(data.adj_t)
        if conf['norm']:
            data.adj_t = gcn_norm(data.adj_t, add_self_loops=False)
    else:  # (data, ptr, in_channels, out_channels) of an earlier build(): reuse the preprocessed graph
        data, ptr, in_channels, out_channels = data
        raw = None
    if host_resident:  # the reference's layout: graph + features in pinned host memory
        data = data.pin_memory()
    criterion = torch.nn.CrossEntropyLoss()
    shard = None
    if world_size > 1:  # partitions, CSR rows and history rows are sharded over the ranks
        from .parallel import Shard
        shard = Shard(ptr, rank, world_size)
    train_loader = SubgraphLoader(data, ptr, batch_size=conf['batch_size'], shuffle=shuffle,
                                  num_neighbors=-1, type='train', IB=conf['VR_update'], log=log,
                                  device=device, shard=shard, halo_plans=(transport == 'nccl'))
    # eval_batch_size: partitions merged into one batch of the layer-wise sweeps.  The reference sizes
    # it for the GPU memory of its day (= the training batch size); with the tables HBM-resident a
    # sweep over few large batches computes the same rows with far fewer, fuller launches.
    ebs = eval_batch_size or conf['batch_size']
    if shard is not None and world_size > 1:
        # merged evaluation batches must not straddle two ranks: the largest block size <= ebs that
        # divides every rank's number of partitions
        import math
        g = 0
        for r in range(world_size):
            g = math.gcd(g, shard.part_bounds[r + 1] - shard.part_bounds[r])
        ebs = max(d for d in range(1, min(ebs, g) + 1) if g % d == 0)
    eval_loader = EvalSubgraphLoader(data, ptr, batch_size=ebs, log=log,
                                     device=device, shard=shard, halo_plans=(transport == 'nccl'))
    buffer_size = max(n_id.numel() for _, _, n_id, _, _ in eval_loader) * 2
    kwargs = {}
    if conf['model'][:3] == 'PNA':
        kwargs['deg'] = data.adj_t.storage.rowcount()
    GNN = getattr(models, conf['model'])
    model = GNN(num_nodes=data.num_nodes, in_channels=in_channels, out_channels=out_channels,
                pool_size=conf['pool_size'], buffer_size=buffer_size, device=history_device,
                **conf['architecture'], **kwargs).to(device)
    if shard is not None:
        model.shard_histories(shard, transport=transport)
    groups = [dict(params=list(model.reg_modules.parameters()), weight_decay=conf['reg_weight_decay']),
              dict(params=list(model.nonreg_modules.parameters()), weight_decay=conf['nonreg_weight_decay'])]
    if fused_adam and torch.device(device).type == 'cuda':
        optimizer = FlatAdam(groups, lr=conf['lr'])   # Adam as one launch over flat buffers
    else:
        optimizer = torch.optim.Adam(groups, lr=conf['lr'], capturable=torch.device(device).type == 'cuda')
    max_steps = conf['max_steps'] if conf['max_steps'] != -1 else int(conf['num_parts'] / conf['batch_size'])
    return dict(conf=conf, data=data, raw=raw, ptr=ptr, shard=shard, train_loader=train_loader, eval_loader=eval_loader,
                model=model, optimizer=optimizer, criterion=criterion, max_steps=max_steps,
                in_channels=in_channels, out_channels=out_channels, buffer_size=buffer_size)
