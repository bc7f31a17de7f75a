#!/usr/bin/env python
"""bench.py — edges/s of one GCNII training epoch on the synthetic ogbn-products shape (BASELINE.json
metric, config C3) through this repo's hot path, plus the roofline of the dominant kernel, the
end-to-end figure with host-resident inputs and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode gas|incagg]
                    [--config C3] [--scale S]

A "step" is one training step of the reference's ``mini_train`` loop (main.py:58-92) on one partition
mini-batch: GPU collate (relabel_one_hop + feature gather) -> forward (SpMM + history push/pull per
layer) -> loss -> backward (transposed SpMM) -> Adam.  edges/s = sum over the timed steps of
nnz(batch adjacency) / time; over a whole epoch the batch adjacencies partition nnz(adj_t).

N > 1 (launched by torch.distributed.run, one rank per GPU): the partitions of the one graph are
sharded over the ranks (rank r owns a contiguous block of partitions and those rows of every history
table); each rank trains on batches of its own partitions, halo rows owned by other ranks are fetched
by an all-to-all-v over NCCL (GAS mode) and gradients are all-reduced every step.  A
step at N GPUs is N batches in flight, so per-GPU work per step is fixed ("weak").

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=150)   # one full C3 epoch (150 partitions, batch 1)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="gas", choices=["gas", "incagg"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--scale", type=int, default=1, help="divide nodes and edges (debug only)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="collate inside each step's graph instead of overlapping it with the previous step")
    ap.add_argument("--no-graphs", action="store_true", help="issue every step eagerly from Python")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU halo rows: NVLink peer loads in the gather kernel, or NCCL all-to-all-v")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=None)
    args = ap.parse_args()
    args.cpu_steps_given = args.cpu_steps is not None
    if args.cpu_steps is None:
        args.cpu_steps = 6
    return args


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU baseline (oracle) ----------------------------------------------------------------------------
def full_table(hist, num_nodes: int):
    """CPU copy of a history table as the [N, D] table the oracle keeps.  On a multi-GPU run the model
    holds only this rank's rows [row_offset, row_offset + n) of every table (models/base.py
    shard_histories): those rows are placed at their global position, the rows of other ranks stay
    zero (the CPU leg times arithmetic; its steps use this rank's own partitions)."""
    emb = hist.emb.detach().cpu()
    if emb.size(0) == num_nodes:
        return emb
    lo = int(getattr(hist, "row_offset", 0) or 0)
    out = torch.zeros(num_nodes, emb.size(1), dtype=emb.dtype)
    out[lo:lo + emb.size(0)] = emb
    return out


def cpu_baseline(run, mode: str, n_steps: int, threads: int):
    """Times the CPU restatement of the same training step (oracle/gas.py: C relabel + ATen CSR
    aggregation on the host cores) on `n_steps` partitions of the same graph and weights (this rank's
    own partitions on a multi-GPU run)."""
    from oracle import gas
    torch.set_num_threads(threads)
    gas.SPMM_IMPL = "csr"
    data, ptr, conf = run["data"], run["ptr"], run["conf"]
    rp, col, val = [t.cpu() if t is not None else None for t in data.adj_t.csr()]
    adj = gas.Adj(rp, col, val, data.num_nodes, data.num_nodes)
    x, y, mask = data.x.cpu(), data.y.cpu(), data.train_mask.cpu()
    kwargs = dict(conf["architecture"])
    if conf["model"] == "PNA":
        kwargs["deg"] = adj.rowptr[1:] - adj.rowptr[:-1]
    model = gas.OracleGNN(conf["model"], run["model"].state_dict(), data.num_nodes, run["in_channels"],
                          out_channels=run["out_channels"], dtype=torch.float32, **kwargs)
    for l in range(model.num_layers):
        model.histories[l].emb.copy_(full_table(run["model"].histories[l], data.num_nodes))
        if mode == "incagg":
            model.histories_ag[l].emb.copy_(full_table(run["model"].histories_ag[l], data.num_nodes))
    opt = torch.optim.Adam(model.parameters(), lr=conf["lr"])
    bs = conf["batch_size"]
    shard = run.get("shard")
    first = shard.part_lo if shard is not None else 0
    last = shard.part_hi if shard is not None else ptr.numel() - 1
    edges, t_total = 0, 0.0
    for s in range(n_steps + 1):
        group = [first + (s * bs + j) % (last - first) for j in range(bs)]
        t0 = time.perf_counter()
        b = gas.collate(adj, x, y, mask, ptr, group, within_batch=(mode == "incagg"))
        gas.train_epoch(model, [b], opt, vr=(mode == "incagg"), grad_norm=conf["grad_norm"])
        dt = time.perf_counter() - t0
        if s > 0:  # first step is warm-up
            t_total += dt
            edges += sum(int(adj.rowptr[int(ptr[p + 1])] - adj.rowptr[int(ptr[p])]) for p in group)
    gas.SPMM_IMPL = "gather"
    return edges / t_total, edges, t_total


MODEL_LABEL = {"GCN2": "GCNII", "GCN": "GCN", "APPNP": "APPNP", "GraphSAGE": "GraphSAGE", "PNA": "PNA"}
DATASET_LABEL = {"amazonproducts": "amazon-products"}


def metric_text(conf) -> str:
    """BASELINE.json's metric for the headline config (C3); the same wording for the other configs."""
    return (f"edges/s (train epoch, {MODEL_LABEL.get(conf['model'], conf['model'])}, "
            f"{DATASET_LABEL.get(conf['dataset'], conf['dataset'])}-shape)")


def workload_text(config: str, model: str, mode: str, dataset: str, nodes: int, nnz: int, parts: int,
                  batch: int) -> str:
    """`config.workload`: the same text in both arms (the driver compares the strings)."""
    return (f"{config}: {model} {mode.upper()} training steps on the synthetic {dataset} shape "
            f"({nodes} nodes, nnz(adj_t)={nnz} directed non-zeros incl. self loops; the named 61.9M is "
            f"used as directed nnz after symmetrisation), {parts} parts, batch {batch}, per GPU")


# ---- one measured pass ------------------------------------------------------------------------------
def batch_edges(run, batch_ids):
    rp = run["train_loader"]._rowptr_host
    ptr = run["ptr"]
    return sum(int(rp[int(ptr[b + 1])]) - int(rp[int(ptr[b])]) for b in batch_ids)


def timed_steps_graphed(run, mode, warmup, steps, dist):
    """W warm-up + K timed training steps, each step one replay of the CUDA graph of its batch
    (train.GraphedTrainer; graphs are captured before the timed region, one per partition)."""
    from incagg_gnn_b200.train import GraphedTrainer
    model, opt, loader, conf = run["model"], run["optimizer"], run["train_loader"], run["conf"]
    averager = None
    if dist is not None:
        from incagg_gnn_b200.parallel import make_grad_sync
        averager = make_grad_sync(model, opt, run["shard"], conf["grad_norm"], TRANSPORT)
    tr = GraphedTrainer(model, loader, opt, VR_update=(mode == "incagg"), grad_norm=conf["grad_norm"],
                        averager=averager, pipeline_collate=loader.fixed_batches and not NO_PIPELINE)
    rp_host, ptr = loader._rowptr_host, loader.ptr
    groups = loader._batches_of_epoch()
    tr.warmup(groups[0])
    # kernels of this library per step, counted on one eager step after the warm-up (a graph replay
    # re-launches exactly the captured kernels, which the host-side counter cannot see)
    from incagg_gnn_b200 import _lib
    l0 = _lib.launch_count()
    tr.warmup(groups[0], steps=1)
    timed_steps_graphed.launches_per_step = _lib.launch_count() - l0
    t0 = time.perf_counter()
    if loader.fixed_batches:
        for ids in groups:
            tr.capture(ids)
    torch.cuda.synchronize()
    t_capture = time.perf_counter() - t0

    def stream_of_ids():
        while True:
            for ids in loader._batches_of_epoch():
                yield ids

    it = stream_of_ids()
    tr.run([next(it) for _ in range(warmup)])
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    seq = [next(it) for _ in range(steps)]
    edges = sum(int(rp_host[int(ptr[b + 1])]) - int(rp_host[int(ptr[b])]) for ids in seq for b in ids)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    prof = os.environ.get("INCAGG_PROFILE") == "1"  # ncu --profile-from-start off: timed region only
    if prof:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    ev0.record()
    tr.run(seq)
    ev1.record()
    torch.cuda.synchronize()
    if prof:
        torch.cuda.profiler.stop()
    wall = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    return ev0.elapsed_time(ev1) / 1e3, edges, 0, 0, wall, len(tr.graphs), t_capture


def timed_steps(run, mode, warmup, steps, dist, e2e=False):
    """W warm-up + K timed training steps.  Returns (seconds, edges, h2d_bytes, d2h_bytes)."""
    from incagg_gnn_b200.train import mini_train  # noqa: F401  (same loop body, unrolled for timing)
    from incagg_gnn_b200.utils import dropout
    model, opt, loader, conf = run["model"], run["optimizer"], run["train_loader"], run["conf"]
    vr = mode == "incagg"
    model.train()
    rp_host = loader._rowptr_host

    def stream_of_batches():  # epochs back to back, collate prefetched on the loader's side stream
        while True:
            for sub in loader:
                yield sub

    batches = stream_of_batches()

    def sub_edges(sub):
        return sum(int(rp_host[o + c]) - int(rp_host[o]) for o, c in zip(sub.offset.tolist(), sub.count.tolist()))

    averager = None
    if dist is not None:
        from incagg_gnn_b200.parallel import make_grad_sync
        averager = make_grad_sync(model, opt, run["shard"], conf["grad_norm"], TRANSPORT)
    fused = getattr(averager, "fused", False)
    h2d = d2h = 0

    def one_step(sub):
        nonlocal h2d, d2h
        batch, B, n_id, offset, count = sub
        x, adj_t = batch.x, batch.adj_t
        y, m = batch.y[:B], batch.train_mask[:B]
        if vr:
            out = model.VR_call(x, adj_t, B, n_id, offset, count)["out"]
        else:
            out = model(x, adj_t, B, n_id, offset, count)["out"]
        if averager is not None:
            averager.zero()
        else:
            opt.zero_grad(set_to_none=True)
        w = m.to(out.dtype)
        loss = (torch.nn.functional.cross_entropy(out, y, reduction="none") * w).sum() / w.sum().clamp(min=1.)
        loss.backward()
        if fused:
            averager.step()  # peer-memory all-reduce + Adam, one kernel (parallel.FusedGradSync)
        else:
            if averager is not None:
                averager()  # NCCL all-reduce of the flat gradient buffer (p.grad are views into it)
            if conf["grad_norm"] is not None:
                torch.nn.utils.clip_grad_norm_(model.parameters(), conf["grad_norm"])
            opt.step()
        if e2e:
            lv = float(loss)  # device -> host read of the step's result
            d2h += 4
            # bytes that crossed the host<->device boundary this step (counted from the tensors)
            h2d += x.numel() * 4 + y.numel() * 8 + m.numel() + adj_t.nnz() * 8 + (B + 1) * 8
            D = model.histories[0].embedding_dim
            H = n_id.numel() - B
            L = model.num_layers
            if vr:
                h2d += 2 * L * B * D * 4
            else:
                n_pull = len(model._gas_pull_histories())
                h2d += n_pull * H * D * 4
                d2h += n_pull * B * D * 4
            return lv
        return None

    for _ in range(warmup):
        one_step(next(batches))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    h2d = d2h = 0
    edges = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        sub = next(batches)   # the collate of this batch is part of the timed step
        edges += sub_edges(sub)
        one_step(sub)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    sec = ev0.elapsed_time(ev1) / 1e3
    return sec, edges, h2d, d2h, wall


def spmm_roofline(run, peaks):
    """Dominant kernel = spmm_rows_kernel (forward aggregation of one batch).  Timed alone with CUDA
    events on the launching stream, L2 flushed (256 MB write) before every launch, over 20 different
    partition batches.  Algorithmic bytes per launch: SURVEY.md §8d (int32 indices)."""
    from incagg_gnn_b200 import ops
    loader, model = run["train_loader"], run["model"]
    F = model.hidden_channels
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    tot_b = tot_t = tot_g = 0.0
    parts = loader._parts[:20]  # this rank's own partitions
    n = len(parts)
    from incagg_gnn_b200.sparse import SparseTensor
    for b in parts:
        # relabel directly (no halo-plan exchange: rank 0 measures alone)
        lo, hi = int(loader.ptr[b]), int(loader.ptr[b + 1])
        idx = torch.arange(lo, hi, device="cuda")
        rp, col, val, n_id = ops.relabel_one_hop(loader._rowptr64, loader._col, loader._val, idx, True,
                                                 ws=loader._ws, out_int32=True)
        adj = SparseTensor(rowptr=rp, col=col, value=val, sparse_sizes=(hi - lo, n_id.numel()), is_sorted=True)
        x = torch.randn(adj.size(1), F, device="cuda")
        out = torch.empty(adj.size(0), F, device="cuda")
        plan = adj.plan()  # built once per structure, as in the training step
        ops.spmm_raw(adj.rowptr, adj.col, adj.value, x, "sum", out=out, plan=plan)  # warm the code path
        flush.fill_(b & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.spmm_raw(adj.rowptr, adj.col, adj.value, x, "sum", out=out, plan=plan)
        e1.record()
        torch.cuda.synchronize()
        rows, nnz, rsrc = adj.size(0), adj.nnz(), adj.size(1)
        tot_b += nnz * 8 + (rows + 1) * 4 + rsrc * F * 4 + rows * F * 4
        tot_g += nnz * F * 4
        tot_t += e0.elapsed_time(e1) / 1e3
    achieved = tot_b / tot_t / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"bound": "hbm", "kernel": "spmm_kernel<SUM,F=%d> fwd, one products batch, cold L2" % F,
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
            "bytes_per_launch": int(tot_b / n), "us_per_launch": round(tot_t / n * 1e6, 2),
            # dram__bytes_read + dram__bytes_write per launch, ncu --set full of this kernel on these batches
            # (profiles/r02_step_kernels_ncu_full.txt): 0.82 x the algorithmic bytes, the output stays in L2
            "traffic": 36.6e6,
            # what the counters show (profiles/r02_spmm_variants.md): every edge pulls an F*4-byte row
            # through L2 (bytes below; floor = that traffic at the ~12.4 TB/s the L2 -> SM path sustains);
            # no memory unit is above a third of its peak, the launch is paced by the instruction stream of
            # the gather loop at the occupancy its registers allow
            "l2_to_sm": {"bytes_per_launch": int(tot_g / n), "achieved_GBps": round(tot_g / tot_t / 1e9, 1),
                         "floor_us_at_12400_GBps": round(tot_g / n / 12.4e12 * 1e6, 2)},
            "limiter": "instruction issue of the gather loop (ncu: issue slots 36-51 % busy at 40 warps/SM, "
                       "L2 26-32 %, DRAM 12-14 % of peak)"}


def release_memory():
    """Between the legs of the run: destroy what the previous leg left behind (trainer <-> model cycles
    keep CUDA graphs and pinned buffers alive until a garbage collection) and hand cached device and
    pinned-host blocks back while no graph capture is in progress."""
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()
    except Exception as exc:  # a failed cudaFreeHost is not sticky; say so and go on
        print(f"[bench] pinned-host cache release: {str(exc).splitlines()[0]}", file=sys.stderr)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


_REAL_STDOUT = None
NO_PIPELINE = False   # --no-pipeline: collate inside the step graph instead of one step ahead
TRANSPORT = "p2p"


def _claim_stdout():
    """stdout carries exactly one JSON line: libraries that write to fd 1 on their own (NCCL prints its
    version banner there) are pointed at stderr for the rest of the run."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global NO_PIPELINE, TRANSPORT
    args = parse_args()
    NO_PIPELINE = args.no_pipeline
    TRANSPORT = args.transport
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod

    import incagg_gnn_b200  # noqa: F401
    from incagg_gnn_b200 import _lib
    from incagg_gnn_b200.train import build, mini_test

    vr = args.mode == "incagg"
    # one products-shaped graph (same seed on every rank); partitions, history rows and batches are
    # sharded over the ranks.  A step at N GPUs = N batches in flight (one per rank) + gradient
    # all-reduce, so per-GPU work per step is fixed ("weak") and `value` is the whole-job edges/s.
    run = build(args.config, device=dev, seed=args.seed, scale=args.scale,
                overrides=dict(VR_update=vr), shuffle=True, rank=rank, world_size=world,
                transport=args.transport)
    model = run["model"]
    mini_test(model, run["eval_loader"], VR_update=vr)  # fill the histories (main.py:211-215), untimed
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    graphs = None
    if args.no_graphs or (world > 1 and args.transport == "nccl"):
        l0 = _lib.launch_count()
        sec, edges, _, _, wall = timed_steps(run, args.mode, args.warmup, args.steps, dist)
        launches = _lib.launch_count() - l0
    else:
        sec, edges, _, _, wall, n_graphs, t_cap = timed_steps_graphed(run, args.mode, args.warmup, args.steps, dist)
        launches = timed_steps_graphed.launches_per_step * args.steps   # counted on one eager step, see there
        # (shuffled multi-partition groups never repeat: GraphedTrainer issues those steps eagerly)
        graphs = {"captured": n_graphs, "capture_s": round(t_cap, 2)} if n_graphs else None
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks of the device time; edges summed over ranks
    if dist is not None:
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e = torch.tensor([edges], device=dev, dtype=torch.float64)
        dist.all_reduce(e)
        sec, edges = float(t), float(e)
    value = edges / sec

    # the per-epoch refresh sweep (mini_inference / mini_inference_vr over all partitions), timed alone
    refresh = None
    from incagg_gnn_b200.train import GraphedSweep
    from incagg_gnn_b200.loader import EvalSubgraphLoader

    def time_sweep(loader):
        sweep = GraphedSweep(model, loader, VR_update=vr)
        sweep()  # captures (once)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        sweep()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:  # max over ranks (the sharded sweep barriers between layer phases)
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t)
        return dt

    nnz_all = run["data"].adj_t.nnz()
    t_epoch = nnz_all / value
    if world == 1:
        t_parts = time_sweep(run["eval_loader"])
        # the same sweep with all partitions merged into ONE evaluation batch (the reference sizes its
        # eval batches for a 2021 GPU; the tables and every intermediate of a whole-graph layer fit HBM)
        n_parts = run["ptr"].numel() - 1
        merged = EvalSubgraphLoader(run["data"], run["ptr"], batch_size=n_parts, log=False, device=dev)
        t_sweep = time_sweep(merged)
        del merged
        torch.cuda.empty_cache()
        refresh = {"sweep_s": round(t_sweep, 4), "sweep_per_partition_batches_s": round(t_parts, 4),
                   "train_epoch_s": round(t_epoch, 4),
                   "edges_per_s_epoch_plus_refresh": nnz_all / (t_epoch + t_sweep),
                   "note": "value = training steps only; one epoch of the reference loop = train epoch + one "
                           "layer-wise refresh sweep over all partitions (main.py:226-236); the sweep is one "
                           "CUDA-graph replay (train.GraphedSweep); sweep_s merges all partitions into one "
                           f"evaluation batch (EvalSubgraphLoader batch_size={n_parts}), "
                           "sweep_per_partition_batches_s uses the training batch size as the reference does"}
    elif args.transport == "p2p":
        # sharded tables: every rank sweeps its own partitions, halo rows read from the peers' shards;
        # one CUDA graph per layer phase, a barrier between phases
        t_sweep = time_sweep(run["eval_loader"])
        refresh = {"sweep_s": round(t_sweep, 4), "train_epoch_s": round(t_epoch, 4),
                   "edges_per_s_epoch_plus_refresh": nnz_all / (t_epoch + t_sweep),
                   "note": f"sharded sweep over {world} ranks (each rank its own partitions, per-partition "
                           "evaluation batches), one CUDA graph per layer phase + barrier, max over ranks"}

    release_memory()

    e2e = None
    if not args.no_e2e:
        # End to end through the public API with HOST buffers: the graph (CSR), features, labels and
        # masks live in pinned host memory and are read by the collate kernels through UVA every step
        # (host->device traffic inside the timed region); the loss accumulator is read back to the
        # host after every step (device->host + a synchronisation per step).  History tables stay in
        # HBM (the product's layout); `e2e_host_histories` below is the same loop with the reference's
        # all-host layout (pinned history tables + AsyncIOPool staging).
        from incagg_gnn_b200.train import GraphedTrainer
        data_pack = (run["data"], run["ptr"], run["in_channels"], run["out_channels"])
        run_h = build(args.config, device=dev, seed=args.seed, scale=args.scale,
                      overrides=dict(VR_update=vr), shuffle=True, host_resident=True,
                      history_device="cuda", data=data_pack, rank=rank, world_size=world,
                      transport=args.transport)
        run_h["model"].load_state_dict(model.state_dict(), strict=False)
        mini_test(run_h["model"], run_h["eval_loader"], VR_update=vr)
        ld = run_h["train_loader"]
        averager_h = None
        if dist is not None:
            from incagg_gnn_b200.parallel import make_grad_sync
            averager_h = make_grad_sync(run_h["model"], run_h["optimizer"], run_h["shard"],
                                        run_h["conf"]["grad_norm"], args.transport)
        tr = GraphedTrainer(run_h["model"], ld, run_h["optimizer"], VR_update=vr,
                            grad_norm=run_h["conf"]["grad_norm"], averager=averager_h, pipeline_collate=True)
        groups = ld._batches_of_epoch()
        tr.warmup(groups[0])
        for ids in groups:
            tr.capture(ids)
        torch.cuda.synchronize()
        k = args.steps
        order = [g for _ in range(k // len(groups) + 2) for g in ld._batches_of_epoch()]
        tr.run(order[:min(args.warmup, 5)])
        torch.cuda.synchronize()
        seq = order[5:5 + k]
        rp_host, ptr_h = ld._rowptr_host, ld.ptr
        ed2 = h2d = 0
        F_in = run_h["in_channels"]
        for ids in seq:
            for b_ in ids:
                lo_, hi_ = int(ptr_h[b_]), int(ptr_h[b_ + 1])
                nnz_ = int(rp_host[hi_]) - int(rp_host[lo_])
                ed2 += nnz_
                n_rows = (hi_ - lo_) + ld._known_sizes[(("ib",) if vr else ("gas",)) + tuple(ids)] * (0 if vr else 1)
                # CSR rows (int64 rowptr + int32 col + fp32 val) + gathered x rows + y + mask
                h2d += (hi_ - lo_ + 1) * 8 + nnz_ * 8 + n_rows * (F_in * 4 + 8 + 1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the result of EVERY step (running loss sum, count) is copied to pinned host memory and read
        # by the host; the read of step i-1 overlaps the replay of step i (double-buffered, event-synced)
        host_res = [torch.zeros(2, dtype=torch.float64).pin_memory() for _ in range(2)]
        res_ev = [torch.cuda.Event(), torch.cuda.Event()]
        results = []
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        def read_back(i):
            host_res[i & 1].copy_(tr.acc, non_blocking=True)
            res_ev[i & 1].record()
            if i > 0:
                res_ev[(i - 1) & 1].synchronize()
                results.append(float(host_res[(i - 1) & 1][0]))

        tr.run(seq, after_step=read_back)
        res_ev[(k - 1) & 1].synchronize()
        results.append(float(host_res[(k - 1) & 1][0]))
        ev1.record()
        torch.cuda.synchronize()
        assert len(results) == k and all(r == r for r in results)
        s2 = ev0.elapsed_time(ev1) / 1e3
        if dist is not None:  # max over ranks of the device time, edges / bytes summed over ranks
            dist.barrier()
            t = torch.tensor([s2], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e = torch.tensor([ed2, h2d], device=dev, dtype=torch.float64)
            dist.all_reduce(e)
            s2, ed2, h2d = float(t), float(e[0]), float(e[1])
        e2e = {"value": ed2 / s2, "unit": "edges/s", "h2d_bytes_per_step": int(h2d / k),
               "d2h_bytes_per_step": 16 * world, "steps": k, "ms_per_step": s2 / k * 1e3,
               "layout": "graph (CSR), features, labels and masks in pinned host memory, read through UVA by the "
                         "collate kernels every step; history tables HBM-resident; every step's result (loss "
                         "sum, count) copied to pinned memory and read by the host, the read of step i-1 "
                         "overlapping step i; CUDA-graph replay per batch, the collate graph of step i+1 (the host-memory "
                         "reads) replayed on a side stream while step i computes"}
        del tr, run_h, ld
        release_memory()
    if e2e is not None and world == 1:
        # the reference's all-host layout: history tables in pinned host memory too
        run_h = build(args.config, device=dev, seed=args.seed, scale=args.scale,
                      overrides=dict(VR_update=vr), shuffle=True, host_resident=True,
                      history_device=None, data=data_pack)
        run_h["model"].load_state_dict(model.state_dict(), strict=False)
        mini_test(run_h["model"], run_h["eval_loader"], VR_update=vr)
        torch.cuda.synchronize()
        k2 = min(args.steps, 40)
        # (a) the reference's own protocol: AsyncIOPool slots, steps issued eagerly
        s3, ed3, h2d3, d2h3, _ = timed_steps(run_h, args.mode, min(args.warmup, 5), k2, dist, e2e=True)
        e2e["host_histories"] = {"value": ed3 / s3, "unit": "edges/s", "h2d_bytes_per_step": int(h2d3 / k2),
                                 "d2h_bytes_per_step": int(d2h3 / k2), "steps": k2,
                                 "layout": "reference layout: all history tables in pinned host memory too, "
                                           "AsyncIOPool staging, eager issue"}
        # (b) the same tables, the step as CUDA-graph replays: halo rows gathered out of host memory
        # through UVA, pushes as DMA slice copies; the pulls of step i+1 ride in its collate graph, i.e.
        # they cross PCIe while step i computes (GAS mode; rows pushed by the step in flight are read one
        # step staler than in the sequential loop)
        release_memory()
        try:
            if not run_h["train_loader"].fixed_batches:
                raise RuntimeError("batches are not fixed (shuffled multi-partition groups): steps are eager")
            ld = run_h["train_loader"]
            trh = GraphedTrainer(run_h["model"], ld, run_h["optimizer"], VR_update=vr,
                                 grad_norm=run_h["conf"]["grad_norm"], pipeline_collate=True, host_prefetch=True)
            groups = ld._batches_of_epoch()
            trh.warmup(groups[0])
            for ids in groups:
                trh.capture(ids)
            torch.cuda.synchronize()
            order = [g for _ in range(k2 // len(groups) + 2) for g in ld._batches_of_epoch()]
            trh.run(order[:5])
            torch.cuda.synchronize()
            seq = order[5:5 + k2]
            ed4 = sum(int(ld._rowptr_host[int(ld.ptr[b + 1])]) - int(ld._rowptr_host[int(ld.ptr[b])])
                      for ids in seq for b in ids)
            host_res = torch.zeros(2, dtype=torch.float64).pin_memory()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            trh.run(seq, after_step=lambda i: host_res.copy_(trh.acc, non_blocking=True))
            e1.record()
            torch.cuda.synchronize()
            s4 = e0.elapsed_time(e1) / 1e3
            assert float(host_res[0]) == float(host_res[0])
            e2e["host_histories_graphed"] = {
                "value": ed4 / s4, "unit": "edges/s", "steps": k2, "ms_per_step": s4 / k2 * 1e3,
                "h2d_bytes_per_step": int(h2d3 / k2), "d2h_bytes_per_step": int(d2h3 / k2),
                "layout": "all history tables in pinned host memory; CUDA-graph replay per batch, halo rows "
                          "gathered from host memory through UVA one step ahead (in the collate graph of the "
                          "next batch), pushes as DMA slice copies inside the step graph"}
            del trh
        except Exception as exc:
            e2e["host_histories_graphed"] = {"error": str(exc).splitlines()[0]}
        del run_h
        release_memory()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    roof = spmm_roofline(run, peaks)
    cpu = None
    if not args.no_cpu_baseline and world == 1 and (args.config == "C3" or args.cpu_steps_given):
        # (rank 0 at N = 1 only: at N > 1 the other ranks' processes wait in a barrier on the same cores)
        # (the other configs step over half a graph of 10^8 edges per batch: minutes per CPU step, so
        # their CPU leg runs only when --cpu-steps is given)
        threads = os.cpu_count() or 1
        v, ed, tt = cpu_baseline(run, args.mode, args.cpu_steps, threads)
        cpu = {"value": v, "unit": "edges/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} training steps (partitions 1..{args.cpu_steps}) of the same "
                         f"graph/weights, {ed} edges in {tt:.1f} s, oracle/gas.py fp32 (C relabel port, ATen "
                         f"CSR SpMM forward and transposed-CSR backward)"}
    conf = run["conf"]
    line = {
        "metric": metric_text(conf), "value": value, "unit": "edges/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.config, conf["model"], args.mode, conf["dataset"],
                                             run["data"].num_nodes, run["data"].adj_t.nnz(),
                                             conf["num_parts"], conf["batch_size"]),
                   "mode": args.mode, "scale": args.scale,
                   "l2": "inputs larger than L2: every step reads a different partition (graph + features "
                         "+ 10 history tables = 14 GB per epoch)",
                   "histories": "HBM-resident" + ("" if world == 1 else
                                                  f", sharded by partition over {world} ranks; halo rows via "
                                                  + ("NVLink peer loads inside the gather kernel (CUDA IPC)"
                                                     if args.transport == "p2p" else "NCCL all-to-all-v")
                                                  + (", gradients exchanged through NVLink peer memory inside the "
                                                     "Adam kernel (one graph per step, no collective call)"
                                                     if args.transport == "p2p" and conf["grad_norm"] is None
                                                     and os.environ.get("INCAGG_FUSED_ALLREDUCE", "1") != "0"
                                                     else ", gradients NCCL all-reduce")),
                   "step_issue": ("eager (Python launches)" if graphs is None else
                                  f"CUDA graph per partition batch ({graphs['captured']} graphs captured before the "
                                  f"timed region in {graphs['capture_s']} s; every replay re-runs collate, forward, "
                                  f"history push/pull, backward and Adam"
                                  + ("" if NO_PIPELINE else "; the collate graph of step i+1 is replayed on a side "
                                     "stream while step i computes") + ")")},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        "cpu_baseline": cpu, "edges_timed": edges, "wall_s": wall, "refresh": refresh,
    }
    _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def reference_arm(args, rank, world):
    """The reference's CPU implementation of the path on the host cores.  The reference package itself
    cannot be imported (torch_sparse / torch_geometric absent) and has no CPU-only execution path
    (SURVEY F6g), so this arm times: the reference's OWN compiled relabel op (oracle/_ref/ref_relabel.so,
    built from /root/reference/csrc by oracle/build_ref.sh; the C restatement when that file is absent)
    + the oracle port of the training step (oracle/gas.py) with the aggregation running through ATen's
    multi-threaded CSR kernels (`torch.sparse_csr_tensor @ X`, backward through the cached transposed
    CSR, as torch_sparse does), with all host threads, on the same config, metric and unit.  Nothing of
    the product package is imported: inputs come from oracle/synth.py (same recipe and seeds), the
    hyper-parameters from the YAML table.  Rank 0 only."""
    if rank != 0:
        return
    import yaml
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import gas, synth, relabel as orl
    assert "incagg_gnn_b200" not in sys.modules
    gas.SPMM_IMPL = "csr"
    with open(os.path.join(ROOT, "incagg_gnn_b200", "conf", "configs.yaml")) as f:
        conf = yaml.safe_load(f)[args.config]
    gen_dev = "cuda" if torch.cuda.is_available() else "cpu"   # same generator device as the product arm
    inp = synth.make_inputs(conf["dataset"], seed=args.seed, scale=args.scale, num_parts=conf["num_parts"],
                            device=gen_dev)
    fin, fout, ptr = inp.num_features, inp.num_classes, inp.ptr
    N = inp.rowptr.numel() - 1
    adj = gas.Adj(inp.rowptr, inp.col, None, N, N)
    if conf["loop"]:
        adj = gas.set_diag(adj)
    if conf["norm"]:
        adj = gas.gcn_norm(adj)
    kind = "port"
    relabel_fn = None
    try:
        one_hop, within = orl.ref_ops_in_process()
        relabel_fn = within if args.mode == "incagg" else one_hop
        kind = "reference"
    except Exception as exc:  # prebuilt .so absent or not loadable with this torch: C restatement
        print(f"[reference arm] reference relabel op unavailable ({exc}); using the C port", file=sys.stderr)
    a = conf["architecture"]
    H, L = a["hidden_channels"], a["num_layers"]
    g = torch.Generator().manual_seed(args.seed)

    def glorot(o, i):
        s = (6.0 / (o + i)) ** 0.5
        return (torch.rand(o, i, generator=g) * 2 - 1) * s

    st = {"lins.0.weight": glorot(H, fin), "lins.0.bias": torch.zeros(H),
          "lins.1.weight": glorot(fout, H), "lins.1.bias": torch.zeros(fout)}
    if conf["model"] == "GCN2":
        for l in range(L):
            st[f"convs.{l}.weight1"] = glorot(H, H)
            st[f"convs.{l}.weight2"] = glorot(H, H)
    elif conf["model"] == "GCN":
        st = {}
        dims = [fin] + [H] * (L - 1) + [fout]
        for l in range(L):
            st[f"convs.{l}.lin.weight"] = glorot(dims[l + 1], dims[l])
            st[f"convs.{l}.bias"] = torch.zeros(dims[l + 1])
    elif conf["model"] == "GraphSAGE":
        st = {}
        dims = [fin] + [H] * (L - 1) + [fout]
        for l in range(L):
            st[f"convs.{l}.lin_l.weight"] = glorot(dims[l + 1], dims[l])
            st[f"convs.{l}.lin_l.bias"] = torch.zeros(dims[l + 1])
            st[f"convs.{l}.lin_r.weight"] = glorot(dims[l + 1], dims[l])
    elif conf["model"] == "APPNP":
        pass  # lins.0 / lins.1 only
    elif conf["model"] == "PNA":
        st = {}
        K = len(a["aggregators"]) * len(a["scalers"])
        dims = [fin] + [H] * (L - 1) + [fout]
        for l in range(L):
            for k in range(K):
                st[f"convs.{l}.pre_lins.{k}.weight"] = glorot(dims[l + 1], dims[l])
                st[f"convs.{l}.pre_lins.{k}.bias"] = torch.zeros(dims[l + 1])
                st[f"convs.{l}.post_lins.{k}.weight"] = glorot(dims[l + 1], dims[l + 1])
                st[f"convs.{l}.post_lins.{k}.bias"] = torch.zeros(dims[l + 1])
            st[f"convs.{l}.lin.weight"] = glorot(dims[l + 1], dims[l])
            st[f"convs.{l}.lin.bias"] = torch.zeros(dims[l + 1])
    else:
        raise SystemExit(f"reference arm: no weight initialiser for model {conf['model']}")
    okw = dict(a)
    if conf["model"] == "PNA":
        okw["deg"] = adj.rowptr[1:] - adj.rowptr[:-1]
    model = gas.OracleGNN(conf["model"], st, N, fin, out_channels=fout, dtype=torch.float32, **okw)
    opt = torch.optim.Adam(model.parameters(), lr=conf["lr"])
    vr = args.mode == "incagg"
    bs = conf["batch_size"]
    P = conf["num_parts"]

    def step(s):
        group = [(s * bs + j) % P for j in range(bs)]
        b = gas.collate(adj, inp.x, inp.y, inp.train_mask, ptr, group, within_batch=vr, relabel_fn=relabel_fn)
        gas.train_epoch(model, [b], opt, vr=vr, grad_norm=conf["grad_norm"])
        return sum(int(adj.rowptr[int(ptr[p + 1])] - adj.rowptr[int(ptr[p])]) for p in group)

    # bounded sample: K timed steps after W warm-up steps, both capped so the arm ends within minutes
    steps, warm = min(args.steps, 150), min(args.warmup, 20)
    for s in range(warm):
        step(s)
    edges, t0 = 0, time.perf_counter()
    for s in range(steps):
        edges += step(warm + s)
    sec = time.perf_counter() - t0
    v = edges / sec
    line = {
        "impl": "reference", "metric": metric_text(conf), "value": v,
        "unit": "edges/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.config, conf["model"], args.mode, conf["dataset"], N,
                                             int(adj.col.numel()), P, bs),
                   "mode": args.mode, "scale": args.scale,
                   "arm": "CPU: reference's compiled relabel op + oracle port of the step, ATen CSR SpMM, "
                          f"{threads} threads; {steps} steps sampled"},
        "cpu_baseline": {"value": v, "unit": "edges/s", "cores": threads, "kind": kind,
                         "sample": f"{steps} training steps after {warm} warm-up steps, {edges} edges in "
                                   f"{sec:.1f} s; relabel = "
                                   + ("the reference's own op (oracle/_ref/ref_relabel.so)" if kind == "reference"
                                      else "C restatement (oracle/relabel_oracle.c)")
                                   + ", step = oracle/gas.py fp32 with torch.sparse_csr_tensor @ X"},
        "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


if __name__ == "__main__":
    main()
