#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   1. the sharded layer-wise sweep (mini_inference / mini_inference_vr with halo rows fetched from the
      owning ranks) reproduces the single-GPU tables bit for bit,
   2. one sharded training epoch (GAS and IncAgg) runs in lockstep, leaves identical parameters on every
      rank and a finite loss.
   torchrun --nproc-per-node 2 tools/multi_gpu_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import incagg_gnn_b200  # noqa
from incagg_gnn_b200.train import build, mini_test
from incagg_gnn_b200.parallel import GradAverager

ok = True
import itertools
for transport, vr in itertools.product(("nccl", "p2p"), (False, True)):
    ov = dict(VR_update=vr, num_parts=12)
    sharded = build("C3", device=dev, seed=0, scale=32, overrides=ov, rank=rank, world_size=world, shuffle=False,
                    transport=transport)
    single = build("C3", device=dev, seed=0, scale=32, overrides=ov, shuffle=False)
    single["model"].load_state_dict(sharded["model"].state_dict(), strict=False)
    # identical weights on all ranks
    for p in sharded["model"].parameters():
        dist.broadcast(p.data, 0)
    single["model"].load_state_dict({k: v for k, v in sharded["model"].state_dict().items()}, strict=False)
    out_s = mini_test(sharded["model"], sharded["eval_loader"], VR_update=vr)
    out_1 = mini_test(single["model"], single["eval_loader"], VR_update=vr)
    sh = sharded["shard"]
    same = torch.equal(out_s, out_1[sh.lo:sh.hi])
    for l in range(sharded["model"].num_layers):
        same &= torch.equal(sharded["model"].histories[l].emb, single["model"].histories[l].emb[sh.lo:sh.hi])
        if vr:
            same &= torch.equal(sharded["model"].histories_ag[l].emb, single["model"].histories_ag[l].emb[sh.lo:sh.hi])
    print(f"[rank {rank}] {transport} {'IncAgg' if vr else 'GAS'} sharded sweep == single-GPU sweep: {same}", flush=True)
    ok &= same
    # one training epoch in lockstep
    model, opt = sharded["model"], sharded["optimizer"]
    avg = GradAverager(model.parameters(), sh, flat=getattr(opt, "flat_g", None))
    model.train()
    tot = 0.0
    for batch, B, n_id, offset, count in sharded["train_loader"]:
        out = (model.VR_call if vr else model)(batch.x, batch.adj_t, B, n_id, offset, count)["out"]
        m = batch.train_mask[:B]
        loss = torch.nn.functional.cross_entropy(out[m], batch.y[:B][m])
        avg.zero()
        loss.backward()
        avg()
        opt.step()
        tot += float(loss)
    del loss, out  # a live autograd graph keeps grad accumulators bound to this (legacy) stream
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same_p = torch.equal(flat, ref)
    fin = bool(torch.isfinite(torch.tensor(tot)))
    print(f"[rank {rank}] {transport} {'IncAgg' if vr else 'GAS'} epoch: loss sum {tot:.4f}, params identical across ranks: {same_p}", flush=True)
    ok &= same_p and fin
    if transport == "p2p":
        # CUDA-graph replays with the peer gathers and the NCCL gradient all-reduce captured
        from incagg_gnn_b200.train import GraphedTrainer
        tr = GraphedTrainer(model, sharded["train_loader"], opt, VR_update=vr, averager=avg)
        tr.warmup(sharded["train_loader"]._batches_of_epoch()[0], steps=1)
        r1 = tr.epoch()
        r2 = tr.epoch()
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same_g = torch.equal(flat, ref)
        print(f"[rank {rank}] p2p {'IncAgg' if vr else 'GAS'} graphed epochs: losses {r1['loss']:.4f} {r2['loss']:.4f}, "
              f"{len(tr.graphs)} graphs, params identical: {same_g}", flush=True)
        ok &= same_g and r2['loss'] == r2['loss']
    del sharded, single, model, opt
    torch.cuda.synchronize()
    dist.barrier()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if int(t) != 1:
    sys.exit(1)
if rank == 0:
    print("MULTI_GPU_CHECK OK")
