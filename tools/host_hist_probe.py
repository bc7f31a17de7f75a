#!/usr/bin/env python
"""ms/step of the reference's all-host layout (pinned history tables + AsyncIOPool), eager issue."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200
from incagg_gnn_b200.train import build, mini_test, train_step

dev = torch.device("cuda:0")
run = build("C3", device=dev, seed=0, shuffle=True, host_resident=True, history_device=None)
model, loader, opt, conf = run["model"], run["train_loader"], run["optimizer"], run["conf"]
mini_test(model, run["eval_loader"], VR_update=False)
model.train()
it = iter(loader)
for _ in range(5):
    train_step(model, next(it), opt, False, conf["grad_norm"])
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 30
for _ in range(n):
    ln, _ = train_step(model, next(it), opt, False, conf["grad_norm"])
    float(ln)
torch.cuda.synchronize()
print(f"host-histories eager: {(time.perf_counter() - t0) / n * 1e3:.2f} ms/step")
if len(sys.argv) > 1:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            ln, _ = train_step(model, next(it), opt, False, conf["grad_norm"])
            float(ln)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
