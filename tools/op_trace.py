#!/usr/bin/env python
"""torch.profiler trace of ONE eager training step (C3 by default): which framework-side aten kernels
(fills, adds, copies) are still launched next to the library's own kernels, with input shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200
from incagg_gnn_b200.train import build, train_step
from torch.profiler import profile, ProfilerActivity

config = sys.argv[1] if len(sys.argv) > 1 else "C3"
vr = len(sys.argv) > 2 and sys.argv[2] == "incagg"
dev = torch.device("cuda:0")
run = build(config, device=dev, seed=0, overrides=dict(VR_update=vr))
model, loader, opt, conf = run["model"], run["train_loader"], run["optimizer"], run["conf"]
model.train()
groups = loader._batches_of_epoch()
for ids in groups[:3]:
    train_step(model, loader._collate(list(ids)), opt, vr, conf["grad_norm"])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    train_step(model, loader._collate(list(groups[3])), opt, vr, conf["grad_norm"])
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=60,
                                                          max_shapes_column_width=70))
