#!/usr/bin/env python
"""Kernel-level profile of ONE collate with host-resident inputs (pinned CSR / features / labels /
masks read through UVA): where the host->device time of the e2e step goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200
from incagg_gnn_b200.train import build
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
run = build("C3", device=dev, seed=0, shuffle=True, host_resident=True, history_device="cuda")
loader = run["train_loader"]
groups = loader._batches_of_epoch()
for ids in groups[:3]:
    loader._collate(list(ids))
torch.cuda.synchronize()
# second visit of the same batches: sizes known, no host synchronisation inside
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for ids in groups[:3]:
        loader._collate(list(ids))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
