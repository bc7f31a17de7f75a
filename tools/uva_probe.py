#!/usr/bin/env python
"""How fast do the collate kernels read pinned host memory (UVA zero-copy) on this box?
Gather of scattered feature rows, bulk DMA of a contiguous block, relabel over a host-resident CSR."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200 as tga
from incagg_gnn_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


N, F = 2449029, 100
x_host = torch.randn(N, F).pin_memory()
x_dev = x_host.to(dev)
for rows in (16000, 70000):
    idx = torch.randint(0, N, (rows,), device=dev)
    idx_sorted = idx.sort().values
    out = torch.empty(rows, F, device=dev)
    for name, src, ix in (("uva random", x_host, idx), ("uva sorted", x_host, idx_sorted), ("hbm random", x_dev, idx)):
        us = timeit(lambda: ops.gather_rows(src, ix, out=out))
        print(json.dumps(dict(op=f"gather {rows} x {F * 4} B, {name}", us=round(us, 1), GBps=round(rows * F * 4 / us / 1e3, 1))), flush=True)
blk = x_host[1000000:1000000 + 70000]
dst = torch.empty_like(blk, device=dev)
us = timeit(lambda: dst.copy_(blk, non_blocking=True))
print(json.dumps(dict(op="bulk DMA 28 MB pinned -> HBM", us=round(us, 1), GBps=round(blk.numel() * 4 / us / 1e3, 1))), flush=True)
blk = x_host[1000000:1000000 + 16000]
dst = torch.empty_like(blk, device=dev)
us = timeit(lambda: dst.copy_(blk, non_blocking=True))
print(json.dumps(dict(op="bulk DMA 6.4 MB pinned -> HBM", us=round(us, 1), GBps=round(blk.numel() * 4 / us / 1e3, 1))), flush=True)

# relabel over host-resident vs device-resident CSR (products shape)
from incagg_gnn_b200.synthetic import SHAPES, synthetic_graph
data, ptr = synthetic_graph(SHAPES["products"][0], SHAPES["products"][1], 8, 4, 150, seed=0, device=dev)
adj = data.adj_t
rowptr, col, val = adj.csr()
val = val if val is not None else torch.ones(col.numel(), device=col.device)
rp_h, col_h, val_h = rowptr.long().cpu().pin_memory(), col.int().cpu().pin_memory(), val.float().cpu().pin_memory()
rp_d, col_d, val_d = rp_h.to(dev), col_h.to(dev), val_h.to(dev)
b = 7
idx = torch.arange(int(ptr[b]), int(ptr[b + 1]), device=dev)
for name, (r, c, v) in (("host CSR (UVA)", (rp_h, col_h, val_h)), ("device CSR", (rp_d, col_d, val_d))):
    us = timeit(lambda: ops.relabel_one_hop(r, c, v, idx, True), reps=5)
    print(json.dumps(dict(op=f"relabel_one_hop B={idx.numel()}, {name}", us=round(us, 1))), flush=True)
