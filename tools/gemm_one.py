#!/usr/bin/env python
"""Run one GEMM shape a few times (target of `ncu --set full`): gemm_one.py M N K ta tb [reps]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200
from incagg_gnn_b200 import ops

M, N, K, ta, tb = (int(v) for v in sys.argv[1:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = torch.device("cuda:0")
A = torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)
B = torch.randn(N, K, device=dev) if tb else torch.randn(K, N, device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(reps):
    ops.gemm(A, B, trans_a=bool(ta), trans_b=bool(tb), out=out)
torch.cuda.synchronize()
ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
print("max rel err", float((out.double() - ref).abs().max() / ref.abs().max()))
