#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 3xTF32 GEMM at the GCNII / products shapes vs torch (cuBLAS fp32)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import incagg_gnn_b200
from incagg_gnn_b200 import ops

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, reps=20):
    """Device time per call: `reps` calls captured in one CUDA graph (no host launch overhead)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


shapes = [("fwd x.W1 (NN, Cin)", 16384, 128, 128, False, False, True),
          ("fwd lins0 (NT, relu)", 87000, 128, 100, False, True, False),
          ("bwd g.W^T (NT)", 16384, 128, 128, False, True, True),
          ("bwd dW = h^T g (TN, split-K)", 128, 128, 16384, True, False, False),
          ("bwd dW lins0 (TN, split-K)", 128, 100, 87000, True, False, False),
          ("head (NT, N=47)", 16384, 47, 128, False, True, False)]
for name, M, N, K, ta, tb, use_cin in shapes:
    A = torch.randn(K, M, device=dev) if ta else torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev) if tb else torch.randn(K, N, device=dev)
    cin = torch.randn(M, N, device=dev) if use_cin else None
    out = torch.empty(M, N, device=dev)
    t_ours = timeit(lambda: ops.gemm(A, B, trans_a=ta, trans_b=tb, alpha=0.5, cin=cin, beta=0.5 if use_cin else 0., out=out))
    Am = A.t() if ta else A
    Bm = B.t() if tb else B
    t_torch = timeit(lambda: torch.mm(Am, Bm, out=out))
    flops = 2.0 * M * N * K
    print(json.dumps(dict(gemm=name, M=M, N=N, K=K, us_ours=round(t_ours, 2), us_cublas_fp32=round(t_torch, 2),
                          tflops_ours=round(flops / t_ours / 1e6, 1), tflops_cublas=round(flops / t_torch / 1e6, 1))), flush=True)
