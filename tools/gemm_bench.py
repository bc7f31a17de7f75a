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

# the fused GCNII pairs (ops.gemm_dual): input gradients g.[W1^T | W2^T] and weight gradients [h | x0]^T.g
M, F = 16384, 128
g = torch.randn(M, F, device=dev); h = torch.randn(M, F, device=dev); x0 = torch.randn(M, F, device=dev)
w1 = torch.randn(F, F, device=dev); w2 = torch.randn(F, F, device=dev)
o1 = torch.empty(M, F, device=dev); o2 = torch.empty(M, F, device=dev)
d1 = torch.zeros(F, F, device=dev); d2 = torch.zeros(F, F, device=dev)
t_k = timeit(lambda: ops.gemm_dual("k", h, w1, x0, w2, scale_b=0.3, scale_b2=0.2, cin=h, beta=0.5, cin2=x0, beta2=0.1, relu=True, out=o1))
t_n = timeit(lambda: ops.gemm_dual("n", g, w1, b2=w2, trans_b=True, scale_b=0.3, scale_b2=0.2, cin=g, beta=0.5, cin2=g, beta2=0.1, out=o1, out2=o2))
t_m = timeit(lambda: ops.gemm_dual("m", h, g, a2=x0, trans_a=True, alpha=0.3, alpha2=0.2, cin=d1, beta=1., cin2=d2, beta2=1., out=d1, out2=d2))
print(json.dumps(dict(gemm="GCNII pairs (16384 x 128)", us_dual_k=round(t_k, 2), us_dual_n=round(t_n, 2), us_dual_m=round(t_m, 2))), flush=True)
