#!/usr/bin/env python
"""Per-kernel micro-benchmark of the hot-path kernels at the BASELINE shapes (GPU box only).

Each kernel is timed alone with CUDA events on the launching stream, after warm-up, with an L2 flush
(256 MB write) before every timed launch ("cold") and back to back ("warm").  Algorithmic bytes per
launch follow SURVEY.md §8d.  Prints one JSON line per kernel/shape; the summaries are copied into
profiles/.

    python tools/kernel_bench.py [--shape products] [--parts 150] [--F 128] [--batches 8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timeit(fn, flush, reps=5, cold=True):
    """Device time of one call: the call is captured in a CUDA graph (so multi-launch operations are
    not measured by their host launch overhead) and replayed between events; `cold` flushes L2 first."""
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for i in range(reps):
        if cold:
            flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--parts", type=int, default=None)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--F", type=int, default=128)
    ap.add_argument("--batches", type=int, default=6)
    ap.add_argument("--full", action="store_true", help="also time the full-graph SpMM")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import incagg_gnn_b200 as tga
    from incagg_gnn_b200 import ops
    peak = 6459.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda:0")
    n, e, f, c, parts = tga.SHAPES[args.shape]
    parts = args.parts or parts
    data, ptr = tga.synthetic_graph(n, e, f, c, parts, seed=0, device=dev)
    adj = tga.gcn_norm(tga.set_diag(data.adj_t))
    rowptr64 = adj.rowptr.to(torch.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    F = args.F
    ws = ops.RelabelWorkspace(n, dev)

    def emit(name, nbytes, t_cold, t_warm, **kw):
        print(json.dumps(dict(kernel=name, shape=args.shape, tag=args.tag, F=F, bytes=int(nbytes),
                              us_cold=round(t_cold * 1e6, 2), us_warm=round(t_warm * 1e6, 2),
                              GBps_cold=round(nbytes / t_cold / 1e9, 1), frac_cold=round(nbytes / t_cold / 1e9 / peak, 4),
                              GBps_warm=round(nbytes / t_warm / 1e9, 1), **kw)), flush=True)

    agg = {}
    for b in range(args.batches):
        ids = list(range(b * args.batch, (b + 1) * args.batch))
        idx = torch.cat([torch.arange(int(ptr[i]), int(ptr[i + 1]), device=dev) for i in ids])
        B = idx.numel()
        nnz_b = int(rowptr64[int(ptr[ids[-1] + 1])] - rowptr64[int(ptr[ids[0]])])
        rp, col, val, n_id = ops.relabel_one_hop(rowptr64, adj.col, adj.value, idx, True, ws=ws, out_int32=True, nnz_b=nnz_b)
        H = n_id.numel() - B
        fn = lambda: ops.relabel_one_hop(rowptr64, adj.col, adj.value, idx, True, ws=ws, out_int32=True, nnz_b=nnz_b, known=H)
        t_c, t_w = timeit(fn, flush), timeit(fn, flush, cold=False)
        by = nnz_b * (4 + 4 + 4 + 4) + B * 16 + (B + H) * 8
        agg.setdefault("relabel_one_hop", []).append((by, t_c, t_w))
        x = torch.randn(B + H, F, device=dev)
        out = torch.empty(B, F, device=dev)
        plan = ops.spmm_plan(rp, B, nnz_b)
        fn = lambda: ops.spmm_raw(rp, col, val, x, "sum", out=out, plan=plan)
        by = nnz_b * 8 + (B + 1) * 4 + (B + H) * F * 4 + B * F * 4
        agg.setdefault("spmm_sum_fwd", []).append((by, timeit(fn, flush), timeit(fn, flush, cold=False)))
        fn = lambda: ops.csr_transpose(rp, col, val, B, B + H)
        t_rp, t_col, t_val, _ = fn()
        agg.setdefault("csr_transpose", []).append((nnz_b * 16 + (2 * B + H) * 4, timeit(fn, flush), timeit(fn, flush, cold=False)))
        g = torch.randn(B, F, device=dev)
        gx = torch.empty(B + H, F, device=dev)
        t_plan = ops.spmm_plan(t_rp, B + H, nnz_b)
        fn = lambda: ops.spmm_raw(t_rp, t_col, t_val, g, "sum", out=gx, plan=t_plan)
        by = nnz_b * 8 + (B + H + 1) * 4 + B * F * 4 + (B + H) * F * 4
        agg.setdefault("spmm_sum_bwd", []).append((by, timeit(fn, flush), timeit(fn, flush, cold=False)))
        table = torch.randn(n, F, device=dev)
        halo = n_id[B:].contiguous()
        dst = torch.empty(H, F, device=dev)
        fn = lambda: ops.gather_rows(table, halo, out=dst)
        agg.setdefault("history_pull_gather", []).append((H * (8 + 2 * F * 4), timeit(fn, flush), timeit(fn, flush, cold=False)))
        off, cnt = ptr[torch.tensor(ids)], ptr[torch.tensor(ids) + 1] - ptr[torch.tensor(ids)]
        src = torch.randn(B, F, device=dev)
        fn = lambda: ops.copy_slices(src, table, off, cnt, 1)
        agg.setdefault("history_push_slices", []).append((B * 2 * F * 4, timeit(fn, flush), timeit(fn, flush, cold=False)))
        xs = torch.randn(B, F, device=dev)
        rpb, colb, valb, _ = ops.relabel_one_hop_within_batch(rowptr64, adj.col, adj.value, idx, True, ws=ws, out_int32=True, nnz_b=nnz_b)
        kept = colb.numel()
        fnw = lambda: ops.relabel_one_hop_within_batch(rowptr64, adj.col, adj.value, idx, True, ws=ws, out_int32=True, nnz_b=nnz_b, known=kept)
        agg.setdefault("relabel_within_batch", []).append((nnz_b * 8 + kept * 8 + B * 16, timeit(fnw, flush), timeit(fnw, flush, cold=False)))
        m_in, m_ag = torch.randn(n, F, device=dev), torch.randn(n, F, device=dev)
        o0 = int(off[0])
        planb = ops.spmm_plan(rpb, B, colb.numel())
        fn = lambda: ops.spmm_delta_raw(rpb, colb, valb, xs, m_in[o0:o0 + B], m_ag[o0:o0 + B], None, "sum", out=out, plan=planb)
        by = colb.numel() * 8 + (B + 1) * 4 + 4 * B * F * 4
        agg.setdefault("incagg_delta", []).append((by, timeit(fn, flush), timeit(fn, flush, cold=False)))
        feat = data.x
        fn = lambda: ops.gather_rows(feat, n_id)
        agg.setdefault("feature_gather", []).append(((B + H) * (8 + 2 * feat.size(1) * 4), timeit(fn, flush), timeit(fn, flush, cold=False)))
        del table, m_in, m_ag
    for name, v in agg.items():
        by = sum(a for a, _, _ in v) / len(v)
        emit(name, by, sum(b for _, b, _ in v) / len(v), sum(c for _, _, c in v) / len(v), launches_avg_over=len(v))
    if args.full:
        x = torch.randn(n, F, device=dev)
        out = torch.empty(n, F, device=dev)
        fplan = ops.spmm_plan(adj.rowptr, n, adj.nnz())
        fn = lambda: ops.spmm_raw(adj.rowptr, adj.col, adj.value, x, "sum", out=out, plan=fplan)
        by = adj.nnz() * 8 + (n + 1) * 4 + 2 * n * F * 4
        emit("spmm_sum_full_graph", by, timeit(fn, flush, reps=3), timeit(fn, flush, reps=3, cold=False), nnz=adj.nnz())


if __name__ == "__main__":
    main()
