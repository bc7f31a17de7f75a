#!/usr/bin/env python
"""SpMM kernel variants on BASELINE-shaped batches (GPU box only): the row-per-warp kernel against the
merge-path (edge stream) kernel and its ring / occupancy variants.

Every timed launch is bracketed by CUDA events on the launching stream after an L2 flush (256 MB write,
"cold") and back to back ("warm"); median over the batches.  Algorithmic bytes per launch follow
SURVEY.md §8d (int32 indices).  One JSON line per (case, variant).

    python tools/spmm_bench.py [--shape products] [--batches 12] [--cases fwd,bwd,delta,full]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

VARIANTS = {"rows": -1, "s8x2": 0, "s4x4": 1, "s4x3": 2, "s2x5": 3, "s2x6": 4, "s16x1": 5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--F", type=int, default=128)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--batch-parts", type=int, default=1)
    ap.add_argument("--cases", default="fwd,bwd,delta,full")
    ap.add_argument("--variants", default="rows,s8x2,s4x3")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import incagg_gnn_b200 as tga
    from incagg_gnn_b200 import ops
    from incagg_gnn_b200.sparse import SparseTensor
    peak = 6459.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda:0")
    n, e, f, c, parts = tga.SHAPES[args.shape]
    data, ptr = tga.synthetic_graph(n, e, f, c, parts, seed=0, device=dev)
    adj = tga.gcn_norm(tga.set_diag(data.adj_t))
    rowptr64 = adj.rowptr.to(torch.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ws = ops.RelabelWorkspace(n, dev)
    F = args.F
    cases = args.cases.split(",")
    variants = args.variants.split(",")

    def time_call(fn, cold):
        ts = []
        for i in range(args.reps):
            if cold:
                flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    results = {}

    def run_case(name, make, nbytes_fn):
        """make(variant) -> (callable, out tensor getter); timed for every variant, outputs compared."""
        ref = None
        for vn in variants:
            ops.tune("spmm_stream_variant", VARIANTS[vn])
            fn, get_out = make()
            fn()
            fn()
            torch.cuda.synchronize()
            o1 = get_out().clone()
            fn()
            torch.cuda.synchronize()
            det = bool(torch.equal(o1, get_out()))
            if ref is None:
                ref = o1
                err = 0.0
            else:
                err = float((o1 - ref).abs().max() / ref.abs().max().clamp(min=1e-30))
            tc, tw = time_call(fn, True), time_call(fn, False)
            results.setdefault((name, vn), []).append((tc, tw, nbytes_fn(), err, det))

    n_b = min(args.batches, parts // args.batch_parts)
    for b in range(n_b):
        lo, hi = int(ptr[b * args.batch_parts]), int(ptr[(b + 1) * args.batch_parts])
        idx = torch.arange(lo, hi, device=dev)
        rp, col, val, n_id = ops.relabel_one_hop(rowptr64, adj.col, adj.value, idx, True, ws=ws, out_int32=True)
        B, R = hi - lo, n_id.numel()
        a = SparseTensor(rowptr=rp, col=col, value=val, sparse_sizes=(B, R), is_sorted=True)
        nnz = a.nnz()
        x = torch.randn(R, F, device=dev)
        g = torch.randn(B, F, device=dev)
        if "fwd" in cases:
            out = torch.empty(B, F, device=dev)

            def make():
                a.drop_caches()
                plan = a.plan()
                return (lambda: ops.spmm_raw(a.rowptr, a.col, a.value, x, "sum", out=out, plan=plan)), (lambda: out)
            run_case("fwd", make, lambda: nnz * 8 + (B + 1) * 4 + R * F * 4 + B * F * 4)
        if "bwd" in cases:  # transposed product over the in-batch source rows, ReLU gate in the epilogue
            gx = torch.empty(R, F, device=dev)
            t_rowptr, t_col, t_val = a.t_csr()
            nnz_p = int(t_rowptr[B])

            def make():
                a.__dict__.pop('_t_prefix_plans', None)
                plan = a.t_plan_prefix(B)
                return (lambda: ops.spmm_raw(t_rowptr, t_col, t_val, g, "sum", rows=B, out=gx[:B], plan=plan,
                                             gate=x)), (lambda: gx[:B])
            run_case("bwd_prefix_gated", make, lambda: nnz_p * 8 + (B + 1) * 4 + 3 * B * F * 4)
        if "delta" in cases:  # IncAgg: A_BB (x - M_in) + M_ag
            rpw, colw, valw, _ = ops.relabel_one_hop_within_batch(rowptr64, adj.col, adj.value, idx, True, ws=ws,
                                                                  out_int32=True)
            aw = SparseTensor(rowptr=rpw, col=colw, value=valw, sparse_sizes=(B, B), is_sorted=True)
            m_in, m_ag = torch.randn(B, F, device=dev), torch.randn(B, F, device=dev)
            outd = torch.empty(B, F, device=dev)
            xb = x[:B].contiguous()

            def make():
                aw.drop_caches()
                plan = aw.plan()
                return (lambda: ops.spmm_delta_raw(aw.rowptr, aw.col, aw.value, xb, m_in, m_ag, None, "sum",
                                                   out=outd, plan=plan)), (lambda: outd)
            run_case("delta", make, lambda: aw.nnz() * 8 + (B + 1) * 4 + 4 * B * F * 4)
        del a, x, g
    if "full" in cases:
        a = SparseTensor(rowptr=adj.rowptr, col=adj.col, value=adj.value, sparse_sizes=(n, n), is_sorted=True)
        x = torch.randn(n, F, device=dev)
        out = torch.empty(n, F, device=dev)

        def make():
            a.drop_caches()
            plan = a.plan()
            return (lambda: ops.spmm_raw(a.rowptr, a.col, a.value, x, "sum", out=out, plan=plan)), (lambda: out)
        run_case("full_graph", make, lambda: a.nnz() * 8 + (n + 1) * 4 + 2 * n * F * 4)
    for (name, vn), rs in results.items():
        k = len(rs)
        tc = sorted(r[0] for r in rs)[k // 2]
        tw = sorted(r[1] for r in rs)[k // 2]
        nb = sum(r[2] for r in rs) / k
        print(json.dumps(dict(case=name, variant=vn, shape=args.shape, F=F, launches=k, bytes=int(nb),
                              us_cold=round(tc * 1e6, 2), us_warm=round(tw * 1e6, 2),
                              GBps_cold=round(nb / tc / 1e9, 1), frac_cold=round(nb / tc / 1e9 / peak, 4),
                              frac_warm=round(nb / tw / 1e9 / peak, 4),
                              max_rel_diff_vs_first=max(r[3] for r in rs),
                              deterministic=all(r[4] for r in rs))), flush=True)


if __name__ == "__main__":
    main()
