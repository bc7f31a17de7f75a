"""Parity of every CUDA kernel family against the CPU oracle, called through the C ABI
(incagg_gnn_b200.ops -> ctypes -> libincagg_b200.so).

Bars: bit-exact for relabel / transpose structure / gather / scatter / slice copies / min-max
argument indices; <= 1e-5 relative (fp32, vs an fp64 oracle) for aggregation outputs.
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # the tolerance BASELINE.json's north_star states for fp32 aggregation outputs


@pytest.fixture(scope="module")
def ops(cuda):
    import incagg_gnn_b200  # noqa: F401
    from incagg_gnn_b200 import ops as _ops
    return _ops


def _rand_csr(rng, rows, cols, maxdeg, long_rows=0, long_len=0):
    deg = rng.integers(0, maxdeg + 1, rows)
    if rows > 10:
        deg[rng.integers(0, rows, rows // 6)] = 0
    for r in range(long_rows):
        deg[rng.integers(0, rows)] = long_len
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, cols, rowptr[-1]).astype(np.int64)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    return rowptr, col, val


def _dev(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)


def _rel_err(got, ref):
    return float(np.abs(got.astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


# ---- SpMM -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("F", [1, 7, 32, 40, 64, 100, 128, 130, 256, 602])
def test_spmm_matches_oracle(ops, cuda, reduce, F):
    rng = np.random.default_rng(F * 7 + len(reduce))
    rows, cols = 300, 500
    rowptr, col, val = _rand_csr(rng, rows, cols, 20)
    X = rng.standard_normal((cols, F)).astype(np.float32)
    out, arg = ops.spmm_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda),
                            _dev(X, cuda), reduce, return_arg=True)
    ref, ref_arg = oracle.spmm(rowptr, col, val, X, reduce, dtype=np.float64, return_arg=True)
    assert _rel_err(out.cpu().numpy(), ref) <= RTOL
    if reduce in ("min", "max"):
        # exact values (one multiply, no accumulation) and exact argument indices
        ref32 = oracle.spmm(rowptr, col, val, X, reduce, dtype=np.float32)
        assert np.array_equal(out.cpu().numpy(), ref32)
        assert np.array_equal(arg.cpu().numpy().astype(np.int64), ref_arg)


@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("F", [24, 128, 130])
def test_spmm_gated_epilogue(ops, cuda, reduce, F):
    """incagg_spmm_csr_gated: the plain result, zeroed where gate <= 0 - bit-identical to masking the
    ungated kernel's output afterwards (long rows and a prefix of the rows included)."""
    rng = np.random.default_rng(F + len(reduce))
    rowptr, col, val = _rand_csr(rng, 300, 500, 20, long_rows=2, long_len=3000)
    X = _dev(rng.standard_normal((500, F)).astype(np.float32), cuda)
    gate_full = _dev(rng.standard_normal((300, F + 6)).astype(np.float32), cuda)
    gate = gate_full[:, :F]                       # a view with a leading dimension wider than F
    rp, c, v = _dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda)
    plain = ops.spmm_raw(rp, c, v, X, reduce)
    gated = ops.spmm_raw(rp, c, v, X, reduce, gate=gate)
    want = torch.where(gate > 0, plain, torch.zeros_like(plain))
    # (the unaligned gate view selects the 8-byte-vector kernel, the plain call may run the merge-path
    # kernel: same zero pattern, values equal to rounding)
    assert torch.equal(gated == 0, want == 0)
    assert float((gated - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # same layout for both calls -> the same kernel -> bit-identical
    gate_al = gate.contiguous() if F % 4 == 0 else gate
    gated2 = ops.spmm_raw(rp, c, v, X, reduce, gate=gate_al)
    if F % 4 == 0:
        assert torch.equal(gated2, want)
    # first 100 rows only, into a slice of a larger buffer (the backward over the in-batch rows)
    buf = torch.full((300, F), 7.0, device=cuda)
    ops.spmm_raw(rp[:101], c, v, X, reduce, rows=100, out=buf[:100], gate=gate_al)
    plain100 = ops.spmm_raw(rp[:101], c, v, X, reduce, rows=100)
    assert torch.equal(buf[:100], torch.where(gate_al[:100] > 0, plain100, torch.zeros_like(plain100)))
    assert float(buf[100:].min()) == 7.0
    assert float((buf[:100] - want[:100]).abs().max()) <= 1e-5 * float(want.abs().max())


def test_spmm_relu_input_backward_matches_separate_mask(cuda):
    """autograd: spmm(..., relu_input=True) returns A^T g masked by [x > 0], for full and prefix rows."""
    import incagg_gnn_b200 as tga
    from incagg_gnn_b200.sparse import spmm
    g = torch.Generator(device="cpu").manual_seed(5)
    n_dst, n_src, F, B = 200, 350, 64, 120
    dense = (torch.rand(n_dst, n_src, generator=g) < 0.05).float() * torch.rand(n_dst, n_src, generator=g)
    row, col = dense.nonzero(as_tuple=True)
    adj = tga.SparseTensor(row=row.to(cuda), col=col.to(cuda), value=dense[row, col].to(cuda),
                           sparse_sizes=(n_dst, n_src), is_sorted=True)
    x0 = torch.randn(n_src, F, generator=g).to(cuda)
    go = torch.randn(n_dst, F, generator=g).to(cuda)
    for grad_rows in (None, B):
        xa = x0.clone().requires_grad_(True)
        spmm(adj, xa, grad_rows=grad_rows, relu_input=True).backward(go)
        xb = x0.clone().requires_grad_(True)
        spmm(adj, xb, grad_rows=grad_rows).backward(go)
        rows = n_src if grad_rows is None else grad_rows
        want = torch.where(x0[:rows] > 0, xb.grad[:rows], torch.zeros_like(xb.grad[:rows]))
        assert torch.equal(xa.grad[:rows], want)


def _csr_from_deg(rng, deg, cols):
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, cols, rowptr[-1]).astype(np.int64)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    return rowptr, col, val


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("F", [68, 128, 200, 604])
def test_spmm_merge_path_kernel_edge_cases(ops, cuda, variant, F):
    """The merge-path (edge stream) kernel that serves wide sum / mean products: rows cut by piece
    boundaries, one row that spans many pieces, empty rows at the front / between / at the end and at
    piece boundaries, pieces of exactly 32 edges, a structure without edges, unweighted edges, a gate,
    the delta form - against the fp64 oracle, and bit-identical from run to run (deterministic)."""
    ops.tune("spmm_stream_variant", variant)
    try:
        rng = np.random.default_rng(1000 * variant + F)
        cols = 3000
        X = rng.standard_normal((cols, F)).astype(np.float32)
        Xd = _dev(X, cuda)
        shapes = {
            "ragged": np.concatenate([[0, 0, 0], rng.integers(0, 40, 4000), [0, 0]]),
            "giant": np.concatenate([rng.integers(0, 6, 50), [150000], rng.integers(0, 6, 50), [0]]),
            "all_32": np.full(6000, 32),
            "tiny": np.array([0, 3, 0, 1, 0]),
            "one_row": np.array([70000]),
            "sparse_rows": (rng.random(20000) < 0.02).astype(np.int64) * 50,
        }
        for name, deg in shapes.items():
            rowptr, col, val = _csr_from_deg(rng, deg, cols)
            rp, c, v = _dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda)
            for reduce, vv, vd in (("sum", val, v), ("mean", None, None)):
                out = ops.spmm_raw(rp, c, vd, Xd, reduce)
                ref = oracle.spmm(rowptr, col, vv, X, reduce, dtype=np.float64)
                assert _rel_err(out.cpu().numpy(), ref) <= RTOL, (name, reduce)
                again = ops.spmm_raw(rp, c, vd, Xd, reduce)
                assert torch.equal(out, again), (name, reduce, "not deterministic")
            # gate in the epilogue = masking the plain result
            rows = len(deg)
            gate = _dev(rng.standard_normal((rows, F)).astype(np.float32), cuda)
            plain = ops.spmm_raw(rp, c, v, Xd, "sum")
            assert torch.equal(ops.spmm_raw(rp, c, v, Xd, "sum", gate=gate),
                               torch.where(gate > 0, plain, torch.zeros_like(plain))), name
        # no edges at all
        z = ops.spmm_raw(torch.zeros(41, dtype=torch.int32, device=cuda),
                         torch.zeros(0, dtype=torch.int32, device=cuda), None, Xd, "sum")
        assert z.shape == (40, F) and float(z.abs().max()) == 0.
        # delta form on a square in-batch structure
        B = 2500
        rowptr, col, val = _csr_from_deg(rng, np.concatenate([[0], rng.integers(0, 60, B - 2), [0]]), B)
        m_in = rng.standard_normal((B, F)).astype(np.float32)
        m_ag = rng.standard_normal((B, F)).astype(np.float32)
        got = ops.spmm_delta_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda),
                                 _dev(X[:B], cuda), _dev(m_in, cuda), _dev(m_ag, cuda), None, "sum")
        ref = oracle.spmm_delta(rowptr, col, val, X[:B], m_in, m_ag, "sum", dtype=np.float64)
        assert _rel_err(got.cpu().numpy(), ref) <= RTOL
    finally:
        ops.tune("spmm_stream_variant", -2)


def test_spmm_merge_path_equals_row_kernel_on_a_products_sized_batch(ops, cuda):
    """Both SpMM kernels on one full-size C3-like batch (16 K rows, ~0.4 M edges, F = 128): equal to
    rounding, each equal to the fp64 oracle within 1e-5."""
    rng = np.random.default_rng(3)
    B, R, F = 16327, 62000, 128
    deg = np.minimum(rng.zipf(1.6, B) + 8, 900)
    rowptr, col, val = _csr_from_deg(rng, deg, R)
    val = np.abs(val) * 0.1
    X = rng.standard_normal((R, F)).astype(np.float32)
    rp, c, v, Xd = _dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda), _dev(X, cuda)
    ref = oracle.spmm(rowptr, col, val, X, "sum", dtype=np.float64)
    outs = {}
    try:
        for variant in (-1, 0):
            ops.tune("spmm_stream_variant", variant)
            outs[variant] = ops.spmm_raw(rp, c, v, Xd, "sum").cpu().numpy()
            assert _rel_err(outs[variant], ref) <= RTOL
    finally:
        ops.tune("spmm_stream_variant", -2)
    assert _rel_err(outs[0], outs[-1].astype(np.float64)) <= 1e-6


@pytest.mark.parametrize("has_val", [True, False])
def test_spmm_without_values_and_empty(ops, cuda, has_val):
    rng = np.random.default_rng(11)
    rowptr, col, val = _rand_csr(rng, 64, 64, 5)
    X = rng.standard_normal((64, 24)).astype(np.float32)
    v = val if has_val else None
    out = ops.spmm_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32),
                       _dev(v, cuda) if has_val else None, _dev(X, cuda), "sum")
    assert _rel_err(out.cpu().numpy(), oracle.spmm(rowptr, col, v, X, "sum", np.float64)) <= RTOL
    # zero rows / zero edges
    z = ops.spmm_raw(torch.zeros(1, dtype=torch.int32, device=cuda), torch.zeros(0, dtype=torch.int32, device=cuda),
                     None, _dev(X, cuda), "sum")
    assert z.shape == (0, 24)
    rp0 = torch.zeros(9, dtype=torch.int32, device=cuda)
    z = ops.spmm_raw(rp0, torch.zeros(0, dtype=torch.int32, device=cuda), None, _dev(X, cuda), "max")
    assert z.shape == (8, 24) and float(z.abs().max()) == 0.


@pytest.mark.parametrize("reduce", ["sum", "max"])
@pytest.mark.parametrize("F", [40, 128, 602])
def test_spmm_long_rows_bucket(ops, cuda, reduce, F):
    """Rows above the long-row threshold go through the CTA-per-row kernel."""
    rng = np.random.default_rng(F)
    rowptr, col, val = _rand_csr(rng, 200, 4000, 30, long_rows=3, long_len=5000)
    X = rng.standard_normal((4000, F)).astype(np.float32)
    out, arg = ops.spmm_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda),
                            _dev(X, cuda), reduce, return_arg=True)
    ref, ref_arg = oracle.spmm(rowptr, col, val, X, reduce, dtype=np.float64, return_arg=True)
    assert _rel_err(out.cpu().numpy(), ref) <= RTOL
    if reduce == "max":
        assert np.array_equal(arg.cpu().numpy().astype(np.int64), ref_arg)


@pytest.mark.parametrize("reduce", ["sum", "min"])
def test_spmm_with_prebuilt_plan_is_identical(ops, cuda, reduce):
    """A cached degree-bucket plan (incagg_spmm_plan) gives bit-identical results to the temporary
    plan of a plain call, run to run (giant rows are combined in a fixed order)."""
    rng = np.random.default_rng(21)
    rowptr, col, val = _rand_csr(rng, 500, 3000, 50, long_rows=3, long_len=7000)
    X = _dev(rng.standard_normal((3000, 96)).astype(np.float32), cuda)
    rp, c, v = _dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda)
    plan = ops.spmm_plan(rp, 500, col.size)
    a = ops.spmm_raw(rp, c, v, X, reduce, plan=plan)
    for _ in range(3):
        assert torch.equal(a, ops.spmm_raw(rp, c, v, X, reduce, plan=plan))
        assert torch.equal(a, ops.spmm_raw(rp, c, v, X, reduce))
    ref = oracle.spmm(rowptr, col, val, X.cpu().numpy(), reduce, np.float64)
    assert _rel_err(a.cpu().numpy(), ref) <= RTOL


def test_spmm_unaligned_views_and_leading_dimension(ops, cuda):
    """Column-cropped history views (ld > F) and 4-byte-aligned-only bases use the narrower paths."""
    rng = np.random.default_rng(2)
    rowptr, col, val = _rand_csr(rng, 100, 100, 8)
    big = torch.from_numpy(rng.standard_normal((100, 75)).astype(np.float32)).to(cuda)
    for lo, hi in [(0, 64), (1, 65), (2, 42), (3, 10)]:
        X = big[:, lo:hi]
        out = ops.spmm_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda), X, "sum")
        ref = oracle.spmm(rowptr, col, val, X.cpu().numpy(), "sum", np.float64)
        assert _rel_err(out.cpu().numpy(), ref) <= RTOL


@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("F", [40, 128, 100])
@pytest.mark.parametrize("use_nid", [False, True])
def test_spmm_delta_matches_oracle(ops, cuda, reduce, F, use_nid):
    rng = np.random.default_rng(F + use_nid)
    B, N = 150, 400
    rowptr, col, val = _rand_csr(rng, B, B, 12)
    x = rng.standard_normal((B, F)).astype(np.float32)
    hist_in = rng.standard_normal((N, F + 8)).astype(np.float32)
    hist_ag = rng.standard_normal((N, F + 8)).astype(np.float32)
    v = val if reduce == "sum" else None
    if use_nid:
        n_id = rng.permutation(N)[:B].astype(np.int64)
        m_in, m_ag = hist_in[n_id], hist_ag[n_id]
        out = ops.spmm_delta_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32),
                                 _dev(v, cuda) if v is not None else None, _dev(x, cuda), _dev(hist_in, cuda),
                                 _dev(hist_ag, cuda), _dev(n_id, cuda), reduce)
    else:
        m_in, m_ag = hist_in[40:40 + B], hist_ag[40:40 + B]
        out = ops.spmm_delta_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32),
                                 _dev(v, cuda) if v is not None else None, _dev(x, cuda),
                                 _dev(hist_in, cuda)[40:40 + B, :F], _dev(hist_ag, cuda)[40:40 + B, :F], None, reduce)
    ref = oracle.spmm_delta(rowptr, col, v, x, m_in, m_ag, reduce, np.float64)
    assert _rel_err(out.cpu().numpy(), ref) <= RTOL


@pytest.mark.parametrize("F", [16, 64, 100])
def test_spmm_multi_matches_separate_passes(ops, cuda, F):
    rng = np.random.default_rng(F)
    rowptr, col, val = _rand_csr(rng, 120, 200, 10, long_rows=1, long_len=3000)
    reducers = ["sum", "mean", "min", "max"]
    X = rng.standard_normal((200, 4 * F)).astype(np.float32)
    out = ops.spmm_multi_raw(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda),
                             _dev(X, cuda), F, reducers)
    ref = oracle.spmm_multi(rowptr, col, val, X, F, reducers, np.float64)
    assert _rel_err(out.cpu().numpy(), ref) <= RTOL


def test_minmax_backward_routes_to_arg(ops, cuda):
    rng = np.random.default_rng(9)
    rowptr, col, val = _rand_csr(rng, 80, 60, 6)
    X = rng.standard_normal((60, 20)).astype(np.float32)
    rp, c, v = _dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32), _dev(val, cuda)
    out, arg = ops.spmm_raw(rp, c, v, _dev(X, cuda), "max", return_arg=True)
    g = rng.standard_normal((80, 20)).astype(np.float32)
    gx = ops.spmm_minmax_bwd_raw(c, v, arg, _dev(g, cuda), 60).cpu().numpy()
    _, ref_arg = oracle.spmm(rowptr, col, val, X, "max", return_arg=True)
    ref = np.zeros((60, 20))
    for i in range(80):
        for f in range(20):
            e = ref_arg[i, f]
            if e >= 0:
                ref[col[e], f] += float(val[e]) * float(g[i, f])
    assert _rel_err(gx, ref) <= RTOL


# ---- dense GEMM (tcgen05 3xTF32) -------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 128, 100), (1000, 47, 128), (77, 64, 32),
                                   (130, 200, 260), (5, 3, 7), (256, 128, 4096), (100, 128, 20000)])
@pytest.mark.parametrize("trans_a,trans_b", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32x3_matches_fp64(ops, cuda, M, N, K, trans_a, trans_b):
    """All four operand layouts, ragged sizes, split-K; fp32-level accuracy (<= 1e-5 of max |out|, the
    north_star tolerance; typically ~1e-6) although the products run on TF32 tensor cores."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(K, N, generator=g)
    a = (A.t().contiguous() if trans_a else A).to(cuda)
    b = (B.t().contiguous() if trans_b else B).to(cuda)
    out = ops.gemm(a, b, trans_a=trans_a, trans_b=trans_b)
    ref = A.double() @ B.double()
    err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err <= RTOL, err
    assert err <= 3e-6, err   # what 3xTF32 actually delivers


def test_gemm_epilogue_and_views(ops, cuda):
    g = torch.Generator().manual_seed(3)
    M, N, K = 333, 128, 128
    x = torch.randn(M, K, generator=g).to(cuda)
    w = torch.randn(N, K, generator=g).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    cin = torch.randn(M, N, generator=g).to(cuda)
    out = ops.gemm(x, w, trans_b=True, alpha=0.3, cin=cin, beta=0.7, bias=bias, relu=True)
    ref = torch.relu(0.3 * (x.double() @ w.double().t()) + 0.7 * cin.double() + bias.double())
    assert float((out.double() - ref).abs().max() / ref.abs().max()) <= RTOL
    # strided views: column-cropped A (ld > K, 16-byte aligned and not), output into a slice
    big = torch.randn(M, 200, generator=g).to(cuda)
    for lo in (0, 4, 1):
        xa = big[:, lo:lo + K]
        o = ops.gemm(xa, w, trans_b=True)
        r = xa.double() @ w.double().t()
        assert float((o.double() - r).abs().max() / r.abs().max()) <= RTOL
    # ReLU-backward gate in the epilogue: zero where gate <= 0 (full and ragged n-tiles, split-K)
    for (m_, n_, k_) in ((333, 128, 47), (200, 100, 64), (96, 128, 4096)):
        a_ = torch.randn(m_, k_, generator=g).to(cuda)
        b_ = torch.randn(k_, n_, generator=g).to(cuda)
        gate = torch.randn(m_, n_, generator=g).to(cuda)
        o = ops.gemm(a_, b_, gate=gate)
        r = (a_.double() @ b_.double()) * (gate > 0)
        assert float((o.double() - r).abs().max() / r.abs().max()) <= RTOL
        assert torch.equal(o == 0, ~(gate > 0) | (o == 0)) and float(o[gate <= 0].abs().max()) == 0.
    dst = torch.zeros(M, 256, device=cuda)
    ops.gemm(x, w, trans_b=True, out=dst[:, 128:])
    r = x.double() @ w.double().t()
    assert float((dst[:, 128:].double() - r).abs().max() / r.abs().max()) <= RTOL
    assert float(dst[:, :128].abs().max()) == 0.


def test_gemm_dual_modes(ops, cuda):
    """K- / N- / M-concatenated GEMMs (the fused GCNII layer) against fp64."""
    g = torch.Generator().manual_seed(12)
    M, C = 700, 128
    h = torch.randn(M, C, generator=g).to(cuda)
    x0 = torch.randn(M, C, generator=g).to(cuda)
    w1 = torch.randn(C, C, generator=g).to(cuda)
    w2 = torch.randn(C, C, generator=g).to(cuda)
    gr = torch.randn(M, C, generator=g).to(cuda)
    c1, c2, e1, e2 = 0.45, 0.05, 0.54, 0.06
    d = lambda t: t.double()
    rel = lambda o, r: float((d(o) - r).abs().max() / r.abs().max())
    out = ops.gemm_dual("k", h, w1, x0, w2, scale_b=c1, scale_b2=c2, cin=h, beta=e1, cin2=x0, beta2=e2, relu=True)
    ref = torch.relu(c1 * d(h) @ d(w1) + c2 * d(x0) @ d(w2) + e1 * d(h) + e2 * d(x0))
    assert rel(out, ref) <= RTOL
    gh, gx0 = ops.gemm_dual("n", gr, w1, b2=w2, trans_b=True, scale_b=c1, scale_b2=c2, cin=gr, beta=e1, cin2=gr, beta2=e2)
    assert rel(gh, c1 * d(gr) @ d(w1).t() + e1 * d(gr)) <= RTOL
    assert rel(gx0, c2 * d(gr) @ d(w2).t() + e2 * d(gr)) <= RTOL
    for rows in (700, 20000):   # without and with split-K
        hh = torch.randn(rows, C, generator=g).to(cuda)
        xx = torch.randn(rows, C, generator=g).to(cuda)
        gg = torch.randn(rows, C, generator=g).to(cuda)
        gw1, gw2 = ops.gemm_dual("m", hh, gg, a2=xx, trans_a=True, alpha=c1, alpha2=c2)
        assert rel(gw1, c1 * d(hh).t() @ d(gg)) <= RTOL
        assert rel(gw2, c2 * d(xx).t() @ d(gg)) <= RTOL


@pytest.mark.parametrize("G,rows,M,N", [(10, 16384, 128, 128), (3, 300, 128, 100), (16, 5000, 64, 47), (1, 70000, 100, 128)])
def test_gemm_group_accumulates_every_problem(ops, cuda, G, rows, M, N):
    """incagg_gemm_tf32x3_group: G problems D[g] += alpha[g] A[g]^T B[g] in one launch (the weight gradients
    of all GCNII layers), with and without split-K, ragged tiles, strided destinations."""
    g = torch.Generator().manual_seed(G * 31 + rows)
    d = lambda t: t.double()
    As = [torch.randn(rows, M, generator=g).to(cuda) for _ in range(G)]
    Bs = [torch.randn(rows, N, generator=g).to(cuda) for _ in range(G)]
    flat = torch.randn(G, M, N + 4, generator=g).to(cuda)
    outs = [flat[i, :, :N] for i in range(G)]
    old = [o.clone() for o in outs]
    alphas = [0.1 * (i + 1) for i in range(G)]
    ops.gemm_group(As, Bs, outs, alphas, trans_a=True, beta=1.0)
    for i in range(G):
        ref = d(old[i]) + alphas[i] * d(As[i]).t() @ d(Bs[i])
        assert float((d(outs[i]) - ref).abs().max() / ref.abs().max()) <= RTOL, i
    keep = flat[:, :, N:].clone()
    ops.gemm_group(As, Bs, outs, alphas, trans_a=True, beta=0.0)
    for i in range(G):
        ref = alphas[i] * d(As[i]).t() @ d(Bs[i])
        assert float((d(outs[i]) - ref).abs().max() / ref.abs().max()) <= RTOL, i
    assert torch.equal(flat[:, :, N:], keep)       # the columns beside the destinations are untouched
    ops.check_device_errors()


@pytest.mark.parametrize("M,C,N", [(700, 128, 128), (333, 128, 100), (20000, 96, 47), (260, 64, 192)])
def test_gemm_dual_n_accumulates_second_output(ops, cuda, M, C, N):
    """N-concatenated pair with acc2: D2 += alpha2 A (s2 B2) + beta2 Cin2 (the x_0 gradient of the GCNII
    layers accumulated in the epilogue): vector and ragged n-tiles, the pair kernel (N <= 128) and the
    two-set general kernel (N = 192), aligned and unaligned destinations."""
    g = torch.Generator().manual_seed(M + N)
    d = lambda t: t.double()
    rel = lambda o, r: float((d(o) - r).abs().max() / r.abs().max())
    gr = torch.randn(M, C, generator=g).to(cuda)
    w1 = torch.randn(N, C, generator=g).to(cuda)
    w2 = torch.randn(N, C, generator=g).to(cuda)
    cin = torch.randn(M, N, generator=g).to(cuda)
    for off in (0, 1):   # 16-byte aligned destination / not
        store = torch.randn(M, N + 8, generator=g).to(cuda)
        acc = store[:, off:off + N]
        old = acc.clone()
        o1, o2 = ops.gemm_dual("n", gr, w1, b2=w2, trans_b=True, scale_b=0.4, scale_b2=0.1, cin=cin, beta=0.5,
                               cin2=cin, beta2=0.25, out2=acc, acc2=True)
        assert o2.data_ptr() == acc.data_ptr()
        assert rel(o1, 0.4 * d(gr) @ d(w1).t() + 0.5 * d(cin)) <= RTOL
        assert rel(acc, d(old) + 0.1 * d(gr) @ d(w2).t() + 0.25 * d(cin)) <= RTOL
        assert torch.equal(store[:, :off], store[:, :off]) and float(store[:, off + N:].abs().max()) > 0


def test_x0_grad_sink_matches_autograd(ops, cuda):
    """Five GCNII dense blocks sharing x_0: gradients with the epilogue-accumulating sink equal the
    plain autograd accumulation (<= 1e-6 relative: same products, different order of the sums)."""
    from incagg_gnn_b200.nn import GCN2Conv, X0GradSink
    g = torch.Generator().manual_seed(21)
    B, H, C = 600, 300, 128
    convs = [GCN2Conv(C, alpha=0.1, theta=0.5, layer=i + 1, shared_weights=False).to(cuda) for i in range(5)]
    base = torch.randn(B + H, C, generator=g).to(cuda)
    gout = torch.randn(B, C, generator=g).to(cuda)

    def run(with_sink):
        x_full = base.clone().requires_grad_(True)
        y = x_full * 1.0                      # stands for the first Linear
        sink = X0GradSink() if with_sink else None
        x_0 = sink.join(y) if with_sink else y
        x0b = x_0[:B]
        h = x_0[:B] * 0.5 + x_0[B:B + B // 2].repeat(2, 1)[:B]   # some function of all of x_0
        for conv in convs:
            conv.zero_grad(set_to_none=True)
            h = conv.forward_after_propagate(h, x0b, relu=True, x0_sink=sink)
        h.backward(gout)
        return x_full.grad.clone(), [c.weight2.grad.clone() for c in convs]

    gx_ref, gw_ref = run(False)
    gx, gw = run(True)
    assert float((gx - gx_ref).abs().max() / gx_ref.abs().max()) <= 1e-6
    for a, b in zip(gw, gw_ref):
        assert torch.equal(a, b)
    ops.check_device_errors()


# ---- transpose --------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(50, 70, 6, 0, 0), (300, 40, 40, 0, 0), (64, 20, 4, 2, 40000)])
def test_csr_transpose_is_bit_exact_and_ordered(ops, cuda, shape):
    rows, cols, maxdeg, nlong, llen = shape
    rng = np.random.default_rng(rows)
    rowptr, col, val = _rand_csr(rng, rows, cols, maxdeg, nlong, llen)
    t_rowptr, t_col, t_val, t_perm = ops.csr_transpose(_dev(rowptr, cuda, torch.int32), _dev(col, cuda, torch.int32),
                                                       _dev(val, cuda), rows, cols, want_perm=True)
    e_rowptr, e_col, e_val, e_perm = oracle.csr_transpose(rowptr, col, val, cols)
    assert np.array_equal(t_rowptr.cpu().numpy(), e_rowptr)
    assert np.array_equal(t_perm.cpu().numpy(), e_perm)
    assert np.array_equal(t_col.cpu().numpy(), e_col)
    assert np.array_equal(t_val.cpu().numpy(), e_val)


def test_backward_spmm_through_transpose(ops, cuda):
    """grad_X = A^T grad_out via the transposed CSR equals the dense formula."""
    from incagg_gnn_b200.sparse import SparseTensor
    rng = np.random.default_rng(4)
    rowptr, col, val = _rand_csr(rng, 90, 140, 9)
    adj = SparseTensor(rowptr=_dev(rowptr, cuda), col=_dev(col, cuda), value=_dev(val, cuda), sparse_sizes=(90, 140))
    x = torch.from_numpy(rng.standard_normal((140, 36)).astype(np.float32)).to(cuda).requires_grad_(True)
    g = torch.from_numpy(rng.standard_normal((90, 36)).astype(np.float32)).to(cuda)
    for reduce in ("sum", "mean"):
        x.grad = None
        (adj.matmul(x, reduce=reduce) * g).sum().backward()
        from scipy.sparse import csr_matrix
        a = csr_matrix((val.astype(np.float64), col, rowptr), shape=(90, 140)).toarray()
        gg = g.cpu().numpy().astype(np.float64)
        if reduce == "mean":
            gg = gg / np.maximum(np.diff(rowptr), 1)[:, None]
        assert _rel_err(x.grad.cpu().numpy(), a.T @ gg) <= RTOL


# ---- rows: gather / scatter / slices -----------------------------------------------------------------
@pytest.mark.parametrize("D,dtype", [(128, torch.float32), (40, torch.float32), (100, torch.float32),
                                     (7, torch.float32), (1, torch.int64), (3, torch.int32)])
@pytest.mark.parametrize("pinned", [False, True])
def test_gather_scatter_rows_bit_exact(ops, cuda, D, dtype, pinned):
    g = torch.Generator().manual_seed(D)
    N, n = 1000, 333
    src = (torch.randn(N, D, generator=g) * 100).to(dtype)
    idx = torch.randperm(N, generator=g)[:n]
    s = src.pin_memory() if pinned else src.to(cuda)
    out = ops.gather_rows(s, idx.to(cuda))
    assert torch.equal(out.cpu(), src[idx])
    dst = torch.zeros(N, D, dtype=dtype)
    d = dst.pin_memory() if pinned else dst.to(cuda)
    ops.scatter_rows(out, idx.to(cuda), d)
    torch.cuda.synchronize()
    exp = torch.zeros(N, D, dtype=dtype)
    exp[idx] = src[idx]
    assert torch.equal(d.cpu(), exp)
    # empty index
    assert ops.gather_rows(s, torch.empty(0, dtype=torch.int64, device=cuda)).shape[0] == 0


@pytest.mark.parametrize("pinned", [False, True])
def test_copy_slices_both_directions(ops, cuda, pinned):
    g = torch.Generator().manual_seed(1)
    N, D = 500, 48
    table = torch.randn(N, D, generator=g)
    t = table.clone().pin_memory() if pinned else table.to(cuda)
    offset = torch.tensor([10, 300, 120, 499, 0])
    count = torch.tensor([50, 7, 0, 1, 3])
    total = int(count.sum())
    packed = torch.zeros(total + 5, D, device=cuda)
    assert ops.copy_slices(t, packed, offset, count, 0) == total
    exp = torch.cat([table[o:o + c] for o, c in zip(offset.tolist(), count.tolist())])
    torch.cuda.synchronize()
    assert torch.equal(packed[:total].cpu(), exp)
    new = torch.randn(total, D, generator=g)
    ops.copy_slices(new.to(cuda), t, offset, count, 1)
    torch.cuda.synchronize()
    exp_t = table.clone()
    s = 0
    for o, c in zip(offset.tolist(), count.tolist()):
        exp_t[o:o + c] = new[s:s + c]
        s += c
    assert torch.equal(t.cpu(), exp_t)
    # many slices (> one launch table) on the device path
    if not pinned:
        k = 150
        off = torch.arange(k) * 3
        cnt = torch.ones(k, dtype=torch.int64) * 2
        pk = torch.zeros(2 * k, D, device=cuda)
        ops.copy_slices(t, pk, off, cnt, 0)
        exp = torch.cat([exp_t[o:o + 2] for o in off.tolist()])
        assert torch.equal(pk.cpu(), exp)
    # bounds are checked like the reference's "Invalid index"
    from incagg_gnn_b200._lib import IncAggError
    with pytest.raises(IncAggError):
        ops.copy_slices(t, packed, torch.tensor([490]), torch.tensor([20]), 0)


# ---- relabel ----------------------------------------------------------------------------------------
def _golden_cases():
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "relabel_golden.npz"))
    for k in range(int(d["num_cases"])):
        p = f"c{k}_"
        val = d[p + "in_value"] if p + "in_value" in d.files else None
        yield (str(d[p + "fn"]), bool(d[p + "bipartite"]), d[p + "in_rowptr"], d[p + "in_col"], val,
               d[p + "in_idx"], d[p + "out_rowptr"], d[p + "out_col"],
               d[p + "out_value"] if val is not None else None, d[p + "out_n_id"])


def test_relabel_matches_reference_golden_vectors(ops, cuda):
    """GPU relabel vs the outputs of the reference's own compiled op, through the re-registered
    torch.ops.torch_geometric_autoscale.* names (int64 in / int64 out like the reference)."""
    ns = torch.ops.torch_geometric_autoscale
    n = 0
    for fn, bip, rowptr, col, val, idx, e_rowptr, e_col, e_val, e_nid in _golden_cases():
        r, c, v, nid = getattr(ns, fn)(_dev(rowptr, cuda), _dev(col, cuda), _dev(val, cuda) if val is not None else None,
                                       _dev(idx, cuda), bip)
        assert r.dtype == torch.int64 and c.dtype == torch.int64
        assert np.array_equal(r.cpu().numpy(), e_rowptr), (fn, bip, n)
        assert np.array_equal(c.cpu().numpy(), e_col), (fn, bip, n)
        assert np.array_equal(nid.cpu().numpy(), e_nid), (fn, bip, n)
        if val is not None:
            assert np.array_equal(v.cpu().numpy(), e_val)
        else:
            assert v is None
        n += 1
    assert n == 56


@pytest.mark.parametrize("within", [False, True])
@pytest.mark.parametrize("col32", [False, True])
def test_relabel_random_graphs_vs_oracle(ops, cuda, within, col32):
    rng = np.random.default_rng(17 + within)
    N = 20000
    rowptr, col, val = _rand_csr(rng, N, N, 40, long_rows=4, long_len=3000)
    fn_gpu = ops.relabel_one_hop_within_batch if within else ops.relabel_one_hop
    fn_cpu = oracle.relabel_one_hop_within_batch if within else oracle.relabel_one_hop
    d_rowptr, d_val = _dev(rowptr, cuda), _dev(val, cuda)
    d_col = _dev(col, cuda, torch.int32 if col32 else torch.int64)
    for trial in range(3):  # the workspace table must be left clean by every call
        idx = (rng.integers(0, N, 3000) if trial == 1 else
               np.concatenate([np.arange(5000, 7000), rng.permutation(5000)[:500]])).astype(np.int64)
        got = fn_gpu(d_rowptr, d_col, d_val, _dev(idx, cuda), trial != 2, out_int32=(trial == 0))
        exp = fn_cpu(rowptr, col, val, idx, trial != 2)
        for g, e in zip(got, exp):
            assert np.array_equal(g.cpu().numpy().astype(e.dtype), e)


def test_relabel_cpu_tensors_raise(ops, cuda):
    with pytest.raises(RuntimeError):
        ops.relabel_one_hop(torch.tensor([0, 1]), torch.tensor([0]), None, torch.tensor([0]))


# ---- async staging ------------------------------------------------------------------------------------
def test_read_write_async_semantics(ops, cuda):
    """read_async / write_async (csrc/cuda/async_cuda.cu:14-165): slices then indexed rows; the same
    argument checks raise."""
    g = torch.Generator().manual_seed(3)
    N, D = 800, 64
    table = torch.randn(N, D, generator=g).pin_memory()
    offset, count = torch.tensor([100, 400]), torch.tensor([30, 20])
    index = torch.randperm(N, generator=g)[:77]
    dst = torch.zeros(200, D, device=cuda)
    buf = torch.empty(200, D).pin_memory()
    side = torch.cuda.Stream(cuda)
    with torch.cuda.stream(side):
        torch.ops.torch_geometric_autoscale.read_async(table, offset, count, index, dst, buf)
    torch.ops.torch_geometric_autoscale.synchronize()
    exp = torch.cat([table[100:130], table[400:420], table[index]])
    assert torch.equal(dst[:127].cpu(), exp)
    with pytest.raises(RuntimeError, match="non-default"):
        ops.read_async(table, offset, count, index, dst, buf)
    with pytest.raises(RuntimeError, match="too small"):
        with torch.cuda.stream(side):
            ops.read_async(table, offset, count, torch.arange(500), dst, buf)
    src = torch.randn(50, D, generator=g).to(cuda)
    with torch.cuda.stream(side):
        side.wait_stream(torch.cuda.current_stream())
        torch.ops.torch_geometric_autoscale.write_async(src, offset, count, table)
    side.synchronize()
    assert torch.equal(table[100:130], src[:30].cpu()) and torch.equal(table[400:420], src[30:].cpu())


def test_async_io_pool_fifo(ops, cuda):
    """AsyncIOPool state machine (pool.py:64-123): more pulls than slots, FIFO order preserved."""
    from incagg_gnn_b200 import AsyncIOPool
    g = torch.Generator().manual_seed(5)
    table = torch.randn(300, 32, generator=g).pin_memory()
    pool = AsyncIOPool(pool_size=2, buffer_size=64, embedding_dim=32).to(cuda)
    idxs = [torch.randperm(300, generator=g)[:40] for _ in range(5)]
    empty_o = torch.tensor([5]), torch.tensor([10])
    for ix in idxs:
        pool.async_pull(table, empty_o[0], empty_o[1], ix)
    for ix in idxs:
        out = pool.synchronize_pull()[:50].clone()
        pool.free_pull()
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), torch.cat([table[5:15], table[ix]]))
    x = torch.randn(20, 32, generator=g).to(cuda)
    pool.async_push(x, torch.tensor([100]), torch.tensor([20]), table)
    pool.synchronize_push()
    assert torch.equal(table[100:120], x.cpu())


def test_gather_rows_sharded_matches_unsharded(ops, cuda):
    """Row gather out of a table split by row ranges over several memories (the multi-GPU history
    pull; here all shards live on one GPU): bit-exact with the gather out of the whole table."""
    g = torch.Generator().manual_seed(8)
    N, D = 1000, 96
    table = torch.randn(N, D, generator=g).to(cuda)
    bounds = [0, 130, 130, 700, 1000]                      # one empty shard
    shards = [table[bounds[i]:bounds[i + 1]].clone() for i in range(4)]
    idx = torch.randperm(N, generator=g)[:400].to(cuda)
    out = torch.full((400, D), float("nan"), device=cuda)
    ops.gather_rows_sharded(shards, bounds, idx, out)
    assert torch.equal(out, table[idx])
    # ids owned by no shard (a window of the id space): zero rows, and the device error word is set
    ops.check_device_errors()
    out2 = torch.full((400, D), float("nan"), device=cuda)
    ops.gather_rows_sharded(shards[2:], bounds[2:], idx, out2)
    inside = (idx >= 130)
    assert torch.equal(out2[inside], table[idx[inside]]) and float(out2[~inside].abs().max()) == 0.
    with pytest.raises(RuntimeError, match="row index"):
        ops.check_device_errors()


def test_relu_bwd_colsum_and_masked_ce(ops, cuda):
    g = torch.Generator().manual_seed(4)
    G = torch.randn(3000, 128, generator=g).to(cuda)
    Y = torch.randn(3000, 128, generator=g).to(cuda)
    gm, cs = ops.relu_bwd_colsum(G, Y)
    ref = G * (Y > 0)
    assert torch.equal(gm, ref)
    assert float((cs.double() - ref.double().sum(0)).abs().max() / ref.double().sum(0).abs().max()) <= RTOL
    _, cs2 = ops.relu_bwd_colsum(G[:, :40].contiguous())
    r2 = G[:, :40].double().sum(0)
    assert float((cs2.double() - r2).abs().max() / r2.abs().max()) <= RTOL
    # scalar-column path: 47 columns (the classifier head), a strided unaligned view, many row blocks
    big = torch.randn(90000, 64, generator=g).to(cuda)
    for view, yv in ((big[:, 1:48], None), (big[:, 3:50], big[:, 10:57])):
        assert ops.colsum_supported(view)
        gm3, cs3 = ops.relu_bwd_colsum(view, yv)
        r3 = view if yv is None else view * (yv > 0)
        assert torch.equal(gm3, r3)
        assert float((cs3.double() - r3.double().sum(0)).abs().max() / r3.double().sum(0).abs().max()) <= RTOL
    # addend on the first rows (the x_0 gradient sink) + accumulation into an existing buffer
    addend = torch.randn(1000, 128, generator=g).to(cuda)
    into = torch.randn(128, generator=g).to(cuda)
    before = into.clone()
    gm5, none5 = ops.relu_bwd_colsum(G, Y, add=addend, colsum_into=into)
    r5 = G.clone()
    r5[:1000] += addend
    r5 = r5 * (Y > 0)
    assert none5 is None and torch.equal(gm5, r5)
    r5s = before.double() + r5.double().sum(0)
    assert float((into.double() - r5s).abs().max() / r5s.abs().max()) <= RTOL
    gm6, cs6 = ops.relu_bwd_colsum(G, None, add=addend)      # no mask: g + addend and its column sums
    r6 = G.clone()
    r6[:1000] += addend
    assert torch.equal(gm6, r6)
    assert float((cs6.double() - r6.double().sum(0)).abs().max() / r6.double().sum(0).abs().max()) <= RTOL
    _, cs4 = ops.relu_bwd_colsum(big)          # 352 row blocks through the parallel finish
    assert float((cs4.double() - big.double().sum(0)).abs().max() / big.double().sum(0).abs().max()) <= RTOL
    # masked_cross_entropy_grad: the same numbers without an autograd node
    from incagg_gnn_b200.nn import masked_cross_entropy_grad
    # masked cross-entropy: value and gradient vs torch
    logits = torch.randn(2000, 47, generator=g).to(cuda).requires_grad_(True)
    y = torch.randint(0, 47, (2000,), generator=g).to(cuda)
    mask = (torch.rand(2000, generator=g) < 0.6).to(cuda)
    from incagg_gnn_b200.nn import masked_cross_entropy
    loss, out3 = masked_cross_entropy(logits, y, mask)
    (loss * 2.0).backward()
    l2 = logits.detach().double().requires_grad_(True)
    ref_loss = torch.nn.functional.cross_entropy(l2[mask], y[mask])
    (ref_loss * 2.0).backward()
    assert abs(float(loss) - float(ref_loss)) <= RTOL * abs(float(ref_loss))
    assert float(out3[2]) == float(mask.sum())
    assert float((logits.grad.double() - l2.grad).abs().max() / l2.grad.abs().max()) <= 1e-5
    o3, dl = masked_cross_entropy_grad(logits, y, mask)
    assert torch.equal(o3, out3) and torch.equal(dl * 2.0, logits.grad)
    # the same computation as three calls (count | gradient + partials | loss value + running sums)
    cnt = ops.mask_count(mask)
    dl2, ws = ops.masked_ce_rows(logits.detach(), y, mask, cnt)
    acc = torch.tensor([1.5, 2.0], dtype=torch.float64, device=cuda)
    o3b = ops.masked_ce_finish(ws, logits.size(0), cnt, acc)
    assert torch.equal(dl2, dl) and torch.equal(o3b, o3)
    assert acc.tolist() == [1.5 + float(o3[0]), 2.0 + float(o3[2])]
    # empty mask: zero loss, zero gradient
    loss0, o0 = masked_cross_entropy(logits.detach(), y, torch.zeros_like(mask))
    assert float(loss0) == 0. and float(o0[2]) == 0.


def test_flat_adam_matches_torch_adam(ops, cuda):
    """The fused one-launch Adam (incagg_adam_step) follows torch.optim.Adam step by step (two groups
    with different weight decay, as main.py:196-201 builds them)."""
    from incagg_gnn_b200.train import FlatAdam
    g = torch.Generator().manual_seed(6)
    shapes = [(64, 32), (32,), (48, 64), (7,)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    pa = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    pb = [torch.nn.Parameter(t.clone().to(cuda).double()) for t in init]
    ours = FlatAdam([dict(params=pa[:2], weight_decay=0.01), dict(params=pa[2:], weight_decay=0.0)], lr=0.01)
    ref = torch.optim.Adam([dict(params=pb[:2], weight_decay=0.01), dict(params=pb[2:], weight_decay=0.0)], lr=0.01)
    for step in range(6):
        grads = [torch.randn(*s, generator=g) for s in shapes]
        ours.zero_grad()
        for p, gr in zip(pa, grads):
            p.grad.add_(gr.to(cuda))          # accumulate into the flat views, as autograd does
        for p, gr in zip(pb, grads):
            p.grad = gr.to(cuda).double()
        ours.step()
        ref.step()
        for a, b in zip(pa, pb):
            assert float((a.detach().double() - b.detach()).abs().max()) <= 2e-6, step
    assert float(ours.step_t) == 6.0
    assert pa[0].data_ptr() == ours.flat_p.data_ptr() and pa[0].grad.data_ptr() == ours.flat_g.data_ptr()


def test_copy_slices_tma_bulk_path(ops, cuda):
    """Dense device<->device slices large enough for the cp.async.bulk (TMA) kernel: ragged chunk tails,
    many slices, both directions, bit-exact."""
    g = torch.Generator().manual_seed(9)
    N, D = 60000, 128
    table = torch.randn(N, D, generator=g).to(cuda)
    offset = torch.tensor([20000, 30000, 7, 45000, 59990])   # disjoint ranges (partitions never overlap)
    count = torch.tensor([4097, 123, 16385, 1, 10])          # chunk = 32 rows of 512 B: ragged tails
    total = int(count.sum())
    packed = torch.zeros(total, D, device=cuda)
    ops.copy_slices(table, packed, offset, count, 0)
    exp = torch.cat([table[o:o + c] for o, c in zip(offset.tolist(), count.tolist())])
    assert torch.equal(packed, exp)
    new = torch.randn(total, D, generator=g).to(cuda)
    t2 = table.clone()
    ops.copy_slices(new, t2, offset, count, 1)
    ref = table.clone()
    s = 0
    for o, c in zip(offset.tolist(), count.tolist()):
        ref[o:o + c] = new[s:s + c]
        s += c
    assert torch.equal(t2, ref)
    # > 64 slices in one call
    k = 200
    off = torch.arange(k) * 250
    cnt = torch.full((k,), 100, dtype=torch.int64)
    pk = torch.zeros(100 * k, D, device=cuda)
    ops.copy_slices(table, pk, off, cnt, 0)
    assert torch.equal(pk, torch.cat([table[o:o + 100] for o in off.tolist()]))


def test_out_of_range_ids_are_flagged_not_silently_skipped(ops, cuda):
    """The reference raises on an id outside its table (index_select, emb[n_id] = x).  The kernels
    record it in the device error word: a gathered row is zeros (not uninitialised memory), a scattered
    row is skipped, relabel treats the id as a node without edges - and ops.check_device_errors()
    raises."""
    ops.check_device_errors()  # clean to start with
    table = torch.arange(40, dtype=torch.float32, device=cuda).reshape(10, 4)
    idx = torch.tensor([3, 12, -1, 9], device=cuda)
    out = ops.gather_rows(table, idx)
    assert torch.equal(out[0], table[3]) and torch.equal(out[3], table[9])
    assert float(out[1].abs().sum()) == 0. and float(out[2].abs().sum()) == 0.
    with pytest.raises(RuntimeError, match="row index"):
        ops.check_device_errors()
    ops.check_device_errors()  # the word was reset
    dst = torch.zeros(10, 4, device=cuda)
    ops.scatter_rows(torch.ones(4, 4, device=cuda), idx, dst)
    assert float(dst.sum()) == 8. and float(dst[3].sum()) == 4. and float(dst[9].sum()) == 4.
    with pytest.raises(RuntimeError, match="row index"):
        ops.check_device_errors()
    # relabel with a batch id outside the graph
    rowptr = torch.tensor([0, 2, 4, 6, 8, 10, 12], device=cuda)
    col = torch.tensor([1, 5, 0, 2, 1, 3, 2, 4, 3, 5, 4, 0], device=cuda)
    rp, c, v, n_id = ops.relabel_one_hop(rowptr, col, None, torch.tensor([1, 77, 2], device=cuda), True)
    assert rp.tolist() == [0, 2, 2, 4] and n_id.tolist()[:3] == [1, 77, 2]
    with pytest.raises(RuntimeError, match="node id"):
        ops.check_device_errors()
    # and a clean call afterwards is clean
    r = ops.relabel_one_hop(rowptr, col, None, torch.tensor([1, 2], device=cuda), True)
    assert r[0].tolist() == [0, 2, 4] and r[1].tolist() == [2, 1, 0, 3] and r[3].tolist() == [1, 2, 0, 3]
    ops.check_device_errors()


def test_spmm_multi_autograd_matches_separate_aggregations(cuda):
    """sparse.spmm_multi (one launch forward; one transposed SpMM + one min/max scatter backward) against
    K separate spmm(..., reduce=) calls with their own autograd: values bit-identical for min / max,
    within 1e-6 for sum / mean (the multi launch uses the row kernel, a plain wide sum may not);
    gradients within 1e-6."""
    import incagg_gnn_b200 as tga
    from incagg_gnn_b200.sparse import spmm, spmm_multi
    g = torch.Generator(device="cpu").manual_seed(9)
    n_dst, n_src, F = 300, 700, 64
    dense = (torch.rand(n_dst, n_src, generator=g) < 0.03).float() * torch.rand(n_dst, n_src, generator=g)
    dense[7] = 0.  # an empty row
    row, col = dense.nonzero(as_tuple=True)
    adj = tga.SparseTensor(row=row.to(cuda), col=col.to(cuda), value=dense[row, col].to(cuda),
                           sparse_sizes=(n_dst, n_src), is_sorted=True)
    for reducers in (["sum", "mean", "min", "max"], ["max", "sum"], ["mean"], ["min", "max", "max", "sum", "mean", "sum"]):
        K = len(reducers)
        hs0 = torch.randn(n_src, K * F, generator=g).to(cuda)
        go = torch.randn(n_dst, K * F, generator=g).to(cuda)
        a = hs0.clone().requires_grad_(True)
        out = spmm_multi(adj, a, F, reducers)
        out.backward(go)
        b = hs0.clone().requires_grad_(True)
        ref = torch.cat([spmm(adj, b[:, k * F:(k + 1) * F], reduce=r) for k, r in enumerate(reducers)], 1)
        ref.backward(go)
        for k, r in enumerate(reducers):
            o, e = out[:, k * F:(k + 1) * F], ref[:, k * F:(k + 1) * F]
            if r in ("min", "max"):
                assert torch.equal(o, e), (reducers, k)
            else:
                assert float((o - e).abs().max()) <= 1e-6 * float(e.abs().max()), (reducers, k)
        assert float((a.grad - b.grad).abs().max()) <= 1e-6 * float(b.grad.abs().max()), reducers
