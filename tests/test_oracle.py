"""Pins the CPU oracle: against the golden vectors produced by the reference's own compiled relabel
op (tests/golden/relabel_golden.npz), the known answers of SURVEY.md §4, the reference .so itself when
it is present (build container only), and cross-checks of the SpMM restatements."""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "relabel_golden.npz")


def _golden_cases():
    d = np.load(GOLDEN)
    for k in range(int(d["num_cases"])):
        p = f"c{k}_"
        val = d[p + "in_value"] if p + "in_value" in d.files else None
        yield (str(d[p + "fn"]), bool(d[p + "bipartite"]), d[p + "in_rowptr"], d[p + "in_col"], val,
               d[p + "in_idx"], d[p + "out_rowptr"], d[p + "out_col"],
               d[p + "out_value"] if val is not None else None, d[p + "out_n_id"])


def test_relabel_oracle_matches_reference_golden_vectors():
    n = 0
    for fn, bip, rowptr, col, val, idx, e_rowptr, e_col, e_val, e_nid in _golden_cases():
        r, c, v, nid = getattr(oracle, fn)(rowptr, col, val, idx, bip)
        assert np.array_equal(r, e_rowptr), (fn, bip)
        assert np.array_equal(c, e_col), (fn, bip)
        assert np.array_equal(nid, e_nid), (fn, bip)
        if val is not None:
            assert np.array_equal(v, e_val)
        n += 1
    assert n == 56


def test_relabel_known_answers_from_survey():
    rowptr = [0, 2, 4, 6, 8, 10, 12]
    col = [1, 5, 0, 2, 1, 3, 2, 4, 3, 5, 4, 0]
    val = np.arange(12, dtype=np.float32)
    r, c, v, n = oracle.relabel_one_hop(rowptr, col, val, [1, 2], True)
    assert r.tolist() == [0, 2, 4] and c.tolist() == [2, 1, 0, 3]
    assert v.tolist() == [2, 3, 4, 5] and n.tolist() == [1, 2, 0, 3]
    r, c, v, n = oracle.relabel_one_hop(rowptr, col, None, [1, 2], False)
    assert r.tolist() == [0, 2, 4, 4, 4] and v is None
    r, c, v, n = oracle.relabel_one_hop_within_batch(rowptr, col, val, [1, 2], True)
    assert r.tolist() == [0, 1, 2] and c.tolist() == [1, 0] and v.tolist() == [3, 4] and n.tolist() == [1, 2]
    r, c, v, n = oracle.relabel_one_hop_within_batch(rowptr, col, None, [1, 2], False)
    assert r.tolist() == [0, 1, 2, 2, 2]


@pytest.mark.skipif(not (oracle.ref_available() and os.path.isdir("/root/reference")),
                    reason="reference .so only exists in the build container")
def test_relabel_oracle_matches_reference_so_live():
    rng = np.random.default_rng(7)
    n = 400
    deg = rng.integers(0, 10, n)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    col = rng.integers(0, n, rowptr[-1])
    val = rng.random(rowptr[-1]).astype(np.float32)
    idx = rng.integers(0, n, 50)
    a = oracle.relabel_one_hop(rowptr, col, val, idx, False)
    b = oracle.ref_relabel("relabel_one_hop", rowptr, col, val, idx, False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def _rand_csr(rng, rows, cols, maxdeg):
    deg = rng.integers(0, maxdeg + 1, rows)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, cols, rowptr[-1]).astype(np.int64)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    return rowptr, col, val


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
def test_spmm_restatements_agree(reduce):
    """numpy restatement vs the sequential C restatement vs scipy (sum)."""
    import ctypes
    from oracle.relabel import lib, _p
    rng = np.random.default_rng(3)
    rowptr, col, val = _rand_csr(rng, 70, 90, 9)
    X = rng.standard_normal((90, 13)).astype(np.float32)
    out_np, arg_np = oracle.spmm(rowptr, col, val, X, reduce, return_arg=True)
    out_c = np.empty((70, 13), np.float32)
    arg_c = np.empty((70, 13), np.int64)
    code = {"sum": 0, "mean": 1, "min": 2, "max": 3}[reduce]
    lib().oracle_spmm_csr(code, _p(rowptr), _p(col), _p(val), _p(X), 13, _p(out_c), 13, _p(arg_c), 70, 13)
    np.testing.assert_allclose(out_np, out_c, rtol=1e-5, atol=1e-6)
    if reduce in ("min", "max"):
        assert np.array_equal(arg_np, arg_c)
        assert (out_c[np.diff(rowptr) == 0] == 0).all()
    if reduce == "sum":
        from scipy.sparse import csr_matrix
        ref = csr_matrix((val.astype(np.float64), col, rowptr), shape=(70, 90)) @ X.astype(np.float64)
        np.testing.assert_allclose(out_np, ref, rtol=1e-5, atol=1e-5)


def test_transpose_restatement():
    rng = np.random.default_rng(5)
    rowptr, col, val = _rand_csr(rng, 40, 30, 7)
    t_rowptr, t_col, t_val, perm = oracle.csr_transpose(rowptr, col, val, 30)
    from scipy.sparse import csr_matrix
    a = csr_matrix((val, col, rowptr), shape=(40, 30)).toarray()
    b = csr_matrix((t_val, t_col, t_rowptr), shape=(30, 40)).toarray()
    np.testing.assert_allclose(a.T, b, rtol=1e-6, atol=1e-6)
    for c in range(30):  # original edge order inside every transposed row
        seg = perm[t_rowptr[c]:t_rowptr[c + 1]]
        assert (np.diff(seg) > 0).all()


def _toy_problem(kind, L=3, hidden=16, seed=0):
    import torch
    from oracle import gas
    g = torch.Generator().manual_seed(seed)
    N, P, Fin, C = 240, 6, 12, 5
    rng = np.random.default_rng(seed)
    deg = rng.integers(1, 8, N)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    col = rng.integers(0, N, rowptr[-1])
    adj = gas.Adj(torch.from_numpy(rowptr), torch.from_numpy(col), None, N, N)
    # sort columns inside rows and drop duplicates is not required by any op
    adj = gas.gcn_norm(gas.set_diag(adj))
    x = torch.randn(N, Fin, generator=g)
    y = torch.randint(0, C, (N,), generator=g)
    mask = torch.rand(N, generator=g) < 0.6
    ptr = torch.arange(0, N + 1, N // P)
    def rnd(*s):
        return torch.randn(*s, generator=g) * 0.3
    st = {}
    if kind == 'GCN2':
        st = {'lins.0.weight': rnd(hidden, Fin), 'lins.0.bias': rnd(hidden),
              'lins.1.weight': rnd(C, hidden), 'lins.1.bias': rnd(C)}
        for l in range(L):
            st[f'convs.{l}.weight1'] = rnd(hidden, hidden)
            st[f'convs.{l}.weight2'] = rnd(hidden, hidden)
    elif kind == 'GCN':
        dims = [Fin] + [hidden] * (L - 1) + [C]
        for l in range(L):
            st[f'convs.{l}.lin.weight'] = rnd(dims[l + 1], dims[l])
            st[f'convs.{l}.bias'] = rnd(dims[l + 1])
    elif kind == 'APPNP':
        st = {'lins.0.weight': rnd(hidden, Fin), 'lins.0.bias': rnd(hidden),
              'lins.1.weight': rnd(C, hidden), 'lins.1.bias': rnd(C)}
    model = gas.OracleGNN(kind, st, N, Fin, hidden, C, L, dtype=torch.float64, shared_weights=False)
    return gas, model, adj, x, y, mask, ptr, P


@pytest.mark.parametrize("kind", ["GCN2", "GCN", "APPNP"])
def test_incagg_equals_gas_right_after_refresh(kind):
    """Size-independent property of the path (README of the reference): while M_in / M_ag
    are consistent with the weights, A_BB (x_B - M_in[B]) + M_ag[B] equals the full aggregation over
    batch + halo, so the IncAgg step and the GAS step produce the same logits."""
    import torch
    gas, model, adj, x, y, mask, ptr, P = _toy_problem(kind)
    eval_batches = [gas.collate(adj, x, y, mask, ptr, [b]) for b in range(P)]
    model.mini_inference(eval_batches, vr=True)
    for b in (0, 3):
        full = gas.collate(adj, x, y, mask, ptr, [b, (b + 2) % P])
        ib = gas.collate(adj, x, y, mask, ptr, [b, (b + 2) % P], within_batch=True)
        with torch.no_grad():
            o_vr = model.VR_forward(ib)
            # the sweep's logits are the full-neighbourhood result for every model
            torch.testing.assert_close(o_vr, model.out[ib.n_id[:ib.batch_size]], rtol=1e-9, atol=1e-9)
            if kind == 'GCN':
                # GCN's GAS step reads layer-l outputs from histories[l+1], where the sweep wrote
                # them, so it agrees too.  GCN2 / APPNP GAS steps read histories[l] (the fork as
                # written) and APPNP-GAS does L+1 propagations: not comparable.
                o_gas = model.forward(full)
                torch.testing.assert_close(o_vr, o_gas[:ib.batch_size], rtol=1e-9, atol=1e-9)


def test_csr_baseline_matches_the_definition():
    """The timed CPU baseline runs sum / mean aggregation through ATen's CSR kernels (forward) and the
    cached transposed CSR (backward); same values and gradients as the gather + index_add definition."""
    import torch
    from oracle import gas
    g = torch.Generator().manual_seed(5)
    rows, cols, F = 300, 500, 24
    deg = torch.randint(0, 12, (rows,), generator=g)
    rowptr = torch.zeros(rows + 1, dtype=torch.int64)
    rowptr[1:] = deg.cumsum(0)
    col = torch.randint(0, cols, (int(rowptr[-1]),), generator=g)
    val = torch.rand(col.numel(), generator=g)
    x = torch.randn(cols, F, generator=g, dtype=torch.float64)
    w = torch.randn(rows, F, generator=g, dtype=torch.float64)
    for reduce in ('sum', 'mean'):
        for v in (val, None):
            adj = gas.Adj(rowptr, col, v, rows, cols)
            res = {}
            for impl in ('gather', 'csr'):
                gas.SPMM_IMPL = impl
                xi = x.clone().requires_grad_(True)
                out = gas.spmm(adj, xi, reduce)
                (out * w).sum().backward()
                res[impl] = (out.detach(), xi.grad)
            gas.SPMM_IMPL = 'gather'
            assert torch.allclose(res['gather'][0], res['csr'][0], rtol=1e-12, atol=1e-12)
            assert torch.allclose(res['gather'][1], res['csr'][1], rtol=1e-12, atol=1e-12)


def test_reference_arm_inputs_equal_the_product_inputs():
    """oracle/synth.py (the generator of `bench.py --impl reference`, which must not import the product)
    produces the product generator's graph, features, labels and masks bit for bit."""
    import torch
    import incagg_gnn_b200 as tga
    from oracle import synth
    for name, scale, parts in (('products', 64, 8), ('arxiv', 8, 10)):
        a = synth.make_inputs(name, seed=3, scale=scale, num_parts=parts)
        d, f, c = tga.get_data('', name, seed=3, scale=scale, num_parts=parts)
        rp, col, _ = d.adj_t.csr()
        assert torch.equal(rp, a.rowptr) and torch.equal(col, a.col)
        assert torch.equal(d.x, a.x) and torch.equal(d.y, a.y) and torch.equal(d.train_mask, a.train_mask)
        assert (f, c) == (a.num_features, a.num_classes)
        from incagg_gnn_b200.metis import block_ptr
        assert torch.equal(block_ptr(d.num_nodes, parts), a.ptr)
    assert synth.SHAPES == tga.SHAPES
