"""Host-side logic that runs without a GPU: containers, preprocessing, synthetic inputs, partition
helpers and the loud failures of the product path when there is no CUDA device."""
import numpy as np
import pytest
import torch

import incagg_gnn_b200 as tga
from oracle import gas


def test_sparse_tensor_container():
    row = torch.tensor([2, 0, 1, 0, 2])
    col = torch.tensor([1, 2, 0, 0, 2])
    val = torch.arange(5, dtype=torch.float32)
    a = tga.SparseTensor(row=row, col=col, value=val, sparse_sizes=(3, 3))
    rp, c, v = a.csr()
    assert rp.tolist() == [0, 2, 3, 5] and c.tolist() == [0, 2, 0, 1, 2] and v.tolist() == [3, 1, 2, 0, 4]
    assert a.rowptr.dtype == torch.int32 and a.col.dtype == torch.int32
    assert a.nnz() == 5 and a.size(0) == 3 and a.sparse_sizes() == (3, 3)
    assert a.storage.row().tolist() == [0, 0, 1, 2, 2] and a.storage.rowcount().tolist() == [2, 1, 2]
    b = a.set_value(None)
    assert b.value is None and b.col is a.col
    m = a.masked_select_nnz(torch.tensor([True, False, True, True, False]))
    assert m.nnz() == 3 and m.csr()[0].tolist() == [0, 1, 2, 3]
    with pytest.raises(RuntimeError):
        a @ torch.randn(3, 4)  # CPU tensors: no fallback


def test_preprocess_matches_oracle():
    data, ptr = tga.synthetic_graph(3000, 30000, 8, 4, 6, seed=3)
    rp, col, _ = data.adj_t.csr()
    o = gas.gcn_norm(gas.set_diag(gas.Adj(rp, col, None, 3000, 3000)))
    p = tga.gcn_norm(tga.set_diag(data.adj_t))
    prp, pcol, pval = p.csr()
    assert torch.equal(prp, o.rowptr) and torch.equal(pcol, o.col)
    torch.testing.assert_close(pval, o.val, rtol=1e-6, atol=1e-7)
    # rows of a normalised symmetric graph with self loops: sum_j a_ij * sqrt(d_j / d_i) == 1
    deg = torch.zeros(3000).index_add_(0, p.storage.row(), torch.ones(p.nnz()))
    s = torch.zeros(3000).index_add_(0, p.storage.row(), pval * deg[pcol].sqrt()) / deg.sqrt()
    torch.testing.assert_close(s, torch.ones(3000), rtol=1e-4, atol=1e-4)


def test_to_symmetric_and_set_diag():
    a = tga.SparseTensor(row=torch.tensor([0, 0, 1]), col=torch.tensor([1, 2, 0]), sparse_sizes=(3, 3))
    s = tga.to_symmetric(a)
    assert s.csr()[0].tolist() == [0, 2, 3, 4] and s.csr()[1].tolist() == [1, 2, 0, 0]
    d = tga.set_diag(s)
    assert d.csr()[1].tolist() == [0, 1, 2, 0, 1, 0, 2]


def test_synthetic_graph_shape_and_determinism():
    d1, ptr = tga.synthetic_graph(5000, 60000, 16, 7, 10, seed=1)
    d2, _ = tga.synthetic_graph(5000, 60000, 16, 7, 10, seed=1)
    assert torch.equal(d1.adj_t.col, d2.adj_t.col) and torch.equal(d1.x, d2.x)
    assert d1.adj_t.nnz() == 60000 and ptr.tolist()[-1] == 5000 and ptr.numel() == 11
    rp, col, _ = d1.adj_t.csr()
    row = d1.adj_t.storage.row()
    assert bool((row != col).all())                       # no self loops
    key = row * 5000 + col
    assert torch.unique(key).numel() == key.numel()       # no duplicates
    assert set((col * 5000 + row).tolist()) == set(key.tolist())  # symmetric
    assert int(d1.y.max()) < 7 and d1.train_mask.dtype == torch.bool
    # a partition's halo is a few times the partition, not the whole graph
    b = gas.collate(gas.Adj(rp, col, None, 5000, 5000), d1.x, d1.y, d1.train_mask, ptr, [3])
    assert b.batch_size == 500 and 0 < b.n_id.numel() - 500 < 4500


def test_metis_and_permute():
    from incagg_gnn_b200.metis import metis_available, block_ptr
    data, ptr = tga.synthetic_graph(6000, 60000, 4, 3, 6, seed=2)
    # a pre-clustered synthetic graph: identity permutation, block boundaries (no METIS call)
    perm, p2 = tga.metis(data.adj_t, 6, log=False)
    assert torch.equal(perm, torch.arange(6000)) and torch.equal(p2, ptr)
    # hide the structure behind a random relabelling; real METIS (libmetis_static.a through
    # csrc/metis_shim.c, the call torch_sparse.partition makes) must find it again
    assert metis_available(), "libincagg_metis.so not built"
    shuffle = torch.randperm(6000, generator=torch.Generator().manual_seed(5))
    shuf = tga.permute(data, shuffle, log=False)
    assert torch.equal(shuf.x, data.x[shuffle]) and shuf.adj_t.nnz() == data.adj_t.nnz()
    perm3, p3 = tga.metis(shuf.adj_t, 6, log=False)
    assert sorted(perm3.tolist()) == list(range(6000))
    assert p3[0] == 0 and p3[-1] == 6000 and bool((p3[1:] > p3[:-1]).all())
    sizes = (p3[1:] - p3[:-1]).float()
    assert float(sizes.max() / sizes.mean()) < 1.1                      # balanced parts
    d3 = tga.permute(shuf, perm3, log=False)

    def cut(adj, bounds):
        r, c = adj.storage.row(), adj.storage.col()
        return float((torch.bucketize(r, bounds[1:], right=True) != torch.bucketize(c, bounds[1:], right=True)).float().mean())

    assert cut(d3.adj_t, p3) < 0.25 < 0.7 < cut(shuf.adj_t, block_ptr(6000, 6))   # planted p_inter = 0.15
    # permuted adjacency: edge (i, j) of the new graph is edge (perm[i], perm[j]) of the old one
    r, c = d3.adj_t.storage.row(), d3.adj_t.storage.col()
    old = set((shuf.adj_t.storage.row() * 6000 + shuf.adj_t.storage.col()).tolist())
    assert set((perm3[r] * 6000 + perm3[c]).tolist()) == old
    # deterministic
    perm4, p4 = tga.metis(shuf.adj_t, 6, log=False)
    assert torch.equal(perm3, perm4) and torch.equal(p3, p4)


def test_no_cpu_fallback_on_the_product_path():
    data, ptr = tga.synthetic_graph(600, 4000, 4, 3, 3, seed=0)
    with pytest.raises(RuntimeError):
        tga.SubgraphLoader(data, ptr)  # CPU graph, no device: refuses
    h = tga.History(10, 4, device='cpu')
    with pytest.raises(RuntimeError):
        h.pull(torch.tensor([1, 2]))
    from incagg_gnn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gather_rows(torch.randn(4, 4), torch.tensor([0]))
    with pytest.raises(RuntimeError):
        ops.spmm_raw(torch.zeros(2, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), None, torch.randn(2, 4))


def test_config_table_matches_reference_yaml_values():
    from incagg_gnn_b200.train import CONFIGS
    c3 = CONFIGS['C3']
    assert c3['model'] == 'GCN2' and c3['num_parts'] == 150 and c3['batch_size'] == 1
    assert c3['architecture']['num_layers'] == 5 and c3['architecture']['hidden_channels'] == 128
    assert c3['architecture']['shared_weights'] is False and c3['lr'] == 0.001
    assert CONFIGS['C2']['architecture']['alpha'] == 0.1 and CONFIGS['C2']['grad_norm'] == 1.0
    assert tga.SHAPES['products'][0] == 2_449_029 and tga.SHAPES['reddit'][2] == 602


def test_utils():
    logits = torch.tensor([[2., 1.], [0., 3.], [1., 0.]])
    y = torch.tensor([0, 1, 1])
    assert abs(tga.compute_micro_f1(logits, y) - 2 / 3) < 1e-9
    assert tga.index2mask(torch.tensor([1, 3]), 5).tolist() == [False, True, False, True, False]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the reference path, no GPU needed) prints one JSON
    line with the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--scale", "128",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    assert line["metric"].startswith("edges/s") and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["warmup"] == 1 and line["steps"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("C3")


def test_gen_masks_and_edge_dropout():
    torch.manual_seed(0)
    y = torch.randint(0, 4, (400,))
    tr, va, te = tga.gen_masks(y, train_per_class=5, val_per_class=7, num_splits=3)
    assert tr.shape == (400, 3) and bool((tr.sum(0) == 20).all()) and bool((va.sum(0) == 28).all())
    assert not bool((tr & va).any()) and bool((tr | va | te).all()) and not bool((te & (tr | va)).any())
    for c in range(4):
        assert bool((tr[y == c].sum(0) == 5).all())
    adj = tga.SparseTensor(row=torch.tensor([0, 0, 1, 2]), col=torch.tensor([1, 2, 0, 0]), sparse_sizes=(3, 3))
    assert tga.dropout(adj, 0.0) is adj and tga.dropout(adj, 0.5, training=False) is adj
    d = tga.dropout(adj, 0.5)
    assert d.nnz() <= 4 and d.sparse_sizes() == (3, 3)
    w = adj.set_value(torch.ones(4))
    dw = tga.dropout(w, 0.5)
    assert dw.nnz() == 4 and set(dw.value.tolist()) <= {0.0, 2.0}
    y2 = torch.tensor([[1., 0.], [0., 1.], [1., 1.]])
    lg = torch.tensor([[2., -1.], [1., 3.], [-1., 2.]])
    # tp = 3, predicted = 4, actual = 4 -> p = r = 0.75
    assert abs(tga.compute_micro_f1(lg, y2) - 0.75) < 1e-9


def test_node_record_packing_roundtrip():
    """Labels + masks packed into 16-byte records (one host gather per step instead of four) unpack to
    the original columns, for any subset of rows and for 2-D narrow attributes."""
    import torch
    from incagg_gnn_b200.loader import pack_node_records, unpack_node_records
    g = torch.Generator().manual_seed(3)
    n = 1000
    fields = [('y', torch.randint(-5, 47, (n,), generator=g)),
              ('train_mask', torch.rand(n, generator=g) < 0.6),
              ('val_mask', torch.rand(n, generator=g) < 0.2),
              ('w', torch.randn(n, generator=g)),
              ('pair', torch.randint(0, 2 ** 15, (n, 2), generator=g).to(torch.int16)),
              ('flag', torch.randint(0, 255, (n,), generator=g).to(torch.uint8))]
    table, layout = pack_node_records(fields)
    assert table.dtype == torch.uint8 and table.size(0) == n and table.size(1) % 16 == 0
    for (k, dt, shp, o, w) in layout:
        assert o % w == 0                    # natural alignment
    idx = torch.randperm(n, generator=g)[:137]
    got = unpack_node_records(table[idx], layout)
    for k, v in fields:
        assert got[k].dtype == v.dtype and got[k].shape == v[idx].shape
        assert torch.equal(got[k], v[idx]), k


def test_fixed_batches_rule():
    """Per-batch graph capture needs a fixed set of batches: single partitions or sequential groups."""
    from incagg_gnn_b200.loader import SubgraphLoader
    class L:  # the property only reads these two attributes
        fixed_batches = SubgraphLoader.fixed_batches
    for bs, shuffle, want in [(1, True, True), (1, False, True), (4, False, True), (4, True, False)]:
        l = L(); l.batch_size, l.shuffle = bs, shuffle
        assert l.fixed_batches is want


def test_hydra_style_overrides_and_yaml_table():
    """conf/configs.yaml + `++key=value` overrides (the reference runs under hydra, main.py:112-122)."""
    from incagg_gnn_b200.train import CONFIGS, apply_overrides, load_configs
    c = apply_overrides(CONFIGS['C3'], ['++architecture.hidden_channels=64', 'lr=0.1', '++grad_norm=null',
                                        '++architecture.aggregators=[sum, max]'])
    assert c['architecture']['hidden_channels'] == 64 and c['lr'] == 0.1 and c['grad_norm'] is None
    assert c['architecture']['aggregators'] == ['sum', 'max']
    assert CONFIGS['C3']['architecture']['hidden_channels'] == 128      # the table itself is untouched
    assert load_configs(overrides=['batch_size=3'])['C1']['batch_size'] == 3
    with pytest.raises(ValueError):
        apply_overrides(CONFIGS['C3'], ['nonsense'])


def test_bench_cpu_baseline_accepts_sharded_history_tables():
    """bench.cpu_baseline() on the run dict of a world_size=2 rank: the model holds only its own rows of
    every history table (round 1's multi-GPU bench died here with a shape error)."""
    import importlib
    import os
    import sys
    from types import SimpleNamespace
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    from incagg_gnn_b200.parallel import Shard
    from incagg_gnn_b200.train import CONFIGS
    conf = dict(CONFIGS['C3'])
    conf['num_parts'] = 8
    data, fin, fout = tga.get_data('', 'products', seed=0, scale=256, num_parts=8)
    data.adj_t = tga.gcn_norm(tga.set_diag(data.adj_t), add_self_loops=False)
    from incagg_gnn_b200.metis import block_ptr
    ptr = block_ptr(data.num_nodes, 8)
    a = conf['architecture']
    H, L = a['hidden_channels'], a['num_layers']
    g = torch.Generator().manual_seed(0)
    state = {'lins.0.weight': torch.randn(H, fin, generator=g) * 0.1, 'lins.0.bias': torch.zeros(H),
             'lins.1.weight': torch.randn(fout, H, generator=g) * 0.1, 'lins.1.bias': torch.zeros(fout)}
    for l in range(L):
        state[f'convs.{l}.weight1'] = torch.randn(H, H, generator=g) * 0.1
        state[f'convs.{l}.weight2'] = torch.randn(H, H, generator=g) * 0.1
    for rank in (0, 1):
        shard = Shard(ptr, rank, 2)
        hist = [SimpleNamespace(emb=torch.randn(shard.num_local, H, generator=g), row_offset=shard.lo)
                for _ in range(L)]
        model = SimpleNamespace(histories=hist, histories_ag=hist, state_dict=lambda: state)
        run = dict(data=data, ptr=ptr, conf=conf, model=model, in_channels=fin, out_channels=fout, shard=shard)
        for mode in ('gas', 'incagg'):
            v, edges, sec = bench.cpu_baseline(run, mode, 2, 2)
            assert v > 0 and edges > 0
    full = bench.full_table(hist[0], data.num_nodes)
    assert full.shape == (data.num_nodes, H) and torch.equal(full[shard.lo:shard.hi], hist[0].emb)
    assert float(full[:shard.lo].abs().sum()) == 0.


def test_x0_grad_sink_join_adds_the_collected_gradient():
    """nn.X0GradSink.join (pure autograd, no kernel): the buffer the layers filled is added to the head
    rows of the gradient that reaches x_0, once, and the sink is emptied."""
    import torch
    from incagg_gnn_b200.nn import X0GradSink
    torch.manual_seed(0)
    x = torch.randn(7, 4, requires_grad=True)
    sink = X0GradSink()
    y = sink.join(x * 2.0)
    sink.buf = torch.ones(3, 4)          # what the layers' GEMM epilogues would have accumulated (B = 3)
    (y * torch.arange(7.).view(7, 1)).sum().backward()
    want = torch.arange(7.).view(7, 1).expand(7, 4).clone()
    want[:3] += 1.0
    assert torch.equal(x.grad, 2.0 * want)
    assert sink.buf is None
    # no contribution collected: the gradient passes through unchanged
    x2 = torch.randn(5, 4, requires_grad=True)
    X0GradSink().join(x2).sum().backward()
    assert torch.equal(x2.grad, torch.ones(5, 4))


def test_colsum_supported_layouts():
    import torch
    from incagg_gnn_b200 import ops
    assert ops.colsum_supported(torch.zeros(10, 128))
    assert ops.colsum_supported(torch.zeros(10, 47))            # scalar columns (<= 256)
    assert ops.colsum_supported(torch.zeros(10, 64)[:, 1:48])   # unaligned strided view
    assert not ops.colsum_supported(torch.zeros(10, 300)[:, :299])
    assert not ops.colsum_supported(torch.zeros(10, 128).double())
    assert not ops.colsum_supported(torch.zeros(128, 10).t())   # column-major
