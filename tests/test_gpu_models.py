"""End-to-end parity of the hot path behind the reference's Python API (SubgraphLoader ->
ScalableGNN.__call__ / VR_call -> mini_inference[_vr]) against the CPU oracle (oracle/gas.py, fp64)
on seeded down-scaled twins of the BASELINE shapes.

Bars: collate outputs, history tables pushed by slices and pulled rows are compared bit-exactly where
they are copies; aggregation outputs / logits / refreshed tables / per-step and per-epoch losses to
<= 1e-5 relative (north_star's fp32 tolerance).  Dropout is 0 (the RNG streams of CPU and GPU differ).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _rel(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def _setup(cuda, config, scale, overrides=None, history_device='cuda', num_parts=None, force_metis=False):
    import incagg_gnn_b200  # noqa: F401
    from incagg_gnn_b200.train import build
    from oracle import gas
    ov = dict(overrides or {})
    arch = dict(ov.pop('architecture', {}))
    from incagg_gnn_b200.train import CONFIGS
    a = dict(CONFIGS[config]['architecture'])
    a.update(arch)
    if 'dropout' in a:
        a['dropout'] = 0.0
    ov['architecture'] = a
    if num_parts is not None:
        ov['num_parts'] = num_parts
    run = build(config, device=cuda, seed=0, scale=scale, overrides=ov, data_device='cpu', shuffle=False,
                history_device=history_device, force_metis=force_metis)
    conf, raw = run['conf'], run['raw']
    # oracle-side inputs: same raw graph, same preprocessing, on the CPU
    rp, col, _ = raw.adj_t.csr()
    adj = gas.Adj(rp, col, None, raw.num_nodes, raw.num_nodes)
    if conf['loop']:
        adj = gas.set_diag(adj)
    if conf['norm']:
        adj = gas.gcn_norm(adj)
    kwargs = dict(a)
    if conf['model'] == 'PNA':
        kwargs['deg'] = adj.rowptr[1:] - adj.rowptr[:-1]
    omodel = gas.OracleGNN(conf['model'], run['model'].state_dict(), raw.num_nodes, run['in_channels'],
                           out_channels=run['out_channels'], dtype=torch.float64, **kwargs)
    return run, gas, omodel, adj, raw


def _oracle_batches(gas, adj, raw, ptr, groups, within):
    return [gas.collate(adj, raw.x, raw.y, raw.train_mask, ptr, g, within_batch=within) for g in groups]


def _groups(num_parts, batch_size):
    return [list(range(i, min(i + batch_size, num_parts))) for i in range(0, num_parts, batch_size)]


def test_preprocess_and_collate_match_oracle(cuda):
    run, gas, omodel, adj, raw = _setup(cuda, 'C3', 64, num_parts=6)
    g_rp, g_col, g_val = run['data'].adj_t.csr()
    assert torch.equal(g_rp.cpu(), adj.rowptr) and torch.equal(g_col.cpu(), adj.col)
    assert _rel(g_val, adj.val) <= 1e-6
    ptr = run['ptr']
    for within, loader in ((False, run['eval_loader']),):
        for (data, B, n_id, offset, count), ob in zip(loader, _oracle_batches(gas, adj, raw, ptr, _groups(6, 1), within)):
            rp, c, v = data.adj_t.csr()
            assert B == ob.batch_size
            assert torch.equal(n_id.cpu(), ob.n_id)
            assert torch.equal(rp.cpu(), ob.adj.rowptr) and torch.equal(c.cpu(), ob.adj.col)
            assert torch.equal(data.x.cpu(), ob.x) and torch.equal(data.y.cpu(), ob.y)
            assert torch.equal(data.train_mask.cpu(), ob.train_mask)
            assert torch.equal(offset, ob.offset) and torch.equal(count, ob.count)


@pytest.mark.parametrize("config,scale,parts,bs", [('C3', 64, 6, 1), ('C1', 8, 6, 3), ('C2', 16, 8, 4),
                                                   ('C4', 64, 8, 4)])
def test_incagg_refresh_and_epoch_match_oracle(cuda, config, scale, parts, bs):
    """mini_inference_vr tables + logits, then one IncAgg training epoch: per-step losses."""
    from incagg_gnn_b200.train import mini_train, mini_test
    ov = dict(VR_update=True, batch_size=bs)
    if config == 'C1':
        ov['architecture'] = dict(hidden_channels=512)  # IncAgg GCN needs F_in <= hidden (gcn.py:355)
    if config == 'C4':
        ov['architecture'] = dict(hidden_channels=640)  # F_in = 602 <= hidden; smaller than 1024 for CPU speed
    run, gas, omodel, adj, raw = _setup(cuda, config, scale, ov, num_parts=parts)
    model, ptr = run['model'], run['ptr']
    out = mini_test(model, run['eval_loader'], VR_update=True)
    ev = _oracle_batches(gas, adj, raw, ptr, _groups(parts, bs), False)
    o_out = omodel.mini_inference(ev, vr=True)
    assert _rel(out, o_out) <= RTOL
    for l in range(model.num_layers):
        assert _rel(model.histories[l].emb, omodel.histories[l].emb) <= RTOL, f'M_in[{l}]'
        assert _rel(model.histories_ag[l].emb, omodel.histories_ag[l].emb) <= RTOL, f'M_ag[{l}]'
    # training epoch (sequential batch order on both sides)
    tr = _oracle_batches(gas, adj, raw, ptr, _groups(parts, bs), True)
    conf = run['conf']
    o_opt = torch.optim.Adam(omodel.parameters(), lr=conf['lr'])
    o_res = gas.train_epoch(omodel, tr, o_opt, vr=True, grad_norm=conf['grad_norm'])
    res = mini_train(model, run['train_loader'], run['criterion'], run['optimizer'], run['max_steps'],
                     grad_norm=conf['grad_norm'], VR_update=True)
    assert abs(res['loss'] - o_res['loss']) <= RTOL * abs(o_res['loss']), (res['loss'], o_res['loss'])


@pytest.mark.parametrize("config,scale,parts,bs", [('C3', 64, 6, 1), ('C1', 8, 6, 3), ('C2', 16, 8, 4),
                                                   ('C4', 64, 8, 4), ('C5', 256, 8, 4), ('C5fused', 256, 8, 4)])
def test_gas_sweep_and_epoch_match_oracle(cuda, config, scale, parts, bs):
    """mini_inference logits + histories, then one GAS training epoch (push_and_pull every layer)."""
    from incagg_gnn_b200.train import mini_train, mini_test
    ov = dict(VR_update=False, batch_size=bs)
    if config == 'C4':
        ov['architecture'] = dict(hidden_channels=256)
    if config == 'C5':  # all five aggregators of the north star, two scalers (std: per-aggregator path)
        ov['architecture'] = dict(hidden_channels=64, aggregators=['sum', 'mean', 'min', 'max', 'std'],
                                  scalers=['identity', 'amplification'])
    fused_case = config == 'C5fused'
    if fused_case:  # the training step through the fused multi-aggregator launch + its backward
        config = 'C5'
        ov['architecture'] = dict(hidden_channels=64, aggregators=['sum', 'mean', 'min', 'max'],
                                  scalers=['identity', 'amplification', 'attenuation'])
    run, gas, omodel, adj, raw = _setup(cuda, config, scale, ov, num_parts=parts)
    model, ptr = run['model'], run['ptr']
    out = mini_test(model, run['eval_loader'], VR_update=False)
    ev = _oracle_batches(gas, adj, raw, ptr, _groups(parts, bs), False)
    o_out = omodel.mini_inference(ev, vr=False)
    assert _rel(out, o_out) <= RTOL
    for l in range(1, model.num_layers):
        assert _rel(model.histories[l].emb, omodel.histories[l].emb) <= RTOL, f'histories[{l}]'
    conf = run['conf']
    o_opt = torch.optim.Adam(omodel.parameters(), lr=conf['lr'])
    o_res = gas.train_epoch(omodel, ev, o_opt, vr=False, grad_norm=conf['grad_norm'])
    res = mini_train(model, run['train_loader'], run['criterion'], run['optimizer'], run['max_steps'],
                     grad_norm=conf['grad_norm'], VR_update=False)
    assert abs(res['loss'] - o_res['loss']) <= RTOL * abs(o_res['loss']), (res['loss'], o_res['loss'])
    # histories after the epoch: pushed rows are what the oracle pushed.  Looser than RTOL: the rows
    # were produced by weights that went through Adam steps (first steps are sign-like, so fp32
    # rounding of tiny gradients moves single weights by ~lr); the loss bar above is the parity gate.
    # (three scalers of the fused PNA case: more near-zero gradient components whose sign decides the
    # first Adam steps, hence the wider bar)
    after = 5e-2 if fused_case else 2e-3
    for l in range(model.num_layers):
        assert _rel(model.histories[l].emb, omodel.histories[l].emb) <= after, f'histories[{l}] after epoch'


def test_pinned_host_histories_with_async_pool_match_hbm_resident(cuda):
    """The reference's layout (pinned-host tables + AsyncIOPool) and the HBM-resident layout give the
    same losses and tables."""
    from incagg_gnn_b200.train import mini_train, mini_test
    res = {}
    for hd in ('cuda', None):
        for vr in (True, False):
            run, *_ = _setup(cuda, 'C3', 64, dict(VR_update=vr), history_device=hd, num_parts=6)
            model = run['model']
            assert (model.pool is not None) == (hd is None)
            mini_test(model, run['eval_loader'], VR_update=vr)
            r = mini_train(model, run['train_loader'], run['criterion'], run['optimizer'], run['max_steps'],
                           VR_update=vr)
            torch.cuda.synchronize()
            res[(hd, vr)] = (r['loss'], [h.emb.cpu().clone() for h in model.histories])
    for vr in (True, False):
        a, b = res[('cuda', vr)], res[(None, vr)]
        assert abs(a[0] - b[0]) <= 1e-6 * abs(a[0])
        for x, y in zip(a[1], b[1]):
            assert _rel(x, y) <= 1e-6


def test_full_size_products_batch_properties(cuda):
    """BASELINE-size properties that need no oracle run: on the full products-shaped graph the GPU
    relabel of a partition (a) maps every column back to the original global id, (b) lists each halo id
    once and in first-seen order, and SpMM is linear: A(ax + by) = a Ax + b Ay."""
    import incagg_gnn_b200 as tga
    from incagg_gnn_b200 import ops
    data, ptr = tga.synthetic_graph(*tga.SHAPES['products'][:4], 150, seed=0, device=cuda)
    adj = tga.gcn_norm(tga.set_diag(data.adj_t))
    assert adj.nnz() == tga.SHAPES['products'][1] + adj.size(0)
    rowptr = adj.rowptr.to(torch.int64)
    idx = torch.arange(int(ptr[7]), int(ptr[8]), device=cuda)
    rp, col, val, n_id = ops.relabel_one_hop(rowptr, adj.col, adj.value, idx, True, out_int32=True)
    B = idx.numel()
    lo, hi = int(rowptr[idx[0]]), int(rowptr[idx[-1] + 1])
    assert torch.equal(n_id[col.long()], adj.col[lo:hi].long())           # (a)
    assert torch.equal(val, adj.value[lo:hi])
    assert torch.unique(n_id).numel() == n_id.numel()                     # (b) no duplicates
    halo_cols = col[col >= B].long()
    first_pos = torch.full((n_id.numel(),), col.numel(), device=cuda, dtype=torch.int64)
    first_pos.scatter_reduce_(0, halo_cols, torch.nonzero(col >= B).squeeze(1), 'amin')
    fp = first_pos[B:]
    assert bool((fp[1:] > fp[:-1]).all())                                 # first-seen order
    x = torch.randn(n_id.numel(), 128, device=cuda)
    y = torch.randn(n_id.numel(), 128, device=cuda)
    lhs = ops.spmm_raw(rp, col, val, 2.0 * x - 3.0 * y)
    rhs = 2.0 * ops.spmm_raw(rp, col, val, x) - 3.0 * ops.spmm_raw(rp, col, val, y)
    assert _rel(lhs, rhs) <= RTOL


@pytest.mark.parametrize("vr", [False, True])
def test_graphed_trainer_matches_eager(cuda, vr):
    """Replaying the per-batch CUDA graphs gives the same losses, weights and history tables as issuing
    the steps eagerly (two epochs, fixed batch order)."""
    from incagg_gnn_b200.train import GraphedTrainer, mini_train, mini_test
    res = {}
    for graphed in (False, True):
        run, *_ = _setup(cuda, 'C3', 64, dict(VR_update=vr), num_parts=6)
        model = run['model']
        mini_test(model, run['eval_loader'], VR_update=vr)
        losses = []
        # one eager step on batch 0 first in both variants (creates the optimizer state, which must
        # exist before a step can be captured)
        tr = GraphedTrainer(model, run['train_loader'], run['optimizer'], VR_update=vr)
        tr.warmup(run['train_loader']._batches_of_epoch()[0], steps=1)
        if graphed:
            for _ in range(2):
                losses.append(tr.epoch()['loss'])
            assert len(tr.graphs) == 6
        else:
            for _ in range(2):
                losses.append(mini_train(model, run['train_loader'], run['criterion'], run['optimizer'],
                                         run['max_steps'], VR_update=vr)['loss'])
        torch.cuda.synchronize()
        res[graphed] = (losses, [p.detach().clone() for p in model.parameters()],
                        [h.emb.clone() for h in model.histories])
    for a, b in zip(res[False][0], res[True][0]):
        assert abs(a - b) <= 1e-6 * abs(a), (a, b)
    for a, b in zip(res[False][1], res[True][1]):
        assert _rel(a, b) <= 1e-5
    for a, b in zip(res[False][2], res[True][2]):
        assert _rel(a, b) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('vr', [False, True])
def test_pipelined_collate_matches_plain_replay(cuda, vr):
    """pipeline_collate=True (collate graph of step i+1 replayed on a side stream while step i runs)
    gives bit-identical weights, losses and history tables to the one-graph-per-step replay: same
    kernels on the same data, only their placement on streams differs."""
    from incagg_gnn_b200.train import GraphedTrainer, mini_test
    res = {}
    for pipelined in (False, True):
        run, *_ = _setup(cuda, 'C3', 64, dict(VR_update=vr), num_parts=6)
        model = run['model']
        mini_test(model, run['eval_loader'], VR_update=vr)
        tr = GraphedTrainer(model, run['train_loader'], run['optimizer'], VR_update=vr,
                            pipeline_collate=pipelined)
        groups = run['train_loader']._batches_of_epoch()
        tr.warmup(groups[0], steps=1)
        seq = [groups[i % 6] for i in (0, 3, 1, 1, 5, 2, 4, 0, 0, 3)]   # repeats exercise buffer reuse
        tr.run(seq)
        torch.cuda.synchronize()
        res[pipelined] = (tr.acc.clone(), [p.detach().clone() for p in model.parameters()],
                          [h.emb.clone() for h in model.histories])
    assert torch.equal(res[False][0], res[True][0])
    for a, b in zip(res[False][1], res[True][1]):
        assert torch.equal(a, b)
    for a, b in zip(res[False][2], res[True][2]):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize('bs,vr', [(1, False), (1, True), (2, False)])
def test_host_resident_collate_matches_device_resident(cuda, bs, vr):
    """Inputs in pinned host memory (graph staged by DMA for single partitions, features gathered
    through UVA, labels + masks gathered as one packed record) collate to exactly the batches the
    HBM-resident loader produces."""
    from incagg_gnn_b200.train import build
    subs = {}
    for host in (False, True):
        run = build('C3', device=cuda, seed=0, scale=64, shuffle=False, host_resident=host,
                    history_device='cuda', overrides=dict(VR_update=vr, num_parts=6, batch_size=bs))
        ld = run['train_loader']
        if host:
            assert ld._host_graph and ld._packed is not None and len(ld._packed_fields) >= 2
        subs[host] = [ld._collate(list(ids)) for ids in ld._batches_of_epoch()]
        subs[host] += [ld._collate(list(ids)) for ids in ld._batches_of_epoch()]  # sizes known: no sync
    torch.cuda.synchronize()
    assert len(subs[True]) == len(subs[False]) > 0
    for a, b in zip(subs[False], subs[True]):
        assert a.batch_size == b.batch_size
        assert torch.equal(a.n_id, b.n_id)
        assert torch.equal(a.offset, b.offset) and torch.equal(a.count, b.count)
        assert torch.equal(a.data.adj_t.rowptr, b.data.adj_t.rowptr)
        assert torch.equal(a.data.adj_t.col, b.data.adj_t.col)
        assert torch.equal(a.data.adj_t.value, b.data.adj_t.value)
        keys_a = sorted(k for k, v in a.data if isinstance(v, torch.Tensor))
        keys_b = sorted(k for k, v in b.data if isinstance(v, torch.Tensor))
        assert keys_a == keys_b and 'y' in keys_a and 'train_mask' in keys_a
        for k in keys_a:
            assert a.data[k].dtype == b.data[k].dtype and a.data[k].shape == b.data[k].shape, k
            assert torch.equal(a.data[k], b.data[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize('vr', [False, True])
def test_sweep_with_merged_eval_batches_matches_per_partition(cuda, vr):
    """The layer-wise sweep over ONE merged evaluation batch (all partitions) refreshes the same
    tables and logits as the sweep over single partitions: same rows, same per-row edge order; only
    the split of very long rows may differ (fp32 re-association), hence 1e-6 instead of bit-equal."""
    from incagg_gnn_b200.train import build, mini_test
    res = {}
    for merged in (False, True):
        run = build('C3', device=cuda, seed=0, scale=64, shuffle=False,
                    overrides=dict(VR_update=vr, num_parts=6), eval_batch_size=6 if merged else None)
        assert len(run['eval_loader']) == (1 if merged else 6)
        model = run['model']
        out = mini_test(model, run['eval_loader'], VR_update=vr).clone()
        res[merged] = [out] + [h.emb.clone() for h in list(model.histories) + list(model.histories_ag)]
    for a, b in zip(res[False], res[True]):
        assert _rel(a, b) <= 1e-6


def test_metis_partitioned_graph_matches_oracle(cuda):
    """Same parity run on a graph partitioned by the real METIS (non-identity permutation, unequal
    partition sizes): refresh tables, logits and one IncAgg + one GAS epoch."""
    from incagg_gnn_b200.train import mini_train, mini_test
    for vr in (True, False):
        run, gas, omodel, adj, raw = _setup(cuda, 'C3', 64, dict(VR_update=vr), num_parts=6, force_metis=True)
        ptr = run['ptr']
        sizes = (ptr[1:] - ptr[:-1]).tolist()
        assert len(set(sizes)) > 1, "METIS parts are not all the same size"
        model = run['model']
        out = mini_test(model, run['eval_loader'], VR_update=vr)
        ev = _oracle_batches(gas, adj, raw, ptr, _groups(6, 1), False)
        o_out = omodel.mini_inference(ev, vr=vr)
        assert _rel(out, o_out) <= RTOL
        tr = _oracle_batches(gas, adj, raw, ptr, _groups(6, 1), vr)
        o_opt = torch.optim.Adam(omodel.parameters(), lr=run['conf']['lr'])
        o_res = gas.train_epoch(omodel, tr, o_opt, vr=vr)
        res = mini_train(model, run['train_loader'], run['criterion'], run['optimizer'], run['max_steps'], VR_update=vr)
        assert abs(res['loss'] - o_res['loss']) <= RTOL * abs(o_res['loss']), (res['loss'], o_res['loss'])


@pytest.mark.parametrize("vr", [False, True])
def test_graphed_sweep_matches_eager(cuda, vr):
    from incagg_gnn_b200.train import GraphedSweep, mini_test
    run, *_ = _setup(cuda, 'C3', 64, dict(VR_update=vr), num_parts=6)
    model = run['model']
    out_e = mini_test(model, run['eval_loader'], VR_update=vr).clone()
    tabs_e = [h.emb.clone() for h in list(model.histories) + list(model.histories_ag)]
    for h in list(model.histories) + list(model.histories_ag):
        h.emb.zero_()
    sweep = GraphedSweep(model, run['eval_loader'], VR_update=vr)
    sweep()
    for h in list(model.histories) + list(model.histories_ag):
        h.emb.zero_()
    out_g = sweep()   # pure replay
    torch.cuda.synchronize()
    assert torch.equal(out_g, out_e)
    for a, b in zip(tabs_e, [h.emb for h in list(model.histories) + list(model.histories_ag)]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("vr", [False, True])
def test_graphed_steps_on_pinned_host_tables_match_the_pool_path(cuda, vr):
    """Pinned-host history tables: the CUDA-graph step (halo rows gathered from host memory through UVA,
    pushes as DMA slice copies, no pool slots) gives the losses and tables of the reference's
    AsyncIOPool protocol issued eagerly; with host_prefetch (pulls one step ahead, GAS) it still trains
    and stays close (rows pushed by the step in flight are read one step staler)."""
    from incagg_gnn_b200.train import GraphedTrainer, mini_train, mini_test
    res = {}
    for mode in ("pool", "graphed", "prefetch"):
        if mode == "prefetch" and vr:
            continue
        run, *_ = _setup(cuda, 'C3', 64, dict(VR_update=vr), history_device=None, num_parts=6)
        model = run['model']
        assert model.pool is not None and not model.histories[0].emb.is_cuda
        mini_test(model, run['eval_loader'], VR_update=vr)
        tr = GraphedTrainer(model, run['train_loader'], run['optimizer'], VR_update=vr,
                            pipeline_collate=(mode != "pool"), host_prefetch=(mode == "prefetch"))
        if mode == "pool":
            model._direct_host = False   # the reference's protocol, eager
        # one eager step on the first batch in every variant (scratch buffers exist before a capture,
        # and all variants start the compared epochs from the same state)
        tr.warmup(run['train_loader']._batches_of_epoch()[0], steps=1)
        if mode == "pool":
            losses = [mini_train(model, run['train_loader'], run['criterion'], run['optimizer'],
                                 run['max_steps'], VR_update=vr)['loss'] for _ in range(2)]
        else:
            losses = [tr.epoch()['loss'] for _ in range(2)]
        torch.cuda.synchronize()
        if model.pool is not None:
            model.pool.synchronize_push()
        res[mode] = (losses, [h.emb.clone() for h in model.histories])
    for a, b in zip(res["pool"][0], res["graphed"][0]):
        assert abs(a - b) <= 1e-6 * abs(a), (a, b)
    for x, y in zip(res["pool"][1], res["graphed"][1]):
        assert _rel(x, y) <= 1e-5
    if not vr:
        for a, b in zip(res["pool"][0], res["prefetch"][0]):
            assert b == b and abs(a - b) <= 5e-2 * abs(a), (a, b)
