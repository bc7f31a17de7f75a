"""Pure-torch CPU restatement of the reference's GAS / IncAgg runtime and models — TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference package cannot be
imported here (torch_sparse / torch_geometric / ipdb / hydra are absent) and holds no tests or golden
vectors; every function cites the reference lines it restates and the third-party conv semantics are
the published PyG / torch_sparse definitions.

Decisions forced by reference defects: the synchronous push_and_pull branch
(models/base.py:411-426) is the semantic definition (the async pool only moves bytes); the
history-slot map is the fork's as written (GCN/SAGE push layer-l output to histories[l+1], GCN2 /
APPNP / PNA to histories[l]); ``forward_after_propagate(h, x_0)`` is everything in stock
``GCN2Conv.forward`` after ``propagate``.

Runs in fp32 (the baseline that is timed) or fp64 (the tolerance anchor for tests).
"""
from itertools import product
from typing import Dict, List, NamedTuple, Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from . import relabel as _rl

EPS = 1e-5


# ---- sparse --------------------------------------------------------------------------------------
class Adj(NamedTuple):
    rowptr: Tensor  # int64 [rows+1]
    col: Tensor     # int64 [nnz]
    val: Optional[Tensor]
    rows: int
    cols: int

    def row(self):
        deg = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.rows), deg)


# How sum / mean aggregation is executed.  'gather' is the definition the parity tests use (gather the
# messages, index_add them: any dtype, plain autograd).  'csr' is what the timed CPU baseline uses: the
# same product through ATen's CSR kernels (`torch.sparse_csr_tensor @ X`, multi-threaded), with the
# backward pass as a product with the transposed CSR that is built once per structure and reused by
# every layer - what torch_sparse does (csr2csc cached in the SparseTensor's storage).  The two agree
# to rounding (tests/test_oracle.py::test_csr_baseline_matches_the_definition).
SPMM_IMPL = 'gather'
_CSR_CACHE: Dict[int, tuple] = {}


def _csr_pair(adj: Adj, dtype):
    key = (adj.rowptr.data_ptr(), adj.col.data_ptr(), 0 if adj.val is None else adj.val.data_ptr(), dtype)
    hit = _CSR_CACHE.get(key)
    if hit is not None and hit[0] is adj.rowptr:
        return hit[1], hit[2]
    val = torch.ones(adj.col.numel(), dtype=dtype) if adj.val is None else adj.val.to(dtype)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')   # "sparse CSR support is in beta state"
        a = torch.sparse_csr_tensor(adj.rowptr, adj.col, val, size=(adj.rows, adj.cols))
        at = a.t().to_sparse_csr()
    if len(_CSR_CACHE) > 64:
        _CSR_CACHE.clear()
    _CSR_CACHE[key] = (adj.rowptr, a, at)
    return a, at


class _CsrMatmul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, a, at):
        ctx.at = at
        return a @ x

    @staticmethod
    def backward(ctx, g):
        return ctx.at @ g.contiguous(), None, None


def spmm(adj: Adj, x: Tensor, reduce: str = 'sum') -> Tensor:
    """torch_sparse.matmul(adj, x, reduce) with autograd w.r.t. x (edge values are constants)."""
    if SPMM_IMPL == 'csr' and reduce in ('sum', 'add', 'mean') and adj.col.numel() > 0:
        a, at = _csr_pair(adj, x.dtype)
        out = _CsrMatmul.apply(x, a, at)
        if reduce == 'mean':
            deg = (adj.rowptr[1:] - adj.rowptr[:-1]).clamp(min=1).to(x.dtype)
            out = out / deg.unsqueeze(1)
        return out
    row = adj.row()
    msg = x[adj.col]
    if adj.val is not None:
        msg = msg * adj.val.to(x.dtype).unsqueeze(1)
    out = x.new_zeros((adj.rows, x.size(1)))
    if reduce in ('sum', 'add', 'mean'):
        out = out.index_add(0, row, msg)
        if reduce == 'mean':
            deg = (adj.rowptr[1:] - adj.rowptr[:-1]).clamp(min=1).to(x.dtype)
            out = out / deg.unsqueeze(1)
        return out
    if reduce in ('min', 'max'):
        idx = row.unsqueeze(1).expand_as(msg)
        return out.scatter_reduce(0, idx, msg, 'amin' if reduce == 'min' else 'amax', include_self=False)
    raise ValueError(reduce)


def strip_values(adj: Adj) -> Adj:
    return Adj(adj.rowptr, adj.col, None, adj.rows, adj.cols)


def select_edges(adj: Adj, batch_size: int, aggregate_combined: bool) -> Adj:
    """gcn.py:117-141: combined mask is all-true; otherwise keep in-batch edges, same shape."""
    if aggregate_combined:
        return adj
    row, col = adj.row(), adj.col
    m = (row < batch_size) & (col < batch_size)
    counts = torch.bincount(row[m], minlength=adj.rows)
    rowptr = torch.zeros(adj.rows + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return Adj(rowptr, col[m], adj.val[m] if adj.val is not None else None, adj.rows, adj.cols)


# ---- preprocessing (inputs of the path; main.py:147-151) --------------------------------------------
def set_diag(adj: Adj) -> Adj:
    row, col = adj.row(), adj.col
    keep = row != col
    n = min(adj.rows, adj.cols)
    d = torch.arange(n)
    r, c = torch.cat([row[keep], d]), torch.cat([col[keep], d])
    v = None
    if adj.val is not None:
        v = torch.cat([adj.val[keep], torch.ones(n, dtype=adj.val.dtype)])
    key, perm = torch.sort(r * adj.cols + c, stable=True)
    counts = torch.bincount(key // adj.cols, minlength=adj.rows)
    rowptr = torch.zeros(adj.rows + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return Adj(rowptr, key % adj.cols, v[perm] if v is not None else None, adj.rows, adj.cols)


def gcn_norm(adj: Adj) -> Adj:
    row = adj.row()
    val = adj.val if adj.val is not None else torch.ones(adj.col.numel(), dtype=torch.float32)
    deg = torch.zeros(adj.rows, dtype=val.dtype).index_add_(0, row, val)
    dis = deg.pow(-0.5)
    dis[dis == float('inf')] = 0.
    return Adj(adj.rowptr, adj.col, dis[row] * val * dis[adj.col], adj.rows, adj.cols)


# ---- History (history.py:9-74) ---------------------------------------------------------------------
class History:
    def __init__(self, num_embeddings: int, embedding_dim: int, dtype=torch.float32):
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.emb = torch.zeros(num_embeddings, embedding_dim, dtype=dtype)

    def pull(self, n_id: Optional[Tensor] = None) -> Tensor:  # history.py:33-39
        return self.emb if n_id is None else self.emb.index_select(0, n_id)

    def push(self, x, n_id=None, offset=None, count=None):  # history.py:41-65
        x = x.detach()
        if n_id is None and x.size(0) != self.num_embeddings:
            raise ValueError
        elif n_id is None:
            self.emb.copy_(x)
        elif offset is None or count is None:
            self.emb[n_id] = x
        else:
            src_o = 0
            for dst_o, c in zip(offset.tolist(), count.tolist()):
                self.emb[dst_o:dst_o + c] = x[src_o:src_o + c]
                src_o += c


def push_slices(table: Tensor, x: Tensor, offset, count):
    """pool.async_push == write_async (async_cuda.cu:139-163): table[o:o+c] = x[s:s+c], zero-padded
    to the table width where the models do so (gcn.py:355-359)."""
    x = x.detach()
    if x.size(1) < table.size(1):
        xp = x.new_zeros((x.size(0), table.size(1)))
        xp[:, :x.size(1)] = x
        x = xp
    s = 0
    for o, c in zip(offset.tolist(), count.tolist()):
        table[o:o + c] = x[s:s + c]
        s += c


def pull_slices_and_index(table: Tensor, offset, count, index: Tensor) -> Tensor:
    """pool.async_pull == read_async (async_cuda.cu:61-111): slices first, then indexed rows."""
    parts = [table[o:o + c] for o, c in zip(offset.tolist(), count.tolist())]
    parts.append(table.index_select(0, index))
    return torch.cat(parts, 0)


# ---- collate (loader.py:172-214) -------------------------------------------------------------------
class SubData(NamedTuple):
    x: Tensor
    y: Tensor
    train_mask: Tensor
    adj: Adj
    batch_size: int
    n_id: Tensor
    offset: Tensor
    count: Tensor


def collate(graph: Adj, x: Tensor, y: Tensor, train_mask: Tensor, ptr: Tensor, batch_ids: List[int],
            within_batch: bool = False, relabel_fn=None) -> SubData:
    """compute_subgraph (GAS, loader.py:172-192) / compute_subgraph_IB (IncAgg, :194-214).
    `relabel_fn(rowptr, col, value, idx, bipartite)` on torch tensors replaces the C restatement (the
    timed reference arm passes the reference's own compiled op, oracle/_ref/ref_relabel.so)."""
    batch_id = torch.tensor(batch_ids)
    n_id = torch.cat([torch.arange(int(ptr[b]), int(ptr[b + 1])) for b in batch_ids])
    batch_size = n_id.numel()
    offset = ptr[batch_id]
    count = ptr[batch_id + 1] - ptr[batch_id]
    if relabel_fn is not None:
        rp, c, v, n_id = relabel_fn(graph.rowptr, graph.col, graph.val, n_id, True)
        adj = Adj(rp, c, v, batch_size, n_id.numel())
    else:
        fn = _rl.relabel_one_hop_within_batch if within_batch else _rl.relabel_one_hop
        val = None if graph.val is None else graph.val.numpy()
        rp, c, v, nid = fn(graph.rowptr.numpy(), graph.col.numpy(), val, n_id.numpy(), True)
        n_id = torch.from_numpy(np.ascontiguousarray(nid))
        adj = Adj(torch.from_numpy(rp), torch.from_numpy(c), None if v is None else torch.from_numpy(v),
                  batch_size, n_id.numel())
    return SubData(x.index_select(0, n_id), y.index_select(0, n_id), train_mask.index_select(0, n_id),
                   adj, batch_size, n_id, offset, count)


# ---- models ----------------------------------------------------------------------------------------
class OracleGNN:
    """Base of the restated models: parameters come from a state_dict with the product's (= PyG's)
    parameter names; histories as in models/base.py:67-81."""

    def __init__(self, kind: str, state: Dict[str, Tensor], num_nodes: int, in_channels: int,
                 hidden_channels: int, out_channels: int, num_layers: int, dtype=torch.float32,
                 alpha: float = 0.1, theta: float = 0.5, shared_weights: bool = True,
                 batch_norm: bool = False, residual: bool = False, linear: bool = False,
                 aggregators=None, scalers=None, deg: Optional[Tensor] = None, **_):
        self.kind, self.dtype = kind, dtype
        self.p = {k: v.detach().clone().to('cpu').to(dtype).requires_grad_(True)
                  for k, v in state.items() if v.dtype.is_floating_point and 'histories' not in k
                  and 'running_' not in k}
        self.num_nodes, self.in_channels, self.out_channels = num_nodes, in_channels, out_channels
        self.hidden_channels, self.num_layers = hidden_channels, num_layers
        self.alpha, self.theta, self.shared_weights = alpha, theta, shared_weights
        self.residual, self.linear = residual, linear
        assert not batch_norm, 'the oracle covers batch_norm=False configurations'
        width = out_channels if kind == 'APPNP' else hidden_channels  # appnp.py:24
        self.width = width
        self.histories = [History(num_nodes, width, dtype) for _ in range(num_layers)]
        self.histories_ag = [History(num_nodes, width, dtype) for _ in range(num_layers)]
        self.out = torch.zeros(num_nodes, out_channels, dtype=dtype)
        self.aggregators, self.scalers = aggregators, scalers
        if kind == 'PNA':
            d = deg.to(torch.float)
            self.avg_deg_log = (d + 1).log().mean().item()

    def parameters(self):
        return list(self.p.values())

    # -- dense pieces ------------------------------------------------------------------------------
    def lin(self, name: str, x: Tensor) -> Tensor:
        out = x @ self.p[f'{name}.weight'].t()
        if f'{name}.bias' in self.p:
            out = out + self.p[f'{name}.bias']
        return out

    def gcn2_after(self, l: int, h: Tensor, x_0: Tensor) -> Tensor:
        """GCN2Conv.forward after propagate [upstream PyG]; layer = l+1 (gcn2.py:46-48)."""
        beta = float(np.log(self.theta / (l + 1) + 1))
        x = h * (1 - self.alpha)
        x0 = self.alpha * x_0[:x.size(0)]
        w1 = self.p[f'convs.{l}.weight1']
        if self.shared_weights:
            out = x + x0
            return torch.addmm(out, out, w1, beta=1. - beta, alpha=beta)
        w2 = self.p[f'convs.{l}.weight2']
        out = torch.addmm(x, x, w1, beta=1. - beta, alpha=beta)
        return out + torch.addmm(x0, x0, w2, beta=1. - beta, alpha=beta)

    def gcn_dense(self, l: int, h: Tensor) -> Tensor:  # conv.lin (+ bias), gcn.py:242-244
        out = h @ self.p[f'convs.{l}.lin.weight'].t()
        if f'convs.{l}.bias' in self.p:
            out = out + self.p[f'convs.{l}.bias']
        return out

    def sage_dense(self, l: int, agg: Tensor, x_root: Tensor) -> Tensor:  # graphsage.py:637-642
        out = self.lin(f'convs.{l}.lin_l', agg)
        if f'convs.{l}.lin_r.weight' in self.p:
            out = out + x_root[:agg.size(0)] @ self.p[f'convs.{l}.lin_r.weight'].t()
        return out

    def pna_conv(self, l: int, x: Tensor, adj: Adj) -> Tensor:  # pna.py:56-84
        deg = (adj.rowptr[1:] - adj.rowptr[:-1]).to(x.dtype).view(-1, 1)
        out = 0
        for k, (aggr, scaler) in enumerate(product(self.aggregators, self.scalers)):
            h = self.lin(f'convs.{l}.pre_lins.{k}', x).relu()
            if aggr in ('std', 'var'):  # PyG's definition (an extension: torch_sparse has no std)
                m1, m2 = spmm(adj, h, 'mean'), spmm(adj, h * h, 'mean')
                var = (m2 - m1 * m1).relu()
                h = var if aggr == 'var' else (var + 1e-5).sqrt()
            else:
                h = spmm(adj, h, aggr)
            h = self.lin(f'convs.{l}.post_lins.{k}', h)
            if scaler == 'amplification':
                h = h * ((deg + 1).log() / self.avg_deg_log)
            elif scaler == 'attenuation':
                h = h * (self.avg_deg_log / ((deg + 1).log() + EPS))
            out = out + h
        return out + self.lin(f'convs.{l}.lin', x)[:adj.rows]

    def conv(self, l: int, x: Tensor, adj: Adj, x_0: Optional[Tensor] = None) -> Tensor:
        if self.kind == 'GCN':  # GCNConv: A (x W) + b
            h = spmm(adj, x @ self.p[f'convs.{l}.lin.weight'].t())
            return h + self.p[f'convs.{l}.bias'] if f'convs.{l}.bias' in self.p else h
        if self.kind == 'GraphSAGE':
            return self.sage_dense(l, spmm(strip_values(adj), x, 'mean'), x)
        if self.kind == 'GCN2':
            return self.gcn2_after(l, spmm(adj, x), x_0)
        if self.kind == 'PNA':
            return self.pna_conv(l, x, adj)
        raise ValueError(self.kind)

    def post(self, h: Tensor, x: Tensor) -> Tensor:
        if self.residual and h.size(-1) == x.size(-1):
            h = h + x[:h.size(0)]
        return h.relu()

    # -- push_and_pull, synchronous branch (models/base.py:411-426) ----------------------------------
    def push_and_pull(self, hist: History, x, batch_size, n_id, offset, count):
        hist.push(x[:batch_size], n_id[:batch_size], offset, count)
        h = hist.pull(n_id[batch_size:])
        return torch.cat([x[:batch_size], h], 0)

    # -- GAS step (dropout = 0 / eval: the oracle is deterministic) ----------------------------------
    def forward(self, b: SubData, aggregate_combined: bool = True) -> Tensor:
        x, B = b.x.to(self.dtype), b.batch_size
        args = (B, b.n_id, b.offset, b.count)
        adj = select_edges(b.adj, B, aggregate_combined) if self.kind != 'PNA' else b.adj
        L = self.num_layers
        if self.kind in ('GCN', 'GraphSAGE'):  # gcn.py:97-205, graphsage.py:110-366
            if self.linear:
                x = self.lin('lins.0', x).relu()
            for l in range(L - 1):
                x = self.post(self.conv(l, x, adj), x)
                x = self.push_and_pull(self.histories[l + 1], x, *args)
            h = self.conv(L - 1, x, adj)
            if not self.linear:
                return h
            return self.lin('lins.1', self.post(h, x))
        if self.kind == 'GCN2':  # gcn2.py:78-185
            x = x_0 = self.lin('lins.0', x).relu()
            for l in range(L - 1):
                h = self.conv(l, x, adj, x_0)
                x = (h + x[:h.size(0)] if self.residual else h).relu()
                x = self.push_and_pull(self.histories[l], x, *args)
            h = self.conv(L - 1, x, adj, x_0)
            x = (h + x[:h.size(0)] if self.residual else h).relu()
            return self.lin('lins.1', x)
        if self.kind == 'APPNP':  # appnp.py:44-106: one propagation per history, then one more
            x = self.lin('lins.1', self.lin('lins.0', x).relu())
            x_0 = x[:adj.rows]
            for hist in self.histories:
                x = (1 - self.alpha) * spmm(adj, x) + self.alpha * x_0
                x = self.push_and_pull(hist, x, *args)
            return (1 - self.alpha) * spmm(adj, x) + self.alpha * x_0
        if self.kind == 'PNA':  # pna.py:138-158 (bns has L-1 entries, so zip stops at L-1)
            for l in range(L - 1):
                x = self.post(self.conv(l, x, adj), x)
                x = self.push_and_pull(self.histories[l], x, *args)
            return self.conv(L - 1, x, adj)
        raise ValueError(self.kind)

    # -- IncAgg step -----------------------------------------------------------------------------------
    def _tables(self, l: int, b: SubData, width: int):
        """pool.async_pull(histories[l].emb, offset, count, empty)[:B, :F] (base.py:318-323)."""
        empty = torch.empty(0, dtype=torch.int64)
        m_in = pull_slices_and_index(self.histories[l].emb, b.offset, b.count, empty)[:b.batch_size, :width]
        m_ag = pull_slices_and_index(self.histories_ag[l].emb, b.offset, b.count, empty)[:b.batch_size, :width]
        return m_in, m_ag

    def VR_forward(self, b: SubData) -> Tensor:
        x, B, adj, L = b.x.to(self.dtype), b.batch_size, b.adj, self.num_layers
        if self.kind in ('GCN', 'GraphSAGE'):  # gcn.py:209-279, graphsage.py:539-707
            if self.linear:
                x = self.lin('lins.0', x).relu()
            h = None
            for l in range(L):
                x = x[:B]
                m_in, m_ag = self._tables(l, b, x.shape[1])
                if self.kind == 'GCN':
                    h = self.gcn_dense(l, spmm(adj, x - m_in) + m_ag)
                else:  # delta term: mean over the IN-BATCH degree (graphsage.py:634)
                    h = self.sage_dense(l, spmm(strip_values(adj), x - m_in, 'mean') + m_ag, x)
                if l < L - 1:
                    x = self.post(h, x)
            if not self.linear:
                return h
            return self.lin('lins.1', self.post(h, x))
        if self.kind == 'GCN2':  # gcn2.py:187-323
            x = x_0 = self.lin('lins.0', x).relu()
            for l in range(L):
                if l == L - 1:
                    x = x[:B]
                m_in, m_ag = self._tables(l, b, x.shape[1])
                h = self.gcn2_after(l, spmm(adj, x - m_in) + m_ag, x_0)
                x = (h + x[:h.size(0)] if self.residual else h).relu()
            return self.lin('lins.1', x)
        if self.kind == 'APPNP':  # appnp.py:108-137
            x = self.lin('lins.1', self.lin('lins.0', x[:B]).relu())
            x_0 = x[:adj.rows]
            for l in range(L):
                m_in, m_ag = self._tables(l, b, x.shape[1])
                x = (1 - self.alpha) * (spmm(adj, x - m_in) + m_ag) + self.alpha * x_0
            return x
        raise NotImplementedError('PNA has no working IncAgg path in the reference (SURVEY F10)')

    # -- layer-wise sweeps -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward_layer(self, l: int, x: Tensor, adj: Adj, state: dict) -> Tensor:
        L = self.num_layers
        if self.kind in ('GCN', 'GraphSAGE'):  # gcn.py:282-332
            if l == 0 and self.linear:
                x = self.lin('lins.0', x).relu()
            h = self.conv(l, x, adj)
            if l < L - 1 or self.linear:
                h = self.post(h, x)
            if self.linear:
                h = self.lin('lins.1', h)
            return h
        if self.kind == 'GCN2':  # gcn2.py:325-374
            if l == 0:
                x = x_0 = self.lin('lins.0', x).relu()
                state['x_0'] = x_0[:adj.rows]
            h = self.conv(l, x, adj, state['x_0'])
            x = self.post(h, x)
            if l == L - 1:
                x = self.lin('lins.1', x)
            return x
        if self.kind == 'APPNP':  # appnp.py:140-166
            if l == 0:
                x = x_0 = self.lin('lins.1', self.lin('lins.0', x).relu())
                state['x_0'] = x_0[:adj.rows]
            return (1 - self.alpha) * spmm(adj, x) + self.alpha * state['x_0']
        if self.kind == 'PNA':  # pna.py:281-295
            h = self.conv(l, x, adj)
            if l < L - 1:
                h = self.post(h, x)
            return h
        raise ValueError(self.kind)

    def _m_in0(self, x: Tensor) -> Tensor:
        if self.kind == 'GCN2':
            return self.lin('lins.0', x).relu()  # gcn2.py:452
        if self.kind == 'APPNP':
            return self.lin('lins.1', self.lin('lins.0', x).relu())  # appnp.py:249-251
        return x  # gcn.py:353, graphsage.py:880

    def _m_ag(self, adj: Adj, x: Tensor) -> Tensor:
        if self.kind == 'GraphSAGE':
            return spmm(strip_values(adj), x, 'mean')  # graphsage.py:897-898
        return spmm(adj, x)

    @torch.no_grad()
    def mini_inference(self, batches: List[SubData], vr: bool = False) -> Tensor:
        """models/base.py:509-603 (vr=False) and <model>.mini_inference_vr (vr=True; gcn.py:335-410,
        gcn2.py:432-507, appnp.py:228-314, graphsage.py:862-960)."""
        states = [dict() for _ in batches]
        H, Hag, L = self.histories, self.histories_ag, self.num_layers
        for b, st in zip(batches, states):
            x = b.x.to(self.dtype)
            out = self.forward_layer(0, x, b.adj, st)[:b.batch_size]
            if vr:
                m_in0 = self._m_in0(x)
                push_slices(Hag[0].emb, self._m_ag(b.adj, m_in0), b.offset, b.count)
                push_slices(H[0].emb, m_in0[:b.batch_size], b.offset, b.count)
            push_slices(H[1].emb if L > 1 else self.out, out, b.offset, b.count)
        for i in range(1, L - 1):
            for b, st in zip(batches, states):
                x = pull_slices_and_index(H[i].emb, b.offset, b.count, b.n_id[b.batch_size:])
                if vr:
                    push_slices(Hag[i].emb, self._m_ag(b.adj, x), b.offset, b.count)
                out = self.forward_layer(i, x, b.adj, st)[:b.batch_size]
                push_slices(H[i + 1].emb, out, b.offset, b.count)
        if L > 1:
            for b, st in zip(batches, states):
                x = pull_slices_and_index(H[-1].emb, b.offset, b.count, b.n_id[b.batch_size:])
                out = self.forward_layer(L - 1, x, b.adj, st)[:b.batch_size]
                if vr:
                    push_slices(Hag[-1].emb, self._m_ag(b.adj, x), b.offset, b.count)
                push_slices(self.out, out, b.offset, b.count)
        return self.out


def train_epoch(model: OracleGNN, batches: List[SubData], optimizer, vr: bool, grad_norm=None):
    """mini_train (main.py:47-96) with total_loss / total_examples initialised."""
    total_loss = total_examples = 0.
    losses = []
    for b in batches:
        mask = b.train_mask[:b.batch_size]
        if int(mask.sum()) == 0:
            continue
        out = model.VR_forward(b) if vr else model.forward(b)
        optimizer.zero_grad()
        loss = F.cross_entropy(out[mask], b.y[:b.batch_size][mask])
        loss.backward()
        if grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(model.parameters(), grad_norm)
        optimizer.step()
        losses.append(float(loss.detach()))
        total_loss += float(loss.detach()) * int(mask.sum())
        total_examples += int(mask.sum())
    return {'loss': total_loss / max(total_examples, 1), 'losses': losses}
