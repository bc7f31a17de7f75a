"""A minimal ``SparseTensor`` with the surface the reference's models and loader use
(SURVEY.md §8b "implicit third-party interface"): torch_sparse is not a dependency here.

Storage is device-resident CSR with **int32** rowptr/col (half the index traffic of torch_sparse's
int64) and optional fp32 values; the transposed CSR that the backward SpMM walks is built on
first use by the counting-sort transpose kernel and cached on the object, so a batch structure
that is reused (layers of one step, epochs of a fixed partition) pays for it once.

    adj @ x, adj.matmul(x, reduce=)       -> hand-written SpMM kernels (ops.spmm_raw), with autograd
    adj.csr(), adj.storage.row()/col()/value()/rowcount(), size(), sparse_sizes(), nnz(),
    set_value(), to(), t()
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops


class _Storage:
    def __init__(self, owner: "SparseTensor"):
        self._o = owner

    def rowptr(self) -> Tensor:
        return self._o.rowptr.to(torch.int64)

    def row(self) -> Tensor:
        o = self._o
        counts = (o.rowptr[1:] - o.rowptr[:-1]).to(torch.int64)
        return torch.repeat_interleave(torch.arange(o.size(0), device=o.device), counts)

    def col(self) -> Tensor:
        return self._o.col.to(torch.int64)

    def value(self) -> Optional[Tensor]:
        return self._o.value

    def rowcount(self) -> Tensor:
        o = self._o
        return (o.rowptr[1:] - o.rowptr[:-1]).to(torch.int64)


class SparseTensor:
    def __init__(self, rowptr: Optional[Tensor] = None, row: Optional[Tensor] = None,
                 col: Optional[Tensor] = None, value: Optional[Tensor] = None,
                 sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False,
                 trust_data: bool = False):
        assert col is not None
        if rowptr is None:
            assert row is not None and sparse_sizes is not None
            M = int(sparse_sizes[0])
            row = row.to(torch.int64)
            if not is_sorted:
                # sort by (row, col) like torch_sparse's constructor
                key = row * int(sparse_sizes[1]) + col.to(torch.int64)
                perm = torch.argsort(key, stable=True)
                row, col = row[perm], col[perm]
                value = value[perm] if value is not None else None
            counts = torch.bincount(row, minlength=M)
            rowptr = torch.zeros(M + 1, dtype=torch.int64, device=col.device)
            torch.cumsum(counts, 0, out=rowptr[1:])
        M = rowptr.numel() - 1
        if sparse_sizes is None:
            N = int(col.max()) + 1 if col.numel() > 0 else 0
            sparse_sizes = (M, max(M, N))
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        # rowptr may carry the non-bipartite padding (extra rows with no edges)
        assert rowptr.numel() - 1 == self._sizes[0], "rowptr does not match sparse_sizes"
        self.rowptr = rowptr.to(torch.int32).contiguous()
        self.col = col.to(torch.int32).contiguous()
        self.value = value.contiguous() if value is not None else None
        self.storage = _Storage(self)
        self._t = None  # cached (t_rowptr, t_col, t_val)
        self._plan = None    # cached SpMM degree-bucket plan of this structure
        self._t_plan = None  # ... and of the transposed structure

    # ---- shape -------------------------------------------------------------------------
    @property
    def device(self):
        return self.col.device

    def size(self, dim: int) -> int:
        return self._sizes[dim]

    def sizes(self):
        return list(self._sizes)

    def sparse_sizes(self) -> Tuple[int, int]:
        return self._sizes

    def nnz(self) -> int:
        return self.col.numel()

    def csr(self):
        """(rowptr, col, value) with int64 indices, as torch_sparse returns them."""
        return self.rowptr.to(torch.int64), self.col.to(torch.int64), self.value

    def csr32(self):
        return self.rowptr, self.col, self.value

    def to(self, device, non_blocking: bool = False) -> "SparseTensor":
        device = torch.device(device)
        if device == self.device:
            return self
        out = SparseTensor.__new__(SparseTensor)
        out._sizes = self._sizes
        out.rowptr = self.rowptr.to(device, non_blocking=non_blocking)
        out.col = self.col.to(device, non_blocking=non_blocking)
        out.value = self.value.to(device, non_blocking=non_blocking) if self.value is not None else None
        out.storage = _Storage(out)
        out._t = None
        out._plan = out._t_plan = None
        return out

    def cuda(self):
        return self.to("cuda")

    def pin_memory(self) -> "SparseTensor":
        out = SparseTensor.__new__(SparseTensor)
        out._sizes = self._sizes
        out.rowptr = self.rowptr.cpu().pin_memory()
        out.col = self.col.cpu().pin_memory()
        out.value = self.value.cpu().pin_memory() if self.value is not None else None
        out.storage = _Storage(out)
        out._t = None
        out._plan = out._t_plan = None
        return out

    def set_value(self, value: Optional[Tensor], layout: Optional[str] = None) -> "SparseTensor":
        if value is None:
            # the value-less view is cached so that its transposed CSR / plans are built once
            if self.value is None:
                return self
            cached = self.__dict__.get('_stripped')
            if cached is not None:
                return cached
        out = SparseTensor.__new__(SparseTensor)
        out._sizes = self._sizes
        out.rowptr, out.col = self.rowptr, self.col
        out.value = value.contiguous() if value is not None else None
        out.storage = _Storage(out)
        out._t = None
        out._plan, out._t_plan = self._plan, self._t_plan  # plans depend on the structure only
        if self._t is not None and value is None:
            out._t = (self._t[0], self._t[1], None)
        if value is None:
            self.__dict__['_stripped'] = out
        return out

    def masked_select_nnz(self, mask: Tensor, layout: Optional[str] = None) -> "SparseTensor":
        row = self.storage.row()[mask]
        col = self.col[mask]
        val = self.value[mask] if self.value is not None else None
        return SparseTensor(row=row, col=col, value=val, sparse_sizes=self._sizes, is_sorted=True)

    # ---- transposed view for the backward pass ------------------------------------------
    def t_csr(self):
        """CSR of A^T: (t_rowptr [cols+1], t_col, t_val), built once by the transpose kernel."""
        if self._t is None:
            t_rowptr, t_col, t_val, _ = ops.csr_transpose(self.rowptr, self.col, self.value,
                                                         self._sizes[0], self._sizes[1])
            self._t = (t_rowptr, t_col, t_val)
        return self._t

    def drop_caches(self) -> None:
        """Forget the derived structures (transposed CSR, plans).  A captured step keeps recomputing
        them inside its graph; dropping the Python references hands their memory back to the pool."""
        self._t = self._plan = self._t_plan = None
        self.__dict__.pop('_t_prefix_plans', None)
        self.__dict__.pop('_stripped', None)
        self.__dict__.pop('_t_mean_val', None)

    def t_mean_values(self) -> Tensor:
        """Values of the transposed structure for the backward pass of a MEAN aggregation:
        d mean_i / d x_j = val_ij / max(deg_i, 1), i.e. the transposed entry (j, i) carries
        val_ij / deg_i.  Built once per structure (one gather), so the backward is a plain transposed
        sum-SpMM instead of a per-step division of the incoming gradient by the degrees."""
        v = self.__dict__.get('_t_mean_val')
        if v is None:
            t_rowptr, t_col, t_val = self.t_csr()
            inv = 1.0 / (self.rowptr[1:] - self.rowptr[:-1]).clamp(min=1).to(torch.float32)
            v = inv[t_col.long()]
            if t_val is not None:
                v = v * t_val
            self.__dict__['_t_mean_val'] = v
        return v

    def plan(self) -> Tensor:
        """SpMM plan of this structure (one small launch, cached)."""
        if getattr(self, '_plan', None) is None:
            self._plan = ops.spmm_plan(self.rowptr, self._sizes[0], self.col.numel())
        return self._plan

    def t_plan(self) -> Tensor:
        if getattr(self, '_t_plan', None) is None:
            t_rowptr = self.t_csr()[0]
            self._t_plan = ops.spmm_plan(t_rowptr, self._sizes[1], self.col.numel())
        return self._t_plan

    def t_plan_prefix(self, rows: int) -> Tensor:
        """Plan of the first `rows` rows of the transposed structure (backward over the in-batch
        source rows only)."""
        cache = self.__dict__.setdefault('_t_prefix_plans', {})
        if rows not in cache:
            cache[rows] = ops.spmm_plan(self.t_csr()[0], rows, -1)
        return cache[rows]

    def t(self) -> "SparseTensor":
        t_rowptr, t_col, t_val = self.t_csr()
        out = SparseTensor.__new__(SparseTensor)
        out._sizes = (self._sizes[1], self._sizes[0])
        out.rowptr, out.col, out.value = t_rowptr, t_col, t_val
        out.storage = _Storage(out)
        out._t = (self.rowptr, self.col, self.value)
        out._plan, out._t_plan = self._t_plan, self._plan
        return out

    # ---- products ------------------------------------------------------------------------
    def matmul(self, x: Tensor, reduce: str = "sum", grad_rows: Optional[int] = None) -> Tensor:
        return spmm(self, x, reduce=reduce, grad_rows=grad_rows)

    def __matmul__(self, x: Tensor) -> Tensor:
        return spmm(self, x, reduce="sum")

    def __repr__(self):
        return (f"SparseTensor(rows={self._sizes[0]}, cols={self._sizes[1]}, nnz={self.nnz()}, "
                f"value={'fp32' if self.value is not None else None}, device={self.device})")


class _SpMM(torch.autograd.Function):
    """reduce(A, x) with grad only w.r.t. x (edge values carry no gradient in the reference's
    pipelines: gcn_norm weights are constants).  ``relu_input``: x is the output of a ReLU that its
    producer fused into its own epilogue and whose backward it leaves to this node: the gradient
    w.r.t. x is returned already multiplied by [x > 0] (one kernel, sum / mean)."""

    @staticmethod
    def forward(ctx, x: Tensor, adj: SparseTensor, reduce: str, grad_rows: Optional[int],
                relu_input: bool = False):
        ctx.adj, ctx.reduce, ctx.n_src, ctx.grad_rows = adj, reduce, x.size(0), grad_rows
        ctx.relu_input = bool(relu_input)
        if reduce in ("min", "max"):
            out, arg = ops.spmm_raw(adj.rowptr, adj.col, adj.value, x, reduce, rows=adj.size(0),
                                    return_arg=True, plan=adj.plan())
            ctx.save_for_backward(arg, x if relu_input else None)
            return out
        ctx.save_for_backward(None, x if relu_input else None)
        return ops.spmm_raw(adj.rowptr, adj.col, adj.value, x, reduce, rows=adj.size(0),
                            plan=adj.plan())

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        adj, reduce = ctx.adj, ctx.reduce
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        grad_out = grad_out.contiguous()
        arg, x = ctx.saved_tensors
        if reduce in ("min", "max"):
            gx = ops.spmm_minmax_bwd_raw(adj.col, adj.value, arg, grad_out, ctx.n_src)
            if ctx.relu_input:
                gx = torch.ops.aten.threshold_backward(gx, x, 0.)
            return gx, None, None, None, None
        mean = reduce == "mean"
        gate = None
        if ctx.relu_input:
            if x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1:
                gate = x
            else:  # layout the kernel's epilogue does not read: mask afterwards
                gx = _transposed_product(adj, grad_out, ctx.n_src, ctx.grad_rows, mean=mean)
                rows = gx.size(0) if ctx.grad_rows is None else min(ctx.grad_rows, gx.size(0))
                gx[:rows] = torch.ops.aten.threshold_backward(gx[:rows], x[:rows].to(gx.dtype), 0.)
                return gx, None, None, None, None
        return (_transposed_product(adj, grad_out, ctx.n_src, ctx.grad_rows, gate=gate, mean=mean),
                None, None, None, None)


def _transposed_product(adj: SparseTensor, grad_out: Tensor, n_src: int, grad_rows: Optional[int],
                        gate: Optional[Tensor] = None, mean: bool = False):
    """grad_x = A^T grad_out.  With grad_rows = k only the first k source rows are computed (the
    rest of x was a constant, e.g. pulled history rows) and the remainder is returned as zeros.
    `gate` (= x when x is the output of a fused ReLU): the result is zeroed where gate <= 0."""
    t_rowptr, t_col, t_val = adj.t_csr()
    if mean:  # mean aggregation: the 1 / deg factor rides in the transposed values
        t_val = adj.t_mean_values()
    t_plan = adj.t_plan()
    if grad_rows is None or grad_rows >= n_src:
        return ops.spmm_raw(t_rowptr, t_col, t_val, grad_out, "sum", rows=n_src, plan=t_plan, gate=gate)
    # Rows >= grad_rows of x are constants (history rows placed there by push_and_pull, whose
    # backward reads only the first `grad_rows` rows of this gradient): they are neither computed nor
    # zero-filled (a [B+H, F] fill per layer per step otherwise).
    gx = torch.empty((n_src, grad_out.size(1)), dtype=grad_out.dtype, device=grad_out.device)
    ops.spmm_raw(t_rowptr, t_col, t_val, grad_out, "sum", rows=grad_rows, out=gx[:grad_rows],
                 plan=adj.t_plan_prefix(grad_rows), gate=gate)
    return gx


def spmm(adj: SparseTensor, x: Tensor, reduce: str = "sum", grad_rows: Optional[int] = None,
         relu_input: bool = False) -> Tensor:
    """torch_sparse.matmul / torch_geometric.utils.spmm replacement (graphsage.py:30,634)."""
    if reduce == "add":
        reduce = "sum"
    if x.size(0) < adj.size(1):
        raise RuntimeError(f"spmm: x has {x.size(0)} rows but the adjacency has {adj.size(1)} columns")
    return _SpMM.apply(x, adj, reduce, grad_rows, relu_input)


class _SpMMDelta(torch.autograd.Function):
    """h = reduce(A, x - M_in) + M_ag, fused (gcn2.py:255).  M_in / M_ag are constants."""

    @staticmethod
    def forward(ctx, x, adj, m_in, m_ag, n_id, reduce, relu_input=False):
        ctx.adj, ctx.reduce, ctx.n_src = adj, reduce, x.size(0)
        ctx.save_for_backward(x if (relu_input and x.dtype == torch.float32 and x.dim() == 2
                                    and x.stride(1) == 1) else None)
        ctx.relu_input = bool(relu_input)
        return ops.spmm_delta_raw(adj.rowptr, adj.col, adj.value, x, m_in, m_ag, n_id, reduce,
                                  rows=adj.size(0), plan=adj.plan())

    @staticmethod
    def backward(ctx, grad_out):
        adj = ctx.adj
        grad_out = grad_out.contiguous()
        (gate,) = ctx.saved_tensors
        if ctx.relu_input and gate is None:
            raise RuntimeError('spmm_delta(relu_input=True) needs a row-major float32 x')
        return (_transposed_product(adj, grad_out, ctx.n_src, None, gate=gate, mean=ctx.reduce == "mean"),
                None, None, None, None, None, None)


def spmm_delta(adj: SparseTensor, x: Tensor, m_in: Tensor, m_ag: Tensor,
               n_id: Optional[Tensor] = None, reduce: str = "sum", relu_input: bool = False) -> Tensor:
    """Fused incremental-aggregation update  A_BB (x - M_in) + M_ag  (one kernel instead of the
    reference's sub + SpMM + add + two clones)."""
    return _SpMMDelta.apply(x, adj, m_in, m_ag, n_id, reduce, relu_input)


class _SpMMMulti(torch.autograd.Function):
    """K slabs of ``hs`` ([n_src, K*F]) reduced with K reducers in ONE launch (PNA, pna.py:66-84 runs K
    separate passes).  Backward: the sum / mean slabs go through the transposed CSR - one launch per run
    of adjacent sum / mean slabs, mean gradients pre-divided by the row degree - and the min / max slabs
    are routed to their winning source rows (one atomic-scatter launch per run of adjacent min / max
    slabs) with the edge indices the forward launch recorded."""

    @staticmethod
    def forward(ctx, hs: Tensor, adj: SparseTensor, F: int, reducers):
        reducers = tuple('sum' if r == 'add' else r for r in reducers)
        need_arg = any(r in ('min', 'max') for r in reducers) and hs.requires_grad
        ctx.adj, ctx.F, ctx.reducers, ctx.n_src = adj, F, reducers, hs.size(0)
        if need_arg:
            out, arg = ops.spmm_multi_raw(adj.rowptr, adj.col, adj.value, hs, F, list(reducers),
                                          rows=adj.size(0), plan=adj.plan(), return_arg=True)
            ctx.save_for_backward(arg)
        else:
            out = ops.spmm_multi_raw(adj.rowptr, adj.col, adj.value, hs, F, list(reducers),
                                     rows=adj.size(0), plan=adj.plan())
            ctx.save_for_backward(None)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        adj, F, reducers, n_src = ctx.adj, ctx.F, ctx.reducers, ctx.n_src
        (arg,) = ctx.saved_tensors
        g = g.contiguous()
        K = len(reducers)
        linear = [r in ('sum', 'mean') for r in reducers]
        gx = torch.empty((n_src, K * F), dtype=g.dtype, device=g.device)
        if not all(linear):
            gx.zero_()  # the min / max scatter accumulates
        k = 0
        t_rowptr = t_col = t_val = t_plan = None
        while k < K:
            e = k
            while e < K and linear[e] == linear[k]:
                e += 1
            cols = slice(k * F, e * F)
            if linear[k]:
                if t_rowptr is None:
                    t_rowptr, t_col, t_val = adj.t_csr()
                    t_plan = adj.t_plan()
                # adjacent slabs with the same transposed values share a launch: sum slabs use val,
                # mean slabs val / deg (SparseTensor.t_mean_values)
                j = k
                while j < e:
                    j2 = j
                    while j2 < e and reducers[j2] == reducers[j]:
                        j2 += 1
                    sub = slice(j * F, j2 * F)
                    tv = adj.t_mean_values() if reducers[j] == 'mean' else t_val
                    ops.spmm_raw(t_rowptr, t_col, tv, g[:, sub], "sum", rows=n_src, out=gx[:, sub], plan=t_plan)
                    j = j2
            else:
                ops.spmm_minmax_bwd_raw(adj.col, adj.value, arg[:, cols], g[:, cols], n_src, out=gx[:, cols])
            k = e
        return gx, None, None, None


def spmm_multi(adj: SparseTensor, hs: Tensor, F: int, reducers) -> Tensor:
    """[rows, K*F] = slab k of ``hs`` reduced over the neighbours with reducers[k], one launch, with
    autograd w.r.t. ``hs``."""
    if hs.size(0) < adj.size(1):
        raise RuntimeError(f"spmm_multi: hs has {hs.size(0)} rows but the adjacency has {adj.size(1)} columns")
    return _SpMMMulti.apply(hs, adj, F, tuple(reducers))
