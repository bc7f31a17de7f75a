"""Conv layers the reference takes from torch_geometric (not a dependency here), restated on top of
the hand-written SpMM kernels.  Parameter names follow PyG so state_dicts line up:
``GCNConv.lin/.bias``, ``GCN2Conv.weight1/.weight2``, ``SAGEConv.lin_l/.lin_r``.

Semantics follow SURVEY.md §8a "Conv dense parts" (upstream PyG definitions) plus the two methods of
the reference's locally patched GCN2Conv:
``forward_after_propagate(h, x_0)`` = everything in ``GCN2Conv.forward`` after ``propagate``;
``forward_no_neighbor(x, x_0)`` = the same with ``h = x``.
"""
import contextlib
import math
import os
from typing import Optional

import torch
from torch import Tensor
from torch.nn import Parameter

from . import ops
from .sparse import SparseTensor, spmm


# ---- dense transforms on the tensor cores ---------------------------------------------------------
def _grad_buffer(param) -> Optional[Tensor]:
    """The persistent gradient buffer of a parameter whose ``.grad`` is a view into a flat buffer
    (train.FlatAdam marks its parameters): weight-gradient GEMMs then accumulate straight into it through
    their epilogue (Cin = D = grad, beta = 1) and return no gradient to autograd, which saves one
    elementwise ``grad += new`` launch per parameter per step."""
    if getattr(param, '_flat_grad', False) and param.grad is not None:
        return param.grad
    return None


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b (optionally ReLU) on the tcgen05 3xTF32 GEMM; backward = two more GEMMs of the
    same kernel (input gradient g W, weight gradient g^T x with deterministic split-K)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu, relu_input=False, x0_sink=None):
        # relu_input: x is the output of a ReLU whose backward mask was deferred to this node (its input
        # gradient GEMM gates its epilogue by [x > 0]).  x0_sink: the gradients that the GCNII layers
        # collected for this Linear's output (nn.X0GradSink) join the incoming gradient in the fused
        # ReLU-backward / bias-gradient kernel.
        y = ops.gemm(x, weight, trans_b=True, bias=bias, relu=relu)
        ctx.relu = relu
        ctx.relu_input, ctx.x0_sink = bool(relu_input), x0_sink
        ctx.save_for_backward(x, weight, y if relu else None)
        ctx.has_bias = bias is not None
        ctx.weight_param, ctx.bias_param = weight, bias
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        g = g.contiguous()
        gx = gw = gb = None
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        side = _WGRAD['stream']
        wbuf = _grad_buffer(ctx.weight_param) if ctx.needs_input_grad[1] else None
        bbuf = _grad_buffer(ctx.bias_param) if want_gb else None
        gate = x if ctx.relu_input else None
        add = None
        if ctx.x0_sink is not None:
            add, ctx.x0_sink.buf = ctx.x0_sink.buf, None
        if (side is not None and not ctx.relu and add is None and ctx.needs_input_grad[0] and wbuf is not None
                and (not want_gb or (bbuf is not None and ops.colsum_supported(g)))):
            # Only the input gradient continues the backward chain: the weight and bias gradients (three
            # launches + the accumulation into the flat gradient buffer) run beside it on the side stream.
            side.wait_stream(torch.cuda.current_stream(g.device))
            gx = ops.gemm(g, weight, gate=gate)
            if _WGRAD_GROUP:   # joins the grouped launch at the end of the backward pass
                _WGRAD['pending'].append((g, x, wbuf, 1.))
                if _WGRAD['ready'] is None:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(g.device))
                    _WGRAD['ready'] = ev
            with torch.cuda.stream(side):
                if not _WGRAD_GROUP:
                    ops.gemm(g, x, trans_a=True, cin=wbuf, beta=1., out=wbuf, ws_slot=1)
                if want_gb:
                    _, cs = ops.relu_bwd_colsum(g, None)
                    bbuf.add_(cs)
                    _WGRAD['keep'].append(cs)
            _WGRAD['keep'].append((g, x))
            return gx, None, None, None, None, None
        if (want_gb and ops.colsum_supported(g) and (not ctx.relu or ops.colsum_supported(y))
                and (add is None or ops.colsum_supported(add))):
            # (sink +) ReLU backward and the bias gradient in one pass over g; the bias gradient lands in
            # the flat gradient buffer directly when there is one
            g, gb = ops.relu_bwd_colsum(g, y if ctx.relu else None, add=add, colsum_into=bbuf)
        else:
            if add is not None:
                g = g.clone()
                g[:add.size(0)].add_(add)
            if ctx.relu:
                g = torch.ops.aten.threshold_backward(g, y, 0.)  # ReLU backward, one kernel
            if want_gb:
                gb = g.sum(0)
        if ctx.needs_input_grad[0]:
            gx = ops.gemm(g, weight, gate=gate)           # [M,N] x [N,K]
        if ctx.needs_input_grad[1]:
            buf = _grad_buffer(ctx.weight_param)
            if buf is not None:                           # accumulate in the GEMM epilogue
                ops.gemm(g, x, trans_a=True, cin=buf, beta=1., out=buf)
            else:
                gw = ops.gemm(g, x, trans_a=True)         # g^T x : [N,M] x [M,K]
        return gx, gw, gb, None, None, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, relu: bool = False,
           relu_input: bool = False, x0_sink=None) -> Tensor:
    if x.dim() != 2:
        assert x0_sink is None
        return linear(x.reshape(-1, x.size(-1)), weight, bias, relu,
                      relu_input).reshape(*x.shape[:-1], weight.size(0))
    return _LinearTC.apply(x, weight, bias, relu, relu_input, x0_sink)


class Linear(torch.nn.Linear):
    """torch.nn.Linear whose product runs on the hand-written tcgen05 GEMM (same parameters, same
    initialisation).  ``forward(x, relu=True)`` fuses the activation into the GEMM epilogue."""

    def forward(self, x: Tensor, relu: bool = False, relu_input: bool = False, x0_sink=None) -> Tensor:
        """relu_input: x is a ReLU output whose backward mask this layer applies (the producer was called
        with defer_relu_bwd=True).  x0_sink: see X0GradSink."""
        return linear(x, self.weight, self.bias, relu, relu_input, x0_sink)


# Weight gradients on a side stream -----------------------------------------------------------------
# Inside train.forward_backward the M-concatenated weight-gradient GEMMs of the GCNII layers
# ([h | x0]^T g, accumulated into the flat gradient buffer) are issued on a side stream: nothing in the
# rest of the backward pass reads them, so they overlap the SpMM^T / input-gradient chain (a GEMM call is
# bound by one CTA per SM moving its tiles, the SpMM by the L2 gather path: they share an SM well).
# The operands are kept alive until the join at the end of the backward pass.
# The GCNII layers do not launch theirs one by one: they queue the problems (2 per layer, all of one
# shape) and the block's exit issues them as ONE grouped split-K launch on the side stream, after the
# last layer's input-gradient GEMM - a weight-gradient GEMM per layer holds every SM's shared memory for
# ~25 us and the chain's own GEMM of that layer had to wait for it (35 instead of 22 us per layer).
_WGRAD = {'stream': None, 'keep': [], 'pending': [], 'ready': None}
_WGRAD_STREAMS = {}
_WGRAD_GROUP = os.environ.get('INCAGG_WGRAD_GROUP', '1') != '0'   # (A/B switch)


def _flush_weight_grads(side):
    """Issue the queued GCNII weight-gradient problems: D += alpha * A^T B, 16 per launch."""
    pend, _WGRAD['pending'] = _WGRAD['pending'], []
    ready, _WGRAD['ready'] = _WGRAD['ready'], None
    if not pend:
        return
    side.wait_event(ready)
    with torch.cuda.stream(side):
        shapes = {}
        for item in pend:   # (A [K, M], B [K, N], D [M, N], alpha)
            shapes.setdefault((tuple(item[0].shape), tuple(item[1].shape)), []).append(item)
        # the largest bucket first: it is the one that must be out of the way before the tail begins
        for items in sorted(shapes.values(), key=lambda it: -len(it) * it[0][0].numel()):
            for i in range(0, len(items), 16):
                chunk = items[i:i + 16]
                ops.gemm_group([c[0] for c in chunk], [c[1] for c in chunk], [c[2] for c in chunk],
                               [c[3] for c in chunk], trans_a=True, beta=1., ws_slot=1)
    _WGRAD['keep'].append(pend)


@contextlib.contextmanager
def weight_grads_on_side_stream(device):
    device = torch.device(device)
    if device.type != 'cuda' or os.environ.get('INCAGG_WGRAD_STREAM', '1') == '0':
        yield
        return
    side = _WGRAD_STREAMS.get(device)
    if side is None:
        side = _WGRAD_STREAMS[device] = torch.cuda.Stream(device)
    _WGRAD['stream'] = side
    try:
        yield
        _flush_weight_grads(side)
    finally:
        _WGRAD['stream'] = None
        _WGRAD['pending'], _WGRAD['ready'] = [], None
        torch.cuda.current_stream(device).wait_stream(side)
        _WGRAD['keep'].clear()


class X0GradSink:
    """Where the layers of one step accumulate d loss / d x_0[:B].  Every GCNII layer reads x_0
    (gcn2.py:121), so autograd would add L gradients of [B, F] with L - 1 elementwise launches, pad the sum
    to the [B + H] rows of x_0 (fill + copy) and add it to the gradient that arrives through layer 0's
    propagation.  With a sink the input-gradient GEMM of each layer adds into one buffer in its epilogue
    (first layer of the backward pass: plain store) and returns no x_0 gradient to autograd.  The buffer
    joins the gradient that reaches x_0 through layer 0's propagation either inside the fused ReLU-backward
    / bias-gradient kernel of the Linear that produced x_0 (``Linear.forward(..., x0_sink=sink)``: no
    launch at all) or, for any other producer, in ``X0GradSink.join`` (one in-place add).  One sink per
    forward pass."""

    def __init__(self):
        self.buf = None

    def join(self, x0: Tensor) -> Tensor:
        return _X0Join.apply(x0, self)


class _X0Join(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, sink):
        ctx.sink = sink
        return x0.view_as(x0)

    @staticmethod
    def backward(ctx, g):
        buf, ctx.sink.buf = ctx.sink.buf, None
        if buf is None:
            return g, None
        if g is None:   # x_0 reached the loss only through the layers' x_0 operands
            raise RuntimeError('X0GradSink.join: x_0 must also be the input of the first layer')
        # g is the fresh output of layer 0's transposed SpMM (nobody else holds it): add in place
        g[:buf.size(0)].add_(buf)
        return g, None


def side_stream_of_weight_grads():
    """The side stream of the enclosing ``weight_grads_on_side_stream`` block (None outside one): work
    that nothing in the rest of the backward pass reads may ride on it; the block joins it on exit."""
    return _WGRAD['stream']


class _GCN2Dense(torch.autograd.Function):
    """The dense half of GCN2Conv after the propagation, fused around the tensor-core GEMM:
        s   = (1-a) h + a x0
        out = (1-b) s + b ((1-a) h W1 + a x0 W2)          (W2 = W1 when weights are shared)
    Unshared weights (the reference's products config):
      forward  ONE launch: K-concatenated GEMM  [h | x0] · [c1 W1 ; c2 W2]  with the  (1-b)s  term and the
               ReLU in the epilogue (h and x0 are the Cin operands);
      backward TWO launches (+ one split-K reduce): N-concatenated  g · [c1 W1^T | c2 W2^T]  whose
               epilogues add the  (1-b)(1-a) g  /  (1-b) a g  terms -> dh, dx0;  M-concatenated
               [h | x0]^T · g -> dW1, dW2.
    (The reference path issues ~10 elementwise / addmm launches forward and ~20 backward here.)"""

    @staticmethod
    def forward(ctx, h, x0, w1, w2, a, b, relu, out_full=None, defer_relu_bwd=False, x0_sink=None):
        # defer_relu_bwd: the ReLU is applied here, but its backward mask is left to the one consumer
        # of the output (an SpMM with relu_input=True, which gates its input gradient by [out > 0] in
        # its epilogue): this node then receives an already masked gradient.
        # out_full: a [B + H, F] buffer whose tail rows (pulled history) are filled by someone else;
        # the GEMM writes its B rows into the head and the whole buffer is the output, so the next
        # layer's input needs no concatenation.  Only the head rows carry gradient.
        rows = h.size(0)
        dst = out_full[:rows] if out_full is not None else None
        if w2 is None:
            s = torch.lerp(h, x0, a)
            out = ops.gemm(s, w1, alpha=b, cin=s, beta=1. - b, relu=relu, out=dst)
            keep = out if out_full is None else out_full   # (a saved view of a dirty base is rejected)
            ctx.save_for_backward(h, x0, w1, w2, keep if (relu and not defer_relu_bwd) else None, s)
        else:
            out = ops.gemm_dual("k", h, w1, x0, w2, scale_b=b * (1. - a), scale_b2=b * a,
                                cin=h, beta=(1. - b) * (1. - a), cin2=x0, beta2=(1. - b) * a, relu=relu,
                                out=dst)
            keep = out if out_full is None else out_full
            ctx.save_for_backward(h, x0, w1, w2, keep if (relu and not defer_relu_bwd) else None, None)
        ctx.a, ctx.b, ctx.relu, ctx.shared = a, b, (relu and not defer_relu_bwd), w2 is None
        ctx.w1_param, ctx.w2_param = w1, w2
        ctx.rows = rows
        ctx.x0_sink = x0_sink if w2 is not None else None
        if out_full is not None:
            ctx.mark_dirty(out_full)
            return out_full
        return out

    @staticmethod
    def backward(ctx, g):
        h, x0, w1, w2, out, s = ctx.saved_tensors
        a, b = ctx.a, ctx.b
        g = g[:ctx.rows].contiguous()
        if ctx.relu:
            out = out[:ctx.rows]
            g = torch.ops.aten.threshold_backward(g, out, 0.)  # ReLU backward, one kernel
        gh = gx0 = gw1 = gw2 = None
        g_ready = None
        if _WGRAD['stream'] is not None and _WGRAD_GROUP and not ctx.shared:
            # the grouped weight-gradient launch may start as soon as the LAST layer's g exists: it then
            # runs beside that layer's input-gradient GEMM and transposed SpMM and is out of the way when
            # the bandwidth-bound tail of the backward pass (first Linear) begins
            g_ready = torch.cuda.Event()
            g_ready.record(torch.cuda.current_stream(g.device))
        if ctx.shared:
            # out = (1-b) s + b s W1 ;  ds = (1-b) g + b g W1^T ; dh = (1-a) ds ; dx0 = a ds
            ds = ops.gemm(g, w1, trans_b=True, alpha=b, cin=g, beta=1. - b)
            gh, gx0 = (1. - a) * ds, a * ds
            gw1 = ops.gemm(s, g, trans_a=True, alpha=b)
        else:
            sink = ctx.x0_sink
            if sink is not None and ctx.needs_input_grad[1]:
                # d x_0 accumulates in the sink through the GEMM epilogue (first contribution: store)
                first = sink.buf is None
                if first:
                    sink.buf = torch.empty_like(g)
                gh, _ = ops.gemm_dual("n", g, w1, b2=w2, trans_b=True, scale_b=b * (1. - a), scale_b2=b * a,
                                      cin=g, beta=(1. - b) * (1. - a), cin2=g, beta2=(1. - b) * a,
                                      out2=sink.buf, acc2=not first)
            else:
                gh, gx0 = ops.gemm_dual("n", g, w1, b2=w2, trans_b=True, scale_b=b * (1. - a), scale_b2=b * a,
                                        cin=g, beta=(1. - b) * (1. - a), cin2=g, beta2=(1. - b) * a)
            b1, b2 = _grad_buffer(ctx.w1_param), _grad_buffer(ctx.w2_param)
            if b1 is not None and b2 is not None:         # accumulate into the flat gradient buffers
                side = _WGRAD['stream']
                if side is not None and _WGRAD_GROUP:
                    # queued: one grouped launch for all layers when the backward pass has been issued
                    _WGRAD['pending'].append((h, g, b1, b * (1. - a)))
                    _WGRAD['pending'].append((x0, g, b2, b * a))
                    _WGRAD['ready'] = g_ready
                elif side is not None:
                    side.wait_stream(torch.cuda.current_stream(g.device))
                    with torch.cuda.stream(side):
                        ops.gemm_dual("m", h, g, a2=x0, trans_a=True, alpha=b * (1. - a), alpha2=b * a,
                                      cin=b1, beta=1., cin2=b2, beta2=1., out=b1, out2=b2, ws_slot=1)
                    _WGRAD['keep'].append((h, x0, g))
                else:
                    ops.gemm_dual("m", h, g, a2=x0, trans_a=True, alpha=b * (1. - a), alpha2=b * a,
                                  cin=b1, beta=1., cin2=b2, beta2=1., out=b1, out2=b2)
            else:
                gw1, gw2 = ops.gemm_dual("m", h, g, a2=x0, trans_a=True, alpha=b * (1. - a), alpha2=b * a)
        return gh, gx0, gw1, gw2, None, None, None, None, None, None


class _MaskedCE(torch.autograd.Function):
    """Mean cross-entropy over the masked rows (main.py:80) with its gradient produced by the same
    fused kernels (3 launches instead of ~20 tiny ones)."""

    @staticmethod
    def forward(ctx, logits, y, mask):
        out3, dl = ops.masked_ce_raw(logits, y, mask)
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(out3)
        return out3[1], out3

    @staticmethod
    def backward(ctx, g_loss, _g_out3):
        (dl,) = ctx.saved_tensors
        return dl * g_loss, None, None


def masked_cross_entropy(logits: Tensor, y: Tensor, mask: Tensor):
    """-> (mean CE over rows with mask, [loss sum, mean, count]) ; logits [B, C] fp32, y int64."""
    return _MaskedCE.apply(logits, y, mask)


def masked_cross_entropy_grad(logits: Tensor, y: Tensor, mask: Tensor):
    """-> ([loss sum, mean, count], d mean-loss / d logits) without an autograd node: the training loop
    feeds the gradient straight into ``logits.backward`` (``loss.backward()`` would first materialise a
    ones tensor and multiply the gradient by it: two launches for nothing)."""
    return ops.masked_ce_raw(logits.detach(), y, mask)


def glorot_(w: Tensor) -> Tensor:
    a = math.sqrt(6.0 / (w.size(-2) + w.size(-1)))
    with torch.no_grad():
        return w.uniform_(-a, a)


class GCNConv(torch.nn.Module):
    """out = A (x W) + b   (PyG GCNConv with normalize=False; reference gcn.py:63)."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False, bias: bool = True):
        super().__init__()
        if normalize:
            raise NotImplementedError('normalisation is applied once to the graph (main.py:151)')
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = Linear(in_channels, out_channels, bias=False)
        self.bias = Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.lin.weight)
        if self.bias is not None:
            torch.nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, adj_t: SparseTensor, grad_rows: Optional[int] = None) -> Tensor:
        x = self.lin(x)
        out = spmm(adj_t, x, reduce='sum', grad_rows=grad_rows)
        if self.bias is not None:
            out = out + self.bias
        return out


class GCN2Conv(torch.nn.Module):
    """GCNII layer (PyG GCN2Conv, normalize=False) with the reference's two extra entry points."""

    def __init__(self, channels: int, alpha: float, theta: float = None, layer: int = None,
                 shared_weights: bool = True, normalize: bool = False):
        super().__init__()
        if normalize:
            raise NotImplementedError('normalisation is applied once to the graph (main.py:151)')
        self.channels = channels
        self.alpha = alpha
        self.beta = 1.
        if theta is not None or layer is not None:
            assert theta is not None and layer is not None
            self.beta = math.log(theta / layer + 1)
        self.weight1 = Parameter(torch.empty(channels, channels))
        if shared_weights:
            self.register_parameter('weight2', None)
        else:
            self.weight2 = Parameter(torch.empty(channels, channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight1)
        if self.weight2 is not None:
            glorot_(self.weight2)

    def forward_after_propagate(self, h: Tensor, x_0: Tensor, relu: bool = False,
                                out_full: Optional[Tensor] = None, defer_relu_bwd: bool = False,
                                x0_sink: Optional[X0GradSink] = None) -> Tensor:
        """Everything of PyG's GCN2Conv.forward after ``propagate``:
            x = (1-alpha) h ; x_0 = alpha x_0[:B]
            shared:   out = x + x_0 ; out = (1-beta) out + beta out W1
            unshared: out = (1-beta) x + beta x W1 + (1-beta) x_0 + beta x_0 W2
        evaluated by the fused tensor-core block (``relu=True`` also applies the activation that
        follows in the models when there is no batch norm / residual in between)."""
        if x_0.size(0) != h.size(0):  # (a no-op slice would still cost a zero-fill + copy in backward)
            x_0 = x_0[:h.size(0)]
        return _GCN2Dense.apply(h.contiguous(), x_0.contiguous(), self.weight1, self.weight2,
                                float(self.alpha), float(self.beta), relu, out_full,
                                bool(defer_relu_bwd and relu), x0_sink)

    def forward_no_neighbor(self, x: Tensor, x_0: Tensor, relu: bool = False,
                            x0_sink: Optional[X0GradSink] = None, defer_relu_bwd: bool = False) -> Tensor:
        return self.forward_after_propagate(x, x_0, relu, None, defer_relu_bwd, x0_sink)

    def forward(self, x: Tensor, x_0: Tensor, adj_t: SparseTensor,
                grad_rows: Optional[int] = None, relu: bool = False,
                out_full: Optional[Tensor] = None, relu_input: bool = False,
                defer_relu_bwd: bool = False, x0_sink: Optional[X0GradSink] = None) -> Tensor:
        """relu_input: x is the ReLU output of a layer called with defer_relu_bwd=True (the two flags
        come in pairs: producer defers, this layer's SpMM applies the mask in its backward)."""
        h = spmm(adj_t, x, reduce='sum', grad_rows=grad_rows, relu_input=relu_input)
        return self.forward_after_propagate(h, x_0, relu, out_full, defer_relu_bwd, x0_sink)


class SAGEConv(torch.nn.Module):
    """out = lin_l(mean_j x_j) + lin_r(x_root)   (PyG SAGEConv, aggr='mean', normalize=False)."""

    def __init__(self, in_channels: int, out_channels: int, normalize: bool = False,
                 root_weight: bool = True, bias: bool = True, aggr: str = 'mean'):
        super().__init__()
        if normalize:
            raise NotImplementedError
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggr = aggr
        self.normalize = normalize
        self.root_weight = root_weight
        self.project = False
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        self.lin_r = Linear(in_channels, out_channels, bias=False) if root_weight else None

    def reset_parameters(self):
        self.lin_l.reset_parameters()
        if self.lin_r is not None:
            self.lin_r.reset_parameters()

    def forward(self, x: Tensor, adj_t: SparseTensor, grad_rows: Optional[int] = None) -> Tensor:
        adj = adj_t.set_value(None) if adj_t.value is not None else adj_t
        out = spmm(adj, x, reduce=self.aggr, grad_rows=grad_rows)
        out = self.lin_l(out)
        if self.lin_r is not None:
            out = out + self.lin_r(x[:adj_t.size(0)])
        return out
