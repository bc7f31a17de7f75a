"""Tensor-level wrappers over the C ABI (include/incagg_b200.h) plus the autograd functions and
the re-registration of the reference's operator names.

Every function here launches the hand-written sm_100a kernels on the current CUDA stream.  Nothing
falls back to PyTorch or to the CPU oracle: non-CUDA inputs raise.

Reference op names re-registered (csrc/relabel.cpp:40-42, csrc/async.cpp:44-48):
    torch.ops.torch_geometric_autoscale.{relabel_one_hop, relabel_one_hop_within_batch,
                                         read_async, write_async, synchronize}
"""
from typing import Optional, Tuple

import os

import torch
from torch import Tensor

from . import _lib
from ._lib import lib, check, ptr, REDUCE

# Count of our kernel-launching C-ABI calls (bench.py reports it as `gpu_launches`).
LAUNCHES = {"calls": 0}


_raw_stream = torch._C._cuda_getCurrentRawStream
_cur_dev = torch._C._cuda_getDevice


def _stream() -> int:
    # raw handle of the current stream of the current device (the Python Stream object of
    # torch.cuda.current_stream() costs ~15 us per call, which matters at ~60 launches per step)
    return _raw_stream(_cur_dev())


def check_device_errors(reset: bool = True) -> None:
    """Raises RuntimeError if a kernel met an index outside its table since the last check (the
    reference raises at the faulty index_select / index_put; kernels record it in a device-side word,
    write zeros for a gathered row and skip a scattered one).  Synchronises with the device."""
    import ctypes
    word = ctypes.c_int32(0)
    check(lib.incagg_device_errors(ctypes.cast(ctypes.pointer(word), ctypes.c_void_p), int(reset)))
    if word.value:
        what = []
        if word.value & 1:
            what.append("gather / scatter row index outside the table")
        if word.value & 2:
            what.append("relabel: batch node id outside [0, num_nodes)")
        if word.value & 4:
            what.append("fused gradient all-reduce: a peer rank did not arrive within ~10 s")
        raise RuntimeError("incagg_b200 device error: " + "; ".join(what))


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("incagg_b200: expected a CUDA tensor (there is no CPU fallback)")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"incagg_b200: expected float32, got {t.dtype}")
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.stride(-1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))


# --------------------------------------------------------------------------------------------
# SpMM
# --------------------------------------------------------------------------------------------
TUNE = {"spmm_stream_variant": 0, "spmm_stream_min_f": 1, "gemm_dual_m_ctas": 2, "gemm_bn64_min_tiles": 3}


def tune(key: str, value: int) -> None:
    """Experiment knobs of the kernels (include/incagg_b200.h INCAGG_TUNE_*).  Plans built before a
    knob that changes the warp partition was set must be rebuilt (SparseTensor.drop_caches)."""
    check(lib.incagg_tune_set(TUNE[key], int(value)))


for _kv in filter(None, os.environ.get("INCAGG_TUNE", "").split(",")):   # INCAGG_TUNE="key=value,..." (experiments)
    tune(_kv.split("=")[0].strip(), int(_kv.split("=")[1]))


# The SpMM kernels keep the partial sums of rows that are split over several CTAs / warps in one
# scratch area per device, so two SpMM launches must not run concurrently.  Launches on one stream
# serialise by themselves; when a call arrives on another stream than the previous one, the new
# stream first waits for everything issued so far on the old one.  (Inside a CUDA-graph capture the
# step's SpMMs are all issued on the capturing stream.)
_SPMM_LAST = {}


def _order_spmm() -> None:
    dev = _cur_dev()
    st = _raw_stream(dev)
    last = _SPMM_LAST.get(dev)
    if last is not None and last != st and not torch.cuda.is_current_stream_capturing():
        ev = torch.cuda.Event()
        ev.record(torch.cuda.ExternalStream(last, device=dev))
        torch.cuda.current_stream(dev).wait_event(ev)
    _SPMM_LAST[dev] = st


def spmm_plan(rowptr: Tensor, rows: Optional[int] = None, nnz: int = -1) -> Tensor:
    """Degree-bucket plan of a CSR structure (built once, reused by every SpMM over it)."""
    _require_cuda(rowptr)
    assert rowptr.dtype == torch.int32
    n_rows = rowptr.numel() - 1 if rows is None else rows
    nbytes = lib.incagg_spmm_plan_bytes(n_rows, nnz)
    plan = torch.empty(nbytes, dtype=torch.uint8, device=rowptr.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_spmm_plan(ptr(rowptr), n_rows, nnz, ptr(plan), nbytes, _stream()))
    return plan


def spmm_raw(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, reduce: str = "sum",
             rows: Optional[int] = None, out: Optional[Tensor] = None,
             return_arg: bool = False, plan: Optional[Tensor] = None, gate: Optional[Tensor] = None):
    """out[i] = reduce_e val[e] * x[col[e]] over a CSR with int32 rowptr/col (device).
    `gate` ([rows, F] float32, sum / mean only): out is zeroed where gate <= 0 in the kernel's epilogue."""
    _require_cuda(rowptr, col, val, x)
    assert rowptr.dtype == torch.int32 and col.dtype == torch.int32
    squeeze = x.dim() == 1
    x = _f32c(x)
    n_rows = rowptr.numel() - 1 if rows is None else rows
    F = x.size(1)
    if out is None:
        out = torch.empty((n_rows, F), dtype=torch.float32, device=x.device)
    arg = None
    if return_arg and reduce in ("min", "max"):
        arg = torch.empty((n_rows, F), dtype=torch.int32, device=x.device)
    LAUNCHES["calls"] += 1
    _order_spmm()
    if gate is not None:
        if reduce not in ("sum", "mean") or gate.dtype != torch.float32 or gate.dim() != 2 \
                or gate.stride(1) != 1 or gate.size(0) < n_rows or gate.size(1) < F:
            raise RuntimeError("spmm: gate must be a row-major float32 [rows, >= F] tensor (sum / mean)")
        _require_cuda(gate)
        check(lib.incagg_spmm_csr_gated(REDUCE[reduce], ptr(rowptr), ptr(col), ptr(val), ptr(x), _ld(x),
                                        ptr(out), _ld(out), n_rows, F, ptr(plan), ptr(gate), gate.stride(0),
                                        _stream()))
        return out.squeeze(1) if squeeze else out
    check(lib.incagg_spmm_csr(REDUCE[reduce], ptr(rowptr), ptr(col), ptr(val), ptr(x), _ld(x),
                              ptr(out), _ld(out), ptr(arg), F if arg is not None else 0, n_rows, F,
                              ptr(plan), _stream()))
    if squeeze:
        out = out.squeeze(1)
    return (out, arg) if return_arg else out


def spmm_delta_raw(rowptr, col, val, x, m_in, m_ag, n_id=None, reduce="sum", rows=None, out=None,
                   plan=None):
    """out = A (x - M_in[g]) + M_ag[g]; g = n_id if given (history tables read in place)."""
    _require_cuda(rowptr, col, val, x, m_in, m_ag, n_id)
    x = _f32c(x)
    n_rows = rowptr.numel() - 1 if rows is None else rows
    F = x.size(1)
    assert m_in.dtype == torch.float32 and m_ag.dtype == torch.float32
    assert m_in.stride(1) == 1 and m_ag.stride(1) == 1 and m_in.size(1) >= F and m_ag.size(1) >= F
    if n_id is not None:
        assert n_id.dtype == torch.int64 and n_id.is_contiguous()
    if out is None:
        out = torch.empty((n_rows, F), dtype=torch.float32, device=x.device)
    LAUNCHES["calls"] += 1
    _order_spmm()
    check(lib.incagg_spmm_delta(REDUCE[reduce], ptr(rowptr), ptr(col), ptr(val), ptr(x), _ld(x),
                                ptr(m_in), m_in.stride(0), ptr(m_ag), m_ag.stride(0), ptr(n_id),
                                ptr(out), _ld(out), n_rows, F, ptr(plan), _stream()))
    return out


def spmm_multi_raw(rowptr, col, val, x, F: int, reducers, rows=None, out=None, plan=None, return_arg=False):
    """x is [n_src, K*F]; slab k reduced with reducers[k] ('sum'|'mean'|'min'|'max').  return_arg: also
    the winning edge of every min / max column ([rows, K*F] int32; -1 elsewhere)."""
    _require_cuda(rowptr, col, val, x)
    x = _f32c(x)
    K = len(reducers)
    assert x.size(1) == K * F
    n_rows = rowptr.numel() - 1 if rows is None else rows
    if out is None:
        out = torch.empty((n_rows, K * F), dtype=torch.float32, device=x.device)
    import ctypes
    red = (ctypes.c_int32 * K)(*[REDUCE[r] for r in reducers])
    LAUNCHES["calls"] += 1
    _order_spmm()
    if return_arg:
        arg = torch.empty((n_rows, K * F), dtype=torch.int32, device=x.device)
        check(lib.incagg_spmm_multi_arg(ptr(rowptr), ptr(col), ptr(val), ptr(x), _ld(x), ptr(out), _ld(out),
                                        ptr(arg), K * F, n_rows, F, K, ctypes.cast(red, ctypes.c_void_p), ptr(plan),
                                        _stream()))
        return out, arg
    check(lib.incagg_spmm_multi(ptr(rowptr), ptr(col), ptr(val), ptr(x), _ld(x), ptr(out), _ld(out),
                                n_rows, F, K, ctypes.cast(red, ctypes.c_void_p), ptr(plan), _stream()))
    return out


def spmm_minmax_bwd_raw(col, val, arg, grad_out, n_src: int, out: Optional[Tensor] = None):
    """Routes min / max gradients to the winning source rows.  `out` ([n_src, F] view, row-major,
    ZERO-filled by the caller): accumulate there instead of into a fresh tensor."""
    grad_out = _f32c(grad_out)
    F = grad_out.size(1)
    grad_x = out if out is not None else torch.zeros((n_src, F), dtype=torch.float32, device=grad_out.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_spmm_minmax_bwd(ptr(col), ptr(val), ptr(arg), arg.stride(0), ptr(grad_out),
                                     _ld(grad_out), ptr(grad_x), _ld(grad_x), grad_out.size(0), F,
                                     _stream()))
    return grad_x


def csr_transpose(rowptr: Tensor, col: Tensor, val: Optional[Tensor], rows: int, cols: int,
                  want_perm: bool = False):
    """CSR of A^T (deterministic entry order). Returns (t_rowptr, t_col, t_val, t_perm|None)."""
    _require_cuda(rowptr, col, val)
    assert rowptr.dtype == torch.int32 and col.dtype == torch.int32
    nnz = col.numel()
    dev = rowptr.device
    t_rowptr = torch.empty(cols + 1, dtype=torch.int32, device=dev)
    t_col = torch.empty(nnz, dtype=torch.int32, device=dev)
    t_val = torch.empty(nnz, dtype=torch.float32, device=dev) if val is not None else None
    t_perm = torch.empty(nnz, dtype=torch.int32, device=dev) if want_perm else None
    ws_bytes = lib.incagg_csr_transpose_workspace_bytes(rows, cols, nnz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    LAUNCHES["calls"] += 1
    check(lib.incagg_csr_transpose(ptr(rowptr), ptr(col), ptr(val), rows, cols, nnz, ptr(t_rowptr),
                                   ptr(t_col), ptr(t_val), ptr(t_perm), ptr(ws), ws_bytes, _stream()))
    return t_rowptr, t_col, t_val, t_perm


# --------------------------------------------------------------------------------------------
# dense X·W on the tensor cores (tcgen05, 3xTF32)
# --------------------------------------------------------------------------------------------
_GEMM_WS = {}
_GEMM_WS_RETIRED = []


def _gemm_workspace(device, nbytes: int, slot: int = 0) -> Tensor:
    """Split-K scratch, one per (device, slot): GEMMs issued on different streams use different slots."""
    ws = _GEMM_WS.get((device, slot))
    if ws is None or ws.numel() < nbytes:
        # Always the 256 MB cap the callers clamp to: the buffer is allocated once (by an eager warm-up
        # step) and never replaced, so the address a captured CUDA graph has baked in stays valid.  A
        # buffer that had to be replaced all the same is kept alive for the graphs that reference it.
        if ws is not None:
            _GEMM_WS_RETIRED.append(ws)
        ws = _GEMM_WS[(device, slot)] = torch.empty(max(nbytes, 1 << 28), dtype=torch.uint8, device=device)
    return ws


def _rowmajor(t: Tensor) -> Tensor:
    if t.dim() != 2 or t.dtype != torch.float32:
        raise RuntimeError("gemm: expected 2-D float32 tensors")
    if t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def gemm(a: Tensor, b: Tensor, trans_a: bool = False, trans_b: bool = False, alpha: float = 1.0,
         cin: Optional[Tensor] = None, beta: float = 0.0, bias: Optional[Tensor] = None,
         relu: bool = False, out: Optional[Tensor] = None, ws_slot: int = 0,
         gate: Optional[Tensor] = None) -> Tensor:
    """out[M,N] = alpha * op(a) @ op(b) + beta * cin + bias (+ReLU) with the tcgen05 3xTF32 kernel.
    trans_a: `a` is stored [K, M]; trans_b: `b` is stored [N, K] (a Linear weight).  ws_slot: which
    split-K scratch buffer to use (GEMMs issued on a side stream must not share the main stream's).
    gate ([M, N]): out is zeroed where gate <= 0 (ReLU backward in the epilogue; excludes cin / bias / relu)."""
    if gate is not None:
        assert cin is None and bias is None and not relu
        cin = gate
    _require_cuda(a, b, cin, bias, out)
    a, b = _rowmajor(a), _rowmajor(b)
    M, K = (a.size(1), a.size(0)) if trans_a else (a.size(0), a.size(1))
    N, Kb = (b.size(0), b.size(1)) if trans_b else (b.size(1), b.size(0))
    if K != Kb:
        raise RuntimeError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if cin is not None:
        cin = _rowmajor(cin)
    ws_bytes = 0
    ws = None
    if K >= 512 and M * N <= (1 << 20):  # long reduction, few output tiles: split-K partials
        ws_bytes = min(lib.incagg_gemm_workspace_bytes(M, N, K), 1 << 28)
        ws = _gemm_workspace(a.device, ws_bytes, ws_slot)
        ws_bytes = ws.numel()
    LAUNCHES["calls"] += 1
    check(lib.incagg_gemm_tf32x3(int(trans_a), int(trans_b), M, N, K, ptr(a), _ld(a), ptr(b), _ld(b),
                                 float(alpha), ptr(cin), _ld(cin) if cin is not None else 0, float(beta),
                                 ptr(bias), int(bool(relu)) | (4 if gate is not None else 0), ptr(out), _ld(out),
                                 ptr(ws), ws_bytes, _stream()))
    return out


def gemm_dual(mode: str, a: Tensor, b: Tensor, a2: Optional[Tensor] = None, b2: Optional[Tensor] = None,
              trans_a: bool = False, trans_b: bool = False, alpha: float = 1.0, alpha2: float = 1.0,
              scale_b: float = 1.0, scale_b2: float = 1.0, cin: Optional[Tensor] = None, beta: float = 0.0,
              cin2: Optional[Tensor] = None, beta2: float = 0.0, relu: bool = False,
              out: Optional[Tensor] = None, out2: Optional[Tensor] = None, ws_slot: int = 0,
              acc2: bool = False):
    """Two GEMMs sharing an operand in one launch (include/incagg_b200.h, incagg_gemm_tf32x3_dual):
    mode 'k': out = alpha (a @ (scale_b b) + a2 @ (scale_b2 b2)) + beta cin + beta2 cin2
    mode 'n': out = alpha a @ (scale_b b) + beta cin ; out2 = alpha2 a @ (scale_b2 b2) + beta2 cin2
              (acc2: out2 += ... instead of out2 = ..., the gradient accumulation of a shared input)
    mode 'm': out = alpha op(a) @ b ; out2 = alpha2 op(a2) @ b      (split-K)."""
    code = {"k": 1, "n": 2, "m": 3}[mode]
    assert not acc2 or (mode == "n" and out2 is not None)
    a, b = _rowmajor(a), _rowmajor(b)
    a2 = _rowmajor(a2) if a2 is not None else None
    b2 = _rowmajor(b2) if b2 is not None else None
    M, K = (a.size(1), a.size(0)) if trans_a else (a.size(0), a.size(1))
    N = b.size(0) if trans_b else b.size(1)
    K2 = 0
    if mode == "k":
        K2 = a2.size(0) if trans_a else a2.size(1)
    dev = a.device
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    if mode in ("n", "m") and out2 is None:
        out2 = torch.empty((M, N), dtype=torch.float32, device=dev)
    cin = _rowmajor(cin) if cin is not None else None
    cin2 = _rowmajor(cin2) if cin2 is not None else None
    ws, ws_bytes = None, 0
    if mode == "m":
        ws_bytes = min(lib.incagg_gemm_workspace_bytes(M, N, K), 1 << 28)
        ws = _gemm_workspace(dev, ws_bytes, ws_slot)
        ws_bytes = ws.numel()
    LAUNCHES["calls"] += 1
    check(lib.incagg_gemm_tf32x3_dual(
        code, int(trans_a), int(trans_b), M, N, K, K2, ptr(a), _ld(a), ptr(a2), _ld(a2) if a2 is not None else 0,
        ptr(b), _ld(b), ptr(b2), _ld(b2) if b2 is not None else 0, float(alpha), float(alpha2), float(scale_b),
        float(scale_b2), ptr(cin), _ld(cin) if cin is not None else 0, float(beta), ptr(cin2),
        _ld(cin2) if cin2 is not None else 0, float(beta2), int(bool(relu)) | (2 if acc2 else 0), ptr(out), _ld(out), ptr(out2),
        _ld(out2) if out2 is not None else 0, ptr(ws), ws_bytes, _stream()))
    return (out, out2) if mode in ("n", "m") else out


def gemm_group(a_list, b_list, outs, alphas, trans_a: bool = False, trans_b: bool = False, beta: float = 0.0,
               ws_slot: int = 0) -> None:
    """outs[g] = alphas[g] * op(a_list[g]) @ op(b_list[g]) + beta * outs[g] for <= 16 problems of one shape,
    ONE launch (include/incagg_b200.h, incagg_gemm_tf32x3_group); split-K for long reductions."""
    import ctypes
    n = len(a_list)
    assert 1 <= n <= 16 and len(b_list) == n and len(outs) == n and len(alphas) == n
    _require_cuda(*a_list, *b_list, *outs)
    a_list = [_rowmajor(t) for t in a_list]
    b_list = [_rowmajor(t) for t in b_list]
    a, b = a_list[0], b_list[0]
    M, K = (a.size(1), a.size(0)) if trans_a else (a.size(0), a.size(1))
    N = b.size(0) if trans_b else b.size(1)
    for t, u, o in zip(a_list, b_list, outs):
        assert t.shape == a.shape and u.shape == b.shape and tuple(o.shape) == (M, N)
        assert o.dtype == torch.float32 and o.stride(1) == 1
    PtrArr, LdArr, FArr = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_float * n
    ws_bytes = min(lib.incagg_gemm_workspace_bytes(M * n, N, K), 1 << 28)
    ws = _gemm_workspace(a.device, ws_bytes, ws_slot)
    LAUNCHES["calls"] += 1
    check(lib.incagg_gemm_tf32x3_group(
        n, int(trans_a), int(trans_b), M, N, K,
        PtrArr(*[t.data_ptr() for t in a_list]), LdArr(*[_ld(t) for t in a_list]),
        PtrArr(*[t.data_ptr() for t in b_list]), LdArr(*[_ld(t) for t in b_list]),
        FArr(*[float(x) for x in alphas]), float(beta),
        PtrArr(*[t.data_ptr() for t in outs]), LdArr(*[_ld(t) for t in outs]), ptr(ws), ws.numel(), _stream()))


def relu_bwd_colsum(g: Tensor, y: Optional[Tensor] = None, add: Optional[Tensor] = None,
                    colsum_into: Optional[Tensor] = None):
    """((g [+ add on its first rows]) * (y > 0), its column sums) in one pass; y=None: no mask.  float4 path
    for 16-byte-aligned rows with cols % 4 == 0, scalar columns otherwise (<= 256 columns:
    colsum_supported).  colsum_into: the column sums are ADDED to this buffer (a bias gradient view of the
    flat gradient buffer) instead of being returned."""
    _require_cuda(g, y, add, colsum_into)
    g = _rowmajor(g)
    rows, cols = g.shape
    gm = None
    if y is not None or add is not None:
        y = _rowmajor(y) if y is not None else None
        add = _rowmajor(add) if add is not None else None
        gm = torch.empty_like(g)
    if colsum_into is not None:
        assert colsum_into.dtype == torch.float32 and colsum_into.is_contiguous() and colsum_into.numel() == cols
        colsum = colsum_into
    else:
        colsum = torch.empty(cols, dtype=torch.float32, device=g.device)
    nb = lib.incagg_colsum_workspace_bytes(rows, cols)
    ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=g.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_relu_bwd_colsum_ex(ptr(g), _ld(g), ptr(y), _ld(y) if y is not None else 0, rows, cols,
                                        ptr(gm), _ld(gm) if gm is not None else 0, ptr(add),
                                        _ld(add) if add is not None else 0, add.size(0) if add is not None else 0,
                                        ptr(colsum), int(colsum_into is not None), ptr(ws), ws.numel(), _stream()))
    return (gm if gm is not None else g), (None if colsum_into is not None else colsum)


def colsum_supported(t: Tensor) -> bool:
    if not (t.dim() == 2 and t.dtype == torch.float32 and t.stride(1) == 1 and 0 < t.size(1) <= 1024):
        return False
    vec = t.size(1) % 4 == 0 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0
    return vec or t.size(1) <= 256


def masked_ce_raw(logits: Tensor, y: Tensor, mask: Tensor):
    """-> (out3 = [loss sum, mean loss, count], dlogits) for the rows with mask != 0."""
    _require_cuda(logits, y, mask)
    logits = _rowmajor(logits)
    assert y.dtype == torch.int64 and y.is_contiguous()
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
    m8 = m8.contiguous()
    rows, C = logits.shape
    dl = torch.empty_like(logits)
    out3 = torch.empty(3, dtype=torch.float32, device=logits.device)
    nb = lib.incagg_masked_ce_workspace_bytes(rows)
    ws = torch.empty(nb, dtype=torch.uint8, device=logits.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_masked_ce(ptr(logits), _ld(logits), ptr(y), ptr(m8), rows, C, ptr(dl), _ld(dl), ptr(out3),
                               ptr(ws), nb, _stream()))
    return out3, dl


def _mask_bytes(mask: Tensor) -> Tensor:
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
    return m8.contiguous()


def mask_count(mask: Tensor) -> Tensor:
    """Number of selected rows as a device float[1] (first third of masked_ce_raw)."""
    _require_cuda(mask)
    m8 = _mask_bytes(mask)
    count = torch.empty(1, dtype=torch.float32, device=mask.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_mask_count(ptr(m8), m8.numel(), ptr(count), _stream()))
    return count


def masked_ce_rows(logits: Tensor, y: Tensor, mask: Tensor, count: Tensor):
    """-> (dlogits, workspace holding the per-block loss partials) given the row count (mask_count)."""
    _require_cuda(logits, y, mask, count)
    logits = _rowmajor(logits)
    assert y.dtype == torch.int64 and y.is_contiguous() and count.dtype == torch.float32
    m8 = _mask_bytes(mask)
    rows, C = logits.shape
    dl = torch.empty_like(logits)
    nb = lib.incagg_masked_ce_workspace_bytes(rows)
    ws = torch.empty(nb, dtype=torch.uint8, device=logits.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_masked_ce_rows(ptr(logits), _ld(logits), ptr(y), ptr(m8), rows, C, ptr(count), ptr(dl), _ld(dl),
                                    ptr(ws), nb, _stream()))
    return dl, ws


def masked_ce_finish(ws: Tensor, rows: int, count: Tensor, acc: Optional[Tensor] = None) -> Tensor:
    """-> out3 = [loss sum, mean loss, count]; acc (float64[2], optional): acc += [loss sum, count]."""
    _require_cuda(ws, count, acc)
    assert acc is None or (acc.dtype == torch.float64 and acc.numel() >= 2 and acc.is_contiguous())
    out3 = torch.empty(3, dtype=torch.float32, device=ws.device)
    LAUNCHES["calls"] += 1
    check(lib.incagg_masked_ce_finish(ptr(ws), int(rows), ptr(count), ptr(out3), ptr(acc), _stream()))
    return out3


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, n_first: int, lr: float, beta1: float, beta2: float,
              eps: float, wd_first: float, wd_rest: float, step: Tensor, arrivals: Tensor) -> None:
    """One fused Adam update over flat fp32 buffers (include/incagg_b200.h, incagg_adam_step)."""
    _require_cuda(p, g, m, v, step, arrivals)
    LAUNCHES["calls"] += 1
    check(lib.incagg_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(n_first), float(lr), float(beta1),
                               float(beta2), float(eps), float(wd_first), float(wd_rest), ptr(step), ptr(arrivals),
                               _stream()))


def allreduce_adam_step(stage_views, signal_views, rank: int, g: Tensor, p: Tensor, m: Tensor, v: Tensor,
                        n_first: int, lr: float, beta1: float, beta2: float, eps: float, wd_first: float,
                        wd_rest: float, step: Tensor, arrivals: Tensor) -> None:
    """Gradient all-reduce over NVLink peer memory fused with the Adam update (incagg_allreduce_adam_step).
    stage_views[r] / signal_views[r]: rank r's staging / signal tensors (peers opened through CUDA IPC)."""
    import ctypes
    _require_cuda(g, p, m, v, step, arrivals, *stage_views, *signal_views)
    W = len(stage_views)
    st = (ctypes.c_void_p * W)(*[t.data_ptr() for t in stage_views])
    sg = (ctypes.c_void_p * W)(*[t.data_ptr() for t in signal_views])
    LAUNCHES["calls"] += 1
    check(lib.incagg_allreduce_adam_step(ctypes.cast(st, ctypes.c_void_p), ctypes.cast(sg, ctypes.c_void_p),
                                         int(rank), W, ptr(g), ptr(p), ptr(m), ptr(v), p.numel(), int(n_first),
                                         float(lr), float(beta1), float(beta2), float(eps), float(wd_first),
                                         float(wd_rest), ptr(step), ptr(arrivals), _stream()))


# --------------------------------------------------------------------------------------------
# rows: gather / scatter / slices
# --------------------------------------------------------------------------------------------
def _row_view(t: Tensor):
    """(pointer, leading dimension in bytes, row bytes, rows) of a 1-D/2-D row-major tensor."""
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.dim() > 2:
        t = t.reshape(t.size(0), -1)
    assert t.stride(1) == 1 or t.size(1) == 1, "rows must be contiguous"
    es = t.element_size()
    ld = (t.stride(0) if t.size(0) > 1 else t.size(1)) * es
    return t, ld, t.size(1) * es, t.size(0)


def _device_accessible(t: Tensor) -> bool:
    return t.is_cuda or t.is_pinned()


def gather_rows(src: Tensor, idx: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """out[i] = src[idx[i]].  src: CUDA or pinned-host tensor; idx: CUDA int64; out: CUDA."""
    if not _device_accessible(src):
        raise RuntimeError("gather_rows: src must be a CUDA or pinned host tensor")
    _require_cuda(idx)
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    n = idx.numel()
    if out is None:
        out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=idx.device)
    _require_cuda(out)
    s, s_ld, row_bytes, s_rows = _row_view(src)
    d, d_ld, d_row_bytes, d_rows = _row_view(out)
    assert d_row_bytes == row_bytes and d_rows >= n
    LAUNCHES["calls"] += 1
    check(lib.incagg_gather_rows(ptr(s), s_ld, s_rows, ptr(idx), n, ptr(d), d_ld, row_bytes, _stream()))
    return out


def gather_rows_sharded(shards, bounds, idx: Tensor, out: Tensor) -> Tensor:
    """out[i] = table[idx[i]] where the table is split by rows over `shards` (shard s holds global rows
    [bounds[s], bounds[s+1])); shards may be peer-GPU memory mapped through CUDA IPC."""
    import ctypes
    _require_cuda(idx, out)
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    W = len(shards)
    ref = next(t for t in shards if t is not None and t.numel() > 0)
    es = ref.element_size()
    ld = ref.stride(0) * es if ref.dim() > 1 else es
    row_bytes = (ref.size(1) if ref.dim() > 1 else 1) * es
    d, d_ld, d_row_bytes, d_rows = _row_view(out)
    assert d_row_bytes == row_bytes and d_rows >= idx.numel()
    ptrs = (ctypes.c_void_p * W)(*[(t.data_ptr() if t is not None and t.numel() > 0 else None) for t in shards])
    bnd = (ctypes.c_int64 * (W + 1))(*[int(b) for b in bounds])
    LAUNCHES["calls"] += 1
    check(lib.incagg_gather_rows_sharded(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(bnd, ctypes.c_void_p),
                                         W, ld, ptr(idx), idx.numel(), ptr(d), d_ld, row_bytes, _stream()))
    return out


def scatter_rows(src: Tensor, idx: Tensor, dst: Tensor) -> None:
    """dst[idx[i]] = src[i].  dst: CUDA or pinned-host tensor."""
    if not _device_accessible(dst):
        raise RuntimeError("scatter_rows: dst must be a CUDA or pinned host tensor")
    _require_cuda(src, idx)
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    n = idx.numel()
    s, s_ld, row_bytes, s_rows = _row_view(src)
    d, d_ld, d_row_bytes, d_rows = _row_view(dst)
    assert d_row_bytes == row_bytes and s_rows >= n
    LAUNCHES["calls"] += 1
    check(lib.incagg_scatter_rows(ptr(s), s_ld, ptr(idx), n, ptr(d), d_ld, d_rows, row_bytes, _stream()))


def copy_slices(src: Tensor, dst: Tensor, offset, count, direction: int) -> int:
    """direction 0 (pull): dst packed <- src[offset_i : offset_i + count_i];
    direction 1 (push): dst[offset_i : +count_i] <- src packed.  offset / count: host int64
    tensors or python lists.  Returns the number of packed rows."""
    import ctypes
    if not (_device_accessible(src) and _device_accessible(dst)):
        raise RuntimeError("copy_slices: tensors must be CUDA or pinned host tensors")
    off = offset.tolist() if isinstance(offset, Tensor) else list(offset)
    cnt = count.tolist() if isinstance(count, Tensor) else list(count)
    if len(off) != len(cnt):
        raise RuntimeError("Size mismatch")
    k = len(off)
    s, s_ld, row_bytes, s_rows = _row_view(src)
    d, d_ld, d_row_bytes, d_rows = _row_view(dst)
    # the packed side may be wider than the strided side's crop: rows are min width
    rb = min(row_bytes, d_row_bytes)
    o_arr = (ctypes.c_int64 * k)(*off)
    c_arr = (ctypes.c_int64 * k)(*cnt)
    LAUNCHES["calls"] += 1
    check(lib.incagg_copy_slices(ptr(s), s_ld, s_rows, ptr(d), d_ld, d_rows,
                                 ctypes.cast(o_arr, ctypes.c_void_p),
                                 ctypes.cast(c_arr, ctypes.c_void_p), k, rb, direction, _stream()))
    return int(sum(cnt))


# --------------------------------------------------------------------------------------------
# relabel
# --------------------------------------------------------------------------------------------
class RelabelWorkspace:
    """Direct-address table over global node ids used by the GPU relabel (one per graph/device).
    Not shareable between concurrent relabel calls."""

    def __init__(self, num_nodes: int, device):
        self.num_nodes = int(num_nodes)
        self.device = torch.device(device)
        nbytes = lib.incagg_relabel_workspace_bytes(self.num_nodes)
        self.buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        self.counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.counts_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        self.last_count = None  # data-dependent output size of the last relabel call
        with torch.cuda.device(self.device):
            check(lib.incagg_relabel_workspace_init(ptr(self.buf), self.num_nodes, _stream()))


_WS_CACHE = {}


def _workspace_for(num_nodes: int, device) -> RelabelWorkspace:
    key = (int(num_nodes), str(device))
    ws = _WS_CACHE.get(key)
    if ws is None:
        ws = _WS_CACHE[key] = RelabelWorkspace(num_nodes, device)
    return ws


def _relabel(within: bool, rowptr: Tensor, col: Tensor, value: Optional[Tensor], idx: Tensor,
             bipartite: bool, ws: Optional[RelabelWorkspace], out_int32: bool,
             nnz_b: Optional[int], known: Optional[int] = None, window=None):
    """`known`: the data-dependent output size of this very batch from an earlier call (number of
    halo ids for relabel_one_hop, number of kept edges for the within-batch variant).  With it the call
    needs no device->host readback, i.e. no synchronisation (and can be captured in a CUDA graph)."""
    # the global CSR may live in HBM or in pinned host memory (read through UVA); idx on the device
    _require_cuda(idx)
    for t in (rowptr, col, value):
        if t is not None and not _device_accessible(t):
            raise RuntimeError("relabel: the graph must be in CUDA or pinned host memory "
                               "(there is no CPU fallback)")
    if rowptr.dtype != torch.int64:
        raise RuntimeError("relabel: rowptr must be int64")
    if col.dtype not in (torch.int32, torch.int64):
        raise RuntimeError("relabel: col must be int32 or int64")
    if value is not None:
        if value.dim() != 1:
            raise RuntimeError("Value tensor must be one-dimensional")
        if value.dtype != torch.float32:
            raise RuntimeError("relabel: only float32 edge values are supported on the GPU path")
    idx = idx.contiguous()
    dev = idx.device
    N = rowptr.numel() - 1
    B = idx.numel()
    # `window = (row_lo, edge_lo, num_nodes)`: rowptr / col / value hold only rows [row_lo, row_lo + len)
    # and their edges [edge_lo, ...) of the global CSR (a staged copy of one contiguous partition).  The
    # kernels read rowptr[v], rowptr[v + 1] and col / value[rowptr[v] .. rowptr[v + 1]) for v in idx only,
    # so they are handed the base addresses the full arrays would have.
    p_rowptr, p_col, p_val = ptr(rowptr), ptr(col), ptr(value)
    if window is not None:
        row_lo, edge_lo, N = (int(v) for v in window)
        p_rowptr -= row_lo * 8
        p_col -= edge_lo * col.element_size()
        if p_val is not None:
            p_val -= edge_lo * 4
    if ws is None:
        ws = _workspace_for(N, dev)
    st = _stream()
    if nnz_b is None:
        LAUNCHES["calls"] += 1
        check(lib.incagg_relabel_degree_sum(p_rowptr, ptr(idx), B, N, ptr(ws.counts), ptr(ws.buf), st))
        nnz_b = int(ws.counts[0].item())
    odt = torch.int32 if out_int32 else torch.int64
    ow = 4 if out_int32 else 8
    out_rowptr = torch.empty(B + 1, dtype=odt, device=dev)
    out_col = torch.empty(nnz_b, dtype=odt, device=dev)
    out_val = torch.empty(nnz_b, dtype=torch.float32, device=dev) if value is not None else None
    cw = 4 if col.dtype == torch.int32 else 8
    LAUNCHES["calls"] += 1
    if within:
        check(lib.incagg_relabel_one_hop_within_batch(
            p_rowptr, p_col, cw, p_val, ptr(idx), B, N, nnz_b, ptr(out_rowptr),
            ptr(out_col), ow, ptr(out_val), ptr(ws.counts), ptr(ws.buf), st))
        if known is None:
            ws.counts_host.copy_(ws.counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            nnz_out = int(ws.counts_host[1])
        else:
            nnz_out = int(known)
        ws.last_count = nnz_out
        out_col = out_col[:nnz_out]
        if out_val is not None:
            out_val = out_val[:nnz_out]
        n_id = idx
        if not bipartite:
            # reference quirk (relabel_cpu.cpp:208-211): |n_id_map| = number of DISTINCT batch ids
            pad = int(torch.unique(idx).numel())
            out_rowptr = torch.cat([out_rowptr, out_rowptr.new_full((pad,), nnz_out)])
        return out_rowptr, out_col, out_val, n_id
    n_id_buf = torch.empty(B + (min(nnz_b, N) if known is None else int(known)), dtype=torch.int64,
                           device=dev)
    check(lib.incagg_relabel_one_hop(
        p_rowptr, p_col, cw, p_val, ptr(idx), B, N, nnz_b, ptr(out_rowptr), ptr(out_col),
        ow, ptr(out_val), ptr(n_id_buf), ptr(ws.counts), ptr(ws.buf), st))
    if known is None:
        ws.counts_host.copy_(ws.counts, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        H = int(ws.counts_host[0])
    else:
        H = int(known)
    ws.last_count = H
    n_id = n_id_buf[:B + H]
    if not bipartite:
        out_rowptr = torch.cat([out_rowptr, out_rowptr.new_full((H,), nnz_b)])
    return out_rowptr, out_col, out_val, n_id


def relabel_one_hop(rowptr, col, value, idx, bipartite: bool = True, ws=None,
                    out_int32: bool = False, nnz_b: Optional[int] = None, known: Optional[int] = None,
                    window=None):
    """GPU relabel_one_hop (csrc/cpu/relabel_cpu.cpp:3-108), bit-exact."""
    return _relabel(False, rowptr, col, value, idx, bipartite, ws, out_int32, nnz_b, known, window)


def relabel_one_hop_within_batch(rowptr, col, value, idx, bipartite: bool = True, ws=None,
                                 out_int32: bool = False, nnz_b: Optional[int] = None,
                                 known: Optional[int] = None, window=None):
    """GPU relabel_one_hop_within_batch (csrc/cpu/relabel_cpu.cpp:111-214), bit-exact."""
    return _relabel(True, rowptr, col, value, idx, bipartite, ws, out_int32, nnz_b, known, window)


# --------------------------------------------------------------------------------------------
# async host<->device history transfer (read_async / write_async / synchronize)
# --------------------------------------------------------------------------------------------
_PENDING_READS = []  # events of read_async calls not yet synchronised (FIFO, like Thread::results)


def read_async(src: Tensor, offset: Optional[Tensor], count: Optional[Tensor], index: Tensor,
               dst: Tensor, buffer: Optional[Tensor] = None) -> None:
    """dst[0:sum(count)] <- src[offset_i : +count_i] slices, then dst[sum : sum+|index|] <- src[index].

    Same contract and argument checks as csrc/cuda/async_cuda.cu:14-115, enqueued on the current
    (non-default) stream.  Differences in mechanism, not in result: the slices are DMA copies, the
    indexed rows are gathered by a kernel that reads the pinned source through UVA (no CPU
    index_select, no bounce buffer: `buffer` is accepted for signature compatibility and only
    validated), and nothing synchronises the stream.
    """
    if src.is_cuda:
        raise RuntimeError("Source tensor must be a CPU tensor")
    if not dst.is_cuda:
        raise RuntimeError("Target tensor must be a CUDA tensor")
    if buffer is not None:
        if buffer.is_cuda:
            raise RuntimeError("Buffer tensor must be a CPU tensor")
        if not buffer.is_pinned():
            raise RuntimeError("Buffer tensor must be pinned")
        if not buffer.is_contiguous():
            raise RuntimeError("Buffer tensor must be contiguous")
    if not src.is_contiguous():
        raise RuntimeError("Source tensor must be contiguous")
    if not dst.is_contiguous():
        raise RuntimeError("Target tensor must be contiguous")
    if index.dim() != 1:
        raise RuntimeError("Index tensor must be one-dimensional")
    if not src.is_pinned():
        raise RuntimeError("Source tensor must be pinned")
    numel = 0
    if offset is not None:
        if count is None:
            raise RuntimeError("Count tensor is undefined")
        if offset.is_cuda or count.is_cuda:
            raise RuntimeError("Offset/Count tensor must be a CPU tensor")
        if offset.dim() != 1 or count.dim() != 1:
            raise RuntimeError("Offset/Count tensor must be one-dimensional")
        if offset.numel() != count.numel():
            raise RuntimeError("Size mismatch")
        numel = int(count.sum())
    n_index = index.numel()
    if buffer is not None and numel + n_index > buffer.size(0):
        raise RuntimeError("Buffer tensor size too small")
    if numel + n_index > dst.size(0):
        raise RuntimeError("Target tensor size too small")
    stream = torch.cuda.current_stream(dst.device)
    if stream == torch.cuda.default_stream(dst.device):
        raise RuntimeError("Asynchronous read requires a non-default CUDA stream")
    with torch.cuda.device(dst.device):
        if offset is not None and numel > 0:
            copy_slices(src, dst, offset, count, 0)
        if n_index > 0:
            idx_dev = index if index.is_cuda else index.to(dst.device, dtype=torch.int64, non_blocking=True)
            gather_rows(src, idx_dev.contiguous(), out=dst[numel:numel + n_index])
        ev = torch.cuda.Event()
        ev.record(stream)
    _PENDING_READS.append(ev)


def write_async(src: Tensor, offset: Tensor, count: Tensor, dst: Tensor) -> None:
    """dst[offset_i : +count_i] <- src[s : s+count_i] (csrc/cuda/async_cuda.cu:117-165), D2H DMA
    copies on the current non-default stream; no trailing stream synchronise."""
    if not src.is_cuda:
        raise RuntimeError("Source tensor must be a CUDA tensor")
    if offset.is_cuda or count.is_cuda:
        raise RuntimeError("Offset/Count tensor must be a CPU tensor")
    if dst.is_cuda:
        raise RuntimeError("Target tensor must be a CPU tensor")
    if not dst.is_pinned():
        raise RuntimeError("Target tensor must be pinned")
    if not src.is_contiguous():
        raise RuntimeError("Index tensor must be contiguous")
    if not dst.is_contiguous():
        raise RuntimeError("Target tensor must be contiguous")
    if offset.dim() != 1 or count.dim() != 1:
        raise RuntimeError("Offset/Count tensor must be one-dimensional")
    if offset.numel() != count.numel():
        raise RuntimeError("Size mismatch")
    stream = torch.cuda.current_stream(src.device)
    if stream == torch.cuda.default_stream(src.device):
        raise RuntimeError("Asynchronous write requires a non-default CUDA stream")
    with torch.cuda.device(src.device):
        copy_slices(src, dst, offset, count, 1)


def synchronize() -> None:
    """Wait for the oldest outstanding read_async (Thread::synchronize, csrc/thread.h:64-69)."""
    if _PENDING_READS:
        _PENDING_READS.pop(0).synchronize()


# --------------------------------------------------------------------------------------------
# torch.ops.torch_geometric_autoscale.* registration (the reference's plugin surface)
# --------------------------------------------------------------------------------------------
_REGISTERED = False


def register_reference_ops() -> None:
    """Expose the five reference op names backed by the CUDA library.  Idempotent."""
    global _REGISTERED
    if _REGISTERED:
        return
    try:
        lib_def = torch.library.Library("torch_geometric_autoscale", "DEF")
    except RuntimeError:
        _REGISTERED = True  # namespace already defined (e.g. the reference's own .so is loaded)
        return
    lib_def.define("relabel_one_hop(Tensor rowptr, Tensor col, Tensor? value, Tensor idx, bool bipartite)"
                   " -> (Tensor, Tensor, Tensor?, Tensor)")
    lib_def.define("relabel_one_hop_within_batch(Tensor rowptr, Tensor col, Tensor? value, Tensor idx, "
                   "bool bipartite) -> (Tensor, Tensor, Tensor?, Tensor)")
    lib_def.define("read_async(Tensor src, Tensor? offset, Tensor? count, Tensor index, Tensor(a!) dst, "
                   "Tensor buffer) -> ()")
    lib_def.define("write_async(Tensor src, Tensor offset, Tensor count, Tensor(a!) dst) -> ()")
    lib_def.define("synchronize() -> ()")

    def _rl(rowptr, col, value, idx, bipartite):
        return relabel_one_hop(rowptr, col, value, idx, bipartite)

    def _rlw(rowptr, col, value, idx, bipartite):
        return relabel_one_hop_within_batch(rowptr, col, value, idx, bipartite)

    def _cpu_reject(*args, **kwargs):
        raise RuntimeError("incagg_b200: this operator runs on CUDA tensors only (no CPU fallback)")

    lib_def.impl("relabel_one_hop", _rl, "CUDA")
    lib_def.impl("relabel_one_hop_within_batch", _rlw, "CUDA")
    lib_def.impl("relabel_one_hop", _cpu_reject, "CPU")
    lib_def.impl("relabel_one_hop_within_batch", _cpu_reject, "CPU")
    lib_def.impl("read_async", lambda src, offset, count, index, dst, buffer:
                 read_async(src, offset, count, index, dst, buffer), "CompositeExplicitAutograd")
    lib_def.impl("write_async", lambda src, offset, count, dst: write_async(src, offset, count, dst),
                 "CompositeExplicitAutograd")
    lib_def.impl("synchronize", lambda: synchronize(), "CompositeExplicitAutograd")
    register_reference_ops._lib = lib_def  # keep alive
    _REGISTERED = True
