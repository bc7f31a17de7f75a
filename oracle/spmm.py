"""numpy restatement of torch_sparse.matmul(reduce=sum|mean|min|max), the IncAgg delta update and
the PNA multi-aggregator pass.  Test infrastructure only (see oracle/__init__.py); PARITY UNPINNED
(third-party semantics):

  sum   out[i] = sum_e val[e] * X[col[e]]
  mean  sum / max(deg_i, 1)
  min/max  elementwise over val[e] * X[col[e]]; empty row -> 0; arg = first winning edge position
           (CSR order), -1 for empty rows
"""
import numpy as np


def spmm(rowptr, col, val, X, reduce="sum", dtype=np.float32, return_arg=False):
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    X = np.asarray(X, dtype)
    squeeze = X.ndim == 1
    if squeeze:
        X = X[:, None]
    rows, F = rowptr.size - 1, X.shape[1]
    deg = np.diff(rowptr)
    msgs = X[col]
    if val is not None:
        msgs = msgs * np.asarray(val, dtype)[:, None]
    out = np.zeros((rows, F), dtype)
    arg = np.full((rows, F), -1, np.int64)
    if reduce in ("sum", "add", "mean"):
        row = np.repeat(np.arange(rows), deg)
        np.add.at(out, row, msgs)
        if reduce == "mean":
            out = out / np.maximum(deg, 1).astype(dtype)[:, None]
    elif reduce in ("min", "max"):
        nz = np.nonzero(deg)[0]
        if nz.size:
            red = np.minimum if reduce == "min" else np.maximum
            out[nz] = red.reduceat(msgs, rowptr[nz], axis=0)
            if return_arg:
                for i in nz:  # small cases only
                    seg = msgs[rowptr[i]:rowptr[i + 1]]
                    a = seg.argmin(0) if reduce == "min" else seg.argmax(0)
                    arg[i] = rowptr[i] + a
    else:
        raise ValueError(reduce)
    if squeeze:
        out = out[:, 0]
    return (out, arg) if return_arg else out


def spmm_delta(rowptr, col, val, x, m_in, m_ag, reduce="sum", dtype=np.float32):
    """h = reduce(A, x - M_in) + M_ag   (gcn.py:241, gcn2.py:255, appnp.py:122, graphsage.py:634)."""
    x = np.asarray(x, dtype)
    F = x.shape[1]
    d = x - np.asarray(m_in, dtype)[:x.shape[0], :F]
    rows = len(rowptr) - 1
    return spmm(rowptr, col, val, d, reduce, dtype) + np.asarray(m_ag, dtype)[:rows, :F]


def spmm_multi(rowptr, col, val, X, F, reducers, dtype=np.float32):
    """Slab k of X (columns kF..(k+1)F) reduced with reducers[k] (pna.py:66-84, K separate passes)."""
    X = np.asarray(X, dtype)
    return np.concatenate([spmm(rowptr, col, val, X[:, k * F:(k + 1) * F], r, dtype)
                           for k, r in enumerate(reducers)], axis=1)


def csr_transpose(rowptr, col, val, cols):
    """CSR of A^T with entries of a transposed row in original edge order (stable counting sort)."""
    rowptr = np.asarray(rowptr, np.int64)
    col = np.asarray(col, np.int64)
    rows = rowptr.size - 1
    row = np.repeat(np.arange(rows), np.diff(rowptr))
    perm = np.argsort(col, kind="stable")
    t_rowptr = np.zeros(cols + 1, np.int64)
    np.cumsum(np.bincount(col, minlength=cols), out=t_rowptr[1:])
    return t_rowptr, row[perm], (None if val is None else np.asarray(val)[perm]), perm
