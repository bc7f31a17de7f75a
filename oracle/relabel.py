"""ctypes wrappers of liboracle.so and of the compiled reference op (oracle/_ref).  Test
infrastructure only (see oracle/__init__.py)."""
import ctypes
import os
import subprocess
import sys
import tempfile
from ctypes import c_int64, c_void_p, c_int

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "ref_relabel.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile relabel_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "relabel_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, src])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        P = c_void_p
        _lib.oracle_relabel_degree_sum.restype = c_int64
        _lib.oracle_relabel_degree_sum.argtypes = [P, P, c_int64]
        _lib.oracle_relabel_one_hop.restype = c_int64
        _lib.oracle_relabel_one_hop.argtypes = [P, P, P, P, c_int64, c_int64, P, P, P, P]
        _lib.oracle_relabel_one_hop_within_batch.restype = c_int64
        _lib.oracle_relabel_one_hop_within_batch.argtypes = [P, P, P, P, c_int64, P, P, P, P]
        _lib.oracle_spmm_csr.restype = None
        _lib.oracle_spmm_csr.argtypes = [c_int, P, P, P, P, c_int64, P, c_int64, P, c_int64, c_int64]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _i64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


def relabel_one_hop(rowptr, col, value, idx, bipartite=True):
    """-> (out_rowptr, out_col, out_value|None, n_id), int64 numpy arrays, exactly the tuple of
    relabel_one_hop_cpu (csrc/cpu/relabel_cpu.cpp:3-108)."""
    L = lib()
    rowptr, col, idx = _i64(rowptr), _i64(col), _i64(idx)
    val = None if value is None else np.ascontiguousarray(value, dtype=np.float32)
    B, N = idx.size, rowptr.size - 1
    nnz = L.oracle_relabel_degree_sum(_p(rowptr), _p(idx), B)
    out_rowptr = np.empty(B + 1, np.int64)
    out_col = np.empty(nnz, np.int64)
    out_val = None if val is None else np.empty(nnz, np.float32)
    n_ids = np.empty(min(nnz, N) + 1, np.int64)
    H = L.oracle_relabel_one_hop(_p(rowptr), _p(col), _p(val), _p(idx), B, N, _p(out_rowptr),
                                 _p(out_col), _p(out_val), _p(n_ids))
    if H < 0:
        raise MemoryError
    if not bipartite:
        out_rowptr = np.concatenate([out_rowptr, np.full(H, nnz, np.int64)])
    return out_rowptr, out_col, out_val, np.concatenate([idx, n_ids[:H]])


def relabel_one_hop_within_batch(rowptr, col, value, idx, bipartite=True):
    """relabel_one_hop_within_batch_cpu (csrc/cpu/relabel_cpu.cpp:111-214)."""
    L = lib()
    rowptr, col, idx = _i64(rowptr), _i64(col), _i64(idx)
    val = None if value is None else np.ascontiguousarray(value, dtype=np.float32)
    B = idx.size
    nnz = L.oracle_relabel_degree_sum(_p(rowptr), _p(idx), B)
    out_rowptr = np.empty(B + 1, np.int64)
    out_col = np.empty(nnz, np.int64)
    out_val = None if val is None else np.empty(nnz, np.float32)
    distinct = np.zeros(1, np.int64)
    kept = L.oracle_relabel_one_hop_within_batch(_p(rowptr), _p(col), _p(val), _p(idx), B,
                                                 _p(out_rowptr), _p(out_col), _p(out_val), _p(distinct))
    if kept < 0:
        raise MemoryError
    out_col = out_col[:kept].copy()
    out_val = None if out_val is None else out_val[:kept].copy()
    if not bipartite:
        out_rowptr = np.concatenate([out_rowptr, np.full(int(distinct[0]), kept, np.int64)])
    return out_rowptr, out_col, out_val, idx


# ---- the reference's own compiled op (oracle/_ref/ref_relabel.so) ---------------------------------
def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


_REF_SCRIPT = r'''
import sys, numpy as np, torch
torch.ops.load_library(sys.argv[1])
d = np.load(sys.argv[2], allow_pickle=False)
fn = getattr(torch.ops.torch_geometric_autoscale, str(d["fn"]))
val = torch.from_numpy(d["value"]) if d["has_value"] else None
r, c, v, n = fn(torch.from_numpy(d["rowptr"]), torch.from_numpy(d["col"]), val,
                torch.from_numpy(d["idx"]), bool(d["bipartite"]))
np.savez(sys.argv[3], rowptr=r.numpy(), col=c.numpy(), value=(v.numpy() if v is not None else np.zeros(0, np.float32)),
         n_id=n.numpy())
'''


def ref_relabel(fn: str, rowptr, col, value, idx, bipartite=True):
    """Run the REFERENCE's compiled op in a subprocess (its operator names collide with the ones the
    product package registers, so it never shares a process with it)."""
    if not ref_available():
        raise FileNotFoundError(_REF_PATH)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        np.savez(fin, fn=fn, rowptr=_i64(rowptr), col=_i64(col),
                 value=(np.zeros(0, np.float32) if value is None else np.ascontiguousarray(value, np.float32)),
                 has_value=value is not None, idx=_i64(idx), bipartite=bool(bipartite))
        subprocess.check_call([sys.executable, "-c", _REF_SCRIPT, _REF_PATH, fin, fout])
        d = np.load(fout)
        return d["rowptr"], d["col"], (d["value"] if value is not None else None), d["n_id"]


def ref_ops_in_process():
    """The reference's compiled relabel ops loaded into THIS process (torch.ops.load_library).  Only
    for processes that never import the product package, whose operator names are the same ones
    (bench.py --impl reference).  Returns (relabel_one_hop, relabel_one_hop_within_batch) taking and
    returning torch tensors, exactly the reference's signatures (csrc/relabel.cpp:10-38)."""
    import sys
    import torch
    if "incagg_gnn_b200" in sys.modules:
        raise RuntimeError("the product package registers the same operator names; run the reference's "
                           "op in a subprocess (ref_relabel) instead")
    if not ref_available():
        raise FileNotFoundError(_REF_PATH)
    torch.ops.load_library(_REF_PATH)
    ns = torch.ops.torch_geometric_autoscale
    return ns.relabel_one_hop, ns.relabel_one_hop_within_batch
