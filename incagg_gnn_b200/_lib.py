"""ctypes binding of the C-ABI library ``csrc/libincagg_b200.so`` (include/incagg_b200.h).

This is the only module that touches the shared library.  There is no CPU or PyTorch fallback:
if the library is missing the import fails, and every compute entry point requires CUDA tensors.
"""
import ctypes
import os
import re
from ctypes import c_int, c_int32, c_int64, c_size_t, c_void_p, c_char_p, c_float

_HERE = os.path.dirname(os.path.abspath(__file__))
# (INCAGG_B200_LIB: an alternative build of the same library, for kernel experiments)
LIB_PATH = os.environ.get("INCAGG_B200_LIB") or os.path.join(_HERE, "csrc", "libincagg_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "incagg_b200.h")

REDUCE = {"sum": 0, "add": 0, "mean": 1, "min": 2, "max": 3}

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3


class IncAggError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status (the reference raises RuntimeError from
    AT_ASSERTM / AT_ERROR at the same checks)."""

    def __init__(self, code, msg):
        super().__init__(f"incagg_b200 error {code}: {msg}")
        self.code = code


def _load():
    # the library links the CUDA runtime dynamically (libcudart.so.12): importing torch first makes the
    # loader reuse the copy torch ships, whatever LD_LIBRARY_PATH says
    import torch  # noqa: F401
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C incagg_gnn_b200/csrc`). "
            "There is no CPU fallback for the hot path.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

P = c_void_p
_PROTOS = {
    "incagg_version": (c_int, []),
    "incagg_last_error": (c_char_p, []),
    "incagg_launch_count": (c_int64, []),
    "incagg_tune_set": (c_int, [c_int, c_int]),
    "incagg_device_errors": (c_int, [P, c_int]),
    "incagg_allreduce_adam_blocks": (c_int, []),
    "incagg_allreduce_adam_max_ranks": (c_int, []),
    "incagg_allreduce_adam_step": (c_int, [P, P, c_int, c_int, P, P, P, P, c_int64, c_int64, c_float, c_float,
                                           c_float, c_float, c_float, c_float, P, P, P]),
    "incagg_device_info": (c_int, [P, P, P]),
    "incagg_enable_peer_access": (c_int, [c_int]),
    "incagg_spmm_plan_bytes": (c_size_t, [c_int64, c_int64]),
    "incagg_spmm_plan": (c_int, [P, c_int64, c_int64, P, c_size_t, P]),
    "incagg_spmm_csr": (c_int, [c_int, P, P, P, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int32, P, P]),
    "incagg_spmm_csr_gated": (c_int, [c_int, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, c_int64, P]),
    "incagg_spmm_delta": (c_int, [c_int, P, P, P, P, c_int64, P, c_int64, P, c_int64, P, P, c_int64,
                                  c_int64, c_int32, P, P]),
    "incagg_spmm_minmax_bwd": (c_int, [P, P, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int32, P]),
    "incagg_spmm_multi": (c_int, [P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, c_int32, P, P, P]),
    "incagg_spmm_multi_arg": (c_int, [P, P, P, P, c_int64, P, c_int64, P, c_int64, c_int64, c_int32, c_int32, P, P,
                                      P]),
    "incagg_gemm_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "incagg_gemm_tf32x3": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, P, c_int64, P, c_int64, c_float, P,
                                   c_int64, c_float, P, c_int, P, c_int64, P, c_size_t, P]),
    "incagg_gemm_tf32x3_dual": (c_int, [c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, P, c_int64, P,
                                        c_int64, P, c_int64, P, c_int64, c_float, c_float, c_float, c_float,
                                        P, c_int64, c_float, P, c_int64, c_float, c_int, P, c_int64, P, c_int64,
                                        P, c_size_t, P]),
    "incagg_gemm_tf32x3_group": (c_int, [c_int, c_int, c_int, c_int64, c_int64, c_int64, P, P, P, P, P, c_float, P, P,
                                         P, c_size_t, P]),
    "incagg_colsum_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "incagg_relu_bwd_colsum": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, P, c_int64, P, P, c_size_t, P]),
    "incagg_relu_bwd_colsum_ex": (c_int, [P, c_int64, P, c_int64, c_int64, c_int32, P, c_int64, P, c_int64, c_int64, P,
                                          c_int, P, c_size_t, P]),
    "incagg_masked_ce_workspace_bytes": (c_size_t, [c_int64]),
    "incagg_masked_ce": (c_int, [P, c_int64, P, P, c_int64, c_int32, P, c_int64, P, P, c_size_t, P]),
    "incagg_mask_count": (c_int, [P, c_int64, P, P]),
    "incagg_masked_ce_rows": (c_int, [P, c_int64, P, P, c_int64, c_int32, P, P, c_int64, P, c_size_t, P]),
    "incagg_masked_ce_finish": (c_int, [P, c_int64, P, P, P, P]),
    "incagg_adam_step": (c_int, [P, P, P, P, c_int64, c_int64, c_float, c_float, c_float, c_float, c_float,
                                 c_float, P, P, P]),
    "incagg_csr_transpose_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "incagg_csr_transpose": (c_int, [P, P, P, c_int64, c_int64, c_int64, P, P, P, P, P, c_size_t, P]),
    "incagg_gather_rows": (c_int, [P, c_int64, c_int64, P, c_int64, P, c_int64, c_int64, P]),
    "incagg_gather_rows_sharded": (c_int, [P, P, c_int, c_int64, P, c_int64, P, c_int64, c_int64, P]),
    "incagg_scatter_rows": (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int64, c_int64, P]),
    "incagg_copy_slices": (c_int, [P, c_int64, c_int64, P, c_int64, c_int64, P, P, c_int64, c_int64,
                                   c_int, P]),
    "incagg_relabel_workspace_bytes": (c_size_t, [c_int64]),
    "incagg_relabel_workspace_init": (c_int, [P, c_int64, P]),
    "incagg_relabel_degree_sum": (c_int, [P, P, c_int64, c_int64, P, P, P]),
    "incagg_relabel_one_hop": (c_int, [P, P, c_int, P, P, c_int64, c_int64, c_int64, P, P, c_int, P, P,
                                       P, P, P]),
    "incagg_relabel_one_hop_within_batch": (c_int, [P, P, c_int, P, P, c_int64, c_int64, c_int64, P, P,
                                                    c_int, P, P, P, P]),
}


def header_symbols():
    """Names of all functions declared in include/incagg_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(incagg_[a-z0-9_]+)\s*\(", text)))


for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name, None)
    if _fn is None:
        continue  # checked by tests/test_abi.py against the header
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.incagg_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc != 0:
        raise IncAggError(rc, last_error())


def launch_count() -> int:
    """Kernels launched by libincagg_b200.so in this process (bench.py's gpu_launches)."""
    return int(lib.incagg_launch_count())


def ptr(t):
    """Device (or pinned-host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
