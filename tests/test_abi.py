"""The C-ABI library loads and exports every symbol include/incagg_b200.h declares (no compute
calls: runs without a GPU)."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from importlib import import_module
    import incagg_gnn_b200  # noqa: F401
    _lib = import_module("incagg_gnn_b200._lib")
    syms = _lib.header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in the header but not exported"
        assert s in _lib._PROTOS, f"{s} has no ctypes prototype"
    assert _lib.lib.incagg_version() >= 100
    assert _lib.last_error() == "" or isinstance(_lib.last_error(), str)


def test_argument_errors_do_not_need_a_gpu():
    from importlib import import_module
    _lib = import_module("incagg_gnn_b200._lib")
    # negative sizes are rejected before any CUDA call
    rc = _lib.lib.incagg_spmm_csr(0, None, None, None, None, 0, None, 0, None, 0, -1, 4, None, None)
    assert rc == _lib.ERR_INVALID and "negative" in _lib.last_error()
    rc = _lib.lib.incagg_gather_rows(None, 0, 0, None, 5, None, 0, 6, None)
    assert rc == _lib.ERR_INVALID


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "incagg_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f


def test_reference_op_names_are_registered():
    import torch
    import incagg_gnn_b200  # noqa: F401
    ns = torch.ops.torch_geometric_autoscale
    for name in ("relabel_one_hop", "relabel_one_hop_within_batch", "read_async", "write_async",
                 "synchronize"):
        assert hasattr(ns, name)
    # CPU tensors are refused loudly (there is no CPU fallback)
    import pytest
    with pytest.raises(RuntimeError):
        ns.relabel_one_hop(torch.tensor([0, 1]), torch.tensor([0]), None, torch.tensor([0]), True)
