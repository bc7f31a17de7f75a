"""CPU oracle of the IncAgg-GNN propagation hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / the timed CPU baseline.  The product package
(``incagg_gnn_b200/``) never imports it and has no CPU fallback.

Contents
  * ``relabel_oracle.c`` -> ``liboracle.so``: plain-C restatement of ``relabel_one_hop`` /
    ``relabel_one_hop_within_batch`` (reference csrc/cpu/relabel_cpu.cpp:3-214) and of the fp32 CSR
    SpMM reducers.  PINNED: checked against the reference's own compiled op (``oracle/_ref/
    ref_relabel.so`` built by ``build_ref.sh`` from /root/reference) and the known answers of
    SURVEY.md §4 (tests/test_oracle.py, fixtures in tests/golden/).
  * ``spmm.py``: numpy restatement of torch_sparse.matmul sum/mean/min/max (+arg), the IncAgg delta
    and the PNA multi-aggregator pass, in fp32 and fp64.  PARITY UNPINNED: torch_sparse /
    torch_scatter / torch_geometric are third-party dependencies that are neither vendored in the
    reference nor installed here, and the reference pins no version and holds no test vectors for
    them; the restatement follows their published semantics and is anchored on the
    reference's call sites.
  * ``run_ref_async.py`` + ``build_ref_async.sh``: the reference's own ``read_async`` / ``write_async``
    (csrc/async.cpp + csrc/cuda/async_cuda.cu) compiled for sm_100a into ``oracle/_ref/ref_async.so`` and
    run on the GPU box in a subprocess.  PINNED: the product's transfer ops and the restatement in
    ``gas.py`` (``pull_slices_and_index`` / ``push_slices``) reproduce its output byte for byte
    (tests/test_gpu_ref_async.py).
  * ``gas.py``: pure-torch (CPU, fp32/fp64) restatement of History push/pull (history.py:33-65),
    push_and_pull's synchronous branch (models/base.py:411-426), the GAS and IncAgg training steps
    and the layer-wise sweeps of GCN / GCN2 / APPNP / GraphSAGE / PNA (files and lines cited per
    function).  PARITY UNPINNED for the same reason (the reference package cannot be imported
    here: hard imports of torch_sparse, torch_geometric, ipdb, hydra).
"""
from .relabel import (relabel_one_hop, relabel_one_hop_within_batch, ref_relabel, ref_available,
                      build as build_c)  # noqa
from .spmm import spmm, spmm_delta, spmm_multi, csr_transpose  # noqa
