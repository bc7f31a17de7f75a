"""Pins the async host<->device history transfer against the REFERENCE's own compiled op
(oracle/_ref/ref_async.so = csrc/async.cpp + csrc/cuda/async_cuda.cu built for sm_100a in the build
container; it travels to the GPU box with the snapshot).  The reference runs in a subprocess; the
product's read_async / write_async and the oracle restatement must produce the same bytes."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "ref_async.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/ref_async.so not built")
def test_read_write_async_match_the_reference_op(cuda):
    import incagg_gnn_b200  # noqa: F401
    from incagg_gnn_b200 import ops
    from oracle import gas
    rng = np.random.default_rng(11)
    N, D, rows = 5000, 96, 1500
    table = rng.standard_normal((N, D)).astype(np.float32)
    offset = np.array([100, 2000, 4990], np.int64)
    count = np.array([300, 150, 10], np.int64)
    index = rng.permutation(N)[:700].astype(np.int64)
    push = rng.standard_normal((int(count.sum()), D)).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.npz"), os.path.join(td, "out.npz")
        np.savez(fin, table=table, offset=offset, count=count, index=index, push=push, buffer_rows=rows)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_async.py"), REF_SO, fin, fout],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
        ref = np.load(fout)
        ref_pulled, ref_idx_only, ref_table = ref["pulled"], ref["pulled_index_only"], ref["table_after_push"]
    # product
    t = torch.from_numpy(table).pin_memory()
    dst = torch.zeros(rows, D, device=cuda)
    side = torch.cuda.Stream(cuda)
    with torch.cuda.stream(side):
        ops.read_async(t, torch.from_numpy(offset), torch.from_numpy(count), torch.from_numpy(index), dst, None)
    ops.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), ref_pulled)
    dst2 = torch.zeros(rows, D, device=cuda)
    with torch.cuda.stream(side):
        ops.read_async(t, None, None, torch.from_numpy(index), dst2, None)
    ops.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(dst2.cpu().numpy(), ref_idx_only)
    t2 = torch.from_numpy(table.copy()).pin_memory()
    x = torch.from_numpy(push).to(cuda)
    with torch.cuda.stream(side):
        side.wait_stream(torch.cuda.current_stream())
        ops.write_async(x, torch.from_numpy(offset), torch.from_numpy(count), t2)
    side.synchronize()
    assert np.array_equal(t2.numpy(), ref_table)
    # oracle restatement (oracle/gas.py) against the same reference outputs
    o_pull = gas.pull_slices_and_index(torch.from_numpy(table), torch.from_numpy(offset), torch.from_numpy(count),
                                       torch.from_numpy(index)).numpy()
    assert np.array_equal(o_pull, ref_pulled[:o_pull.shape[0]])
    o_table = torch.from_numpy(table.copy())
    gas.push_slices(o_table, torch.from_numpy(push), torch.from_numpy(offset), torch.from_numpy(count))
    assert np.array_equal(o_table.numpy(), ref_table)
