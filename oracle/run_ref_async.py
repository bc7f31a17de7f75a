"""Runs the REFERENCE's own read_async / write_async (oracle/_ref/ref_async.so, compiled from
/root/reference/csrc/async.cpp + csrc/cuda/async_cuda.cu by oracle/build_ref_async.sh) on a GPU and
stores what they produce.  Executed in a subprocess by tests/test_gpu_ref_async.py: the reference's
operator names are the ones the product package registers too, so the two never share a process.
Test infrastructure only.

    python oracle/run_ref_async.py <ref_async.so> <in.npz> <out.npz>
"""
import sys

import numpy as np
import torch

torch.ops.load_library(sys.argv[1])
ns = torch.ops.torch_geometric_autoscale
d = np.load(sys.argv[2])
table = torch.from_numpy(d["table"]).pin_memory()
offset, count, index = torch.from_numpy(d["offset"]), torch.from_numpy(d["count"]), torch.from_numpy(d["index"])
rows = int(d["buffer_rows"])
dst = torch.zeros(rows, table.size(1), device="cuda")
buf = torch.empty(rows, table.size(1)).pin_memory()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    ns.read_async(table, offset, count, index, dst, buf)
ns.synchronize()
torch.cuda.synchronize()
pulled = dst.cpu().numpy()
# index-only pull (offset / count undefined), as ScalableGNN.__call__ issues it (base.py:203-204)
dst2 = torch.zeros(rows, table.size(1), device="cuda")
with torch.cuda.stream(side):
    ns.read_async(table, None, None, index, dst2, buf)
ns.synchronize()
torch.cuda.synchronize()
push = torch.from_numpy(d["push"]).cuda()
table2 = torch.from_numpy(d["table"]).pin_memory()
torch.cuda.synchronize()
with torch.cuda.stream(side):
    ns.write_async(push, offset, count, table2)
torch.cuda.synchronize()
np.savez(sys.argv[3], pulled=pulled, pulled_index_only=dst2.cpu().numpy(), table_after_push=table2.numpy())
