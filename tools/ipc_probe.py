import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = torch.full((1000, 128), float(rank + 1), device=dev)
info = t.untyped_storage()._share_cuda_()
meta = (info, tuple(t.shape), tuple(t.stride()), t.storage_offset())
gathered = [None] * world
dist.all_gather_object(gathered, meta)
views = []
for r in range(world):
    if r == rank:
        views.append(t)
    else:
        inf, shape, stride, off = gathered[r]
        # open the handle in THIS rank's device context (lazy peer access): kernels of this device can
        # then load the peer's HBM directly over NVLink
        st = torch.UntypedStorage._new_shared_cuda(local, *inf[1:])
        v = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, off, shape, stride)
        views.append(v)
torch.cuda.synchronize()
dist.barrier()
peer = views[(rank + 1) % world]
print(rank, "peer view device", peer.device, "ptr", hex(peer.data_ptr()), "value", float(peer[5, 7]), flush=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import incagg_gnn_b200
from incagg_gnn_b200 import ops
idx = torch.arange(0, 1000, 3, device=dev)
out = torch.empty(idx.numel(), 128, device=dev)
ops.gather_rows(peer, idx, out=out)   # kernel on my device reading the peer's HBM over NVLink
torch.cuda.synchronize()
print(rank, "gathered from peer:", float(out.mean()), out.device, flush=True)
dist.barrier()
dist.destroy_process_group()
