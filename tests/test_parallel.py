"""Multi-process (gloo, world_size 2, CPU) tests of the sharding protocol of parallel.py: ownership
map, halo-plan id exchange, the row all-to-all-v and the gradient averaging.  The row gather / scatter
on each side are injected torch stand-ins (the product path uses the CUDA kernels and has no CPU
fallback); what is tested here is the host-side exchange logic."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import incagg_gnn_b200  # noqa: F401
        from incagg_gnn_b200.parallel import Shard, HaloPlan, pull_halo_rows, GradAverager
        N, D, P = 1000, 6, 10
        ptr = torch.arange(0, N + 1, N // P)
        shard = Shard(ptr, rank, world)
        assert shard.part_bounds == [0, 5, 10] and (shard.lo, shard.hi) == (rank * 500, rank * 500 + 500)
        assert shard.owner_of(torch.tensor([0, 499, 500, 999])).tolist() == [0, 0, 1, 1]
        g = torch.Generator().manual_seed(0)
        table = torch.randn(N, D, generator=g)                 # the global table (same on both ranks)
        local = table[shard.lo:shard.hi].clone()               # this rank's shard
        gather = lambda t, idx, out=None: t.index_select(0, idx)   # noqa: E731

        def scatter(src, idx, dst):
            dst[idx] = src

        for trial in range(3):
            gg = torch.Generator().manual_seed(10 * trial + rank)
            halo = torch.randperm(N, generator=gg)[:137 + 20 * rank]   # ids owned by both ranks
            if trial == 2 and rank == 1:
                halo = halo[:0]                                         # one side requests nothing
            plan = HaloPlan(halo, shard)
            assert sum(plan.req_counts) == halo.numel()
            out = torch.full((halo.numel(), D), float("nan"))
            pull_halo_rows(local, plan, out, gather=gather, scatter=scatter)
            assert torch.equal(out, table[halo]), f"rank {rank} trial {trial}"
            # a second layer's pull reuses the plan
            out2 = torch.zeros_like(out)
            pull_halo_rows(local * 2, plan, out2, gather=gather, scatter=scatter)
            assert torch.equal(out2, table[halo] * 2)
        # gradient averaging
        w = torch.nn.Parameter(torch.zeros(3, 2))
        b = torch.nn.Parameter(torch.zeros(5))
        avg = GradAverager([w, b], shard)          # p.grad become views into one flat buffer
        assert w.grad.data_ptr() == avg.flat.data_ptr() and b.grad.numel() == 5
        (w.sum() * float(rank + 1)).backward()     # autograd accumulates in place into the views
        avg()
        assert torch.allclose(w.grad, torch.full((3, 2), 1.5)) and torch.equal(b.grad, torch.zeros(5))
        assert w.grad.data_ptr() == avg.flat.data_ptr()
        avg.zero()
        assert float(avg.flat.abs().sum()) == 0.
        assert shard.steps_per_epoch(1) == 5 and Shard(torch.arange(0, 8), 0, 2).steps_per_epoch(1) == 4
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_and_grad_average_world2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok", ret.get(0)
    assert ret.get(1) == "ok", ret.get(1)


def test_shard_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import incagg_gnn_b200  # noqa: F401
    from incagg_gnn_b200.parallel import Shard
    sh = Shard(torch.tensor([0, 10, 25, 40]), 0, 1)
    assert (sh.lo, sh.hi, sh.num_local) == (0, 40, 40) and list(sh.parts) == [0, 1, 2]
    with pytest.raises(ValueError):
        Shard(torch.tensor([0, 10]), 0, 2)
