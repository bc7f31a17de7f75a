"""Seeded synthetic graphs of the BASELINE shapes for the timed CPU reference arm — TEST / BENCH
INFRASTRUCTURE ONLY (see oracle/__init__.py).

`bench.py --impl reference` must not import the product package, so the input generator of SURVEY.md
§8d is restated here with plain torch.  Same recipe, same order of random draws and same generator seeds
as the product's input generator: for the same (shape, seed, scale, device) both produce the same graph,
features, labels and masks (tests/test_oracle.py::test_reference_arm_inputs_equal_the_product_inputs).
Nothing here is on the product path.
"""
from typing import NamedTuple, Tuple

import torch
from torch import Tensor

# name -> (nodes, directed nnz after symmetrisation, features, classes, parts)   
SHAPES = {
    'flickr': (89_250, 899_756, 500, 7, 24),
    'arxiv': (169_343, 2 * 1_166_243, 128, 40, 80),
    'products': (2_449_029, 61_859_140, 100, 47, 150),
    'reddit': (232_965, 114_615_892, 602, 41, 200),
    'amazonproducts': (1_569_960, 264_339_468, 200, 107, 200),
}
_PRIME = 1_000_003


class Inputs(NamedTuple):
    rowptr: Tensor      # int64 [N+1]
    col: Tensor         # int64 [nnz]
    x: Tensor           # fp32 [N, F]
    y: Tensor           # int64 [N]
    train_mask: Tensor  # bool [N]
    ptr: Tensor         # int64 [parts+1] partition boundaries (contiguous node blocks)
    num_features: int
    num_classes: int


def part_ptr(num_nodes: int, num_parts: int) -> Tensor:
    base, rem = divmod(num_nodes, num_parts)
    sizes = torch.full((num_parts,), base, dtype=torch.int64)
    sizes[:rem] += 1
    ptr = torch.zeros(num_parts + 1, dtype=torch.int64)
    torch.cumsum(sizes, 0, out=ptr[1:])
    return ptr


def _local(u: Tensor, size: Tensor, power: float) -> Tensor:
    loc = torch.minimum((u.pow(power) * size).to(torch.int64), size - 1)
    return (loc * _PRIME) % size


def make_inputs(name: str, seed: int = 0, scale: int = 1, num_parts: int = None, device='cpu',
                p_inter: float = 0.15, skew: float = 1.6) -> Inputs:
    n, e, f, c, parts = SHAPES[name.lower().replace('ogbn-', '')]
    P = parts if num_parts is None else num_parts
    N, E = max(n // scale, P), max(e // scale, 2)
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ptr = part_ptr(N, P)
    ptr_d = ptr.to(device)
    sizes = ptr_d[1:] - ptr_d[:-1]
    n_pairs = int(E // 2 * 1.03) + 16

    def rnd(k):
        return torch.rand(k, generator=g, device=device)

    blk = torch.minimum((rnd(n_pairs) * P).to(torch.int64), torch.tensor(P - 1, device=device))
    src = ptr_d[blk] + _local(rnd(n_pairs), sizes[blk], skew)
    inter = rnd(n_pairs) < p_inter
    off = (-(1.0 - rnd(n_pairs)).log() * 2.0).to(torch.int64) + 1
    sign = torch.where(rnd(n_pairs) < 0.5, -1, 1)
    dblk = torch.where(inter, (blk + sign * off) % P, blk)
    dst = ptr_d[dblk] + _local(rnd(n_pairs), sizes[dblk], skew)
    del blk, inter, off, sign, dblk
    keep = src != dst
    lo, hi = torch.minimum(src, dst)[keep], torch.maximum(src, dst)[keep]
    del src, dst, keep
    key = torch.unique(lo * N + hi)
    want = E // 2
    if key.numel() > want:
        key = key[torch.randperm(key.numel(), generator=g, device=device)[:want]]
    lo, hi = key // N, key % N
    k2, _ = torch.sort(torch.cat([lo, hi]) * N + torch.cat([hi, lo]))
    row, col = k2 // N, k2 % N
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
    torch.cumsum(torch.bincount(row, minlength=N), 0, out=rowptr[1:])
    x = torch.randn(N, f, generator=g, device=device, dtype=torch.float32)
    node_blk = torch.repeat_interleave(torch.arange(P, device=device), sizes)
    y = (node_blk * 7919) % c
    noise = rnd(N) < 0.3
    y = torch.where(noise, (rnd(N) * c).to(torch.int64).clamp_(max=c - 1), y)
    k = min(f, c)
    x[:, :k] += 0.5 * torch.nn.functional.one_hot(y, c)[:, :k].to(x.dtype)
    r = rnd(N)
    return Inputs(rowptr.cpu(), col.cpu(), x.cpu(), y.cpu(), (r < 0.6).cpu(), ptr, f, c)
