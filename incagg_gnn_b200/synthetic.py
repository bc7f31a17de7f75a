"""Seeded synthetic graphs of the BASELINE shapes — there is no network, so these
stand in for ``data.get_data`` (reference data.py:118-145).

A graph is generated *already clustered*: ``num_parts`` near-equal contiguous node blocks stand in
for the METIS output the reference computes before the hot path (main.py:144-145), so ``metis``
returns the identity permutation for it.  Edges: a skewed (power-law-like) source distribution; the
destination is inside the source's block with probability ``1 - p_inter`` and otherwise in a nearby
block (geometric block offset, skewed towards the block's popular nodes), which gives batches a
one-hop halo of a few times the batch size, as METIS partitions of the real datasets do.  The graph
is symmetrised; the named edge count is used as the number of *directed* non-zeros after
symmetrisation (self loops from ``set_diag`` come on top).

Everything is generated with torch on the requested device from ``torch.Generator`` seeds, so the GPU
box and the CPU oracle see identical inputs for identical (shape, seed, scale).
"""
import math
from typing import Tuple

import torch
from torch import Tensor

from .data import Data
from .metis import block_ptr
from .sparse import SparseTensor

# name -> (nodes, directed nnz after symmetrisation, features, classes, parts)
SHAPES = {
    'flickr': (89_250, 899_756, 500, 7, 24),            # C1
    'arxiv': (169_343, 2 * 1_166_243, 128, 40, 80),     # C2 (data.py:59 symmetrises 1.17 M edges)
    'products': (2_449_029, 61_859_140, 100, 47, 150),  # C3
    'reddit': (232_965, 114_615_892, 602, 41, 200),     # C4
    'amazonproducts': (1_569_960, 264_339_468, 200, 107, 200),  # C5
}

_PRIME = 1_000_003


def _skewed_local(u: Tensor, size: Tensor, power: float) -> Tensor:
    """Map uniform u in [0,1) to a skewed local index in [0, size), then scatter it over the block
    with a multiplicative hash so that popular nodes are not adjacent."""
    loc = (u.pow(power) * size).to(torch.int64)
    loc = torch.minimum(loc, size - 1)
    return (loc * _PRIME) % size


def synthetic_graph(num_nodes: int, num_edges: int, num_features: int, num_classes: int,
                    num_parts: int, seed: int = 0, device='cpu', p_inter: float = 0.15,
                    skew: float = 1.6, feature_dtype=torch.float32) -> Tuple[Data, Tensor]:
    """Returns ``(data, ptr)``: ``data.adj_t`` symmetric, unweighted, without self loops; ``x``
    ~ N(0,1); ``y`` block-correlated labels; 60/20/20 random masks; ``ptr`` the block boundaries."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N, P = int(num_nodes), int(num_parts)
    ptr = block_ptr(N, P)
    ptr_d = ptr.to(device)
    sizes = ptr_d[1:] - ptr_d[:-1]
    # ~3 % extra pairs make up for duplicates / self pairs removed below
    n_pairs = int(num_edges // 2 * 1.03) + 16

    def rnd(n):
        return torch.rand(n, generator=g, device=device)

    blk = torch.minimum((rnd(n_pairs) * P).to(torch.int64), torch.tensor(P - 1, device=device))
    src = ptr_d[blk] + _skewed_local(rnd(n_pairs), sizes[blk], skew)
    inter = rnd(n_pairs) < p_inter
    off = (-(1.0 - rnd(n_pairs)).log() * 2.0).to(torch.int64) + 1      # geometric-ish block offset
    sign = torch.where(rnd(n_pairs) < 0.5, -1, 1)
    dblk = torch.where(inter, (blk + sign * off) % P, blk)
    dst = ptr_d[dblk] + _skewed_local(rnd(n_pairs), sizes[dblk], skew)
    del blk, inter, off, sign, dblk
    keep = src != dst
    lo = torch.minimum(src, dst)[keep]
    hi = torch.maximum(src, dst)[keep]
    del src, dst, keep
    key = torch.unique(lo * N + hi)
    want = num_edges // 2
    if key.numel() > want:  # drop a seeded random subset to hit the named count exactly
        sel = torch.randperm(key.numel(), generator=g, device=device)[:want]
        key = key[sel]
    lo, hi = key // N, key % N
    row = torch.cat([lo, hi])
    col = torch.cat([hi, lo])
    k2, _ = torch.sort(row * N + col)
    row, col = k2 // N, k2 % N
    counts = torch.bincount(row, minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    adj_t = SparseTensor(rowptr=rowptr, col=col, value=None, sparse_sizes=(N, N), is_sorted=True)
    adj_t.clustered_parts = P

    x = torch.randn(N, num_features, generator=g, device=device, dtype=torch.float32).to(feature_dtype)
    # labels follow the block id (so that a GNN can learn something) with 30 % noise
    node_blk = torch.repeat_interleave(torch.arange(P, device=device), sizes)
    y = (node_blk * 7919) % num_classes
    noise = rnd(N) < 0.3
    y = torch.where(noise, (rnd(N) * num_classes).to(torch.int64).clamp_(max=num_classes - 1), y)
    # make the features weakly informative about the label
    x[:, :min(num_features, num_classes)] += 0.5 * torch.nn.functional.one_hot(
        y, num_classes)[:, :min(num_features, num_classes)].to(x.dtype)
    r = rnd(N)
    data = Data(x=x, y=y, adj_t=adj_t, train_mask=r < 0.6, val_mask=(r >= 0.6) & (r < 0.8),
                test_mask=r >= 0.8)
    return data, ptr


def get_data(root: str, name: str, seed: int = 0, device='cpu', scale: int = 1,
             num_parts: int = None) -> Tuple[Data, int, int]:
    """``get_data(root, name)`` of the reference (data.py:118-145) for the synthetic twins.
    ``scale`` divides nodes and edges (÷16 twins for CPU-oracle parity runs).  The partition
    boundaries are attached as ``data.ptr`` is NOT set; use ``metis(data.adj_t, num_parts)``."""
    key = name.lower().replace('ogbn-', '')
    if key not in SHAPES:
        raise NotImplementedError(f'no synthetic shape for dataset {name!r}')
    n, e, f, c, parts = SHAPES[key]
    parts = parts if num_parts is None else num_parts
    data, _ = synthetic_graph(max(n // scale, parts), max(e // scale, 2), f, c, parts, seed=seed,
                              device=device)
    return data, f, c
