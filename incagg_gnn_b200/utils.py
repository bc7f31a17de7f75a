"""Driver-loop helpers with the reference's names and semantics
(torch_geometric_autoscale/utils.py:9-73): ``index2mask``, ``compute_micro_f1``, ``gen_masks``, edge
``dropout``.  Tensor programs only (they are inputs / metrics of the hot path, not part of it)."""
from typing import Optional, Tuple

import torch
from torch import Tensor

from .sparse import SparseTensor


def index2mask(idx: Tensor, size: int) -> Tensor:
    """Boolean membership mask of ``idx`` over ``range(size)``."""
    return torch.zeros(size, dtype=torch.bool, device=idx.device).index_fill_(0, idx, True)


def accuracy(logits: Tensor, y: Tensor) -> float:
    """Fraction of rows whose arg-max class equals the label."""
    return int((logits.argmax(dim=-1) == y).sum()) / max(y.size(0), 1)


def multilabel_counts(logits: Tensor, y: Tensor) -> Tuple[int, int, int]:
    """(true positives, predicted positives, actual positives) with the reference's thresholds:
    a logit above 0 predicts the label, a target above 0.5 carries it."""
    pred, true = logits > 0, y > 0.5
    return int((pred & true).sum()), int(pred.sum()), int(true.sum())


def compute_micro_f1(logits: Tensor, y: Tensor, mask: Optional[Tensor] = None) -> float:
    """The metric main.py:239-243 logs per epoch: accuracy for single-label targets (1-D ``y``),
    micro-averaged F1 for multi-label ones; 0.0 when precision or recall is undefined."""
    if mask is not None:
        logits, y = logits[mask], y[mask]
    if y.dim() == 1:
        return accuracy(logits, y)
    tp, n_pred, n_true = multilabel_counts(logits, y)
    if tp == 0 or n_pred == 0 or n_true == 0:
        return 0.
    precision, recall = tp / n_pred, tp / n_true
    return 2 * precision * recall / (precision + recall)


def sharded_micro_f1(logits_local: Tensor, y: Tensor, mask: Tensor, shard) -> float:
    """``compute_micro_f1`` when every rank holds the logits of its own node range only
    (``ScalableGNN._out`` with sharded histories): per-rank counts, one scalar all-reduce."""
    import torch.distributed as dist
    lo, hi = shard.lo, shard.hi
    yl, ml = y[lo:hi].to(logits_local.device), mask[lo:hi].to(logits_local.device)
    lg, yl = logits_local[ml], yl[ml]
    if yl.dim() == 1:
        stats = torch.tensor([float((lg.argmax(-1) == yl).sum()), float(yl.size(0))], device=lg.device)
        if shard.world_size > 1:
            dist.all_reduce(stats, group=shard.group)
        return float(stats[0]) / max(float(stats[1]), 1.)
    stats = torch.tensor([float(v) for v in multilabel_counts(lg, yl)], device=lg.device)
    if shard.world_size > 1:
        dist.all_reduce(stats, group=shard.group)
    tp, n_pred, n_true = stats.tolist()
    if tp == 0 or n_pred == 0 or n_true == 0:
        return 0.
    return 2 * (tp / n_pred) * (tp / n_true) / (tp / n_pred + tp / n_true)


def gen_masks(y: Tensor, train_per_class: int = 20, val_per_class: int = 30,
              num_splits: int = 20) -> Tuple[Tensor, Tensor, Tensor]:
    """``num_splits`` random splits: per class ``train_per_class`` training and ``val_per_class``
    validation nodes, everything else test.  Masks are ``[num_nodes, num_splits]``."""
    n = y.size(0)
    train_mask = torch.zeros(n, num_splits, dtype=torch.bool)
    val_mask = torch.zeros(n, num_splits, dtype=torch.bool)
    cols = torch.arange(num_splits)
    for c in range(int(y.max()) + 1):
        members = torch.nonzero(y == c).view(-1)
        # one independent shuffle of the class members per split
        order = torch.stack([members[torch.randperm(members.numel())] for _ in range(num_splits)], dim=1)
        tr, va = order[:train_per_class], order[train_per_class:train_per_class + val_per_class]
        train_mask[tr, cols.expand_as(tr)] = True
        val_mask[va, cols.expand_as(va)] = True
    return train_mask, val_mask, ~(train_mask | val_mask)


def dropout(adj_t: SparseTensor, p: float, training: bool = True) -> SparseTensor:
    """Edge dropout: weighted graphs get inverted-dropout on the edge values (structure kept),
    unweighted graphs lose each edge with probability ``p``."""
    if p == 0. or not training:
        return adj_t
    value = adj_t.storage.value()
    if value is None:
        keep = torch.rand(adj_t.nnz(), device=adj_t.device) > p
        return adj_t.masked_select_nnz(keep, layout='coo')
    return adj_t.set_value(torch.nn.functional.dropout(value, p=p), layout='coo')
