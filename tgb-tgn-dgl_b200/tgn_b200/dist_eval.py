"""Data-parallel TGB evaluation (reference epoch_utils.py:28-165).

Each positive edge is ranked against its Q pre-generated negatives; the Q scores are
independent given the embeddings, so the negative matrix [B, Q] is split by columns across
the ranks.  Every rank keeps a full replica of the (small) model state, scores its own
column shard, and the two integer counts of the TGB rank formula

    rank_i = 1 + (#{neg > pos_i} + #{neg >= pos_i}) / 2          (tgb Evaluator)

are summed over ranks with ONE all-reduce of 2*B int32 per batch.  The counts are exact
integers, so the MRR is bit-identical to single-GPU evaluation.  The reference's quirks are
kept: negatives are truncated to the shortest list of the batch (epoch_utils.py:48-56) and
the epoch metric is the mean of per-batch means (epoch_utils.py:113,163).

The state update of a batch (update_state + insert, epoch_utils.py:155-157) depends only on
the positives, so every rank applies it redundantly and the replicas never diverge.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def truncate_negatives(neg_lists: Sequence[Sequence[int]]) -> Tensor:
    """epoch_utils.py:48-56: all rows are cut to the shortest negative list of the batch."""
    q = min(len(r) for r in neg_lists)
    return torch.tensor([list(r[:q]) for r in neg_lists], dtype=torch.long)


def shard_columns(neg: Tensor, rank: int, world: int) -> Tensor:
    """Column shard of the [B, Q] negative matrix owned by `rank` (round-robin, so ranks stay
    balanced for any Q); shards are disjoint and cover every column."""
    return neg[:, rank::world].contiguous()


def reduce_counts(gt: Tensor, ge: Tensor, group=None) -> Tuple[Tensor, Tensor]:
    """Sum of the per-positive counts over the ranks of `group` (one all-reduce)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        both = torch.stack([gt, ge]).to(torch.int32)
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
        return both[0], both[1]
    return gt, ge


def reciprocal_ranks(gt: Tensor, ge: Tensor) -> Tensor:
    return 1.0 / (0.5 * (gt.to(torch.float32) + ge.to(torch.float32)) + 1.0)


def eval_batch_dp(count_fn: Callable[[Tensor], Tuple[Tensor, Tensor]], neg: Tensor, rank: int, world: int,
                  group=None) -> Tensor:
    """One evaluation batch.  `count_fn(neg_shard) -> (gt[B], ge[B])` scores the positives against a
    column shard (TGNEngine.eval_batch on the GPU; the CPU oracle in the gloo tests) and applies the
    batch's state update.  Returns the per-positive reciprocal ranks (identical on every rank)."""
    gt, ge = count_fn(shard_columns(neg, rank, world))
    gt, ge = reduce_counts(gt, ge, group)
    return reciprocal_ranks(gt, ge)


def evaluate_dp(engine, batches, rank: int = 0, world: int = 1, group=None, shard_embeddings: bool = False) -> float:
    """epoch loop of test() (epoch_utils.py:28-165) on TGNEngine replicas: `batches` yields
    (src, dst, neg[B,Q], t, msg); returns the epoch MRR (mean of per-batch means).
    shard_embeddings=True: TGNEngine.eval_batch_dp -- every root of a batch is embedded on exactly one rank and
    the decoder-projected rows are all-gathered, instead of every rank embedding every root."""
    per_batch: List[Tensor] = []
    if shard_embeddings:
        # counts all-reduced and the epoch metric accumulated inside the captured step: one host read per epoch
        if hasattr(engine, "mrr_acc"):
            engine.mrr_acc.zero_()
        for src, dst, neg, t, msg in batches:
            engine.eval_batch_dp(src, dst, neg, t, msg, rank, world, group, reduce=True)
        return engine.epoch_mrr()
    for src, dst, neg, t, msg in batches:
        def count_fn(shard, _a=(src, dst, t, msg)):
            _, _, gt, ge = engine.eval_batch(_a[0], _a[1], shard, _a[2], _a[3], want_neg_scores=False)
            return gt, ge
        per_batch.append(eval_batch_dp(count_fn, neg, rank, world, group).mean())
    return float(torch.stack(per_batch).mean())
