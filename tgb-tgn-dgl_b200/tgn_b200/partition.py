"""Host-side layout arithmetic of the owner-partitioned node memory (the state is the
reference's `memory` [N,D] / `last_update` [N], modules/memory_module.py:80-83).

    owner(n)     = n % P          local_row(n) = n // P          shard of rank r = full[r::P]

These helpers are plumbing around the CUDA path (csrc/partition.cu does the per-step work on the
device): the shard <-> full conversions used when state is loaded or exported, and a device-agnostic
statement of the all-reduce row-assembly protocol that the world-size-2 gloo test checks on the CPU
(`tests/test_dist_partition_cpu.py`) -- the same protocol `tgn_part_gather` + one all-reduce implement."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


def owner(n: Tensor, world: int) -> Tensor:
    return n % world


def local_row(n: Tensor, world: int) -> Tensor:
    return torch.div(n, world, rounding_mode="floor")


def rows_per_rank(num_nodes: int, world: int) -> int:
    return (num_nodes + world - 1) // world


def shard(full: Tensor, rank: int, world: int) -> Tensor:
    """Rows of `full` owned by `rank`, padded with zeros to rows_per_rank (every shard has the same shape)."""
    own = full[rank::world]
    out = full.new_zeros((rows_per_rank(full.shape[0], world),) + tuple(full.shape[1:]))
    out[:own.shape[0]] = own
    return out


def interleave(shards: List[Tensor], num_nodes: int) -> Tensor:
    """Inverse of shard(): full[n] = shards[n % P][n // P]."""
    return torch.stack(list(shards), 1).reshape((-1,) + tuple(shards[0].shape[1:]))[:num_nodes].contiguous()


def stage_owned_rows(shard_rows: Tensor, ids: Tensor, rank: int, world: int) -> Tensor:
    """What tgn_part_gather writes: row i = shard[local_row(ids[i])] where this rank owns ids[i], zeros
    elsewhere (ids < 0 -> zeros)."""
    ok = (ids >= 0) & (owner(ids.clamp(min=0), world) == rank)
    out = shard_rows.new_zeros((ids.numel(),) + tuple(shard_rows.shape[1:]))
    out[ok] = shard_rows[local_row(ids[ok], world)]
    return out


def assemble_rows(shard_rows: Tensor, ids: Tensor, rank: int, world: int, group=None) -> Tensor:
    """All ranks end with full[ids]: the owners' disjoint contributions are summed by one all-reduce."""
    import torch.distributed as dist
    staged = stage_owned_rows(shard_rows, ids, rank, world)
    if world > 1:
        dist.all_reduce(staged, group=group)
    return staged


def scatter_owned(shard_rows: Tensor, ids: Tensor, values: Tensor, rank: int, world: int) -> None:
    """Owner-side write-back (tgn_memory_scatter_owned): shard[local_row(n)] = values[i] for owned ids."""
    ok = owner(ids, world) == rank
    shard_rows[local_row(ids[ok], world)] = values[ok]
