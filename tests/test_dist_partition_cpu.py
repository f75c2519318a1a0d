"""CPU (gloo, world_size 2 and 3): the owner-partitioned memory protocol -- shard layout n % P / n // P,
row assembly by one all-reduce of the owners' disjoint contributions, owner-side write-back -- against
plain indexing of the full matrix.  The device kernels (csrc/partition.cu) are checked against the same
layout on the GPU (tests/test_gpu_kernels.py, tests/dist_partition_check.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tgb-tgn-dgl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, D, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tgn_b200 import partition as pt
    g = torch.Generator().manual_seed(7)                     # every rank builds the same full state
    full = torch.randn(N, D, generator=g)
    lu = torch.randint(0, 10_000, (N,), generator=g)
    mem_s, lu_s = pt.shard(full, rank, world), pt.shard(lu, rank, world)
    ok = True
    for step in range(4):
        ids = torch.randint(0, N, (57,), generator=g)
        ids[::9] = -1                                        # rows past the live count / nodes without events
        rows = pt.assemble_rows(mem_s, ids, rank, world)
        lus = pt.assemble_rows(lu_s, ids, rank, world)
        want = torch.where((ids >= 0)[:, None], full[ids.clamp(min=0)], torch.zeros(1))
        ok &= torch.equal(rows, want) and torch.equal(lus, torch.where(ids >= 0, lu[ids.clamp(min=0)], 0))
        # a batch touches some nodes: every rank computes the same new rows, owners write them back
        upd = torch.unique(torch.randint(0, N, (23,), generator=g))
        new = torch.randn(upd.numel(), D, generator=g)
        full[upd] = new
        pt.scatter_owned(mem_s, upd, new, rank, world)
    shards = [torch.empty_like(mem_s) for _ in range(world)]
    dist.all_gather(shards, mem_s)
    ok &= torch.equal(pt.interleave(shards, N), full)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array(int(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 101), (3, 64)])
def test_partitioned_rows_assemble_gloo(tmp_path, world, N):
    mp.spawn(_worker, args=(world, _free_port(), N, 6, str(tmp_path)), nprocs=world, join=True)
    assert all(int(np.load(tmp_path / f"ok_{r}.npy")) == 1 for r in range(world))


def test_shard_layout_roundtrip():
    from tgn_b200 import partition as pt
    full = torch.arange(23 * 3, dtype=torch.float32).view(23, 3)
    for P in (1, 2, 3, 8):
        shards = [pt.shard(full, r, P) for r in range(P)]
        assert all(s.shape[0] == pt.rows_per_rank(23, P) for s in shards)
        assert torch.equal(pt.interleave(shards, 23), full)
        n = torch.arange(23)
        assert torch.equal(torch.stack([shards[int(o)][int(l)] for o, l in zip(pt.owner(n, P), pt.local_row(n, P))]), full)
