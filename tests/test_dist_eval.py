"""CPU (gloo, world_size 2): the data-parallel TGB evaluation host logic -- column sharding of
the negatives, the all-reduce of the two rank counts, the reference's truncation rule -- against
the single-process oracle MRR (bit-identical: the reduced quantities are integers)."""
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


def _scores(seed, B, Q):
    rng = np.random.default_rng(seed)
    pos = rng.random(B).astype(np.float32)
    neg = rng.random((B, Q)).astype(np.float32)
    neg[:, ::7] = pos[:, None]          # exact ties with the positive
    neg[::3, 1::5] = 1.0                # saturated sigmoid outputs
    return pos, neg


def _worker(rank, world, port, B, Q, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tgn_b200 import dist_eval
    res = []
    for batch in range(3):
        pos, neg = _scores(100 + batch, B, Q)
        neg_ids = torch.arange(B * Q).view(B, Q)      # stand-in node ids: column identity

        def count_fn(shard_ids):
            cols = (shard_ids[0] % Q).numpy()          # which columns this rank owns
            v = neg[:, cols]
            return (torch.from_numpy((v > pos[:, None]).sum(1).astype(np.int32)),
                    torch.from_numpy((v >= pos[:, None]).sum(1).astype(np.int32)))
        res.append(dist_eval.eval_batch_dp(count_fn, neg_ids, rank, world).numpy())
    np.save(os.path.join(out_dir, f"rr_{rank}.npy"), np.stack(res))
    dist.destroy_process_group()


@pytest.mark.parametrize("B,Q", [(17, 20), (8, 999), (5, 1)])
def test_dp_eval_counts_gloo_world2(tmp_path, B, Q):
    from oracle import tgn_oracle as orc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), B, Q, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"rr_{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])               # every rank ends with the same ranks
    for batch in range(3):
        pos, neg = _scores(100 + batch, B, Q)
        assert np.array_equal(got[0][batch], orc.mrr_ref(pos, neg).astype(np.float32))


def test_shards_are_a_partition_and_truncation_rule():
    from tgn_b200 import dist_eval
    neg = torch.arange(6 * 11).view(6, 11)
    for world in (1, 2, 3, 8, 16):
        parts = [dist_eval.shard_columns(neg, r, world) for r in range(world)]
        assert sum(p.shape[1] for p in parts) == 11
        assert sorted(torch.cat([p[0] for p in parts]).tolist()) == neg[0].tolist()
    # epoch_utils.py:48-56: rows cut to the shortest list of the batch
    t = dist_eval.truncate_negatives([[1, 2, 3, 4], [5, 6], [7, 8, 9]])
    assert t.tolist() == [[1, 2], [5, 6], [7, 8]]
    gt, ge = torch.tensor([0, 3]), torch.tensor([0, 5])
    assert torch.allclose(dist_eval.reciprocal_ranks(gt, ge), torch.tensor([1.0, 1.0 / 5.0]))
