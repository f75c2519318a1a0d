"""Run under torchrun on >= 2 GPUs (see tests/test_gpu_multi.py): data-parallel TGB evaluation with the
embedding sharded over the ranks (TGNEngine.eval_batch_dp: roots dealt round-robin, decoder-projected rows
all-gathered, negative columns sharded, integer counts all-reduced) against single-replica evaluation of the
same batches: per-batch reciprocal ranks bit-identical up to score ties, identical state on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tgb-tgn-dgl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from oracle import tgn_oracle as orc
    from tgn_b200 import dist_eval, synth
    from tgn_b200.engine import TGNEngine
    for N, Q in ((403, 30), (90, 41)):         # sparse roots (unique + relabel) / dense roots (every node)
        De, D, K, B, steps = 12, 32, 5, 50, 8
        rng = np.random.default_rng(3)
        E = B * steps
        ns = N // 2
        src = np.floor(rng.random(E) ** 2 * ns).astype(np.int64)
        dst = ns + np.floor(rng.random(E) ** 2 * (N - ns)).astype(np.int64)
        t = np.sort(rng.integers(0, 40 * E, E)).astype(np.int64)
        msg = rng.standard_normal((E, De)).astype(np.float32)
        ref = orc.build_model(De, D, N, seed=5)
        with torch.no_grad():
            ref["memory"].time_enc.lin.weight.mul_(0.002)
        ev = dict(src=torch.from_numpy(src), dst=torch.from_numpy(dst), t=torch.from_numpy(t), msg=torch.from_numpy(msg),
                  neg=torch.from_numpy(dst.copy()))
        engs = []
        for _ in range(2):
            eng = TGNEngine(N, De, D, K, B, device=dev, lr=1e-4, dropout=0.0, use_graph=True, log_capacity=E)
            eng.load_state(ref["memory"].state_dict(), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
            eng.set_events(**ev)
            eng.flush_to_eval()
            engs.append(eng)
        single, sharded = engs
        batches = []
        for s in range(steps):
            sl = slice(s * B, (s + 1) * B)
            neg = torch.from_numpy(synth.eval_negatives(src[sl], dst[sl], N, Q, seed=s, dst_lo=ns))
            batches.append(tuple(x.to(dev) for x in (ev["src"][sl], ev["dst"][sl], neg, ev["t"][sl], ev["msg"][sl])))
        for b in batches:
            pos, _, gt, ge = single.eval_batch(*b, want_neg_scores=False)
            rr_single = dist_eval.reciprocal_ranks(gt, ge).clone()
            pos_s, gt_s, ge_s = sharded.eval_batch_dp(*b, rank, world)
            gt_s, ge_s = dist_eval.reduce_counts(gt_s, ge_s)
            rr = dist_eval.reciprocal_ranks(gt_s, ge_s)
            torch.testing.assert_close(pos_s, pos, rtol=1e-5, atol=1e-6)
            # the gathered rows are bit-identical to the locally computed ones (same kernels, same inputs), so are the counts
            assert torch.equal(rr, rr_single), (N, Q, rank)
        torch.cuda.synchronize()
        assert sharded._eval_ctx_dp(B, Q, rank, world).dense == (N == 90)
        assert torch.equal(sharded.last_update, single.last_update) and torch.equal(sharded.e_id, single.e_id)
        torch.testing.assert_close(sharded.memory, single.memory, rtol=1e-5, atol=1e-6)
        sharded.check_device_errors()
    dist.barrier()
    if rank == 0:
        print(f"sharded eval check OK: world={world}, sparse and dense roots, reciprocal ranks identical to one replica", flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
