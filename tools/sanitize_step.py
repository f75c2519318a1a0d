"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): a few eager training steps, one
evaluation batch, the t-CSR builder + sampler, the ring lookup -- every hot-path kernel at least once.
usage: compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_step.py"""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import ops, synth
from tgn_b200.engine import TGNEngine
from tgn_b200.neg_table import SyntheticNegatives, evaluate_table

dev = torch.device("cuda", 0)
N, De, D, K, B, steps = 300, 4, 32, 5, 40, 4
rng = np.random.default_rng(0)
E = B * (steps + 2)
src = rng.integers(0, N // 2, E); dst = rng.integers(N // 2, N, E)
t = np.sort(rng.integers(0, 4000, E)).astype(np.int64)
msg = rng.standard_normal((E, De)).astype(np.float32)
neg = rng.integers(N // 2, N, E)
eng = TGNEngine(N, De, D, K, B, device=dev, lr=1e-3, dropout=0.1, use_graph=False, log_capacity=E)
eng.load_state(*bench.init_state_dicts(De, D, N, seed=1))
ev = {k: torch.from_numpy(np.asarray(v)) for k, v in dict(src=src, dst=dst, t=t, msg=msg, neg=neg).items()}
eng.set_events(**ev)
for _ in range(steps):
    loss = eng.train_step(from_device=True)
print("train loss", float(loss))
eng.flush_to_eval()
sl = slice(steps * B, (steps + 1) * B)
mrr = evaluate_table(eng, ev["src"][sl], ev["dst"][sl], ev["t"][sl], ev["msg"][sl], SyntheticNegatives(20, N // 2, N), B)
print("eval mrr", mrr)
g = ops.tcsr_build(ev["src"].to(dev), ev["dst"].to(dev), ev["t"].to(dev), N)
coarse = ops.tcsr_build_index(g[3])
roots = torch.from_numpy(rng.integers(0, N, 500).astype(np.int32)).to(dev)
rts = torch.from_numpy(rng.integers(0, 4200, 500).astype(np.float32)).to(dev)
for strat in (0, 1):
    out, off, cnt = ops.tcsr_sample(g[0], g[1], g[2], g[3], roots, rts, 10, strat, coarse=coarse)
print("sampled", int(cnt))
# ---- round 2: the large-launch ring lookup (count + emit passes, >= 64 tiles), odd K
for K2, R2 in ((10, 20_000), (7, 30_000)):
    N2 = 5_000
    e2 = torch.randint(-1, 1000, (N2, K2), device=dev)
    nb2 = torch.randint(0, N2, (N2, K2), device=dev)
    t2 = torch.rand((N2, K2), device=dev)
    r2 = torch.randint(0, N2, (R2,), device=dev)
    o = ops.nbr_lookup_raw(r2, nb2, e2, t2, None)
    assert int(o[5]) == int((e2[r2] >= 0).sum())
print("large lookup ok")
# ---- the persistent GEMM kernel (more than two waves of tiles), both precisions, a device row count
A = torch.randn(40_000, 116, device=dev); W = torch.randn(100, 116, device=dev); C = torch.zeros(40_000, 100, device=dev)
md = torch.tensor([39_001], dtype=torch.int32, device=dev)
for prec in (3, 1):
    ops.gemm_batch([ops.gemm_desc(A, W, C, m=40_000, n=100, k=116, lda=116, ldb=116, ldc=100, m_dev=md)], prec)
print("persistent gemm", float(C[:39_001].abs().sum()) > 0)
# ---- sharded evaluation on a dense candidate set (in-place rows), one rank
eng2 = TGNEngine(60, De, D, K, B, device=dev, lr=1e-3, dropout=0.1, use_graph=False, log_capacity=E)
eng2.load_state(*bench.init_state_dicts(De, D, 60, seed=1))
s2 = rng.integers(0, 30, E); d2 = rng.integers(30, 60, E)
eng2.set_events(torch.from_numpy(s2), torch.from_numpy(d2), ev["t"], ev["msg"], torch.from_numpy(d2))
eng2.flush_to_eval()
negs = torch.from_numpy(rng.integers(30, 60, (B, 20)))
pos, gt, ge = eng2.eval_batch_dp(torch.from_numpy(s2[:B]), torch.from_numpy(d2[:B]), negs, ev["t"][:B], ev["msg"][:B], 0, 1)
assert eng2._eval_ctx_dp(B, 20, 0, 1).dense
print("dense eval", float(pos.sum()))
torch.cuda.synchronize()
print("sanitize pass done")
