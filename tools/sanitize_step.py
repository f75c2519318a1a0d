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
torch.cuda.synchronize()
print("sanitize pass done")
