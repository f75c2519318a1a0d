"""Where the t-CSR two-layer training step (configs[3], module path) spends its time: CPU-side profile with
kernel counts.  usage: python tools/tcsr_profile.py"""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import ops, synth
from tgn_b200.tcsr_trainer import TCSRTrainer
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda", 0)
cfg = synth.SHAPES["tgbl-comment"]; B, K = cfg["B"], cfg["K"]
data = synth.synth_events("tgbl-comment", seed=0, max_events=1_000_000)
N, De, E = data["num_nodes"], data["raw_dim"], data["src"].size
s_d, d_d, t_d = (torch.from_numpy(data[k]).to(dev) for k in ("src", "dst", "t"))
indptr, indices, eid, ts = ops.tcsr_build(s_d, d_d, t_d, N, t_sorted=True)
tr = TCSRTrainer(indptr, indices, eid, ts, N, De, bench.HIDDEN, [K, K], False, torch.from_numpy(data["msg"]), device=dev,
                 lr=bench.LR, dropout=0.1, seed=7)
tr.train()
ev = {k: torch.from_numpy(data[k]).to(dev) for k in ("src", "dst", "neg", "t")}
msg = torch.from_numpy(data["msg"]).to(dev)
lo0 = E - 60 * B
def step(i):
    sl = slice(lo0 + i * B, lo0 + (i + 1) * B)
    return tr.train_step(ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], msg[sl])
for i in range(8):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8, 28):
    step(i)
torch.cuda.synchronize()
print(f"ms/step {(time.perf_counter() - t0) / 20 * 1e3:.2f}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(28, 32):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=20, max_name_column_width=60))
