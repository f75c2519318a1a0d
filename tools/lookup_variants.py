"""Experiment: large-launch variants of the ring lookup (TGN_LOOKUP_VARIANT), per-kernel times from CUPTI."""
import os, sys, json, tempfile
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
from tgn_b200 import ops
from torch.profiler import ProfilerActivity, profile
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
N, K = 352_637, 10
nb = torch.randint(0, N, (N, K), generator=g).to(dev); ei = torch.randint(0, 1 << 30, (N, K), generator=g).to(dev)
ei[torch.rand(N, K, generator=g).to(dev) < 0.1] = -1
tt = torch.rand(N, K, generator=g).to(dev)
r2 = torch.randint(0, N, (1_000_000,), generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: ops.nbr_lookup_raw(r2, nb, ei, tt, None)
out = fn()
torch.cuda.synchronize()
c = int(out[5].item())
w = torch.arange(1, c + 1, device=dev, dtype=torch.float64)
sig = [c] + [float((x[:c].double() * w).sum().item()) for x in out[:4]] + [int(out[4].long().sum().item())]
times = []
for _ in range(5):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) * 1e3)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    flush.zero_(); fn(); torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json"); prof.export_chrome_trace(path)
ks = [(e["name"].split("(")[0][-40:], e["dur"]) for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and "elementwise" not in e["name"] and "fill" not in e["name"].lower()]
print("variant", os.environ.get("TGN_LOOKUP_VARIANT", "0"), "us", " ".join(f"{t:.1f}" for t in times), "| kernels", ks, "| sig", sig)
