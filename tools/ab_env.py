"""A/B on one box: ms/step of the device-resident arm under the environment the process was started with."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import synth
from tgn_b200.engine import TGNEngine
dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else bench.WORKLOAD
Bo = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg = synth.SHAPES[name]; B, K = Bo or cfg["B"], cfg["K"]
prefill = bench.PREFILL.get(name, 300_000)
data = synth.synth_events(name, seed=0, max_events=prefill + 2200 * B, batch=B, extend=True)
N, De = data["num_nodes"], data["raw_dim"]
ring = bench.ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
eng = TGNEngine(N, De, bench.HIDDEN, K, B, device=dev, lr=bench.LR, dropout=0.1, use_graph=True,
                log_capacity=data["src"].size, seed=1234, precision=3, fused_zero_grad=True)
eng.load_state(*bench.init_state_dicts(De, bench.HIDDEN, N, seed=1))
eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
eng.prefill(prefill, tuple(torch.from_numpy(x) for x in ring))
eng.train_steps(61)
torch.cuda.synchronize()
res = []
for _ in range(3):
    n = 600
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.train_steps(n); b.record(); torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / n * 1e3)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("TGN_"))
print(f"{name} B={B} [{tag}]: " + " ".join(f"{x:.2f}" for x in res) + f" us/step  loss {float(eng.loss):.4f}", flush=True)
