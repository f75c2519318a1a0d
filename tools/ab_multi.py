"""A/B on one box: 1000 training steps as single-step graph launches vs three-step graphs."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import synth
from tgn_b200.engine import TGNEngine
dev = torch.device("cuda", 0)
cfg = synth.SHAPES[bench.WORKLOAD]; B, K = cfg["B"], cfg["K"]
prefill = 1_000_000
data = synth.synth_events(bench.WORKLOAD, seed=0, max_events=prefill + 8000 * B)
N, De = data["num_nodes"], data["raw_dim"]
eng = TGNEngine(N, De, bench.HIDDEN, K, B, device=dev, lr=bench.LR, dropout=0.1, use_graph=True,
                log_capacity=data["src"].size, seed=1234, precision=3, fused_zero_grad=True)
eng.load_state(*bench.init_state_dicts(De, bench.HIDDEN, N, seed=1))
eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
ring = bench.ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
eng.prefill(prefill, tuple(torch.from_numpy(x) for x in ring))
eng.train_steps(40)
for _ in range(20):
    eng.train_step(from_device=True)
torch.cuda.synchronize()
def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
for rep in range(3):
    t1 = timed(lambda: [eng.train_step(from_device=True) for _ in range(999)])
    t3 = timed(lambda: eng.train_steps(999))
    print(f"rep {rep}: single-step graphs {t1 / 999 * 1e3:.2f} us/step   three-step graphs {t3 / 999 * 1e3:.2f} us/step")
