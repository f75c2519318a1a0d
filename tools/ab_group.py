"""A/B on one box: steps per captured graph (TGNEngine group_size) vs ms/step of the device-resident arm."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import synth
from tgn_b200.engine import TGNEngine
dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else bench.WORKLOAD
cfg = synth.SHAPES[name]; B, K = cfg["B"], cfg["K"]
prefill = bench.PREFILL.get(name, 300_000)
data = synth.synth_events(name, seed=0, max_events=prefill + 3000 * B, batch=B, extend=True)
N, De = data["num_nodes"], data["raw_dim"]
ring = bench.ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
for G in (1, 3, 6, 12, 24):
    eng = TGNEngine(N, De, bench.HIDDEN, K, B, device=dev, lr=bench.LR, dropout=0.1, use_graph=True,
                    log_capacity=data["src"].size, seed=1234, precision=3, fused_zero_grad=True, group_size=G)
    eng.load_state(*bench.init_state_dicts(De, bench.HIDDEN, N, seed=1))
    eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
    eng.prefill(prefill, tuple(torch.from_numpy(x) for x in ring))
    eng.train_steps(1 + 16 * G)
    torch.cuda.synchronize()
    n = 48 * 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.train_steps(n); b.record(); torch.cuda.synchronize()
    print(f"{name} group_size {G}: {a.elapsed_time(b) / n * 1e3:.2f} us/step  loss {float(eng.loss):.4f}", flush=True)
    del eng
