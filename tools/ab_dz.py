"""A/B on one box: training step with the d_z GEMM split-K 1 vs 3."""
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
data = synth.synth_events(bench.WORKLOAD, seed=0, max_events=prefill + 5000 * B)
N, De = data["num_nodes"], data["raw_dim"]
ring = bench.ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
res = {}
for split in (1, 3, 1, 3):
    eng = TGNEngine(N, De, bench.HIDDEN, K, B, device=dev, lr=bench.LR, dropout=0.1, use_graph=True,
                    log_capacity=data["src"].size, seed=1234, precision=3, fused_zero_grad=True)
    eng.dz_split = split
    eng.load_state(*bench.init_state_dicts(De, bench.HIDDEN, N, seed=1))
    eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
    eng.prefill(prefill, tuple(torch.from_numpy(x) for x in ring))
    eng.train_steps(60)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.train_steps(900); b.record(); torch.cuda.synchronize()
    print(f"dz_split {split}: {a.elapsed_time(b) / 900 * 1e3:.2f} us/step  loss {float(eng.loss):.4f}", flush=True)
    del eng
