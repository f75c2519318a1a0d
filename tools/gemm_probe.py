"""Times the dense-path GEMM shapes of one training step on both engines
(fp32 CUDA cores vs tcgen05) with CUDA events; used to drive kernel work."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
from tgn_b200 import ops

dev = "cuda"
shapes = [  # name, m, n, k, trans_a, trans_b, split
    ("gi   x W_ih^T", 5023, 300, 472, False, False, 1),
    ("gh   h W_hh^T", 5023, 300, 100, False, False, 1),
    ("proj x Wn^T", 5023, 400, 100, False, False, 1),
    ("dec  z Ws^T", 200, 100, 100, False, False, 1),
    ("dWih dgi^T x", 300, 472, 5023, True, True, 16),
    ("dWn  dproj^T x", 400, 100, 5023, True, True, 16),
    ("dx   dproj Wn", 5023, 100, 400, False, True, 1),
    ("dxt  dgi Wih[:,t]", 5023, 100, 300, False, True, 1),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, m, n, k, ta, tb, split in shapes:
    if only and only not in name:
        continue
    A = torch.randn((k, m) if ta else (m, k), device=dev)
    B = torch.randn((k, n) if tb else (n, k), device=dev)
    out = torch.zeros(m, n, device=dev)
    res = []
    for prec in (0, 1, 3):
        run = lambda: ops.sgemm(A, B, m=m, n=n, k=k, lda=A.shape[1], ldb=B.shape[1], trans_a=ta, trans_b=tb,
                                out=out, split_k=split, prec=prec)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()          # graph replay: device time without the Python launch cost
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1e3)
    fl = 2.0 * m * n * k
    print(f"{name:20s} m={m:5d} n={n:4d} k={k:5d}  fp32 {res[0]:7.1f} us  tf32 {res[1]:7.1f} us  3xtf32 {res[2]:7.1f} us   ({fl/res[1]/1e6:6.1f} TF/s tf32)")
