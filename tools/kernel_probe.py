"""Launches one hot-path kernel on a large input a few times (for ncu captures / quick timing).
usage: python tools/kernel_probe.py {gemm3|gemm1|gru_pair|tcsr|tcsr_plain|lookup|agg_last|agg_mean}
TGN_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity first (experiment only)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
from tgn_b200 import ops

which = sys.argv[1]
dev = "cuda"
torch.zeros(1, device=dev)
if os.environ.get("TGN_L2_FETCH"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["TGN_L2_FETCH"])))   # cudaLimitMaxL2FetchGranularity
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
    print("L2 fetch granularity", rc, v.value)
g = torch.Generator(device="cpu").manual_seed(0)
N = 352_637
if which == "gru_pair":
    # the step's heaviest launch: gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh in one tgn_gemm_batch
    S, Dx, ldx, D = 5_023, 301, 304, 100
    x = torch.randn(S, ldx, device=dev); h = torch.randn(S, D, device=dev)
    wih = torch.randn(3 * D, ldx, device=dev); whh = torch.randn(3 * D, D, device=dev)
    bi = torch.randn(3 * D, device=dev); bh = torch.randn(3 * D, device=dev)
    gi = torch.empty(S, 3 * D, device=dev); gh = torch.empty(S, 3 * D, device=dev)
    fn = lambda: ops.gemm_batch([
        ops.gemm_desc(x, wih, gi, m=S, n=3 * D, k=Dx, lda=ldx, ldb=ldx, ldc=3 * D, bias=bi),
        ops.gemm_desc(h, whh, gh, m=S, n=3 * D, k=D, lda=D, ldb=D, ldc=3 * D, bias=bh)], 3)
elif which in ("gru_fused", "gru_fused1"):
    # the step's GRUCell forward as one launch (tgn_gru_fused_fwd) on step-sized operands
    S, Dx, ldx, D = 5_023, 301, 304, 100
    x = torch.randn(S, ldx, device=dev); h = torch.randn(S, D, device=dev)
    wih = torch.randn(3 * D, ldx, device=dev); whh = torch.randn(3 * D, D, device=dev)
    bi = torch.randn(3 * D, device=dev); bh = torch.randn(3 * D, device=dev)
    z = torch.empty(S, D, device=dev); gates = torch.empty(S, 4 * D, device=dev)
    fn = lambda: ops.gru_fused_fwd(x, h, wih, whh, bi, bh, dx=Dx, ldx=ldx, ldw=ldx, num=S, prec=3 if which == "gru_fused" else 1, out=z, gates=gates)
elif which.startswith("gemm"):
    # gemm3 / gemm1 [ _ld480 : rows padded to a multiple of 128 bytes ] [ _small : 5,023 rows as in the step ]
    S, Dx, D = (5_023 if "small" in which else 65_536), 472, 100
    ld = 480 if "ld480" in which else Dx
    x = torch.randn(S, ld, device=dev); w = torch.randn(3 * D, ld, device=dev); o = torch.empty(S, 3 * D, device=dev)
    fn = lambda: ops.sgemm(x, w, m=S, n=3 * D, k=Dx, lda=ld, ldb=ld, out=o, prec=3 if which.startswith("gemm3") else 1)
elif which in ("tcsr", "tcsr_plain"):
    deg = 113
    indptr = (torch.arange(N + 1, dtype=torch.int64) * deg).to(torch.int32).to(dev)
    ts = torch.arange(deg, dtype=torch.float32).repeat(N).to(dev)
    indices = torch.randint(0, N, (N * deg,), generator=g, dtype=torch.int32).to(dev)
    eid = torch.arange(N * deg, dtype=torch.int32, device=dev)
    R = 2_000_000
    roots = torch.randint(0, N, (R,), generator=g, dtype=torch.int32).to(dev)
    rts = (torch.rand(R, generator=g) * deg + 12).to(dev)
    coarse = ops.tcsr_build_index(ts) if which == "tcsr" else None
    fn = lambda: ops.tcsr_sample(indptr, indices, eid, ts, roots, rts, 10, coarse=coarse)
elif which == "lookup":
    K = 10
    nb = torch.randint(0, N, (N, K), generator=g).to(dev); ei = torch.randint(0, 1 << 30, (N, K), generator=g).to(dev)
    tt = torch.rand(N, K, generator=g).to(dev)
    r2 = torch.randint(0, N, (1_000_000,), generator=g).to(dev)
    fn = lambda: ops.nbr_lookup_raw(r2, nb, ei, tt, None)
else:
    M, S, W = 1_000_000, 250_000, 472
    msg = torch.randn(M, W, device=dev); idx = torch.randint(0, S, (M,), generator=g).to(dev)
    tm = torch.randint(0, 1000, (M,), generator=g).to(dev)
    fn = (lambda: ops.agg_last(msg, idx, tm, S)) if which == "agg_last" else (lambda: ops.agg_mean(msg, idx, S))
for _ in range(3):
    fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(which, f"{e0.elapsed_time(e1) * 1e2:.1f} us per launch (graph of 10)")
