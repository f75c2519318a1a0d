"""Timeline of ONE steady-state training step as the GPU ran it (CUDA-graph replay, all streams):
kernel start / end times from CUPTI (torch.profiler), relative to the step's first kernel.
This is what the serialised ncu launch list cannot show: which launches overlap, and where the
dependent chain waits.   usage: python tools/step_timeline.py [--prefill N] [--precision 3]"""
import argparse, json, os, sys, tempfile
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tgb-tgn-dgl_b200"))
import bench
from tgn_b200 import synth
from tgn_b200.engine import TGNEngine

ap = argparse.ArgumentParser()
ap.add_argument("--prefill", type=int, default=300_000)
ap.add_argument("--precision", type=int, default=3)
ap.add_argument("--workload", default=bench.WORKLOAD)
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--group", action="store_true", help="profile train_steps() (three steps per captured graph) and print "
                "the window between two consecutive steps INSIDE one graph replay")
ap.add_argument("--dp", action="store_true", help="with --eval: the sharded evaluation step eval_batch_dp (one rank)")
ap.add_argument("--eval", type=int, default=0, help="Q > 0: timeline of one evaluation batch with Q negatives per positive")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:          # torchrun: ONE job with the node memory partitioned by owner (timeline of rank 0)
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfg = synth.SHAPES[a.workload]
B, K = a.batch or cfg["B"], cfg["K"]
data = synth.synth_events(a.workload, seed=0, max_events=a.prefill + 200 * B, batch=B, extend=True)
N, De = data["num_nodes"], data["raw_dim"]
eng = TGNEngine(N, De, bench.HIDDEN, K, B, device=dev, lr=bench.LR, dropout=0.1, use_graph=True,
                log_capacity=data["src"].size, seed=1234, precision=a.precision, fused_zero_grad=True,
                rank=rank, world=world)
eng.load_state(*bench.init_state_dicts(De, bench.HIDDEN, N, seed=1))
eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
ring = bench.ring_after(data["src"][:a.prefill], data["dst"][:a.prefill], data["t"][:a.prefill], K, N)
eng.prefill(a.prefill, tuple(torch.from_numpy(x) for x in ring))
from torch.profiler import ProfilerActivity, profile
if a.eval:
    ev_t = {k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg")}
    def batch(b):
        sl = slice(a.prefill + b * B, a.prefill + (b + 1) * B)
        neg = torch.from_numpy(synth.eval_negatives(data["src"][sl], data["dst"][sl], N, a.eval, seed=1000 + b))
        return tuple(x.to(dev) for x in (ev_t["src"][sl], ev_t["dst"][sl], neg, ev_t["t"][sl], ev_t["msg"][sl]))
    bs = [batch(b) for b in range(12)]
    run = (lambda b: eng.eval_batch_dp(*b, 0, 1, reduce=True)) if a.dp else (lambda b: eng.eval_batch(*b))
    for b in bs[:8]:
        run(b)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for b in bs[8:]:
            run(b)
        torch.cuda.synchronize()
    marker = "nbr_count" if a.dp else "unique_mark"
elif a.group:
    eng.train_steps(61)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.train_steps(12)
        torch.cuda.synchronize()
    marker = "part_select_owned" if world > 1 else "msg_build"
else:
    for _ in range(30):
        eng.train_step(from_device=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(6):
            eng.train_step(from_device=True)
        torch.cuda.synchronize()
    marker = "part_select_owned" if world > 1 else "msg_build"
if rank:
    torch.cuda.synchronize(); dist.barrier(); os._exit(0)
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
starts = [i for i, e in enumerate(ev) if marker in e["name"]]
if a.eval:      # the first unique_mark of every batch: keep starts that are >200 us apart
    keep = [starts[0]]
    for i in starts[1:]:
        if ev[i]["ts"] - ev[keep[-1]]["ts"] > 200:
            keep.append(i)
    starts = keep
lo, hi = starts[-3], starts[-2]
if a.group:     # the closest pair of consecutive steps: both inside one graph replay
    gaps = [ev[starts[i + 1]]["ts"] - ev[starts[i]]["ts"] for i in range(len(starts) - 1)]
    print("step starts (us apart):", " ".join(f"{g:.1f}" for g in gaps))
    i = max(range(len(starts) - 3), key=lambda j: -(ev[starts[j + 3]]["ts"] - ev[starts[j]]["ts"]))
    lo, hi = starts[i], starts[i + 3]      # three consecutive steps, the tightest such window
t0 = ev[lo]["ts"]
streams = sorted({e["args"].get("stream") for e in ev[lo:hi]})
print(f"{a.workload} B={B}{' eval Q=%d' % a.eval if a.eval else ''}: step = {ev[hi]['ts'] - t0:.1f} us between two {marker} launches; streams {streams}")
print(f"{'start':>8s} {'end':>8s} {'dur':>7s}  st  kernel")
for e in ev[lo:hi]:
    name = e["name"].split("(")[0].replace("void ", "").replace("tgn::", "")[:48]
    print(f"{e['ts'] - t0:8.1f} {e['ts'] - t0 + e['dur']:8.1f} {e['dur']:7.1f}  {streams.index(e['args'].get('stream')):2d}  "
          f"{name}  grid={e['args'].get('grid')}")

if world > 1:
    sys.stdout.flush(); torch.cuda.synchronize(); dist.barrier(); os._exit(0)
