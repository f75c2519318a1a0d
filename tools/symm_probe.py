"""2+ ranks under torchrun: does torch symmetric memory work on this box (peer pointers + device barrier,
inside a captured CUDA graph)?  Prints per-call cost of the signal-pad barrier."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
out = torch.zeros(1024, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        hdl.barrier(channel=0); out.copy_(peer); hdl.barrier(channel=1)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        hdl.barrier(channel=0)
        out.copy_(peer)
        hdl.barrier(channel=1)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        g.replay()
    e1.record(); torch.cuda.synchronize()
print(rank, "graph(barrier, peer copy, barrier):", e0.elapsed_time(e1) / 200 * 1e3, "us; out", float(out[0]), flush=True)
dist.barrier(); dist.destroy_process_group()
