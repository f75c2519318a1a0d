#!/usr/bin/env python
"""Headline benchmark: train events/sec of the TGN step (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo (B200 path)
    python bench.py --impl reference --steps K --warmup W    # reference CPU path (oracle port)

Workload at N=1 (BASELINE.json configs[1]): synthetic tgbl-review shape (352,637 nodes,
D_e=1), batch 200, 10 most-recent neighbours, memory/time/embedding dim 100, Adam lr 1e-4,
attention dropout 0.1.  One step = sample -> memory -> attention -> decode -> BCE ->
backward -> Adam -> update_state -> insert on one batch.  The step starts from a mid-epoch
state (neighbour rings filled by the first `--prefill` events) so that every root has its K
neighbours, as in steady-state training.

value  : events/s with the event arrays resident in HBM (CUDA-graph replay, CUDA events).
e2e    : events/s through TGNEngine with HOST batches: per step two H2D copies from pinned
         memory (ids+timestamps, messages) and a D2H read of the loss.
N > 1  : the training step does not shard across batches (batch i+1 reads the memory batch i
         wrote, memory_module.py:126-138): N independent replicas, weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "tgb-tgn-dgl_b200")
for _p in (REPO, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOAD = "tgbl-review"
HIDDEN = 100
LR = 1e-4


def ring_after(src, dst, t, K, N):
    """Neighbour ring (neighbors, e_id, t) after inserting events 0..len-1 in order -- the K
    largest event ids per node (neighbor_loader.py:52-104), built vectorised on the host."""
    E = src.size
    nodes = np.concatenate([dst, src])
    nbrs = np.concatenate([src, dst])
    eid = np.concatenate([np.arange(E), np.arange(E)])
    order = np.lexsort((-eid, nodes))
    nodes, nbrs, eid = nodes[order], nbrs[order], eid[order]
    start = np.flatnonzero(np.r_[True, nodes[1:] != nodes[:-1]])
    rank = np.arange(nodes.size) - np.repeat(start, np.diff(np.r_[start, nodes.size]))
    keep = rank < K
    nb = np.zeros((N, K), np.int64)
    ei = np.full((N, K), -1, np.int64)
    tt = np.full((N, K), -1.0, np.float32)
    nb[nodes[keep], rank[keep]] = nbrs[keep]
    ei[nodes[keep], rank[keep]] = eid[keep]
    tt[nodes[keep], rank[keep]] = t[eid[keep]].astype(np.float32)
    return nb, ei, tt


def init_state_dicts(raw_dim, hidden, num_nodes, seed):
    """Random-init weights with the reference's initialisers (pyg_model_utils.getModel)."""
    from modules.decoder import LinkPredictor
    from modules.emb_module import GraphAttentionEmbedding
    from modules.memory_module import TGNMemory
    from modules.msg_agg import LastAggregator
    from modules.msg_func import IdentityMessage
    torch.manual_seed(seed)
    mem = TGNMemory(1, raw_dim, hidden, hidden, IdentityMessage(raw_dim, hidden, hidden), LastAggregator())
    gnn = GraphAttentionEmbedding(hidden, hidden, raw_dim, mem.time_enc)
    lp = LinkPredictor(hidden)
    strip = {k: v for k, v in mem.state_dict().items() if k not in ("memory", "last_update", "_assoc")}
    return strip, gnn.state_dict(), lp.state_dict()


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons during the timed regions: NVML polled every few
    milliseconds (falls back to one `nvidia-smi` query per sample when pynvml is unusable)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.mx, self.reasons, self.how = [], [], set(), "nvidia-smi"
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h, self.nv, self.how = pynvml.nvmlDeviceGetHandleByIndex(phys), pynvml, "nvml"
        except Exception:
            self.h = None

    def _sample_nvml(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        self.reasons |= {n for n, b in self.BITS.items() if mask & b}

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        r = [c.strip() for c in out.strip().split(",")]
        if len(r) > 8 and r[1].replace(".", "").isdigit():
            self.sm.append(float(r[1]))
            self.mx.append(float(r[2]))
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.reasons |= {n for n, v in zip(names, r[5:9]) if v == "Active"}

    def run(self):
        while not self.stop_flag:
            try:
                self._sample_nvml() if self.h is not None else self._sample_smi()
            except Exception:
                if self.h is not None:
                    self.h = None       # NVML call failed: fall back to nvidia-smi
                    self.how = "nvidia-smi"
            time.sleep(0.005)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.how}


# --------------------------------------------------------------------------- workloads
PREFILL = {"tgbl-review": 1_000_000, "tgbl-wiki": 50_000}


def workload_text(name, data, B, K, prefill):
    return (f"synthetic {name} shape: {data['num_nodes']} nodes, raw_dim {data['raw_dim']}, batch {B}, "
            f"{K} recent nbrs, dim {HIDDEN}, Adam lr {LR}, ring prefilled with {prefill} events")


def load_workload(name, B, prefill, n_batches, seed=0):
    """Event stream of one workload: `prefill` events that only fill the neighbour rings + n_batches batches.
    The wiki shape has 157k events; longer runs extend the stream at the same event rate (synth.extend)."""
    from tgn_b200 import synth
    return synth.synth_events(name, seed=seed, max_events=prefill + n_batches * B, batch=B, extend=True)


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_state(data, prefill, seed=1, dropout=True):
    """Oracle model + neighbour loader in the state the GPU arm starts from (rings after `prefill` events)."""
    from oracle import tgn_oracle as orc
    N, De, K = data["num_nodes"], data["raw_dim"], data["K"]
    model = orc.build_model(De, HIDDEN, N, seed=seed)
    for m in model.values():
        m.train()
    if not dropout:
        model["gnn"].conv.dropout = 0.0
    loader = orc.TorchNeighborLoader(N, K)
    nb, ei, tt = ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
    loader.neighbors, loader.e_id, loader.t = torch.from_numpy(nb), torch.from_numpy(ei), torch.from_numpy(tt)
    loader.cur_e_id = prefill
    opt = torch.optim.Adam(orc.model_parameters(model), lr=LR)
    return model, loader, opt


def run_cpu_reference(data, steps, warmup, prefill, budget_s):
    """The reference's CPU path for the same step: oracle port of modules/* + neighbor_loader
    (Python-dict message store included -- that is the reference's algorithm), torch CPU with
    all host threads.  Bounded by `budget_s` seconds of timed work."""
    from oracle import tgn_oracle as orc
    B = data["batch"]
    model, loader, opt = cpu_reference_state(data, prefill)
    ev = {k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")}
    done, t_timed = 0, 0.0
    for s in range(warmup + steps):
        lo = prefill + s * B
        sl = slice(lo, lo + B)
        t0 = time.perf_counter()
        orc.train_step(model, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], ev["msg"][sl],
                       ev["t"], ev["msg"], dropout=True)
        dt = time.perf_counter() - t0
        if s >= warmup:
            t_timed += dt
            done += 1
            if t_timed > budget_s:
                break
    return done, t_timed


def cpu_baseline_entry(name, B, prefill, steps, budget_s, warmup=5):
    torch.set_num_threads(os.cpu_count() or 1)
    cdata = load_workload(name, B, prefill, steps + warmup + 3)
    done, t_cpu = run_cpu_reference(cdata, steps, warmup, prefill, budget_s=budget_s)
    return {"value": done * B / t_cpu, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
            "ms_per_step": 1e3 * t_cpu / done,
            "sample": f"{done} training steps ({done * B} events) of the same workload (same shape, batch, prefill) on "
                      "the host CPU: oracle port of the reference's torch-CPU path incl. its Python-dict message "
                      "store; scatter_max vectorised (torch_scatter runs it in C++)"}


# --------------------------------------------------------------------------- parity beside the numbers
def parity_leg(name, B, prefill, steps, dev, precision=3, free_steps=None):
    """GPU engine vs the CPU oracle on the SAME first batches of a workload, from the same weights and the
    same prefilled rings, dropout off (the two dropout streams cannot agree).
      synced phase : after every step the oracle's post-Adam weights are copied into the engine, so each
                     step compares the same function on identical inputs (loss, last_update, ring);
      free phase   : `free_steps` more steps without syncing (Adam amplifies rounding-level gradient
                     noise on near-zero gradients into lr-sized weight differences: a looser number).
    Returns the error summary that bench.py prints as `parity` and the tests assert on."""
    from oracle import tgn_oracle as orc
    from tgn_b200.engine import TGNEngine
    free_steps = steps if free_steps is None else free_steps
    data = load_workload(name, B, prefill, steps + free_steps + 2)
    N, De, K = data["num_nodes"], data["raw_dim"], data["K"]
    torch.set_num_threads(os.cpu_count() or 1)
    ref, loader, opt = cpu_reference_state(data, prefill, dropout=False)
    eng = TGNEngine(N, De, HIDDEN, K, B, device=dev, lr=LR, dropout=0.0, use_graph=True,
                    log_capacity=data["src"].size, seed=5, precision=precision)
    eng.load_state(ref["memory"].state_dict(), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    ev = {k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")}
    eng.set_events(**ev)
    eng.prefill(prefill, (loader.neighbors, loader.e_id, loader.t))
    strip = lambda sd: {k: v for k, v in sd.items() if k not in ("memory", "last_update", "_assoc")}
    out = {"workload": workload_text(name, data, B, K, prefill), "steps_synced": steps, "steps_free": free_steps,
           "dropout": 0.0, "max_loss_rel_err": 0.0, "free_running_max_loss_rel_err": 0.0, "last_update_equal": True,
           "memory_max_abs_err": 0.0}
    for s in range(steps + free_steps):
        sl = slice(prefill + s * B, prefill + (s + 1) * B)
        loss = float(eng.train_step(from_device=True))
        loss_ref = orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl],
                                  ev["msg"][sl], ev["t"], ev["msg"], dropout=False)
        rel = abs(loss - loss_ref) / max(1e-12, abs(loss_ref))
        key = "max_loss_rel_err" if s < steps else "free_running_max_loss_rel_err"
        out[key] = max(out[key], rel)
        out["last_update_equal"] &= bool(torch.equal(eng.last_update.cpu(), ref["memory"].last_update))
        if s < steps:
            touched = torch.cat([ev["src"][sl], ev["dst"][sl]]).unique()
            err = (eng.memory[touched.to(dev)].cpu() - ref["memory"].memory.detach()[touched]).abs().max()
            out["memory_max_abs_err"] = max(out["memory_max_abs_err"], float(err))
            eng.load_state(strip(ref["memory"].state_dict()), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    out["ring_equal"] = bool(torch.equal(eng.e_id.cpu(), loader.e_id) and torch.equal(eng.t_ring.cpu(), loader.t))
    out["final_loss"] = {"gpu": loss, "cpu": loss_ref}
    out["memory_max_abs_err_final_free_running"] = float(
        (eng.memory.cpu() - ref["memory"].memory.detach()).abs().max())
    out["note"] = ("synced phase = the parity bar (same function, identical inputs).  Free-running: the TimeEncoder's "
                   "own init has |w| up to 1 on time deltas of 1e5..1e6 s, so an lr-sized (1e-4) difference in one "
                   "weight -- Adam turns a rounding-level sign flip of a near-zero gradient into exactly that -- "
                   "moves cos(w*dt) by O(1): individual memory rows decorrelate while the loss stays within 1e-2")
    return out


_JSON_FD = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# --------------------------------------------------------------------------- the GPU arms of one workload
def gpu_leg(name, B, K_steps, W, prefill, dev, rank, world, precision, clocks=None):
    """Device-resident arm (`value`) and end-to-end arm (`e2e`) of one workload on this rank's GPU.
    Returns (result dict, engine)."""
    import torch.distributed as dist
    from tgn_b200.engine import TGNEngine
    # batches per captured graph / host group (TGNEngine group_size): three, unless another small group size
    # divides K_steps and three does not (the driver's --steps 20 -> groups of four: no leftover single steps)
    G0 = next((g for g in (3, 4, 5, 2) if K_steps % g == 0), 3)
    warm_steps = max(W, 13 * G0 + 1)        # eager calls + the capture of the multi-step graph of every slot group
    warm_dev = warm_steps + 4 * 3 * G0
    # warm-up groups of the e2e arm: 3 eager calls + the capture per slot group; with a K % 3 leftover the slot
    # group the leftover lands on is additionally stepped through singly (4 more passes)
    Wg = max(14, (W + G0 - 1) // G0 + 2) + (13 if K_steps % G0 else 0)
    n_batches = warm_dev + K_steps + G0 * (Wg + K_steps // G0 + 2) + 16
    data = load_workload(name, B, prefill, n_batches, seed=rank)
    N, De, K = data["num_nodes"], data["raw_dim"], data["K"]
    eng = TGNEngine(N, De, HIDDEN, K, B, device=dev, lr=LR, dropout=0.1, use_graph=True,
                    log_capacity=data["src"].size, seed=1234 + rank, precision=precision, fused_zero_grad=True,
                    group_size=G0)
    eng.load_state(*init_state_dicts(De, HIDDEN, N, seed=1))
    ev = {k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")}
    eng.set_events(**ev)
    ring = ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
    eng.prefill(prefill, tuple(torch.from_numpy(a) for a in ring))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm
    eng.train_steps(warm_steps)   # warm-up: eager calls and the captures of the multi-step graphs of all slot groups
    for _ in range(4 * eng.nslots):                   # ... and of the single-step graph of every slot (the K % 3 leftover)
        eng.train_step(from_device=True)
    barrier()
    if clocks is not None:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps(K_steps)      # exactly K_steps batches: three per captured graph + the remainder one by one
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    loss_dev = float(eng.loss)

    # ---------------- end-to-end arm: host batches, H2D every step, loss read back every step
    while eng.cur % eng.group_size:                        # the device arm may have stopped inside a slot group
        eng.train_step(from_device=True, _capture=False)
    pos = eng.events_done
    # Host batches arrive in groups of three (the engine's pipelining unit): per group the three batches are
    # PACKED from the host event arrays into a pinned buffer (inside the timed region: that is the loader's
    # work), ONE pinned H2D copy (on a copy stream while the previous group is still executing), ONE captured
    # graph of three training steps, ONE D2H read of the three losses (read after the next group has been
    # launched, the last ones before the timer stops).  K_steps % 3 leftover steps run one by one.
    G = eng.group_size
    ng, rem = K_steps // G, K_steps % G
    n_groups = Wg + ng + 1
    ring_bufs = 4                                         # pinned staging buffers in rotation (a copy is long done
    host = torch.zeros((ring_bufs, eng.group_nbytes()), dtype=torch.uint8).pin_memory()   # when its buffer comes round)
    n_ev = ev["src"].numel()

    views = [eng.group_views(host[r]) for r in range(ring_bufs)]     # numpy views of the pinned buffers
    src_h, dst_h, neg_h, t_h, msg_h = (data[k] for k in ("src", "dst", "neg", "t", "msg"))

    def pack(g):
        """the host loader: the group's batches go from the host event arrays into the pinned staging buffer"""
        for i, (vs, vd, vn, vt, vm) in enumerate(views[g % ring_bufs]):
            lo = min(pos + (g * G + i) * B, n_ev - B)
            vs[:] = src_h[lo:lo + B]
            vd[:] = dst_h[lo:lo + B]
            vn[:] = neg_h[lo:lo + B]
            vt[:] = t_h[lo:lo + B]
            if De:
                vm[:] = msg_h[lo:lo + B]
        return host[g % ring_bufs]

    eng.stage_group(pack(0), ahead=False)
    # the K_steps % 3 leftover steps of the timed region run one by one on the slots that follow the last whole
    # group: whenever the warm-up passes that slot group it steps through it singly, so their graphs exist too
    c_final = (eng.cur + G * (Wg + ng)) % eng.nslots
    visits = 0
    for g in range(Wg):
        eng.stage_group(pack(g + 1))
        if rem and eng.cur == c_final:
            visits += 1
        if rem and eng.cur == c_final and visits > 4:      # (the first four passes capture this group's own graph)
            for _ in range(G):
                eng.train_step_logged(from_device=False, lookahead=True)
        else:
            eng.train_group_logged()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for g in range(Wg, Wg + ng):
        eng.stage_group(pack(g + 1))
        eng.train_group_logged()
    for _ in range(rem):
        eng.train_step_logged(from_device=False, lookahead=True)
    losses = eng.flush_group_losses()
    loss_host = eng.flush_loss() if rem else losses[-1]
    f1.record()
    barrier()
    if clocks is not None:
        clocks.stop_flag = True
    ms_e2e = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(ms), float(ms_e2e)
    res = {"workload": workload_text(name, data, B, K, prefill),
           "value": world * K_steps * B / (ms / 1e3), "unit": "events/s", "ms_per_step": ms / K_steps,
           "e2e": {"value": world * K_steps * B / (ms_e2e / 1e3), "unit": "events/s",
                   "h2d_bytes_per_step": eng.group_nbytes() // eng.group_size, "d2h_bytes_per_step": 4,
                   "grouping": f"{G} batches per host pack / H2D copy / graph launch / loss read-back; the pack "
                               "(host event arrays -> pinned buffer) is inside the timed region; negatives are "
                               "part of the host event stream (drawn once per epoch)",
                   "ms_per_step": ms_e2e / K_steps},
           "final_loss": {"device_arm": loss_dev, "e2e_arm": loss_host}}
    return res, eng


def bench_module_path(dev, n_events=60_000, B=200):
    """What the UNCHANGED driver script gets: events/s of `epoch_utils.train` (the call at pyg-mem-tgn.py:57) on
    the drop-in modules.  Two numbers on the same data:
      script_path : train() as the script calls it -- it recognises the standard model (model_utils.getModel,
                    Adam, BCEWithLogits) and runs the epoch on TGNEngine attached to the modules' own state;
      module_path : use_engine=False -- one Python call per module per batch, autograd, torch.optim.Adam.
    Synthetic wiki shape, wall clock around the second epoch (host batches, per-batch negative sampling,
    weight / optimizer hand-over and the final synchronisation all inside)."""
    import utils
    from epoch_utils import train as run_train
    from model_utils import getModel, getOptimizer
    from neg_sampler import NegLinkSamplerDest
    from neighbor_loader import LastNeighborLoader
    train_param = {"batch_size": B, "lr": LR, "epoch": 1}
    data, tr, va, te, ns, evaluator, metric = utils.getDataWithDependecyBlock(f"tgbl-wiki@{n_events}", train_param)
    neg_dest_sampler = NegLinkSamplerDest(torch.unique(data.dst))
    assoc = torch.empty(data.num_nodes, dtype=torch.long, device=dev)
    n_ev = len(tr.dataset)
    out = {"workload": f"synthetic tgbl-wiki shape, first {n_events} events (70% train = {n_ev}), batch {B}, "
                       "wall clock of one epoch_utils.train call", "unit": "events/s"}
    for label, use_engine in (("script_path", None), ("module_path", False)):
        loader = LastNeighborLoader(data.num_nodes, size=10, device=dev)
        model = getModel(data.msg.shape[1], HIDDEN, data.num_nodes, dev, gnn_param={"dim_out": HIDDEN})
        opt = getOptimizer(model, LR)
        crit = torch.nn.BCEWithLogitsLoss()
        for _ in range(2 if use_engine is None else 1):      # warm-up epochs (graph captures on the engine path)
            run_train(model, data.msg, tr, loader, neg_dest_sampler, assoc, dev, opt, crit, use_engine=use_engine)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = run_train(model, data.msg, tr, loader, neg_dest_sampler, assoc, dev, opt, crit, use_engine=use_engine)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[label] = {"value": n_ev / dt, "ms_per_step": 1e3 * dt / max(1, (n_ev + B - 1) // B), "loss_sum": float(loss)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=["tgbl-review", "tgbl-wiki"],
                    help="headline workload (BASELINE.json configs[1] = review; configs[0] = wiki)")
    ap.add_argument("--batch", type=int, default=None, help="batch size (default: 200; config/TGN.yml:27 says 2000)")
    ap.add_argument("--prefill", type=int, default=None)
    ap.add_argument("--cpu-steps", type=int, default=400, help="bounded sample of the CPU baseline leg")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-eval-dp", action="store_true")
    ap.add_argument("--only-eval-dp", action="store_true", help="run just the data-parallel evaluation leg (development)")
    ap.add_argument("--eval-replicated", action="store_true", help="eval leg: every rank embeds every root (round-1 design)")
    ap.add_argument("--no-partitioned", action="store_true")
    ap.add_argument("--only-partitioned", action="store_true", help="run just the partitioned-memory leg (development)")
    ap.add_argument("--no-wiki", action="store_true", help="skip the wiki-shape legs (configs[0], batch 200 and 2000)")
    ap.add_argument("--no-module-path", action="store_true")
    ap.add_argument("--no-tcsr", action="store_true", help="skip the configs[3] leg (t-CSR uniform-20, 2 layers)")
    ap.add_argument("--only-tcsr", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-tf32", action="store_true", help="skip the single-pass tf32 variant of the headline workload")
    ap.add_argument("--part-compute", default="owner", choices=["owner", "replicated"],
                    help="partitioned-memory leg: owner-side compute of the memory path, or the replicated round-1 design")
    ap.add_argument("--part-exchange", default="p2p", choices=["p2p", "allreduce"],
                    help="row assembly of the partitioned-memory leg: peer reads over symmetric memory, or NCCL all-reduce")
    ap.add_argument("--eval-batches", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=int, default=3, choices=[1, 3],
                    help="tensor-core GEMM mode: 3 = 3xTF32 (fp32-level accuracy), 1 = single-pass tf32")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 (NCCL prints its version
    # banner there) are sent to stderr, the JSON line goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K_steps, W = args.steps, max(args.warmup, 3)

    from tgn_b200 import synth
    name = args.workload
    cfg = synth.SHAPES[name]
    B, K = args.batch or cfg["B"], cfg["K"]
    prefill = args.prefill if args.prefill is not None else PREFILL[name]
    probe = dict(num_nodes=cfg["N"], raw_dim=cfg["De"])
    config = {"workload": workload_text(name, probe, B, K, prefill),
              "l2": "inputs differ every step (new batch, new ring/memory rows); weights (~1.2 MB) stay L2-resident by design",
              "parallelism": "one process per device; the step does not shard across batches" if world == 1
              else f"{world} independent replicas (step does not shard across batches)"}

    if args.impl == "reference":
        if rank != 0:
            return
        data = load_workload(name, B, prefill, W + K_steps + 2)
        torch.set_num_threads(os.cpu_count() or 1)
        done, t_timed = run_cpu_reference(data, K_steps, W, prefill, budget_s=150.0)
        val = done * B / t_timed
        line = {"impl": "reference", "metric": "train events/sec (TGN step)", "value": val, "unit": "events/s",
                "n_gpus": args.gpus, "steps": done, "warmup": W, "ms_per_step": 1e3 * t_timed / done,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "events/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{done} of {K_steps} requested steps (150 s budget), oracle port of the reference's torch-CPU path"},
                "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if args.only_tcsr:
        r = bench_tcsr_two_layer(dev, rank, world)
        if rank == 0:
            emit(r)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)
        return
    if args.only_eval_dp:
        r = bench_eval_dp(dev, rank, world, args.eval_batches, args.precision, shard=not args.eval_replicated)
        if rank == 0:
            emit(r)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)
        return
    if args.only_partitioned:
        part = bench_partitioned(dev, rank, world, max(60, min(K_steps, 300)), W, args.precision, args.part_exchange,
                                 args.part_compute)
        if rank == 0:
            emit(part)
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)
        return
    clocks = ClockSampler(local_rank)
    head, eng = gpu_leg(name, B, K_steps, W, prefill, dev, rank, world, args.precision, clocks)

    # ---------------- launches per step (counted once, outside the timed region)
    from torch.profiler import ProfilerActivity, profile
    eng_use_graph = eng.use_graph
    eng.use_graph = False
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.train_step(from_device=True)
        torch.cuda.synchronize()
    kern = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()
            and "memset" not in e.name.lower() and "memcpy" not in e.name.lower()]
    ours = [e for e in kern if "tgn::" in e.name]
    per_kernel = {}
    for e in kern:
        key = e.name.split("(")[0][-60:]
        per_kernel.setdefault(key, [0, 0.0])
        per_kernel[key][0] += 1
        per_kernel[key][1] += e.device_time
    top = sorted(per_kernel.items(), key=lambda kv: -kv[1][1])[:6]

    # ---------------- roofline of the dominant kernel, timed inside (eager) steps
    roof = dominant_kernel_roofline(eng, dev)
    eng.use_graph = eng_use_graph
    del eng
    torch.cuda.empty_cache()

    def guarded(fn, *a, **kw):
        """Secondary legs: on one GPU a failure is reported inside the JSON line instead of costing the
        headline; with several ranks an exception must propagate (the others are inside collectives)."""
        if world > 1:
            return fn(*a, **kw)
        try:
            return fn(*a, **kw)
        except Exception as e:   # noqa: BLE001
            return {"error": repr(e)[:300]}

    part = None
    if not args.no_partitioned:
        part = guarded(bench_partitioned, dev, rank, world, max(60, min(K_steps, 300)), W, args.precision,
                       args.part_exchange, args.part_compute)
    eval_dp = None
    if not args.no_eval_dp:
        eval_dp = guarded(bench_eval_dp, dev, rank, world, args.eval_batches, args.precision, not args.eval_replicated)

    tcsr_leg = None
    if not args.no_tcsr:
        tcsr_leg = guarded(bench_tcsr_two_layer, dev, rank, world)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {"metric": "train events/sec (TGN step)", "value": head["value"], "unit": "events/s",
                "n_gpus": world, "steps": K_steps, "warmup": W, "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32x3 (fp32-accurate split, fp32 accumulate)" if args.precision == 3 else "tf32",
                "data": "synthetic", "config": config, "e2e": head["e2e"],
                "gpu_launches": len(ours) * K_steps,
                "launches_per_step": {"tgn_kernels": len(ours), "all_kernels": len(kern)},
                "top_kernels_eager_us": {k: [v[0], round(v[1], 1)] for k, v in top},
                "clocks": clocks.summary(), "final_loss": head["final_loss"],
                "roofline": roof_with_peak(roof, peaks)}
        if part is not None:
            line["partitioned_memory"] = part
        if eval_dp is not None:
            line["eval_dp"] = eval_dp
        if tcsr_leg is not None:
            line["tcsr_two_layer"] = tcsr_leg
        if world == 1 and not args.no_kernel_rooflines:
            line["kernel_rooflines"] = guarded(kernel_rooflines, dev, peaks)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(name, B, prefill, args.cpu_steps, budget_s=20.0)
        if world == 1 and not args.no_parity:
            # the GPU arm and the CPU arm on the same first 20 batches of the headline workload, dropout off
            line["parity"] = guarded(parity_leg, name, B, prefill, 20, dev, args.precision)
        if world == 1 and not args.no_wiki and name != "tgbl-wiki":
            # BASELINE.json configs[0] (the shape the >=50x target is quoted on), at BASELINE's batch 200 and at
            # the batch config/TGN.yml:27 sets (2000); each with its own CPU arm and parity beside it
            wiki = {}
            for wb, wsteps, wcpu in ((200, max(60, min(K_steps, 300)), 120), (2000, max(30, min(K_steps, 90)), 20)):
                def one(wb=wb, wsteps=wsteps, wcpu=wcpu):
                    r, e = gpu_leg("tgbl-wiki", wb, wsteps, W, PREFILL["tgbl-wiki"], dev, 0, 1, args.precision)
                    del e
                    torch.cuda.empty_cache()
                    r["steps"] = wsteps
                    if not args.no_cpu_baseline:
                        r["cpu_baseline"] = cpu_baseline_entry("tgbl-wiki", wb, PREFILL["tgbl-wiki"], wcpu, budget_s=8.0)
                        r["vs_cpu"] = {"device": r["value"] / r["cpu_baseline"]["value"],
                                       "e2e": r["e2e"]["value"] / r["cpu_baseline"]["value"]}
                    if not args.no_parity:
                        r["parity"] = parity_leg("tgbl-wiki", wb, PREFILL["tgbl-wiki"], 20 if wb <= 200 else 6, dev,
                                                 args.precision)
                    return r
                wiki[f"batch_{wb}"] = guarded(one)
            line["wiki"] = wiki
        if world == 1 and not args.no_module_path:
            line["script_path"] = guarded(bench_module_path, dev)
        if world == 1 and not args.no_tf32 and args.precision == 3:
            # north_star allows the GRU / projections in tf32 (2e-2 bar): the same workload with single-pass tf32
            # tensor-core GEMMs, its own parity numbers beside it (the headline stays the fp32-accurate 3xTF32 mode)
            def tf32_leg():
                r, e = gpu_leg(name, B, max(60, min(K_steps, 300)), W, prefill, dev, 0, 1, 1)
                del e
                torch.cuda.empty_cache()
                if not args.no_parity:
                    r["parity"] = parity_leg(name, B, prefill, 10, dev, 1, free_steps=0)
                r["dtype"] = "tf32 (single pass), fp32 accumulate"
                return r
            line["tf32"] = guarded(tf32_leg)
        emit(line)
    if world > 1:
        # captured graphs hold NCCL kernels: leave without the collective shutdown (it can block)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def bench_partitioned(dev, rank, world, steps, warmup, precision, exchange="p2p", compute="owner"):
    """ONE training job with the node memory partitioned by owner over the ranks (BASELINE.json
    configs[2]): synthetic tgbl-coin shape (638,486 nodes), batch 600, K=10.  Node n lives on rank
    n % world; the shards are symmetric memory mapped into every rank, so the rows a step needs are read
    straight out of their owners' HBM over NVLink/NVSwitch (tgn_part_gather_p2p between two device-side
    rank barriers; csrc/partition.cu), gradients are all-reduced, everything else is replicated.
    value = events/s of that one job (strong scaling: the batch does not grow with the ranks)."""
    import torch.distributed as dist
    from tgn_b200 import synth
    from tgn_b200.engine import TGNEngine
    name, prefill = "tgbl-coin", 600_000
    cfg = synth.SHAPES[name]
    B, K = cfg["B"], cfg["K"]
    data = synth.synth_events(name, seed=0, max_events=prefill + (max(warmup, 52) + 4 * 9 + steps + 12) * B)
    N, De = data["num_nodes"], data["raw_dim"]
    eng = TGNEngine(N, De, HIDDEN, K, B, device=dev, lr=LR, dropout=0.1, use_graph=True,
                    log_capacity=data["src"].size, seed=99, precision=precision, rank=rank, world=world,
                    fused_zero_grad=True, part_exchange=exchange, part_compute=compute)
    eng.load_state(*init_state_dicts(De, HIDDEN, N, seed=1))
    eng.set_events(**{k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")})
    ring = ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
    eng.prefill(prefill, tuple(torch.from_numpy(a) for a in ring))
    eng.train_steps(max(warmup, 52))          # eager warm-up calls + the captures of the three-step graphs
    for _ in range(4 * eng.nslots):             # ... and of the single-step graphs (steps % 3 leftovers)
        eng.train_step(from_device=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps(steps)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    rows = int(eng.w.Nb_dev.item())
    owned = int(eng.w.So_dev.item()) if eng.owner_compute else rows
    eng.check_device_errors()
    D = HIDDEN
    # bytes this rank moves over NVLink per step (owner compute): rows published to the P-1 peers, remote rows
    # read for the messages (other endpoints, (P-1)/P of them remote), gradients read by the optimiser
    nvl = (owned * (4 * D + 8) * (world - 1) + owned * 4 * D * (world - 1) // max(world, 1)
           + 4 * eng.n_param * (world - 1 + (1 if rank else 0))) if eng.owner_compute else None
    return {"metric": "train events/sec (TGN step), node memory partitioned by owner", "value": steps * B / (ms / 1e3),
            "compute": ("owner-side: message build + GRU forward/backward only on owned rows" if eng.owner_compute
                        else ("replicated on every rank" if world > 1 else "single GPU")),
            "rows_owned_last_step": owned, "nvlink_bytes_per_step_per_rank": nvl,
            "unit": "events/s", "n_gpus": world, "ms_per_step": ms / steps, "steps": steps, "scaling": "strong",
            "final_loss": float(eng.loss), "rows_per_step": rows,
            "memory_rows_per_gpu": eng.Nloc, "memory_bytes_per_gpu": eng.Nloc * HIDDEN * 4,
            "workload": f"synthetic {name} shape: {N} nodes, raw_dim {De}, batch {B}, {K} recent nbrs, dim {HIDDEN}",
            "exchange": eng.part_exchange,
            "collectives_per_step": ("none (single GPU)" if world == 1 else
                                     "no NCCL call: remote memory rows read and result rows published over the peer "
                                     "mapping, partial gradients summed inside the optimiser kernel out of peer memory; "
                                     "three signal-pad rank barriers" if eng.owner_compute else
                                     "peer-memory row gather between two signal-pad rank barriers, NCCL all-reduce of "
                                     "the flat gradient")}


def bench_eval_dp(dev, rank, world, n_batches, precision, shard=True):
    """TGB evaluation with the negatives scored data-parallel (BASELINE.json configs[4]): synthetic
    tgbl-flight shape (18,143 nodes, D_e=16), 200 positives x 999 negatives per batch, negative
    columns sharded round-robin over the ranks, one all-reduce of the 2*B rank counts per batch
    (tgn_b200/dist_eval.py).  Model replicas are identical on every rank (same seeds), so the MRR
    printed here is the same number for every --gpus N."""
    import torch.distributed as dist
    from tgn_b200 import dist_eval, synth
    from tgn_b200.engine import TGNEngine
    cfg = synth.SHAPES["tgbl-flight"]
    B, K, Q, prefill = cfg["B"], cfg["K"], 999, 300_000
    warm = 6            # 3 eager calls, the graph capture, two replays: all before the timed region
    data = synth.synth_events("tgbl-flight", seed=0, max_events=prefill + (n_batches + warm + 1) * B)
    N, De = data["num_nodes"], data["raw_dim"]
    eng = TGNEngine(N, De, HIDDEN, K, B, device=dev, lr=LR, dropout=0.1, use_graph=True,
                    log_capacity=data["src"].size, seed=7, precision=precision)
    eng.load_state(*init_state_dicts(De, HIDDEN, N, seed=1))
    ev = {k: torch.from_numpy(data[k]) for k in ("src", "dst", "t", "msg", "neg")}
    eng.set_events(**ev)
    ring = ring_after(data["src"][:prefill], data["dst"][:prefill], data["t"][:prefill], K, N)
    eng.prefill(prefill, tuple(torch.from_numpy(a) for a in ring))
    batches = []
    for b in range(n_batches + warm):
        sl = slice(prefill + b * B, prefill + (b + 1) * B)
        neg = torch.from_numpy(synth.eval_negatives(data["src"][sl], data["dst"][sl], N, Q, seed=1000 + b))
        batches.append(tuple(x.to(dev) for x in (ev["src"][sl], ev["dst"][sl], neg, ev["t"][sl], ev["msg"][sl])))
    dist_eval.evaluate_dp(eng, batches[:warm], rank, world, shard_embeddings=shard)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mrr = dist_eval.evaluate_dp(eng, batches[warm:], rank, world, shard_embeddings=shard)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    return {"metric": "TGB eval scored candidate edges/sec (999 negatives/positive, negatives data-parallel)",
            "value": n_batches * B * (1 + Q) / (ms / 1e3), "unit": "edges/s", "n_gpus": world,
            "ms_per_batch": ms / n_batches, "batches": n_batches, "mrr": mrr, "scaling": "strong",
            "workload": f"synthetic tgbl-flight shape: {N} nodes, raw_dim {De}, batch {B}, {Q} negatives, "
                        f"ring prefilled with {prefill} events",
            "embedding": ("sharded: every root embedded on one rank, decoder-projected rows all-gathered" if shard
                          else "replicated: every rank embeds every root"),
            "collective": "none" if world == 1 else (
                ("no NCCL call in the step: every rank writes its projected rows [2, roots/P, 100] fp32 and its 2*B int32 "
                 "rank counts into every rank's symmetric-memory tables (peer stores over NVLink), two device-side rank "
                 "barriers; the counts are summed locally in rank order"
                 if shard and os.environ.get("TGN_EVAL_EXCHANGE", "peer") == "peer" else
                 ("one all-gather of the projected rows [2, roots/P, 100] fp32 + " if shard else "") +
                 "one all-reduce(sum) of 2*B int32 rank counts per batch")),
            "nvlink_bytes_per_batch_per_rank": (None if world == 1 or not shard else
                                                 (world - 1) * (8 * HIDDEN * ((N + world - 1) // world) + 8 * B))}


def bench_tcsr_two_layer(dev, rank, world, steps=40, events=4_000_000):
    """BASELINE.json configs[3]: TGN on the synthetic tgbl-comment shape (994,790 nodes, raw_dim 2), UNIFORM-20
    sampling over the t-CSR graph, TWO attention layers (tgn_b200.tcsr_trainer.TCSRTrainer; t-CSR built on the
    device by tgn_tcsr_build from the first `events` events of the stream).  N > 1: independent replicas.
    Also the statistical check of the uniform draws on the live graph: the normalised position of a draw inside
    its root's candidate range must average 1/2."""
    import torch.distributed as dist
    from tgn_b200 import ops, synth
    from tgn_b200.tcsr_trainer import TCSRTrainer
    cfg = synth.SHAPES["tgbl-comment"]
    B, K = cfg["B"], cfg["K"]
    data = synth.synth_events("tgbl-comment", seed=rank, max_events=events)
    N, De, E = data["num_nodes"], data["raw_dim"], data["src"].size
    s_d, d_d, t_d = (torch.from_numpy(data[k]).to(dev) for k in ("src", "dst", "t"))
    t0 = time.perf_counter()
    indptr, indices, eid, ts = ops.tcsr_build(s_d, d_d, t_d, N, t_sorted=True)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    tr = TCSRTrainer(indptr, indices, eid, ts, N, De, HIDDEN, [K, K], False, torch.from_numpy(data["msg"]), device=dev,
                     lr=LR, dropout=0.1, seed=7 + rank)
    tr.train()
    ev = {k: torch.from_numpy(data[k]).to(dev) for k in ("src", "dst", "neg", "t")}
    msg = torch.from_numpy(data["msg"]).to(dev)
    lo0 = E - (steps + 8) * B                     # the tail of the stream: roots with long histories
    def step(i):
        sl = slice(lo0 + i * B, lo0 + (i + 1) * B)
        return tr.train_step(ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], msg[sl])
    for i in range(8):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tr.sampled_edges = 0
    t0 = time.perf_counter()
    for i in range(8, 8 + steps):
        loss = step(i)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt)
    # uniform-draw statistics on the live graph: hub roots, draws inside [row start, first entry >= t)
    deg = (indptr[1:] - indptr[:-1]).long()
    hubs = torch.topk(deg, 2000).indices.to(torch.int32)
    rts = torch.full((hubs.numel(),), float(t_d[-1]) + 1.0, device=dev)
    (nbr, col, se, sts, dts), off, cnt = ops.tcsr_sample(indptr, indices, eid, ts, hubs, rts, K, ops.SAMPLE_UNIFORM, seed=99,
                                                         coarse=tr.sampler.coarse)
    n = int(cnt.item())
    rows = hubs[col[:n].long()].long()
    lo = indptr[rows].long()
    mean_pos = float(((sts[:n] - ts[lo]) / (ts[lo + deg[rows] - 1] - ts[lo]).clamp(min=1.0)).mean())
    return {"metric": "train events/sec, t-CSR uniform-20 sampling, 2 attention layers (module path)",
            "value": world * steps * B / dt, "unit": "events/s", "n_gpus": world, "ms_per_step": 1e3 * dt / steps,
            "steps": steps, "final_loss": float(loss), "sampled_nbrs_per_s": world * tr.sampled_edges / dt,
            "sampled_edges_per_step": tr.sampled_edges / steps,
            "workload": f"synthetic tgbl-comment shape: {N} nodes, raw_dim {De}, first {E} events ({2 * E} t-CSR entries), "
                        f"batch {B}, uniform-{K} x 2 layers, dim {HIDDEN}",
            "tcsr_build_s": build_s, "parallelism": "independent replicas" if world > 1 else "single GPU",
            "uniform_check": {"hub_roots": int(hubs.numel()), "draws": n,
                              "mean_normalised_time_of_draw": mean_pos,
                              "expected": "~0.5 (event times are uniform over the span, draws uniform over the row)"}}


def dominant_kernel_roofline(eng, dev):
    """The kernel with the largest share of the step is tgn::tgemm_kernel (TMA + tcgen05 GEMM,
    ~45% over its 5 launches); its heaviest forward launch is the GRU gate GEMM pair
    gi[S,3D] = x[S,Dx] W_ih^T + b, gh[S,3D] = h[S,D] W_hh^T + b (one launch).  It is timed on the
    step's OWN operands (the buffers the last training step left behind, L2-warm as in the step):
    a CUDA graph of 20 back-to-back launches, CUDA events around the replay, on the launching
    stream.  Algorithmic work per launch: 2*S*(Dx+D)*3D flops;
    bytes S*(Dx+D)*4 + 3D*(Dx+D)*4 + 2*S*3D*4."""
    from tgn_b200 import ops
    w, p, D = eng.w, eng.p, eng.D
    S = int(w.Nb_dev.item())

    fused = getattr(eng, "fused_gru", False)

    def launch():
        if fused:
            eng._memory_gru(w, w.Nb, w.Nb_dev)       # tgn_gru_fused_fwd: both gate GEMMs + gate math, one launch
            return
        ops.gemm_batch([
            ops.gemm_desc(w.x, eng.flat, w.gi, m=w.Nb, n=3 * D, k=eng.Dx, lda=eng.ldx, ldb=eng.ldx, ldc=3 * D,
                          b_off=eng.off["memory_updater.weight_ih"], bias=p["memory_updater.bias_ih"], m_dev=w.Nb_dev),
            ops.gemm_desc(w.h, eng.flat, w.gh, m=w.Nb, n=3 * D, k=D, lda=D, ldb=D, ldc=3 * D,
                          b_off=eng.off["memory_updater.weight_hh"], bias=p["memory_updater.bias_hh"], m_dev=w.Nb_dev),
        ], eng.prec)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            launch()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3 / reps)
    t = float(np.median(ts))
    Dx = eng.Dx
    flops = 2.0 * S * (Dx + D) * 3 * D
    if fused:   # x, h read (h twice: operand + blend), weights, h' and the 4 saved gate planes written
        nbytes = 4.0 * (S * (Dx + 2 * D) + 3 * D * (Dx + D) + S * 5 * D)
        name = ("tgn::gru_fused_kernel (GRUCell forward: gi = x W_ih^T, gh = h W_hh^T in TMEM, gate math in the "
                "epilogue, one launch)")
    else:
        nbytes = 4.0 * (S * (Dx + D) + 3 * D * (Dx + D) + 2 * S * 3 * D)
        name = "tgn::tgemm_kernel (GRU gate GEMMs gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh, one launch)"
    return {"kernel": name, "rows": S, "seconds": t, "flops": flops, "bytes": nbytes, "launches_timed": reps * 5,
            "prec": eng.prec, "fused": fused}


def ncu_traffic(key):
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic.json")))[key]
        return {"traffic": d["dram_bytes_per_launch"], "traffic_source": d["source"]}
    except Exception:
        return {"traffic": None, "traffic_source": "no ncu --set full digest committed for this kernel"}


def roof_with_peak(r, peaks):
    peak = peaks.get("bf16_tflops", 1590.0)
    ach = r["flops"] / r["seconds"] / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE such launch, from the committed ncu --set full
            # digest (profiles/ncu_traffic.json names the capture it was read from); null when there is none
            **ncu_traffic("gru_fused_instep" if r.get("fused") else "tgemm_gru_pair_instep"),
            "kernel": r["kernel"], "rows_per_launch": r["rows"], "us_per_launch": r["seconds"] * 1e6,
            "launches_timed": r["launches_timed"],
            "algorithmic_flops_per_launch": r["flops"], "algorithmic_bytes_per_launch": r["bytes"],
            "achieved_gbs": r["bytes"] / r["seconds"] / 1e9,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, cuBLAS bf16)" if "bf16_tflops" in peaks
                           else "fallback 1590 (B200_PROFILING.md)",
            "note": "algorithmic flops (one product per multiply-add); the kernel computes in tf32 (dense tf32 peak is "
                    "half the bf16 figure used as denominator)" + (", and issues 3 tensor-core products per algorithmic "
                    "one (3xTF32 split for fp32-level accuracy)" if r["prec"] == 3 else "") +
                    "; at ~5k rows per launch the launch is latency-bound, see kernel_rooflines for the same kernels "
                    "on large launches"}


def _time_launch(fn, reps=8, flush=None):
    ts = []
    for i in range(reps + 2):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def kernel_rooflines(dev, peaks):
    """Large-launch roofline points of the hot-path kernels (the per-step launches at B=200 are
    latency-bound): algorithmic bytes / CUDA-event time, L2 flushed between launches."""
    from tgn_b200 import ops
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf = peaks.get("bf16_tflops", 1590.0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = []
    g = torch.Generator(device="cpu").manual_seed(0)
    # ---- t-CSR most-recent-k sampler: 2M roots over a 40M-entry graph (deg ~ 113), k = 10
    N, deg = 352_637, 113
    indptr = (torch.arange(N + 1, dtype=torch.int64) * deg).to(torch.int32).to(dev)
    nnz = N * deg
    ts = torch.arange(deg, dtype=torch.float32).repeat(N).to(dev)
    indices = torch.randint(0, N, (nnz,), generator=g, dtype=torch.int32).to(dev)
    eid = torch.arange(nnz, dtype=torch.int32, device=dev)
    R, k = 2_000_000, 10
    roots = torch.randint(0, N, (R,), generator=g, dtype=torch.int32).to(dev)
    rts = (torch.rand(R, generator=g) * deg + 12).to(dev)
    coarse = ops.tcsr_build_index(ts)          # skip index, built once per graph (sampler_core.ParallelSampler.__init__)
    t = _time_launch(lambda: ops.tcsr_sample(indptr, indices, eid, ts, roots, rts, k, coarse=coarse), flush=flush)
    nbytes = R * (8 + 4 * 7 + k * 12 + k * 20 + 8)   # indptr pair, ~log2(deg) probes, k entries in, k rows out, root
    out.append({"kernel": "tgn::tcsr_sample_kernel (recent-10, 2M roots, deg 113)", "bound": "hbm",
                "achieved": nbytes / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / t / 1e9 / hbm,
                "sampled_nbrs_per_s": R * k / t, "us": t * 1e6})
    del indices, eid, ts, indptr, coarse
    # ---- the same sampler on the workload BASELINE configs[4] names: one TGB evaluation batch of the flight
    # shape (200 positives x (2 + 999) candidate roots at the batch's time over 18,143 nodes).  The graph is
    # built on the device by tgn_tcsr_build from a 10M-event synthetic stream (also timed: entries/s).
    from tgn_b200 import synth
    fl = synth.synth_events("tgbl-flight", seed=0, max_events=10_000_000)
    Nf, Ef = fl["num_nodes"], fl["src"].size
    s_d, d_d, t_d = (torch.from_numpy(fl[k]).to(dev) for k in ("src", "dst", "t"))
    tb = _time_launch(lambda: ops.tcsr_build(s_d, d_d, t_d, Nf, t_sorted=True), reps=3)
    indptr, indices, eid, ts = ops.tcsr_build(s_d, d_d, t_d, Nf, t_sorted=True)
    out.append({"kernel": "tgn::tcsr_build (flight shape, 10M events -> 20M entries, chronological stream)",
                "bound": "hbm", "achieved": 2 * Ef * 160 / tb / 1e9, "peak": hbm, "unit": "GB/s",
                "frac": 2 * Ef * 160 / tb / 1e9 / hbm, "entries_per_s": 2 * Ef / tb, "us": tb * 1e6,
                "note": "160 B per entry: key generation 24 (12 read, 12 written) + 3 radix passes x 32 (8 histogram "
                        "read, 12 read, 12 written) + emit 40 (12 read, 16 gathered, 12 written)"})
    coarse = ops.tcsr_build_index(ts)
    Bq, Q = 200, 999
    e0 = Ef - Bq
    negs = torch.from_numpy(synth.eval_negatives(fl["src"][e0:], fl["dst"][e0:], Nf, Q)).to(dev)
    roots = torch.cat([s_d[e0:], d_d[e0:], negs.reshape(-1)]).to(torch.int32)
    rts = torch.cat([t_d[e0:], t_d[e0:], t_d[e0:].repeat_interleave(Q)]).to(torch.float32)
    Rf = roots.numel()
    t = _time_launch(lambda: ops.tcsr_sample(indptr, indices, eid, ts, roots, rts, k, coarse=coarse), flush=flush)
    deg_f = 2 * Ef / Nf
    nbytes = Rf * (8 + 4 * int(np.ceil(np.log2(deg_f))) + k * 12 + k * 20 + 8)
    out.append({"kernel": f"tgn::tcsr_sample_kernel (recent-10, one flight-shape eval batch: {Rf} roots, mean deg {deg_f:.0f})",
                "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / t / 1e9 / hbm,
                "sampled_nbrs_per_s": Rf * k / t, "us": t * 1e6,
                "note": "roots of one batch repeat nodes and share the batch time, so most row reads hit L2: "
                        "algorithmic bytes per root exceed the DRAM bytes actually moved"})
    del indices, eid, ts, indptr, coarse, s_d, d_d, t_d, negs, roots, rts
    # ---- neighbour-ring lookup: 1M roots, K = 10
    K = 10
    nb = torch.randint(0, N, (N, K), generator=g).to(dev)
    ei = torch.randint(0, 1 << 30, (N, K), generator=g).to(dev)
    tt = torch.rand(N, K, generator=g).to(dev)
    R2 = 1_000_000
    r2 = torch.randint(0, N, (R2,), generator=g).to(dev)
    t = _time_launch(lambda: ops.nbr_lookup_raw(r2, nb, ei, tt, None), flush=flush)
    nbytes = R2 * (8 + K * 20 + K * 28)
    out.append({"kernel": "tgn::nbr_lookup_kernel (1M roots, K=10)", "bound": "hbm", "achieved": nbytes / t / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": nbytes / t / 1e9 / hbm, "sampled_nbrs_per_s": R2 * K / t,
                "us": t * 1e6})
    del nb, ei, tt
    # ---- aggregators on materialised messages: M = 1M messages of width 472 into S = 250k segments
    M, S, W = 1_000_000, 250_000, 472
    msg = torch.randn(M, W, device=dev)
    idx = torch.randint(0, S, (M,), generator=g).to(dev)
    tm = torch.randint(0, 1000, (M,), generator=g).to(dev)
    t = _time_launch(lambda: ops.agg_last(msg, idx, tm, S), flush=flush)
    nbytes = M * 16 + S * 8 + S * 2 * W * 4
    out.append({"kernel": "tgn::agg_last (1M msgs x 472 -> 250k segments)", "bound": "hbm", "achieved": nbytes / t / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": nbytes / t / 1e9 / hbm, "us": t * 1e6})
    t = _time_launch(lambda: ops.agg_mean(msg, idx, S), flush=flush)
    nbytes = M * W * 4 + M * 8 + S * W * 4
    out.append({"kernel": "tgn::agg_mean (1M msgs x 472 -> 250k segments)", "bound": "hbm", "achieved": nbytes / t / 1e9,
                "peak": hbm, "unit": "GB/s", "frac": nbytes / t / 1e9 / hbm, "us": t * 1e6})
    del msg, idx, tm
    # ---- the GEMM kernel on a large launch: GRU gate GEMM for 262,144 rows
    Sg, Dx, D = 262_144, 472, 100
    x = torch.randn(Sg, Dx, device=dev); wih = torch.randn(3 * D, Dx, device=dev); o = torch.empty(Sg, 3 * D, device=dev)
    for prec in (3, 1):
        t = _time_launch(lambda: ops.sgemm(x, wih, m=Sg, n=3 * D, k=Dx, lda=Dx, ldb=Dx, out=o, prec=prec), flush=flush)
        fl = 2.0 * Sg * Dx * 3 * D
        out.append({"kernel": f"tgn::tgemm_kernel (x W_ih^T, 262144 rows, {'3xTF32' if prec == 3 else 'tf32'})",
                    "bound": "tensor", "achieved": fl / t / 1e12, "peak": tf, "unit": "TFLOP/s",
                    "frac": fl / t / 1e12 / tf, "achieved_gbs": 4.0 * (Sg * Dx + Sg * 3 * D) / t / 1e9, "us": t * 1e6})
    return out


if __name__ == "__main__":
    main()
