"""Drop-in for TGL's pybind11 module `sampler_core` (the "TGL sampler built by
setup.py build_ext" that the reference's README.md:1-5 and config/TGN.yml:1-9
describe; its C++ source is NOT part of the reference tree).

    sampler = ParallelSampler(indptr, indices, eid, ts, num_thread_per_worker,
                              num_workers, num_layers, num_neighbors, recent,
                              prop_time, num_history, window_duration)
    sampler.sample(root_nodes.astype(np.int32), root_ts.astype(np.float32))
    ret = sampler.get_ret()          # list[TemporalGraphBlock], numpy accessors

Sampling runs on the GPU (csrc/tcsr.cu) over a t-CSR graph that is uploaded
once at construction; `sample_device()` is the zero-copy entry for GPU
pipelines.  `num_thread_per_worker` / `num_workers` are accepted for API
compatibility (the CUDA grid replaces the OpenMP team).  Uniform sampling draws
from Philox keyed by (seed, root index, draw) -- reproducible, unlike upstream's
per-thread rand_r.
"""
import time
from typing import List

import numpy as np
import torch

from tgn_b200 import ops


class TemporalGraphBlock:
    """One sampled layer/snapshot; numpy accessors follow TGL's binding."""

    def __init__(self, row, col, eid, ts, dts, nodes, dim_in, dim_out):
        self._row, self._col, self._eid = row, col, eid
        self._ts, self._dts, self._nodes = ts, dts, nodes
        self._dim_in, self._dim_out = int(dim_in), int(dim_out)
        self.ptr_time = self.search_time = self.sample_time = self.tot_time = self.coo_time = 0.0

    def row(self): return self._row
    def col(self): return self._col
    def eid(self): return self._eid
    def ts(self): return self._ts
    def dts(self): return self._dts
    def nodes(self): return self._nodes
    def dim_in(self): return self._dim_in
    def dim_out(self): return self._dim_out


class ParallelSampler:
    def __init__(self, indptr, indices, eid, ts, num_thread_per_worker: int, num_workers: int,
                 num_layers: int, num_neighbors: List[int], recent: bool, prop_time: bool,
                 num_history: int, window_duration: float, device=None, seed: int = 0):
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("sampler_core (B200 build) needs a CUDA device; there is no CPU fallback")
        self.device = dev
        as_t = lambda a, dt: (a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))).to(dev, dt).contiguous()
        self.indptr = as_t(indptr, torch.int32)
        self.indices = as_t(indices, torch.int32)
        self.eid = as_t(eid, torch.int32)
        self.ts = as_t(ts, torch.float32)
        self.coarse = ops.tcsr_build_index(self.ts) if self.ts.numel() else None   # skip index, once per graph
        self.num_threads = num_thread_per_worker * num_workers
        self.num_layers = num_layers
        self.num_neighbors = list(num_neighbors)
        self.recent, self.prop_time = bool(recent), bool(prop_time)
        self.num_history, self.window_duration = int(num_history), float(window_duration)
        self.seed = int(seed)
        self._calls = 0
        self.ret: List[TemporalGraphBlock] = []

    def reset(self):
        """Upstream clears its per-node time pointers; the binary search here is stateless."""
        self.ret = []

    # -- device-level API -------------------------------------------------
    def sample_device(self, roots: torch.Tensor, root_ts: torch.Tensor):
        """Returns, per layer and snapshot, a dict of device tensors
        (nbr, col, eid, ts, dts, root_off, count) -- no host synchronisation."""
        out = []
        cur_nodes, cur_ts = [roots.to(self.device, torch.int32)], [root_ts.to(self.device, torch.float32)]
        strategy = ops.SAMPLE_RECENT if self.recent else ops.SAMPLE_UNIFORM
        for layer in range(self.num_layers):
            nxt_nodes, nxt_ts = [], []
            for h in range(self.num_history):
                rn = cur_nodes[h if layer > 0 else 0]
                rt = cur_ts[h if layer > 0 else 0]
                (nbr, col, eid, ts, dts), off, cnt = ops.tcsr_sample(
                    self.indptr, self.indices, self.eid, self.ts, rn, rt, self.num_neighbors[layer],
                    strategy, offset=-h * self.window_duration, duration=self.window_duration, coarse=self.coarse,
                    seed=self.seed + 1315423911 * self._calls + 2654435761 * (layer * self.num_history + h))
                out.append(dict(roots=rn, root_ts=rt, nbr=nbr, col=col, eid=eid, ts=ts, dts=dts,
                                root_off=off, count=cnt))
                if layer + 1 < self.num_layers:
                    n = int(cnt.item())
                    nxt_nodes.append(torch.cat([rn, nbr[:n]]))
                    nxt_ts.append(torch.cat([rt, rt[col[:n].long()] if self.prop_time else ts[:n]]))
            cur_nodes, cur_ts = nxt_nodes, nxt_ts
        self._calls += 1
        return out

    # -- TGL numpy API ----------------------------------------------------
    def sample(self, root_nodes, root_ts):
        t0 = time.perf_counter()
        roots = torch.as_tensor(np.ascontiguousarray(root_nodes, dtype=np.int32))
        rts = torch.as_tensor(np.ascontiguousarray(root_ts, dtype=np.float32))
        blocks = self.sample_device(roots, rts)
        self.ret = []
        for b in blocks:
            n = int(b["count"].item())
            R = b["roots"].numel()
            nbr = b["nbr"][:n].cpu().numpy()
            blk = TemporalGraphBlock(
                row=np.arange(R, R + n, dtype=np.int32),
                col=b["col"][:n].cpu().numpy(),
                eid=b["eid"][:n].cpu().numpy(),
                ts=np.concatenate([b["root_ts"].cpu().numpy(), b["ts"][:n].cpu().numpy()]),
                dts=np.concatenate([np.zeros(R, np.float32), b["dts"][:n].cpu().numpy()]),
                nodes=np.concatenate([b["roots"].cpu().numpy(), nbr]),
                dim_in=R + n, dim_out=R)
            self.ret.append(blk)
        tot = time.perf_counter() - t0
        for blk in self.ret:
            blk.tot_time = tot
            blk.sample_time = tot

    def get_ret(self) -> List[TemporalGraphBlock]:
        return self.ret
