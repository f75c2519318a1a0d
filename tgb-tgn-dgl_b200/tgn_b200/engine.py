"""Fused TGN training / evaluation step on the sm_100a kernels.

One step = the per-batch flow the reference's four hot-path pieces serve
(pyg_epoch_utils.py:106-137 commented flow == PyG/TGB tgn.py):

    roots = unique(src, dst, neg)            -> bitmap ranking                 (unique.cu)
    neighbour lookup of the roots            -> ring kernel                    (nbr_ring.cu)
    n_id = unique(roots + neighbours), relabel                                 (step.cu)
    z, last_update = memory(n_id)            -> store gather + concat + time-enc + Last
                                                aggregation (msgstore.cu), fused GRUCell: both gate
                                                GEMMs in TMEM + gate math (gemm_tma.cu)
    z = gnn(z, last_update, edges, t, msg)   -> projection / edge GEMMs + softmax core (step.cu); in training
                                                (batch <= 600) the softmax core runs INSIDE the decoder launch
    pos/neg logits, BCE loss                 -> fused decoder forward + loss + d_emb (step.cu)
    backward                                 -> hand-derived, same kernels; every weight
                                                gradient is a split-K GEMM accumulating
                                                straight into the flat gradient buffer
    Adam                                     -> one launch over the flat buffers (dense.cu)
    loader.insert + sampling of batch s+1    -> forked stream, from the START of step s
    memory.update_state                      -> forked stream, after the GRU forward
(stream layout and the timeline it was cut from: DESIGN.md section 5)

No autograd, no allocation and no host synchronisation inside a step: all buffers are
sized by upper bounds (3B roots, 3B*K edges, ...), the true counts stay in device memory,
and the whole step is captured once into a CUDA graph.  State and weights use the
reference's names so that state_dicts move freely between this engine, the drop-in
modules and the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _cabi, ops
from ._cabi import check

_L = _cabi.lib
_p = ops._p
_stream = ops._stream


def _up4(n: int) -> int:
    return (n + 3) // 4 * 4


class TGNEngine:
    def __init__(self, num_nodes: int, raw_dim: int, hidden: int, size_k: int, batch_size: int,
                 device="cuda", lr: float = 1e-4, heads: int = 2, dropout: float = 0.1,
                 log_capacity: int = 1 << 20, seed: int = 0, use_graph: bool = True,
                 precision: int = 3, rank: int = 0, world: int = 1, group=None, fused_zero_grad: bool = False,
                 part_exchange: str = "p2p", share: Optional["TGNEngine"] = None, part_compute: str = "owner",
                 track_metrics: bool = False, group_size: int = 3):
        """group_size: training steps per captured graph / host group (train_steps, train_group_logged); the
        staging slots are three such groups.
        share: another engine of the same model (nodes, dims, K, world) whose weights, Adam moments, node
        memory, neighbour ring, message store, event arrays and cursors this one USES instead of allocating
        its own -- a second step geometry (another batch size, e.g. the tail batch of an epoch) on the same
        training state.  Only one of the engines may be stepping at a time."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TGNEngine runs on CUDA only (no CPU fallback)")
        if hidden % 4 or hidden % heads:
            raise ValueError("hidden must be a multiple of 4 (TMA row alignment) and of heads")
        if precision not in (1, 3):
            raise ValueError("precision: 1 = tf32 tensor cores, 3 = 3xTF32 (fp32-level accuracy)")
        self.dev, self.N, self.De, self.D, self.K, self.B = dev, num_nodes, raw_dim, hidden, size_k, batch_size
        self.H, self.C = heads, hidden // heads
        self.HC = self.H * self.C
        self.Dt = hidden
        self.Dx = 2 * hidden + raw_dim + hidden          # IdentityMessage width (msg_func.py:15)
        self.ldx = _up4(self.Dx)                          # TMA-aligned row strides
        self.Din = self.Dt + raw_dim                      # edge_attr width (emb_module.py:28)
        self.lde = _up4(self.Din)
        self.lr, self.dropout, self.seed, self.use_graph, self.prec = lr, dropout, seed, use_graph, precision
        # owner-partitioned node memory (csrc/partition.cu): node n lives on rank n % world at local
        # row n // world; ring, message store, weights and batches are replicated
        if not (0 <= rank < world):
            raise ValueError("rank must be in [0, world)")
        self.rank, self.world, self.group = rank, world, group
        self.Nloc = (num_nodes + world - 1) // world
        D, Dt, HC = self.D, self.Dt, self.HC
        # ---- parameters: one flat buffer (one Adam launch, one memset for all gradients).
        # (name, logical shape, leading dimension): padded columns stay zero for ever -- their
        # gradients are exact zeros, so Adam never moves them
        table = [("time_enc.lin.weight", (Dt, 1), 1), ("time_enc.lin.bias", (Dt,), None),
                 ("memory_updater.weight_ih", (3 * D, self.Dx), self.ldx),
                 ("memory_updater.weight_hh", (3 * D, D), D),
                 ("memory_updater.bias_ih", (3 * D,), None), ("memory_updater.bias_hh", (3 * D,), None),
                 ("conv.w_node", (4 * HC, D), D), ("conv.b_node", (4 * HC,), None),
                 ("conv.lin_edge.weight", (HC, self.Din), self.lde),
                 ("lin_src.weight", (D, D), D), ("lin_src.bias", (D,), None),
                 ("lin_dst.weight", (D, D), D), ("lin_dst.bias", (D,), None),
                 ("lin_final.weight", (1, D), D), ("lin_final.bias", (1,), None)]
        self._table = table
        self.off: Dict[str, int] = {}
        self.ld: Dict[str, int] = {}
        o = 0
        for name, shp, ld in table:
            self.off[name] = o
            self.ld[name] = ld if ld is not None else shp[0]
            o += _up4(shp[0] * ld if ld is not None else shp[0])
        self.n_param = o
        # gradients, the loss scalar and the embedding-gradient rows share one zero-filled blob
        R, E, Nb = self._bounds(batch_size)
        if share is not None and (share.N, share.De, share.D, share.K, share.H, share.world, share.rank) != \
                (num_nodes, raw_dim, hidden, size_k, heads, world, rank):
            raise ValueError("share: the two engines must describe the same model and partition")
        # Partitioned memory with owner-side compute (world > 1, peer exchange): the gradient blob is symmetric
        # memory -- [replicated part | loss | PARTIAL sums over the rows this rank owns | d_emb] -- because the
        # optimiser of every rank reads rank 0's replicated part and all ranks' partial parts out of peer memory.
        if part_compute not in ("owner", "replicated"):
            raise ValueError("part_compute must be 'owner' or 'replicated'")
        self.owner_compute = world > 1 and part_exchange == "p2p" and part_compute == "owner" and share is None
        if self.owner_compute:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            # THREE blobs in rotation (blob = slot % 3; nine slots): the optimiser of a slow rank may still be
            # reading this rank's gradients of step s while this rank already clears and fills the blob of step
            # s+1; a blob comes round again three steps later, after two more rank barriers.  That removes the
            # barrier between one step's optimiser and the next step's gradient clear.
            self._grad_blobs = []
            for _ in range(3):
                blob = symm_mem.empty((o + 4 + o + Nb * HC,), dtype=torch.float32, device=dev)
                blob.zero_()
                h = symm_mem.rendezvous(blob, grp)
                self._grad_blobs.append(SimpleNamespace(
                    blob=blob, handle=h, rep0=int(h.buffer_ptrs[0]),
                    peer_part=(ctypes.c_void_p * world)(*[int(q) + 4 * (o + 4) for q in h.buffer_ptrs]),
                    grad=blob[:o], loss=blob[o:o + 1], part=blob[o + 4:o + 4 + o], d_emb=blob[o + 4 + o:].view(Nb, HC)))
            g0 = self._grad_blobs[0]
            self.zero_blob, self._grad_rep0, self._peer_grad_part = g0.blob, g0.rep0, g0.peer_part
            self.flat_grad_part, self.d_emb = g0.part, g0.d_emb
            fused_zero_grad = False      # peers read this rank's gradients: a blob is cleared when its turn comes again
        else:
            self.zero_blob = torch.zeros(o + 4 + Nb * HC, device=dev)
            self.flat_grad_part = None
            self.d_emb = self.zero_blob[o + 4:].view(Nb, HC)
        self.flat = torch.zeros(o, device=dev) if share is None else share.flat
        self.flat_grad = self.zero_blob[:o]
        self.loss_acc = self.zero_blob[o:o + 1]
        self.exp_avg = torch.zeros(o, device=dev) if share is None else share.exp_avg
        self.exp_avg_sq = torch.zeros(o, device=dev) if share is None else share.exp_avg_sq
        self.adam_step_dev = torch.zeros(1, device=dev) if share is None else share.adam_step_dev
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        # fused_zero_grad: Adam clears the gradients (and the other per-step accumulators) after using
        # them, so a step starts without a memset; p[...].grad then reads zero after train_step()
        self.fused_zero_grad = fused_zero_grad
        self.p: Dict[str, Tensor] = {}
        for name, shp, ld in table:
            v = self._view(self.flat, name, shp)
            v.requires_grad_()
            v.grad = self._view(self.flat_grad, name, shp)
            self.p[name] = v
        # ---- state
        # this rank's shard (all of it if world == 1).  Partitioned + "p2p": the shards are torch symmetric
        # memory, mapped into every rank (NVLink / NVSwitch peer access): the row assembly reads remote rows
        # straight out of their owners' HBM (tgn_part_gather_p2p) between two device-side rank barriers,
        # instead of staging owned rows and all-reducing 2*Nb*D floats ("allreduce").
        if part_exchange not in ("p2p", "allreduce"):
            raise ValueError("part_exchange must be 'p2p' or 'allreduce'")
        self.part_exchange = part_exchange if world > 1 else "none"
        self._shared_with = share
        if share is not None:
            if share._shared_with is not None:
                raise ValueError("share: pass the engine that owns the state, not one of its siblings")
            share.__dict__.setdefault("_siblings", []).append(self)
            self.part_exchange = share.part_exchange
            for k in ("memory", "last_update", "assoc", "assoc_r", "neighbors", "e_id", "t_ring", "store", "bitmap",
                      "cur_e_id_dev", "log_base_dev", "pos_dev", "step_dev", "_symm_mem", "_symm_lu", "_peer_mem",
                      "_peer_lu"):
                if hasattr(share, k):
                    setattr(self, k, getattr(share, k))
        elif self.part_exchange == "p2p":
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self.memory = symm_mem.empty((self.Nloc, D), dtype=torch.float32, device=dev)
            self.last_update = symm_mem.empty((self.Nloc,), dtype=torch.long, device=dev)
            self.memory.zero_()
            self.last_update.zero_()
            self._symm_mem = symm_mem.rendezvous(self.memory, grp)
            self._symm_lu = symm_mem.rendezvous(self.last_update, grp)
            self._peer_mem = (ctypes.c_void_p * world)(*[int(p) for p in self._symm_mem.buffer_ptrs])
            self._peer_lu = (ctypes.c_void_p * world)(*[int(p) for p in self._symm_lu.buffer_ptrs])
            torch.cuda.synchronize()
            self._symm_mem.barrier(channel=0)
        elif share is None:
            self.memory = torch.zeros((self.Nloc, D), device=dev)
            self.last_update = torch.zeros(self.Nloc, dtype=torch.long, device=dev)
        if share is None:
            self.assoc = torch.zeros(num_nodes, dtype=torch.long, device=dev)
            self.assoc_r = torch.zeros(num_nodes, dtype=torch.long, device=dev)    # rank of a node in the ROOT list of its batch
            self.neighbors = torch.zeros((num_nodes, size_k), dtype=torch.long, device=dev)
            self.e_id = torch.full((num_nodes, size_k), -1, dtype=torch.long, device=dev)
            self.t_ring = torch.full((num_nodes, size_k), -1.0, device=dev)
            self.store = ops.MsgStore(num_nodes, raw_dim, dev, capacity=log_capacity, t_dtype=torch.int64)
            self.bitmap = torch.zeros(_L().tgn_bitmap_bytes(num_nodes) // 4, dtype=torch.int32, device=dev)
            self.cur_e_id_dev = torch.zeros(1, dtype=torch.long, device=dev)   # ring event counter
            self.log_base_dev = torch.zeros(1, dtype=torch.long, device=dev)   # store log position
            self.pos_dev = torch.zeros(1, dtype=torch.long, device=dev)        # dataset cursor
            self.step_dev = torch.zeros(1, dtype=torch.long, device=dev)       # dropout stream
        # host mirrors of the device cursors (a graph replay does not run Python): events_done = log_base_dev
        # (message-store log position, reset by flush_to_eval), ring_pos = cur_e_id_dev (the e_id the NEXT
        # inserted event gets; e_id == row of the resident event arrays)
        self._cur = SimpleNamespace(events_done=0, ring_pos=0, events=None) if share is None else share._cur
        self.side = torch.cuda.Stream(device=dev)    # ring insert + sampling of the NEXT batch
        self.upd = torch.cuda.Stream(device=dev)     # memory / message-store update of this batch
        self.aux = torch.cuda.Stream(device=dev, priority=0)   # edge branch of the attention / small reductions
        self.w = self._alloc_work(R, E, Nb, batch_size)
        if self.owner_compute:
            # every rank's table of the step's rows (h', last_update'): owners publish into it over NVLink
            self.w.z = symm_mem.empty((Nb, D), dtype=torch.float32, device=dev)
            self.w.lu = symm_mem.empty((Nb,), dtype=torch.long, device=dev)
            self.w.z.zero_()
            self.w.lu.zero_()
            hz, hl = symm_mem.rendezvous(self.w.z, grp), symm_mem.rendezvous(self.w.lu, grp)
            self._symm_rows = (hz, hl)
            self._peer_z = (ctypes.c_void_p * world)(*[int(q) for q in hz.buffer_ptrs])
            self._peer_rowlu = (ctypes.c_void_p * world)(*[int(q) for q in hl.buffer_ptrs])
            # workspace of the rows this rank owns
            self.So_b = self._owner_bound(Nb)
            self.wo = self._alloc_work(1, 1, self.So_b, 1)
            torch.cuda.synchronize()
        # Slots of {batch inputs, sampling results} in rotation: while step s runs on slot `cur`, the forked
        # stream already loads and samples batch s+1 into the next slot (software pipelining), and a host
        # loader can copy later batches into slots the running steps do not touch.
        # Nine slots = three groups of three: a host loader can also feed whole groups (stage_group +
        # train_group_logged: one H2D copy, one graph launch and one loss read-back per THREE steps).  With
        # three groups the copy of group g+1 only has to wait for group g-2, so it overlaps group g-1
        # completely and group g starts the moment g-1 ends (two groups left a copy-sized bubble per group).
        # The staging regions of all slots are one allocation so a group's three batches are contiguous.
        if group_size < 1:
            raise ValueError("group_size must be >= 1")
        self.nslots, self.group_size = 3 * group_size, group_size
        De1 = max(raw_dim, 1)
        self._packed = (32 * batch_size + 4 * batch_size * De1 + 15) // 16 * 16      # bytes per staged batch
        self._in_all = torch.zeros(self.nslots * self._packed, dtype=torch.uint8, device=dev)
        self.slots = [self._alloc_slot(R, E, Nb, batch_size, i) for i in range(self.nslots)]
        self.cur = 0
        self.copy_stream = torch.cuda.Stream(device=dev)    # H2D of host-staged batches
        self.loss_stream = torch.cuda.Stream(device=dev)    # D2H of the per-step loss (train_step_logged)
        self._slot_free = [torch.cuda.Event() for _ in range(self.nslots)]    # last step that used the slot is done
        self._slot_ready = [torch.cuda.Event() for _ in range(self.nslots)]   # its staged batch has landed
        self._slot_async = [False] * self.nslots
        self._primed = None      # "device" / "host": slots[cur] holds a staged AND sampled batch
        self._bind(0)
        self.loss_slots = torch.zeros(self.nslots, device=dev)   # loss of the step that trained on slot i (a lagged
        self._loss_views = [self.loss_slots[i] for i in range(self.nslots)]   # host read of step s survives until the
        self.loss = self._loss_views[0]                                        # slot comes round again)
        self.bounds = (R, E, Nb)
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._graph_store_gen = None      # generation of the message-store log the captured graphs point into
        self.training = True
        self.fused_decoder = _L().tgn_dec_fused_smem_bytes(hidden) <= 215 * 1024 and hidden <= 128
        self.fused_gru = hidden % 4 == 0    # tgn_gru_fused_fwd (TMA strides need 16-byte rows)
        # attention forward computed inside the decoder launch (tgn_dec_attn_fused): two heads, <= 10 neighbours
        # (up to ~one decoder pass per SM: at B = 2000 the 500 passes of 148 CTAs would walk what the stand-alone
        # attention kernel spreads over 1,500 CTAs -- measured 0.358 against 0.325 ms per wiki step)
        self.fused_attn_dec = (self.fused_decoder and heads == 2 and size_k <= 10 and self.HC == hidden and
                               hidden <= 128 and (hidden // heads) % 2 == 0 and batch_size <= 600 and
                               os.environ.get("TGN_FUSED_ATTN_DEC", "1") == "1")
        self.dz_split = 3                   # split-K of the d_z GEMM (1 = plain store)
        self.probe = None   # bench.py: {"name": [(start_event, stop_event), ...]} filled in eager steps
        # per-batch AP / AUC of the training logits, accumulated on the device (epoch_utils.py:312-317)
        self.track_metrics = track_metrics
        self.metric_acc = (torch.zeros(3, dtype=torch.float64, device=dev) if share is None
                           else share.metric_acc)

    events_done = property(lambda self: self._cur.events_done,
                           lambda self, v: setattr(self._cur, "events_done", v))
    ring_pos = property(lambda self: self._cur.ring_pos, lambda self, v: setattr(self._cur, "ring_pos", v))
    events = property(lambda self: self._cur.events, lambda self, v: setattr(self._cur, "events", v))

    def _family(self):
        root = self
        while root._shared_with is not None:
            root = root._shared_with
        return [root] + list(getattr(root, "_siblings", []))

    def _advance(self, n_events: int):
        """host mirrors after n_events more events have been queued (graph replays do not run Python)"""
        self._cur.events_done += n_events
        self._cur.ring_pos += n_events
        self.store.size = self._cur.events_done

    def _grow_log(self, need: int):
        """Re-allocates the message-store log.  Captured graphs hold the old ev_* / perm pointers BY VALUE, so
        every graph (of this engine and of the engines sharing its state) is dropped and re-captured."""
        self.store._alloc_log(int(need))
        self._drop_graphs()

    def _drop_graphs(self):
        """Forgets every captured graph of this engine and of the engines sharing its state (they hold device
        addresses and scalars by value); the next calls run eagerly and re-capture."""
        torch.cuda.synchronize()
        for e in self._family():
            e._graphs = {}
            e._graph_store_gen = None

    def _reserve(self, n_events: int):
        """Host-side guard in front of every replay: the log has room for the events about to be appended
        (the kernel would drop them and raise TGN_DEVERR_LOG_OVERFLOW), and the resident event arrays cover
        every e_id the ring is about to hand out (edge features are events['msg'][e_id], epoch_utils.py:224)."""
        self.check_device_errors()
        if any(e._graph_store_gen not in (None, self.store.generation) for e in self._family()):
            # the log arrays moved behind the engine's back (a module-path update_state grew the shared store)
            self._drop_graphs()
        if self.events_done + n_events > self.store.capacity:
            self._grow_log(max(2 * self.store.capacity, self.events_done + n_events))
        if self.events is not None and self.ring_pos + n_events > self.events["src"].numel():
            raise _cabi.TgnError(
                f"ring event id {self.ring_pos + n_events} would pass the {self.events['src'].numel()} resident events: "
                "set_events() must hold EVERY event the neighbour ring can name (train + val + test), e_id == row")

    def check_device_errors(self):
        """Raises if a kernel flagged a contract violation (see tgn_device_errors in include/tgn_b200.h);
        a plain host read, valid for work that has completed."""
        v = _L().tgn_device_errors(1)
        if v:
            names = [n for b, n in ((1, "message-store log overflow (events dropped)"),
                                    (2, "event id outside the resident event arrays"),
                                    (4, "batch above a kernel's sort capacity"),
                                    (8, "partitioned memory: one rank owns more rows of a step than its buffers hold"))
                     if v & b]
            raise _cabi.TgnError("device-side error flag: " + "; ".join(names))

    def epoch_metrics(self, reset: bool = True):
        """(mean AP, mean AUC) over the batches trained since the last reset (track_metrics=True): what the
        reference prints as "ap and auc" at the end of train() (epoch_utils.py:317).  One host read."""
        a = self.metric_acc.tolist()
        if reset:
            self.metric_acc.zero_()
        n = max(a[2], 1.0)
        return a[0] / n, a[1] / n

    def handover(self):
        """Call before stepping an engine that shares this one's state (share=): drops the pre-sampled batch."""
        self._unprime()

    # ------------------------------------------------------------------ layout helpers
    def _bounds(self, B: int, roots: Optional[int] = None):
        R = 3 * B if roots is None else roots
        E = R * self.K
        return R, E, min(self.N, R + E)

    def _owner_bound(self, Nb: int) -> int:
        """Rows of a step one rank is sized for under owner-side compute: twice the even share (ownership is
        node id mod world over essentially random ids); tgn_part_select_owned flags TGN_DEVERR_OWNER_CAP if a
        step ever exceeds it."""
        return min(Nb, max(1024, 2 * ((Nb + self.world - 1) // self.world)))

    def _view(self, flat: Tensor, name: str, shp) -> Tensor:
        o, ld = self.off[name], self.ld[name]
        if len(shp) == 1:
            return flat[o:o + shp[0]]
        return flat[o:o + shp[0] * ld].view(shp[0], ld)[:, :shp[1]]

    def _g(self, name: str) -> int:
        """element offset of a parameter's gradient inside flat_grad (== offset inside flat)"""
        return self.off[name]

    def _alloc_work(self, R: int, E: int, Nb: int, B: int, train: bool = True) -> SimpleNamespace:
        dev, D, HC, H = self.dev, self.D, self.HC, self.H
        f = lambda *s: torch.zeros(s, device=dev)
        i64 = lambda *s: torch.zeros(s, dtype=torch.long, device=dev)
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)
        w = SimpleNamespace(R=R, E=E, Nb=Nb, B=B)
        self._alloc_sample_fields(w, R, E, Nb)
        # memory forward
        w.x, w.h = f(Nb, self.ldx), f(Nb, D)
        w.lu, w.sel_ev, w.sel_dt = i64(Nb), i32(Nb), f(Nb)
        w.sn_m = f(Nb, max(self.Dt, 1)) if train else None
        if self.world > 1:   # staging for the row assembly: [rows of n_id | rows of the other endpoints]
            w.g_rows, w.g_lu = f(2, Nb, D), i64(Nb)
        w.gi, w.gh, w.z, w.gates = f(Nb, 3 * D), f(Nb, 3 * D), f(Nb, D), f(Nb, 4 * D)
        # attention
        w.proj = f(Nb, 4 * HC)
        w.ea, w.sn_e, w.rel = f(E, self.lde), f(E, max(self.Dt, 1)), f(E)
        w.ee, w.alpha, w.emb = f(E, HC), f(E, H), f(Nb, HC)
        if train:
            w.zcat, w.hcat, w.dhcat, w.dzcat = f(3 * B, D), f(3 * B, D), f(3 * B, D), f(3 * B, D)
            w.logits = f(2 * B)
            w.d_proj, w.d_ee, w.d_eat = f(Nb, 4 * HC), f(E, HC), f(E, max(self.Dt, 1))
            w.d_z, w.d_gi, w.d_gh, w.d_x = f(Nb, D), f(Nb, 3 * D), f(Nb, 3 * D), f(Nb, self.ldx)
        return w

    _SAMPLE_FIELDS = ("roots", "R_dev", "nbr_g", "ctr_g", "eid", "t_e", "root_off", "E_dev", "lookup_ws", "n_id",
                      "Nb_dev", "nbr_l", "ctr_l", "own_n", "own_pos", "So_dev")
    _SLOT_FIELDS = _SAMPLE_FIELDS + ("ids_l", "ids_r", "in_i64", "in_ids3", "in_t_i64", "in_t_f32", "in_msg")

    def _alloc_sample_fields(self, w, R: int, E: int, Nb: int):
        dev = self.dev
        f = lambda *s: torch.zeros(s, device=dev)
        i64 = lambda *s: torch.zeros(s, dtype=torch.long, device=dev)
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)
        w.roots, w.R_dev = i64(R), i32(1)
        w.nbr_g, w.ctr_g, w.eid, w.t_e = i64(E), i64(E), i64(E), f(E)
        w.root_off, w.E_dev = i32(R + 1), i32(1)
        w.lookup_ws = i64(max(_L().tgn_nbr_lookup_ws_bytes(R, self.K), 16) // 8)
        w.n_id, w.Nb_dev = i64(Nb), i32(1)
        w.nbr_l, w.ctr_l = i64(E), i64(R)
        # owner-side compute: the rows of n_id this rank owns (picked right behind the sampling, off the chain)
        So = self._owner_bound(Nb) if getattr(self, "owner_compute", False) else 1
        w.own_n, w.own_pos, w.So_dev = i64(So), i64(So), i32(1)

    def _alloc_slot(self, R: int, E: int, Nb: int, B: int, index: int = 0) -> SimpleNamespace:
        dev = self.dev
        sl = SimpleNamespace(R=R, E=E, Nb=Nb, B=B)
        self._alloc_sample_fields(sl, R, E, Nb)
        sl.ids_l = torch.zeros(3 * B, dtype=torch.long, device=dev)
        sl.ids_r = torch.zeros(3 * B, dtype=torch.long, device=dev)      # index of every batch id in the root list
        # one contiguous staging region per slot: [src | dst | neg | t] int64 followed by msg [B, De] float32,
        # so a host batch is ONE H2D copy (stage_packed1); in_i64 / in_msg are views of it
        De1 = max(self.De, 1)
        sl.in_raw = self._in_all[index * self._packed: index * self._packed + 32 * B + 4 * B * De1]
        sl.in_i64 = sl.in_raw[:32 * B].view(torch.long)
        sl.in_ids3, sl.in_t_i64 = sl.in_i64[:3 * B], sl.in_i64[3 * B:]
        sl.in_t_f32 = torch.zeros(B, device=dev)
        sl.in_msg = sl.in_raw[32 * B:].view(torch.float32).view(B, De1)
        return sl

    def _bind(self, idx: int):
        """Points the step workspace (and the in_* attributes) at slot `idx`."""
        sl = self.slots[idx]
        for k in self._SLOT_FIELDS:
            setattr(self.w, k, getattr(sl, k))
        self.in_i64, self.in_ids3, self.in_t_i64 = sl.in_i64, sl.in_ids3, sl.in_t_i64
        self.in_t_f32, self.in_msg = sl.in_t_f32, sl.in_msg
        if getattr(self, "owner_compute", False):      # this step's gradient blob (rotation of three)
            g = self._grad_blobs[idx % 3]
            self.zero_blob, self._grad_rep0, self._peer_grad_part = g.blob, g.rep0, g.peer_part
            self.flat_grad, self.loss_acc, self.flat_grad_part, self.d_emb = g.grad, g.loss, g.part, g.d_emb

    def _next_slot(self) -> int:
        return (self.cur + 1) % self.nslots

    def _unprime(self):
        """Drops a batch that was pre-sampled but not trained on (mode switch, flush, reset)."""
        if self._primed == "device":
            self.pos_dev -= self.B          # the dataset cursor had already moved past it
        self._primed = None

    # ------------------------------------------------------------------ weights / state exchange
    def load_state(self, memory_sd: Dict[str, Tensor], gnn_sd: Dict[str, Tensor], lp_sd: Dict[str, Tensor]):
        """state_dicts of (TGNMemory, GraphAttentionEmbedding, LinkPredictor) in the reference's key names."""
        with torch.no_grad():
            g = lambda sd, k: sd[k].to(self.dev, torch.float32)
            for k in ("time_enc.lin.weight", "time_enc.lin.bias", "memory_updater.weight_ih",
                      "memory_updater.weight_hh", "memory_updater.bias_ih", "memory_updater.bias_hh"):
                self.p[k].copy_(g(memory_sd, k))
            self.p["conv.w_node"].copy_(torch.cat([g(gnn_sd, f"conv.lin_{n}.weight") for n in ("query", "key", "value", "skip")]))
            self.p["conv.b_node"].copy_(torch.cat([g(gnn_sd, f"conv.lin_{n}.bias") for n in ("query", "key", "value", "skip")]))
            self.p["conv.lin_edge.weight"].copy_(g(gnn_sd, "conv.lin_edge.weight"))
            for k in ("lin_src.weight", "lin_src.bias", "lin_dst.weight", "lin_dst.bias", "lin_final.weight", "lin_final.bias"):
                self.p[k].copy_(g(lp_sd, k))
            if "memory" in memory_sd:   # full [N,D] table: every rank keeps the rows it owns
                own = memory_sd["memory"].to(self.dev)[self.rank::self.world]
                self.memory.zero_()
                self.last_update.zero_()
                self.memory[:own.shape[0]].copy_(own)
                self.last_update[:own.shape[0]].copy_(memory_sd["last_update"].to(self.dev)[self.rank::self.world])

    # ---- the drop-in modules as the owners of the state (epoch_utils.train on the engine)
    def attach_modules(self, memory_module, neighbor_loader):
        """Makes the engine step ON the drop-in modules' own state tensors instead of private copies:
        TGNMemory.memory / last_update / message store and LastNeighborLoader.neighbors / e_id / t / _assoc.
        After an engine epoch the modules therefore ARE in the state the reference's loop would have left
        them in (memory in train mode with its messages pending, memory_module.py:126-138) -- no copy, no
        flush; `memory.eval()` then flushes through the module's own code path.  Single GPU, before the
        first step."""
        if self.world != 1 or self._graphs or self._shared_with is not None:
            raise _cabi.TgnError("attach_modules: single-GPU engines only, before the first step")
        m, nl = memory_module, neighbor_loader
        ok = (m.memory.is_cuda and m.memory.is_contiguous() and tuple(m.memory.shape) == (self.N, self.D) and
              m.memory.dtype == torch.float32 and tuple(nl.neighbors.shape) == (self.N, self.K) and
              nl.t.dtype == torch.float32 and nl.e_id.dtype == torch.long)
        if not ok:
            raise _cabi.TgnError("attach_modules: module state does not match the engine's shapes / dtypes")
        st = m.store
        if st.t_dtype != torch.int64:
            st._set_t_dtype(torch.empty(0, dtype=torch.long))
        self.memory, self.last_update = m.memory.detach(), m.last_update
        self.neighbors, self.e_id, self.t_ring, self.assoc = nl.neighbors, nl.e_id, nl.t, nl._assoc
        self.store = st
        self._attached = (m, nl)

    def begin_epoch_on_modules(self):
        """Cursors to the start of an epoch whose state tensors the modules have just reset
        (memory.reset_state(), neighbor_loader.reset_state(): pyg_epoch_utils.py:15-16)."""
        m, nl = self._attached
        self.cur_e_id_dev.fill_(int(nl.cur_e_id))
        self.log_base_dev.fill_(int(m.store.size))
        self.pos_dev.zero_()
        self.events_done, self.ring_pos = int(m.store.size), int(nl.cur_e_id)
        self._primed = None
        self.cur = 0        # every epoch walks the slots from the same start: the captured graphs are keyed by slot

    def end_epoch_on_modules(self):
        m, nl = self._attached
        self._unprime()
        m.store.size = self.events_done
        nl.cur_e_id = self.ring_pos

    def _flat_named(self, flat: Tensor):
        return {name: self._view(flat, name, shp) for name, shp, _ in self._table}

    def _copy_in(self, views, mem_sd, gnn_sd, lp_sd):
        g = lambda sd, k: sd[k].to(self.dev, torch.float32)
        with torch.no_grad():
            for k in ("time_enc.lin.weight", "time_enc.lin.bias", "memory_updater.weight_ih",
                      "memory_updater.weight_hh", "memory_updater.bias_ih", "memory_updater.bias_hh"):
                views[k].copy_(g(mem_sd, k).view(views[k].shape))
            views["conv.w_node"].copy_(torch.cat([g(gnn_sd, f"conv.lin_{n}.weight") for n in ("query", "key", "value", "skip")]))
            views["conv.b_node"].copy_(torch.cat([g(gnn_sd, f"conv.lin_{n}.bias") for n in ("query", "key", "value", "skip")]))
            views["conv.lin_edge.weight"].copy_(g(gnn_sd, "conv.lin_edge.weight"))
            for k in ("lin_src.weight", "lin_src.bias", "lin_dst.weight", "lin_dst.bias", "lin_final.weight", "lin_final.bias"):
                views[k].copy_(g(lp_sd, k).view(views[k].shape))

    def _copy_out(self, views, mem_t, gnn_t, lp_t):
        """inverse of _copy_in: writes the flat views into the tensors of three name->tensor dicts"""
        HC = self.HC
        with torch.no_grad():
            for k in ("time_enc.lin.weight", "time_enc.lin.bias", "memory_updater.weight_ih",
                      "memory_updater.weight_hh", "memory_updater.bias_ih", "memory_updater.bias_hh"):
                mem_t[k].copy_(views[k].view(mem_t[k].shape))
            for i, n in enumerate(("query", "key", "value", "skip")):
                gnn_t[f"conv.lin_{n}.weight"].copy_(views["conv.w_node"][i * HC:(i + 1) * HC])
                gnn_t[f"conv.lin_{n}.bias"].copy_(views["conv.b_node"][i * HC:(i + 1) * HC])
            gnn_t["conv.lin_edge.weight"].copy_(views["conv.lin_edge.weight"])
            for k in ("lin_src.weight", "lin_src.bias", "lin_dst.weight", "lin_dst.bias", "lin_final.weight", "lin_final.bias"):
                lp_t[k].copy_(views[k].view(lp_t[k].shape))

    def sync_from_modules(self, model, optimizer=None):
        """weights (and, when the optimizer has stepped before, torch.optim.Adam's moments and step count)
        of the module triple -> the engine's flat buffers"""
        named = [dict(model[k].named_parameters()) for k in ("memory", "gnn", "link_pred")]
        self._copy_in(self._flat_named(self.flat), *named)
        if optimizer is None:
            return
        params = [p for d in named for p in d.values()]
        if not all(p in optimizer.state and "exp_avg" in optimizer.state[p] for p in params):
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.adam_step_dev.zero_()
            return
        for key, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
            self._copy_in(self._flat_named(flat), *[{n: optimizer.state[p][key] for n, p in d.items()} for d in named])
        self.adam_step_dev.fill_(float(optimizer.state[params[0]]["step"]))

    def sync_to_modules(self, model, optimizer=None):
        named = [dict(model[k].named_parameters()) for k in ("memory", "gnn", "link_pred")]
        self._copy_out(self._flat_named(self.flat), *named)
        if optimizer is None:
            return
        step = float(self.adam_step_dev)
        for d in named:
            for p in d.values():
                st = optimizer.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
                if torch.is_tensor(st["step"]):
                    st["step"].fill_(step)
                else:
                    st["step"] = step
        for key, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
            self._copy_out(self._flat_named(flat), *[{n: optimizer.state[p][key] for n, p in d.items()} for d in named])

    def export_state(self):
        HC = self.HC
        c = lambda k: self.p[k].detach().clone().contiguous()
        mem = {k: c(k) for k in ("time_enc.lin.weight", "time_enc.lin.bias", "memory_updater.weight_ih",
                                 "memory_updater.weight_hh", "memory_updater.bias_ih", "memory_updater.bias_hh")}
        full_mem, full_lu = self.full_memory()
        mem.update(memory=full_mem, last_update=full_lu, _assoc=self.assoc.clone())
        gnn = {"time_enc.lin.weight": mem["time_enc.lin.weight"], "time_enc.lin.bias": mem["time_enc.lin.bias"],
               "conv.lin_edge.weight": c("conv.lin_edge.weight")}
        for i, n in enumerate(("query", "key", "value", "skip")):
            gnn[f"conv.lin_{n}.weight"] = self.p["conv.w_node"].detach()[i * HC:(i + 1) * HC].clone()
            gnn[f"conv.lin_{n}.bias"] = self.p["conv.b_node"].detach()[i * HC:(i + 1) * HC].clone()
        lp = {k: c(k) for k in ("lin_src.weight", "lin_src.bias", "lin_dst.weight", "lin_dst.bias",
                                "lin_final.weight", "lin_final.bias")}
        return mem, gnn, lp

    def full_memory(self):
        """(memory [N,D], last_update [N]) reassembled from the shards (all-gather when partitioned)."""
        if self.world == 1:
            return self.memory.clone(), self.last_update.clone()
        import torch.distributed as dist
        mems = [torch.empty_like(self.memory) for _ in range(self.world)]
        lus = [torch.empty_like(self.last_update) for _ in range(self.world)]
        dist.all_gather(mems, self.memory, group=self.group)
        dist.all_gather(lus, self.last_update, group=self.group)
        from . import partition
        return partition.interleave(mems, self.N), partition.interleave(lus, self.N)

    def _all_reduce(self, t: Tensor):
        import torch.distributed as dist
        dist.all_reduce(t, group=self.group)

    def reset_state(self):
        """memory.reset_state() + neighbor_loader.reset_state() (start of every epoch, pyg_epoch_utils.py:15-16)."""
        self.memory.zero_()
        self.last_update.zero_()
        self.e_id.fill_(-1)
        self.t_ring.fill_(-1)
        self.cur_e_id_dev.zero_()
        self.log_base_dev.zero_()
        self.pos_dev.zero_()
        self.events_done = 0
        self.ring_pos = 0
        self._primed = None
        self.store.reset()

    # ------------------------------------------------------------------ data
    def set_events(self, src: Tensor, dst: Tensor, t: Tensor, msg: Tensor, neg: Optional[Tensor] = None):
        """Device-resident event arrays.  `t` int64, `msg` [E, De].  e_id of the ring == row of these arrays,
        like data.msg[e_id] in the reference (epoch_utils.py:224): they must hold EVERY event the ring can
        name -- when evaluation follows training on the same ring, pass train + val + test in stream order
        (_reserve() refuses to step past their end; the kernel flags TGN_DEVERR_EVENT_RANGE).  Host-staged
        batches (stage_*) feed the step's own batch only; they do not replace these arrays."""
        dev = self.dev
        new = dict(src=src, dst=dst, t=t, msg=msg, neg=neg)
        old = self.events
        if old is not None and all((old[k] is None) == (new[k] is None) and
                                   (new[k] is None or tuple(old[k].shape) == tuple(new[k].shape)) for k in new):
            for k, v in new.items():      # same sizes (a new epoch's negatives): refill in place, the captured
                if v is not None:         # graphs keep pointing at these buffers
                    old[k].copy_(v)
            return
        if old is not None:               # different arrays: the graphs hold the old pointers
            self._drop_graphs()
        self.events = dict(src=src.to(dev, torch.long).contiguous(), dst=dst.to(dev, torch.long).contiguous(),
                           t=t.to(dev, torch.long).contiguous(), msg=msg.to(dev, torch.float32).contiguous(),
                           neg=None if neg is None else neg.to(dev, torch.long).contiguous())
        need = self.events["src"].numel()
        if self.store.capacity < need:
            self._grow_log(need)

    def stage_batch_from_device(self, slot: Optional[int] = None):
        ev, sl = self.events, self.slots[self.cur if slot is None else slot]
        check(_L().tgn_batch_load(_p(ev["src"]), _p(ev["dst"]), _p(ev["neg"]), _p(ev["t"]), _p(ev["msg"]),
                                  self.De, self.B, ev["src"].numel(), _p(self.pos_dev), _p(sl.in_ids3), _p(sl.in_t_i64),
                                  _p(sl.in_t_f32), _p(sl.in_msg), _stream()))

    def stage_packed(self, ids_t: Tensor, msg: Tensor, ahead: bool = False):
        """End-to-end staging: `ids_t` = pinned host int64 [4B] = [src|dst|neg|t], `msg` pinned
        host float32 [B, De]; two H2D copies, the float timestamps are derived on the device.
        ahead=False: the batch the next train_step(from_device=False) trains on;
        ahead=True : the batch AFTER that one (train_step(..., lookahead=True) samples it while it
        trains on the current one -- the prefetching data-loader pattern)."""
        sl = self.slots[self._next_slot() if ahead else self.cur]
        sl.in_i64.copy_(ids_t, non_blocking=True)
        if self.De:
            sl.in_msg.copy_(msg, non_blocking=True)
        # (the float32 timestamps the ring wants are derived inside the step that consumes the batch)

    def packed_nbytes(self) -> int:
        """Size of the single-copy host batch of stage_packed1()."""
        return 32 * self.B + 4 * self.B * max(self.De, 1)

    def pack_host_batch(self, out: Tensor, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor) -> Tensor:
        """Fills `out` (pinned uint8 [packed_nbytes()]) with [src | dst | neg | t] int64 + msg float32."""
        B = self.B
        out[:32 * B].view(torch.long).copy_(torch.cat([src, dst, neg, t]))
        if self.De:
            out[32 * B:].view(torch.float32).view(B, self.De).copy_(msg)
        return out

    def stage_packed1(self, buf: Tensor, ahead: bool = False):
        """End-to-end staging with ONE H2D copy: `buf` = pinned host uint8 [packed_nbytes()] laid out by
        pack_host_batch().  `ahead` as in stage_packed."""
        if not ahead:
            self.slots[self.cur].in_raw.copy_(buf, non_blocking=True)
            return
        # the slot of batch s+1 was last used two steps ago, so this copy runs on its own stream while
        # step s-1 is still executing; train_step waits for it (an event that has usually fired already)
        idx = self._next_slot()
        cs = self.copy_stream
        cs.wait_event(self._slot_free[idx])
        raw = self.slots[idx].in_raw
        if buf.numel() != raw.numel() or buf.dtype != torch.uint8 or not buf.is_contiguous():
            raise _cabi.TgnError("stage_packed1: buf must be a contiguous uint8 tensor of packed_nbytes() bytes")
        check(_L().tgn_memcpy_async(raw.data_ptr(), buf.data_ptr(), raw.numel(), cs.cuda_stream))
        self._slot_ready[idx].record(cs)
        self._slot_async[idx] = True

    def prefill(self, count: int, ring_state=None):
        """Start from a mid-epoch state: the first `count` events of set_events() are taken as
        already seen.  ring_state = (neighbors, e_id, t) host/device tensors of the ring after
        those events (bench.py builds them vectorised); the cursors move to `count`."""
        if ring_state is not None:
            self.neighbors.copy_(ring_state[0])
            self.e_id.copy_(ring_state[1])
            self.t_ring.copy_(ring_state[2])
        self.cur_e_id_dev.fill_(count)
        self.log_base_dev.fill_(count)
        self.pos_dev.fill_(count)
        self.events_done = count
        self.ring_pos = count
        self.store.size = count
        self._primed = None

    def stage_batch(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor, ahead: bool = False):
        """Copies one batch (host or device tensors) into a slot (see stage_packed for `ahead`)."""
        B, sl = self.B, self.slots[self._next_slot() if ahead else self.cur]
        sl.in_ids3[:B].copy_(src, non_blocking=True)
        sl.in_ids3[B:2 * B].copy_(dst, non_blocking=True)
        sl.in_ids3[2 * B:].copy_(neg, non_blocking=True)
        sl.in_t_i64.copy_(t, non_blocking=True)
        sl.in_t_f32.copy_(t, non_blocking=True)
        if self.De:
            sl.in_msg.copy_(msg, non_blocking=True)

    # ------------------------------------------------------------------ pieces of the step
    def _timed(self, name: str, fn):
        """Runs fn(); with self.probe set (eager steps only) brackets it with CUDA events on the
        launching stream so bench.py can read the in-step duration of one launch."""
        if self.probe is None:
            fn()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.probe.setdefault(name, []).append((e0, e1))

    def _sample(self, w, ids: Tensor, ids_l: Tensor, ids_dev: Optional[Tensor] = None, owner_select: bool = False,
                ids_r: Optional[Tensor] = None):
        """roots -> neighbour lookup -> union -> relabel; everything bound-sized, counts on device.
        ids_dev: device count of the valid prefix of `ids` (entries beyond it must be -1: the marking
        kernels ignore them, the relabel stops in front of them)."""
        N, K = self.N, self.K
        L, s = _L(), _stream()
        # ids_r (training slots): the rank of every batch id in the root list, for the attention-fused decoder
        rank_of = _p(self.assoc_r) if ids_r is not None else None
        if ids.numel() <= 8192:   # small id list: marking folded into the single-CTA ranking launch
            check(L.tgn_unique_mark_rank(_p(ids), ids.numel(), _p(self.bitmap), N, _p(w.roots), w.R, rank_of,
                                         _p(w.R_dev), 1, s))
        else:
            check(L.tgn_unique_mark(_p(ids), ids.numel(), None, N, _p(self.bitmap), s))
            check(L.tgn_unique_rank(_p(self.bitmap), N, _p(w.roots), w.R, rank_of, _p(w.R_dev), 1, s))
        if ids_r is not None:
            check(L.tgn_relabel(_p(ids), ids.numel(), _p(ids_dev), _p(self.assoc_r), _p(ids_r), s))
        check(L.tgn_nbr_lookup(_p(w.roots), w.R, _p(w.R_dev), K, N, _p(self.neighbors), _p(self.e_id),
                               _p(self.t_ring), _p(w.nbr_g), _p(w.ctr_g), _p(w.eid), _p(w.t_e),
                               _p(w.root_off), _p(w.E_dev), _p(self.bitmap), _p(w.lookup_ws), s))
        check(L.tgn_unique_rank(_p(self.bitmap), N, _p(w.n_id), w.Nb, _p(self.assoc), _p(w.Nb_dev), 0, s))
        check(L.tgn_relabel3(_p(w.nbr_g), w.E, _p(w.E_dev), _p(w.nbr_l), _p(w.roots), w.R, _p(w.R_dev),
                             _p(w.ctr_l), _p(ids), ids.numel(), _p(ids_dev), _p(ids_l), _p(self.assoc), s))
        if owner_select and self.owner_compute:
            check(L.tgn_part_select_owned(_p(w.n_id), w.Nb, _p(w.Nb_dev), self.rank, self.world, _p(w.own_n),
                                          _p(w.own_pos), w.own_n.numel(), _p(w.So_dev), s))

    def _assemble_rows(self, w, n_id: Tensor, S: int, S_dev: Optional[Tensor]):
        """Partitioned memory: every rank writes the memory rows it owns (rows of n_id and of the
        other endpoints of their stored events) into the zero-filled staging buffer; one all-reduce
        over NVLink/NVSwitch assembles all of them on every rank.  w.g_rows [2,Nb,D], w.g_lu [Nb]."""
        D = self.D
        if S > w.g_rows.shape[1]:
            raise _cabi.TgnError("row assembly: workspace too small")
        if self.part_exchange == "p2p":
            # barrier 0: every rank's owner-side writes of the previous step / batch are complete;
            # barrier 1: every rank has read its rows before anyone's next scatter can start
            self._symm_mem.barrier(channel=0)
            check(_L().tgn_part_gather_p2p(ctypes.byref(self.store.struct()), _p(n_id), S, _p(S_dev), self._peer_mem,
                                           self._peer_lu, D, self.world, _p(w.g_rows),
                                           w.g_rows.data_ptr() + 4 * w.g_rows.shape[1] * D, _p(w.g_lu), _stream()))
            self._symm_mem.barrier(channel=1)
            return
        check(_L().tgn_part_gather(ctypes.byref(self.store.struct()), _p(n_id), S, _p(S_dev), _p(self.memory),
                                   _p(self.last_update), D, self.rank, self.world, _p(w.g_rows),
                                   w.g_rows.data_ptr() + 4 * w.g_rows.shape[1] * D, _p(w.g_lu), None, _stream()))
        self._all_reduce(w.g_rows)
        self._all_reduce(w.g_lu)

    def _scatter_owned(self, n_id: Tensor, new_mem: Tensor, new_lu: Tensor, memory: Tensor, last_update: Tensor,
                       src_rows: Optional[Tensor] = None, num_dev: Optional[Tensor] = None):
        if self.world == 1:
            ops.memory_scatter(n_id, new_mem, new_lu, memory, last_update, src_rows=src_rows, num_dev=num_dev)
        else:
            check(_L().tgn_memory_scatter_owned(_p(n_id), n_id.numel(), _p(num_dev), _p(new_mem), _p(new_lu), 0, _p(src_rows),
                                                self.D, self.rank, self.world, _p(memory), _p(last_update), _stream()))

    def _memory_fwd(self, w, n_id: Tensor, S: int, S_dev: Optional[Tensor]):
        """TGNMemory._get_updated_memory (memory_module.py:152-178): w.z [S,D], w.lu [S]."""
        self._memory_msgs(w, n_id, S, S_dev)
        self._memory_gru(w, S, S_dev)

    def _memory_msgs(self, w, n_id: Tensor, S: int, S_dev: Optional[Tensor]):
        p, D, L, s = self.p, self.D, _L(), _stream()
        if self.world == 1:
            check(L.tgn_msg_build_ld(ctypes.byref(self.store.struct()), _p(n_id), S, _p(S_dev), ops.AGG_LAST,
                                     _p(self.memory), _p(self.last_update), D, _p(p["time_enc.lin.weight"]),
                                     _p(p["time_enc.lin.bias"]), self.Dt, _p(w.x), self.ldx, _p(w.h), _p(w.sn_m),
                                     _p(w.lu), _p(w.sel_ev), _p(w.sel_dt), s))
        else:
            self._assemble_rows(w, n_id, S, S_dev)
            check(L.tgn_msg_build_gathered(ctypes.byref(self.store.struct()), _p(n_id), S, _p(S_dev), _p(w.g_rows),
                                           w.g_rows.data_ptr() + 4 * w.g_rows.shape[1] * D, _p(w.g_lu), D,
                                           _p(p["time_enc.lin.weight"]), _p(p["time_enc.lin.bias"]), self.Dt,
                                           _p(w.x), self.ldx, _p(w.h), _p(w.sn_m), _p(w.lu), _p(w.sel_ev),
                                           _p(w.sel_dt), _stream()))

    def _memory_fwd_owner(self, w):
        """Owner-side TGNMemory._get_updated_memory (memory_module.py:152-178): this rank builds the messages
        (its own rows are local, the other endpoint's row of each message comes out of its owner's shard over
        NVLink) and runs the GRU for the rows of n_id it owns, then publishes h' / last_update' into every
        rank's row table w.z / w.lu.  Rank barrier B makes the tables complete before anyone reads them, and
        orders all remote reads of this step before the owners' memory write-back."""
        wo, D, L, p, Sb = self.wo, self.D, _L(), self.p, self.So_b
        st = ctypes.byref(self.store.struct())
        g_o = wo.g_rows.data_ptr() + 4 * wo.g_rows.shape[1] * D
        if os.environ.get("TGN_PART_FUSED_GATHER", "1") == "1":
            # the message build reads its rows in place, out of the owners' shards (no staging launch)
            check(L.tgn_msg_build_p2p(st, _p(w.own_n), Sb, _p(w.So_dev), self._peer_mem, self._peer_lu, self.world, D,
                                      _p(p["time_enc.lin.weight"]), _p(p["time_enc.lin.bias"]), self.Dt,
                                      _p(wo.x), self.ldx, _p(wo.h), _p(wo.sn_m), _p(wo.lu), _p(wo.sel_ev),
                                      _p(wo.sel_dt), _stream()))
        else:
            check(L.tgn_part_gather_p2p(st, _p(w.own_n), Sb, _p(w.So_dev), self._peer_mem, self._peer_lu, D,
                                        self.world, _p(wo.g_rows), g_o, _p(wo.g_lu), _stream()))
            check(L.tgn_msg_build_gathered(st, _p(w.own_n), Sb, _p(w.So_dev), _p(wo.g_rows), g_o, _p(wo.g_lu), D,
                                           _p(p["time_enc.lin.weight"]), _p(p["time_enc.lin.bias"]), self.Dt,
                                           _p(wo.x), self.ldx, _p(wo.h), _p(wo.sn_m), _p(wo.lu), _p(wo.sel_ev),
                                           _p(wo.sel_dt), _stream()))
        self._memory_gru(wo, Sb, w.So_dev)
        check(L.tgn_part_publish(_p(wo.z), _p(wo.lu), _p(w.own_pos), Sb, _p(w.So_dev), D, self._peer_z,
                                 self._peer_rowlu, self.world, _stream()))
        self._symm_mem.barrier(channel=1)

    def _memory_gru(self, w, S: int, S_dev: Optional[Tensor]):
        p, D, L, s = self.p, self.D, _L(), _stream()
        if self.fused_gru:
            # both gate GEMMs and the gate math in one launch; gi / gh stay in TMEM
            self._timed("gru_gate_gemm", lambda: check(L.tgn_gru_fused_fwd(
                _p(w.x), self.ldx, self.Dx, _p(w.h), D, self.flat.data_ptr() + 4 * self.off["memory_updater.weight_ih"],
                self.ldx, self.flat.data_ptr() + 4 * self.off["memory_updater.weight_hh"],
                _p(p["memory_updater.bias_ih"]), _p(p["memory_updater.bias_hh"]), S, _p(S_dev), self.prec,
                _p(w.z), _p(w.gates), s)))
            return
        self._timed("gru_gate_gemm", lambda: ops.gemm_batch([
            ops.gemm_desc(w.x, self.flat, w.gi, m=S, n=3 * D, k=self.Dx, lda=self.ldx, ldb=self.ldx, ldc=3 * D,
                          b_off=self.off["memory_updater.weight_ih"], bias=p["memory_updater.bias_ih"], m_dev=S_dev),
            ops.gemm_desc(w.h, self.flat, w.gh, m=S, n=3 * D, k=D, lda=D, ldb=D, ldc=3 * D,
                          b_off=self.off["memory_updater.weight_hh"], bias=p["memory_updater.bias_hh"], m_dev=S_dev),
        ], self.prec))
        check(L.tgn_gru_gates_fwd(_p(w.gi), _p(w.gh), _p(w.h), None, S, _p(S_dev), D, _p(w.z), _p(w.gates), s))

    def _edge_branch(self, w, lu: Tensor, train: bool):
        """edge_attr = [cos(w * rel_t + b), msg] and its projection ee = W_edge edge_attr
        (emb_module.py:26-28 + TransformerConv.lin_edge); independent of the GRU output."""
        p, HC, L, s = self.p, self.HC, _L(), _stream()
        ev = self.events
        if ev is None:
            raise _cabi.TgnError("TGNEngine needs set_events(): edge features and times are events['msg'/'t'][e_id] "
                                 "(data.msg[e_id], epoch_utils.py:224) for every e_id the ring holds")
        check(L.tgn_edge_attr_ld(_p(lu), _p(w.nbr_l), _p(ev["t"]), _p(ev["msg"]) if self.De else None, _p(w.eid),
                                 ev["t"].numel(), w.E, _p(w.E_dev), self.De, self.Dt, _p(p["time_enc.lin.weight"]),
                                 _p(p["time_enc.lin.bias"]), self.lde, _p(w.ea), _p(w.sn_e) if train else None,
                                 _p(w.rel), s))
        ops.gemm_batch([ops.gemm_desc(w.ea, self.flat, w.ee, m=w.E, n=HC, k=self.Din, lda=self.lde, ldb=self.lde,
                                      ldc=HC, b_off=self.off["conv.lin_edge.weight"], m_dev=w.E_dev)], self.prec)

    def _node_proj(self, w, z: Tensor):
        p, D, HC = self.p, self.D, self.HC
        ops.gemm_batch([ops.gemm_desc(z, self.flat, w.proj, m=w.Nb, n=4 * HC, k=D, lda=D, ldb=D, ldc=4 * HC,
                                      b_off=self.off["conv.w_node"], bias=p["conv.b_node"], m_dev=w.Nb_dev)], self.prec)

    def _attention_core(self, w, train: bool):
        check(_L().tgn_attn_core_fwd(_p(w.proj), _p(w.nbr_l), _p(w.root_off), _p(w.ctr_l), w.R, _p(w.R_dev),
                                     self.H, self.C, _p(w.ee), self.dropout if train else 0.0, self.seed,
                                     _p(self.step_dev), self.K, _p(w.emb), _p(w.alpha), _stream()))

    def _attention_fwd(self, w, z: Tensor, lu: Tensor, train: bool):
        """GraphAttentionEmbedding.forward (emb_module.py:25-29) for the centres (= roots).  The node projection
        (a GEMM over the rows) and the edge branch (time encoding + edge projection) are independent: the
        projection runs on the auxiliary stream beside the edge branch."""
        main, aux = torch.cuda.current_stream(), self.aux
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            self._node_proj(w, z)
        self._edge_branch(w, lu, train)
        main.wait_stream(aux)
        self._attention_core(w, train)

    def _update_state(self, w):
        """memory.update_state (memory_module.py:126-138).  Train ordering: memory rows first (they are
        the rows the forward just produced: same store, same weights), then the store."""
        B = self.B
        self._scatter_owned(self.in_ids3[:2 * B], w.z, w.lu, self.memory, self.last_update,
                            src_rows=w.ids_l[:2 * B])
        self.store.update(self.in_ids3[:B], self.in_ids3[B:2 * B], self.in_t_i64, self.in_msg,
                          base_dev=self.log_base_dev)

    def _ring_insert(self):
        """neighbor_loader.insert (epoch_utils.py:300).  It needs the batch's ids and times only, and the
        step's compute never reads the ring (the batch was sampled before), so it does not wait for it."""
        B = self.B
        check(_L().tgn_nbr_insert(_p(self.in_ids3), self.in_ids3[B:].data_ptr(), _p(self.in_t_f32), B, 0,
                                  _p(self.cur_e_id_dev), self.K, self.N, _p(self.neighbors), _p(self.e_id),
                                  _p(self.t_ring), _stream()))

    def _decoder_gemm_path(self, w, s):
        """Decoder forward/backward on the batched GEMM for hidden sizes whose weights do not fit
        the fused kernel's shared memory."""
        p, B, D, HC, L = self.p, self.B, self.D, self.HC, _L()
        off, fg, fl = self.off, self.flat_grad, self.flat
        gptr = lambda name: fg.data_ptr() + 4 * off[name]
        check(L.tgn_gather_rows(_p(w.emb), _p(w.ids_l), 3 * B, None, HC, _p(w.zcat), s))
        ops.gemm_batch([
            ops.gemm_desc(w.zcat, fl, w.hcat, m=B, n=D, k=D, lda=D, ldb=D, ldc=D, b_off=off["lin_src.weight"],
                          bias=p["lin_src.bias"]),
            ops.gemm_desc(w.zcat, fl, w.hcat, m=2 * B, n=D, k=D, lda=D, ldb=D, ldc=D, a_off=B * D, c_off=B * D,
                          b_off=off["lin_dst.weight"], bias=p["lin_dst.bias"]),
        ], self.prec)
        check(L.tgn_dec_loss(_p(w.hcat), w.hcat.data_ptr() + 4 * B * D, _p(p["lin_final.weight"]),
                             _p(p["lin_final.bias"]), B, D, _p(self.loss_acc), _p(w.logits),
                             w.dhcat.data_ptr() + 4 * B * D, _p(w.dhcat), gptr("lin_final.weight"),
                             gptr("lin_final.bias"), gptr("lin_src.bias"), gptr("lin_dst.bias"), s))
        ops.gemm_batch([
            # dW_src += dhs^T z_src ; dW_dst += dh^T [z_dst; z_neg]
            ops.gemm_desc(w.dhcat, w.zcat, fg, m=D, n=D, k=B, lda=D, ldb=D, ldc=D, trans_a=True, trans_b=True,
                          mode=1, c_off=off["lin_src.weight"]),
            ops.gemm_desc(w.dhcat, w.zcat, fg, m=D, n=D, k=2 * B, lda=D, ldb=D, ldc=D, trans_a=True, trans_b=True,
                          mode=1, a_off=B * D, b_off=B * D, c_off=off["lin_dst.weight"]),
            # d z_src = dhs W_src ; d [z_dst; z_neg] = dh W_dst
            ops.gemm_desc(w.dhcat, fl, w.dzcat, m=B, n=D, k=D, lda=D, ldb=D, ldc=D, trans_b=True,
                          b_off=off["lin_src.weight"]),
            ops.gemm_desc(w.dhcat, fl, w.dzcat, m=2 * B, n=D, k=D, lda=D, ldb=D, ldc=D, trans_b=True,
                          a_off=B * D, c_off=B * D, b_off=off["lin_dst.weight"]),
        ], self.prec)
        check(L.tgn_scatter_add_rows(_p(w.dzcat), _p(w.ids_l), 3 * B, None, HC, _p(self.d_emb), s))

    def _train_body(self, from_device: bool, pipelined: bool):
        """One step on slot `cur`.  pipelined: slots[cur] is already sampled; the forked stream loads
        (from_device) and samples the next batch into the other slot after the state update."""
        self._bind(self.cur)
        w, p, B, D, HC, L = self.w, self.p, self.B, self.D, self.HC, _L()
        off, fg, fl = self.off, self.flat_grad, self.flat
        main, side, aux, upd = torch.cuda.current_stream(), self.side, self.aux, self.upd
        own = self.owner_compute
        # (owner-side compute: no barrier here.  The previous step's barrier C, in front of its optimiser, came
        # after every rank's memory write-back, so the peers' shards are current; the gradient blobs rotate.)
        if own and os.environ.get("TGN_PART_BARRIER_A") == "1":
            self._symm_mem.barrier(channel=0)
        if not self.fused_zero_grad:
            self.zero_blob.zero_()
        if not pipelined:
            if from_device:
                self.stage_batch_from_device()
            self._sample(w, w.in_ids3, w.ids_l, owner_select=True, ids_r=w.ids_r)
        # ---- forked stream 1, from the start of the step: the batch's events enter the ring, then the
        # NEXT batch is loaded and sampled (single-CTA, latency-bound kernels: ~100 us of them hide
        # behind the whole step instead of trailing the GRU)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if not from_device:     # host-staged batch: int64 stamps -> the ring's float32 (neighbor_loader.py:21)
                self.in_t_f32.copy_(self.in_t_i64)
            self._ring_insert()
            if pipelined:
                nxt = self.slots[self._next_slot()]
                if from_device:
                    self.stage_batch_from_device(self._next_slot())
                self._sample(nxt, nxt.in_ids3, nxt.ids_l, owner_select=True, ids_r=nxt.ids_r)
        # the attention backward accumulates into a zero-filled d_proj (10 MB): cleared here, beside
        # msg_build, instead of in front of the backward on the dependent chain
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            w.d_proj.zero_()
            if self.dz_split > 1:
                w.d_z.zero_()
        if own:
            self._memory_fwd_owner(w)
        else:
            self._timed("msg_build", lambda: self._memory_msgs(w, w.n_id, w.Nb, w.Nb_dev))
        # ---- edge branch of the attention on its own stream: needs last_update / edges only
        aux.wait_stream(main)
        with torch.cuda.stream(aux):
            self._edge_branch(w, w.lu, True)
        if not own:
            self._memory_gru(w, w.Nb, w.Nb_dev)
        # ---- forked stream 2: memory / store update (needs z / last_update of the forward only)
        upd.wait_stream(main)
        with torch.cuda.stream(upd):
            self._update_state(w)
        self._node_proj(w, w.z)
        main.wait_stream(aux)
        if not self.fused_attn_dec:
            self._attention_core(w, True)
        s = _stream()
        # ---- decoder + loss + decoder backward (decoder.py:24-27; BCEWithLogits, pyg-mem-tgn.py:51)
        gptr = lambda name: fg.data_ptr() + 4 * off[name]
        if self.fused_attn_dec:
            # the attention forward of every (event, endpoint) occurrence runs inside the decoder launch: one link
            # of the dependent chain less, no embedding table
            check(L.tgn_dec_attn_fused(_p(w.proj), _p(w.nbr_l), _p(w.root_off), _p(w.ctr_l), w.R, self.H, self.C,
                                       _p(w.ee), self.dropout, self.seed, _p(self.step_dev), self.K, _p(w.alpha),
                                       _p(w.ids_r), _p(w.ids_l), B, _p(p["lin_src.weight"]), _p(p["lin_src.bias"]),
                                       _p(p["lin_dst.weight"]), _p(p["lin_dst.bias"]), _p(p["lin_final.weight"]),
                                       _p(p["lin_final.bias"]), _p(self.loss_acc), _p(w.logits), _p(self.d_emb),
                                       gptr("lin_src.bias"), gptr("lin_dst.bias"), gptr("lin_final.weight"),
                                       gptr("lin_final.bias"), _p(w.zcat), _p(w.dhcat), s))
        elif self.fused_decoder:
            check(L.tgn_dec_fused(_p(w.emb), _p(w.ids_l), B, D, _p(p["lin_src.weight"]), _p(p["lin_src.bias"]),
                                  _p(p["lin_dst.weight"]), _p(p["lin_dst.bias"]), _p(p["lin_final.weight"]),
                                  _p(p["lin_final.bias"]), _p(self.loss_acc), _p(w.logits), _p(self.d_emb),
                                  None, gptr("lin_src.bias"), None,
                                  gptr("lin_dst.bias"), gptr("lin_final.weight"), gptr("lin_final.bias"),
                                  _p(w.zcat), _p(w.dhcat), s))
        if self.fused_decoder:
            # (one CTA per [D,D] gradient walks the whole batch: beyond ~256 events the reduction is split)
            split_d = max(1, min(16, B // 128))
            dec_wgrad = [  # dW_src += g_src^T z_src ; dW_dst += [g_pos; g_neg]^T [z_dst; z_neg]  (aux stream, below)
                ops.gemm_desc(w.dhcat, w.zcat, fg, m=D, n=D, k=B, lda=D, ldb=D, ldc=D, trans_a=True, trans_b=True,
                              mode=2, split_k=split_d, c_off=off["lin_src.weight"]),
                ops.gemm_desc(w.dhcat, w.zcat, fg, m=D, n=D, k=2 * B, lda=D, ldb=D, ldc=D, trans_a=True,
                              trans_b=True, mode=2, split_k=split_d, a_off=B * D, b_off=B * D,
                              c_off=off["lin_dst.weight"]),
            ]
        else:
            self._decoder_gemm_path(w, s)
            dec_wgrad = []
        if self.track_metrics:      # off the chain: the update stream has been idle since the store update
            upd.wait_stream(main)
            with torch.cuda.stream(upd):
                check(L.tgn_ap_auc_accum(_p(w.logits), B, B, _p(self.metric_acc), _stream()))
        # ---- attention backward
        check(L.tgn_attn_core_bwd(_p(w.proj), _p(w.nbr_l), _p(w.root_off), _p(w.ctr_l), w.R, _p(w.R_dev),
                                  self.H, self.C, _p(w.ee), _p(w.alpha), _p(self.d_emb), self.dropout, self.seed,
                                  _p(self.step_dev), self.K, 0, _p(w.d_proj), _p(w.d_ee), s))
        split_e = max(1, min(16, w.E // 512))
        split_n = max(1, min(16, w.Nb // 512))
        # Only d z = d_proj W_node feeds the GRU backward; it gets its own small launch on the main
        # stream.  The weight gradients of the attention (and everything that hangs off them) run
        # beside it on the auxiliary stream: the two launches together fit the 148 SMs.
        aux.wait_stream(main)       # fork point: both launches depend on the attention backward only
        # (38 live row tiles on 148 SMs: the 13 k-blocks of the reduction are split three ways, the partial
        # products are added into the d_z cleared beside msg_build)
        ops.gemm_batch([ops.gemm_desc(w.d_proj, fl, w.d_z, m=w.Nb, n=D, k=4 * HC, lda=4 * HC, ldb=D, ldc=D,
                                      trans_b=True, b_off=off["conv.w_node"], m_dev=w.Nb_dev,
                                      mode=2, split_k=self.dz_split)], self.prec)
        with torch.cuda.stream(aux):
            g = [  # dW_edge += d_ee^T edge_attr ; dW_node += d_proj^T z
                ops.gemm_desc(w.d_ee, w.ea, fg, m=HC, n=self.Din, k=w.E, lda=HC, ldb=self.lde, ldc=self.lde,
                              trans_a=True, trans_b=True, mode=2, split_k=split_e, k_dev=w.E_dev,
                              c_off=off["conv.lin_edge.weight"]),
                ops.gemm_desc(w.d_proj, w.z, fg, m=4 * HC, n=D, k=w.Nb, lda=4 * HC, ldb=D, ldc=D, trans_a=True,
                              trans_b=True, mode=2, split_k=split_n, k_dev=w.Nb_dev, c_off=off["conv.w_node"]),
            ]
            if self.Dt:  # d edge_attr[:, :Dt] = d_ee W_edge[:, :Dt]  (time-encoder gradient)
                g.append(ops.gemm_desc(w.d_ee, fl, w.d_eat, m=w.E, n=self.Dt, k=HC, lda=HC, ldb=self.lde,
                                       ldc=self.Dt, trans_b=True, b_off=off["conv.lin_edge.weight"], m_dev=w.E_dev))
            ops.gemm_batch(g, self.prec)
            if dec_wgrad:
                ops.gemm_batch(dec_wgrad, self.prec)
            # bias gradient of the node projection and the attention-side TimeEncoder gradient
            ops.colsum(w.d_proj, w.Nb, 4 * HC, 4 * HC, fg[off["conv.b_node"]:off["conv.b_node"] + 4 * HC], True,
                       rows_dev=w.Nb_dev)
            if self.Dt:
                check(L.tgn_time_bwd_sin(_p(w.rel), None, w.E, _p(w.E_dev), _p(w.sn_e), self.Dt, _p(w.d_eat), self.Dt,
                                         gptr("time_enc.lin.weight"), gptr("time_enc.lin.bias"), _stream()))
        # ---- GRU backward (torch.nn.GRUCell, memory_module.py:72,172).  Owner-side compute: only on the rows
        # this rank owns (gm = their workspace, rows_dev = how many), gradients into the PARTIAL buffer that the
        # optimiser sums over the ranks
        gm, rows_dev, fgm, rows_b = w, w.Nb_dev, fg, w.Nb
        if own:
            gm, rows_dev, fgm, rows_b = self.wo, w.So_dev, self.flat_grad_part, self.So_b
            check(L.tgn_gather_rows(_p(w.d_z), _p(w.own_pos), rows_b, _p(rows_dev), D, _p(gm.d_z), s))
        gptr_m = lambda name: fgm.data_ptr() + 4 * off[name]
        check(L.tgn_gru_gates_bwd_bias(_p(gm.d_z), _p(gm.gates), _p(gm.h), rows_b, _p(rows_dev), D, _p(gm.d_gi),
                                       _p(gm.d_gh), gptr_m("memory_updater.bias_ih"), gptr_m("memory_updater.bias_hh"), s))
        # One launch, one wave: the GEMM CTAs hold ~193 KB of shared memory (one per SM), so the split-K
        # factor of the two weight gradients is chosen to leave SMs for the row tiles of d x -- launched
        # separately (or with a larger split) d x simply queues behind the weight-gradient CTAs.
        c0 = (2 * D + self.De) & ~3
        tiles_dx = (rows_b + 127) // 128 if self.Dt else 0
        tiles_w = ((3 * D + 127) // 128) * ((self.Dx + 127) // 128) + ((3 * D + 127) // 128) * ((D + 127) // 128)
        split_g = max(1, min(split_n, (148 - tiles_dx) // tiles_w))
        if split_g < 4:
            # larger batches (coin B=600, wiki B=2000): the bound-sized d x tiles alone exceed one wave, and a
            # weight-gradient CTA that walks ALL rows (hundreds of k-blocks) becomes the step's longest launch
            # (196 us measured at B=600).  Give every split <= ~24 k-blocks and accept a second wave.
            split_g = max(split_g, min(16, (rows_b // 32 + 23) // 24))
        g = [
            ops.gemm_desc(gm.d_gi, gm.x, fgm, m=3 * D, n=self.Dx, k=rows_b, lda=3 * D, ldb=self.ldx, ldc=self.ldx,
                          trans_a=True, trans_b=True, mode=2, split_k=split_g, k_dev=rows_dev,
                          c_off=off["memory_updater.weight_ih"]),
            ops.gemm_desc(gm.d_gh, gm.h, fgm, m=3 * D, n=D, k=rows_b, lda=3 * D, ldb=D, ldc=D, trans_a=True,
                          trans_b=True, mode=2, split_k=split_g, k_dev=rows_dev,
                          c_off=off["memory_updater.weight_hh"]),
        ]
        if self.Dt:
            # d x = d_gi W_ih: only the time-encoding columns [2D+De, Dx) are consumed (memory-side
            # TimeEncoder gradient), so only those are computed, from the 16-byte aligned column below them
            g.append(ops.gemm_desc(gm.d_gi, fl, gm.d_x, m=rows_b, n=self.Dx - c0, k=3 * D, lda=3 * D, ldb=self.ldx,
                                   ldc=self.ldx, trans_b=True, b_off=off["memory_updater.weight_ih"] + c0, c_off=c0,
                                   m_dev=rows_dev))
        ops.gemm_batch(g, self.prec)
        if self.Dt:
            check(L.tgn_time_bwd_sin(_p(gm.sel_dt), _p(gm.sel_ev), rows_b, _p(rows_dev), _p(gm.sn_m), self.Dt,
                                     gm.d_x.data_ptr() + 4 * (2 * D + self.De), self.ldx,
                                     gptr_m("time_enc.lin.weight"), gptr_m("time_enc.lin.bias"), s))
        main.wait_stream(aux)
        main.wait_stream(side)
        main.wait_stream(upd)
        if own:
            # rank barrier C: every rank's gradients are final; then one optimiser launch that reads rank 0's
            # replicated part and sums the partial parts of all ranks out of peer memory (no NCCL in the graph)
            self._symm_mem.barrier(channel=2)
            check(L.tgn_adam_finish_peers(_p(self.flat), self._grad_rep0, self._peer_grad_part, self.world,
                                          off["conv.w_node"], _p(self.exp_avg), _p(self.exp_avg_sq), self.n_param, self.lr, 0.9, 0.999,
                                          1e-8, _p(self.adam_step_dev), _p(self.step_dev), _p(self.loss_acc),
                                          self.loss_slots.data_ptr() + 4 * self.cur, _p(self.done_ctr), _stream()))
            return
        if self.world > 1:   # replicated compute: average the gradients so the weight replicas stay bit-identical
            self._all_reduce(self.flat_grad)
            self.flat_grad.mul_(1.0 / self.world)
        check(L.tgn_adam_finish(_p(self.flat), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                                self.n_param, self.lr, 0.9, 0.999, 1e-8, _p(self.adam_step_dev), _p(self.step_dev),
                                _p(self.loss_acc), self.loss_slots.data_ptr() + 4 * self.cur, _p(self.done_ctr), int(self.fused_zero_grad),
                                _p(self.d_emb), self.d_emb.numel(), _stream()))

    def _run(self, key: tuple, body, capture: bool = True):
        if not self.use_graph:
            body()
            return
        g = self._graphs.get(key)
        if g is None and not capture:      # caller does not want a capture (sync + ~10 ms) at this point
            body()
            return
        if g is None:
            # warm-up: the first calls of every configuration run eagerly
            cnt = self._graphs.get(("warm",) + key, 0)
            if cnt < 3:
                self._graphs[("warm",) + key] = cnt + 1
                body()
                return
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            self._graph_store_gen = self.store.generation
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
        g.replay()

    def train_step(self, from_device: bool = True, lookahead: bool = False, _capture: bool = True):
        """One training batch; returns the (device) loss of that batch.

        from_device=True : batches are sliced out of the resident event arrays (set_events).  The
                           step is software-pipelined: while batch i trains, the forked stream loads
                           and samples batch i+1.
        from_device=False: the caller staged the batch with stage_batch() / stage_packed().  With
                           lookahead=True the caller has ALSO staged the following batch
                           (stage_*(…, ahead=True)); it is sampled during this step, and the next call
                           trains on it (prefetching data-loader pattern, used by bench.py's e2e arm).
        """
        mode = "device" if from_device else "host"
        pipelined = from_device or lookahead
        self._reserve(self.B)
        if pipelined:
            if self._primed != mode:           # first pipelined step: sample the current batch now
                self._unprime()
                self._bind(self.cur)
                if from_device:
                    self.stage_batch_from_device()
                self._sample(self.slots[self.cur], self.slots[self.cur].in_ids3, self.slots[self.cur].ids_l,
                             owner_select=True, ids_r=self.slots[self.cur].ids_r)
                self._primed = mode
        else:
            self._unprime()
        main = torch.cuda.current_stream()
        nxt = self._next_slot()
        if pipelined and self._slot_async[nxt]:          # the batch staged ahead on the copy stream has landed
            main.wait_event(self._slot_ready[nxt])
            self._slot_async[nxt] = False
        self._run(("train", from_device, pipelined, self.cur), lambda: self._train_body(from_device, pipelined),
                  capture=_capture)
        self._slot_free[self.cur].record(main)
        self.loss = self._loss_views[self.cur]
        if pipelined:
            self.cur = nxt
        self._advance(self.B)
        return self.loss

    def train_steps(self, n: int):
        """n consecutive training batches from the resident event arrays (set_events).  `group_size` (three)
        consecutive steps are captured as ONE graph: one launch per three steps, no launch gap at the two
        inner step boundaries.  The remainder runs step by step.  Returns the
        (device) loss of the last batch."""
        done = 0
        if n > 0 and self._primed != "device":
            self.train_step(from_device=True)
            done = 1
        G = self.group_size
        while self.use_graph and n - done >= G:
            self._reserve(G * self.B)
            self._run(("train_multi", self.cur), lambda: self._multi_body(True))
            self.cur = (self.cur + G) % self.nslots
            done += G
            self._advance(G * self.B)
            self.loss = self._loss_views[(self.cur - 1) % self.nslots]
        while done < n:          # remainder: replays the single-step graph if it exists, else runs eagerly
            self.train_step(from_device=True, _capture=False)
            done += 1
        return self.loss

    def _multi_body(self, from_device: bool):
        """`group_size` consecutive steps; leaves self.cur where it found it (the caller advances it, so that a
        replay -- which does not run this body -- and a capture / eager run end in the same state)."""
        c0 = self.cur
        for _ in range(self.group_size):
            self._train_body(from_device, True)
            self.cur = self._next_slot()
        self.cur = c0

    # ---- grouped end-to-end feeding: three batches per H2D copy / graph launch / loss read-back
    def group_nbytes(self) -> int:
        return self.group_size * self._packed

    def pack_host_group(self, out: Tensor, batches) -> Tensor:
        """Fills `out` (pinned uint8 [group_nbytes()]) from `group_size` tuples (src, dst, neg, t, msg)."""
        for i, b in enumerate(batches):
            self.pack_host_batch(out[i * self._packed: i * self._packed + self.packed_nbytes()], *b)
        return out

    def group_views(self, buf: Tensor):
        """numpy views of a pinned group buffer (uint8 [group_nbytes()]), one tuple (src, dst, neg, t, msg) per batch
        slot: a host loader fills a group with plain array assignments -- ~20 us per group of four batches instead
        of ~150 us of small tensor ops through pack_host_group (same layout)."""
        if buf.numel() != self.group_nbytes() or buf.dtype != torch.uint8 or not buf.is_contiguous() or buf.is_cuda:
            raise _cabi.TgnError("group_views: buf must be a contiguous host uint8 tensor of group_nbytes() bytes")
        import numpy as np
        B, De, raw = self.B, self.De, buf.numpy()
        views = []
        for i in range(self.group_size):
            o = raw[i * self._packed: i * self._packed + self.packed_nbytes()]
            ids = o[:32 * B].view(np.int64)
            msg = o[32 * B:].view(np.float32).reshape(B, max(De, 1))
            views.append((ids[:B], ids[B:2 * B], ids[2 * B:3 * B], ids[3 * B:], msg))
        return views

    def stage_group(self, buf: Tensor, ahead: bool = True):
        """Stages the NEXT group (ahead=True: the three batches after the group train_group_logged() is about
        to train; the copy runs on the copy stream while the previous group is still executing) or the
        current one (ahead=False, before the first call)."""
        if buf.numel() != self.group_nbytes() or buf.dtype != torch.uint8 or not buf.is_contiguous():
            raise _cabi.TgnError("stage_group: buf must be a contiguous uint8 tensor of group_nbytes() bytes")
        if self.cur % self.group_size:
            raise _cabi.TgnError("stage_group: the slot cursor is not on a group boundary")
        g0 = (self.cur + self.group_size) % self.nslots if ahead else self.cur
        dst = self._in_all.data_ptr() + g0 * self._packed
        if not ahead:
            check(_L().tgn_memcpy_async(dst, buf.data_ptr(), buf.numel(), torch.cuda.current_stream().cuda_stream))
            return
        cs = self.copy_stream
        cs.wait_event(self._slot_free[g0])          # the group that last used these slots has finished
        check(_L().tgn_memcpy_async(dst, buf.data_ptr(), buf.numel(), cs.cuda_stream))
        self._slot_ready[g0].record(cs)
        self._slot_async[g0] = True

    def train_group_logged(self):
        """Trains the `group_size` staged batches (one captured graph) with the NEXT group staged ahead
        (stage_group); returns the list of the previous group's losses (None on the first call)."""
        G, main, ls = self.group_size, torch.cuda.current_stream(), self.loss_stream
        if not hasattr(self, "_gl_pin"):
            self._gl_pin = torch.zeros(2 * G, dtype=torch.float32).pin_memory()
            self._gl_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._gl_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._gl_n = 0
        if self.cur % G:
            raise _cabi.TgnError("train_group_logged: the slot cursor is not on a group boundary")
        if self._primed != "host":                 # first call: sample the group's first batch now
            self._unprime()
            self._bind(self.cur)
            self._sample(self.slots[self.cur], self.slots[self.cur].in_ids3, self.slots[self.cur].ids_l,
                         owner_select=True, ids_r=self.slots[self.cur].ids_r)
            self._primed = "host"
        i = self._gl_n & 1
        main.wait_event(self._gl_ev[i])
        nxt = (self.cur + G) % self.nslots
        if self._slot_async[nxt]:
            main.wait_event(self._slot_ready[nxt])
            self._slot_async[nxt] = False
        c0 = self.cur
        self._reserve(G * self.B)
        self._run(("train_group_host", c0), lambda: self._multi_body(False))
        for k in range(G):
            self._slot_free[c0 + k].record(main)
        self.cur = nxt
        self._advance(G * self.B)
        self.loss = self._loss_views[(nxt - 1) % self.nslots]
        self._gl_done[i].record(main)
        ls.wait_event(self._gl_done[i])
        check(_L().tgn_memcpy_async(self._gl_pin.data_ptr() + 4 * G * i, self.loss_slots.data_ptr() + 4 * c0, 4 * G,
                                    ls.cuda_stream))
        self._gl_ev[i].record(ls)
        self._gl_n += 1
        if self._gl_n < 2:
            return None
        self._gl_ev[i ^ 1].synchronize()
        return self._gl_pin[G * (i ^ 1): G * (i ^ 1) + G].tolist()

    def flush_group_losses(self):
        if not getattr(self, "_gl_n", 0):
            return None
        i, G = (self._gl_n - 1) & 1, self.group_size
        self._gl_ev[i].synchronize()
        self._gl_n = 0
        return self._gl_pin[G * i: G * i + G].tolist()

    def train_step_logged(self, **kw) -> Optional[float]:
        """train_step() with a pipelined loss read-back for logging loops: the device->host copy of
        THIS step's loss is enqueued into pinned memory behind the step, and the value returned is the
        loss of the PREVIOUS call (None on the first one) -- the host never waits for the step it has
        just launched, so the GPU does not idle while the next batch is staged.  flush_loss() returns
        the last step's loss."""
        if not hasattr(self, "_loss_pin"):
            self._loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._loss_pin_ptr = self._loss_pin.data_ptr()
            self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._loss_n = 0
        i = self._loss_n & 1
        main, ls = torch.cuda.current_stream(), self.loss_stream
        main.wait_event(self._loss_ev[i])      # the read of two steps ago (same pinned word) is done: long since
        loss = self.train_step(**kw)
        self._loss_done[i].record(main)
        ls.wait_event(self._loss_done[i])      # the copy runs beside the next step, not between two steps
        check(_L().tgn_memcpy_async(self._loss_pin_ptr + 4 * i, loss.data_ptr(), 4, ls.cuda_stream))
        self._loss_ev[i].record(ls)
        self._loss_n += 1
        if self._loss_n < 2:
            return None
        self._loss_ev[i ^ 1].synchronize()
        return float(self._loss_pin[i ^ 1])

    def flush_loss(self) -> Optional[float]:
        if not getattr(self, "_loss_n", 0):
            return None
        i = (self._loss_n - 1) & 1
        self._loss_ev[i].synchronize()
        self._loss_n = 0        # the log is drained: the next train_step_logged() starts a fresh lag (returns None)
        return float(self._loss_pin[i])

    # ------------------------------------------------------------------ evaluation
    def _eval_ctx(self, B: int, Q: int) -> SimpleNamespace:
        """Static buffers of the evaluation step for one (batch, negatives) shape."""
        ctxs = self.__dict__.setdefault("_ectx", {})
        c = ctxs.get((B, Q))
        if c is None:
            dev, D = self.dev, self.D
            n_ids = B * (2 + Q)
            R, E, Nb = self._bounds(B, roots=min(self.N, n_ids))
            c = SimpleNamespace(B=B, Q=Q)
            c.w = self._alloc_work(R, E, Nb, B, train=False)
            c.w.hs, c.w.hd = torch.empty((Nb, D), device=dev), torch.empty((Nb, D), device=dev)
            c.mw = self._alloc_work(1, 1, 2 * B, 1, train=False)         # memory update of the <= 2B touched nodes
            c.ids = torch.zeros(n_ids, dtype=torch.long, device=dev)      # [src | dst | neg (row-major)]
            c.ids_l = torch.zeros(n_ids, dtype=torch.long, device=dev)
            c.t_i = torch.zeros(B, dtype=torch.long, device=dev)
            c.t_f = torch.zeros(B, device=dev)
            c.msg = torch.zeros((B, max(self.De, 1)), device=dev)
            c.pos, c.negs = torch.zeros(B, device=dev), torch.zeros((B, max(Q, 1)), device=dev)
            c.gt = torch.zeros(B, dtype=torch.int32, device=dev)
            c.ge = torch.zeros(B, dtype=torch.int32, device=dev)
            c.n_upd = torch.zeros(2 * B, dtype=torch.long, device=dev)
            c.U_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            ctxs[(B, Q)] = c
        return c

    def _eval_body(self, c: SimpleNamespace):
        dev, N, D, L = self.dev, self.N, self.D, _L()
        B, Q, w, mw = c.B, c.Q, c.w, c.mw
        self._sample(w, c.ids, c.ids_l)
        s = _stream()
        if self.world == 1:
            check(L.tgn_gather_rows(_p(self.memory), _p(w.n_id), w.Nb, _p(w.Nb_dev), D, _p(w.z), s))
            check(L.tgn_relabel(_p(w.n_id), w.Nb, _p(w.Nb_dev), _p(self.last_update), _p(w.lu), s))
            z_in, lu_in = w.z, w.lu
        else:
            self._assemble_rows(w, w.n_id, w.Nb, w.Nb_dev)
            z_in, lu_in = w.g_rows[0], w.g_lu
        self._attention_fwd(w, z_in, lu_in, False)
        p, off = self.p, self.off
        ops.gemm_batch([
            ops.gemm_desc(w.emb, self.flat, w.hs, m=w.Nb, n=D, k=D, lda=D, ldb=D, ldc=D, b_off=off["lin_src.weight"],
                          bias=p["lin_src.bias"], m_dev=w.Nb_dev),
            ops.gemm_desc(w.emb, self.flat, w.hd, m=w.Nb, n=D, k=D, lda=D, ldb=D, ldc=D, b_off=off["lin_dst.weight"],
                          bias=p["lin_dst.bias"], m_dev=w.Nb_dev),
        ], self.prec)
        check(L.tgn_score_negs(_p(w.hs), _p(w.hd), _p(c.ids_l), c.ids_l.data_ptr() + 8 * B, c.ids_l.data_ptr() + 16 * B,
                               B, Q, D, _p(p["lin_final.weight"]), _p(p["lin_final.bias"]), _p(c.pos), _p(c.negs),
                               _p(c.gt), _p(c.ge), _stream()))
        # eval ordering of update_state: store first, then memory (memory_module.py:135-138)
        src, dst = c.ids[:B], c.ids[B:2 * B]
        self.store.update(src, dst, c.t_i, c.msg, base_dev=self.log_base_dev)
        check(L.tgn_unique_mark_rank(_p(c.ids), 2 * B, _p(self.bitmap), N, _p(c.n_upd), 2 * B, None, _p(c.U_dev), 0,
                                     _stream()))
        self._memory_fwd(mw, c.n_upd, 2 * B, c.U_dev)
        self._scatter_owned(c.n_upd, mw.z, mw.lu, self.memory, self.last_update, num_dev=c.U_dev)
        check(L.tgn_nbr_insert(_p(src), _p(dst), _p(c.t_f), B, 0, _p(self.cur_e_id_dev), self.K, N,
                               _p(self.neighbors), _p(self.e_id), _p(self.t_ring), _stream()))

    @torch.no_grad()
    def eval_batch(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor,
                   want_neg_scores: bool = True):
        """test() body for one batch (epoch_utils.py:28-157): embeddings of the batch's nodes,
        scores of the positives and of the [B,Q] negatives, then eval-mode update_state (store
        first, then memory) and insert.  Returns (pos[B], neg[B,Q] or None, gt[B], ge[B]) with
        gt/ge = number of negatives scoring above / not below the positive (TGB MRR counts).
        `neg` may be any column shard of the full negative matrix (data-parallel evaluation):
        the state update does not depend on it.  All buffers are static per (B, Q) and the step is
        a captured CUDA graph; the returned tensors are overwritten by the next call of that shape."""
        self._unprime()
        B, Q = neg.shape
        c = self._eval_ctx(B, Q)
        self._reserve(B)
        c.ids[:B].copy_(src, non_blocking=True)
        c.ids[B:2 * B].copy_(dst, non_blocking=True)
        if Q:
            c.ids[2 * B:].view(B, Q).copy_(neg, non_blocking=True)
        c.t_i.copy_(t, non_blocking=True)
        c.t_f.copy_(c.t_i)
        if self.De:
            c.msg.copy_(msg, non_blocking=True)
        self._run(("eval", B, Q), lambda: self._eval_body(c))
        self._advance(B)
        return c.pos, (c.negs[:, :Q] if want_neg_scores else None), c.gt, c.ge

    # ---- data-parallel evaluation with the EMBEDDING sharded as well (epoch_utils.py:74-113)
    def _eval_ctx_dp(self, B: int, Q: int, rank: int, P: int, group=None) -> SimpleNamespace:
        ctxs = self.__dict__.setdefault("_ectx", {})
        c = ctxs.get(("dp", B, Q, rank, P))
        if c is not None:
            return c
        dev, D, N = self.dev, self.D, self.N
        n_ids = B * (2 + Q)
        c = SimpleNamespace(B=B, Q=Q, rank=rank, P=P, n_ids=n_ids)
        # the candidates of a batch cover (nearly) every node (flight: 200 x 1001 ids over 18k nodes): every
        # node is a root, no unique / relabel of the 200k ids is needed, rank r embeds the nodes r, r+P, ...
        c.dense = n_ids >= 4 * N
        c.Rall = N if c.dense else min(N, n_ids)
        c.Rm = (c.Rall + P - 1) // P
        R, E, Nb = self._bounds(B, roots=c.Rm)
        if c.dense:
            Nb = N      # dense: the row of node n in every per-node table of the step IS n (no unique / relabel)
        c.w = self._alloc_work(R, E, Nb, B, train=False)
        c.mw = self._alloc_work(1, 1, 2 * B, 1, train=False)
        i64 = lambda *s_: torch.zeros(s_, dtype=torch.long, device=dev)
        i32 = lambda *s_: torch.zeros(s_, dtype=torch.int32, device=dev)
        c.ids, c.ids_g = i64(n_ids), i64(n_ids)
        c.t_i, c.t_f = i64(B), torch.zeros(B, device=dev)
        c.msg = torch.zeros((B, max(self.De, 1)), device=dev)
        c.pos, c.cnt = torch.zeros(B, device=dev), i32(2, B)
        c.gt, c.ge = c.cnt[0], c.cnt[1]              # one tensor: the two counts travel in ONE all-reduce
        c.rr = torch.zeros(B, device=dev)
        c.n_upd, c.U_dev = i64(2 * B), i32(1)
        c.roots_all, c.Rall_dev = i64(c.Rall), i32(1)
        c.my_roots, c.Rm_dev, c.my_l = torch.full((c.Rm,), -1, dtype=torch.long, device=dev), i32(1), i64(c.Rm)
        if c.dense:
            mine = torch.arange(rank, N, P, device=dev)
            c.my_roots[:mine.numel()] = mine
            c.Rm_dev.fill_(mine.numel())
            # identity relabelling: centres are addressed by their global ids, the neighbour lookup's global ids
            # are used as rows directly, every node has a row
            c.w.ctr_l, c.w.R_dev, c.my_l = c.my_roots, c.Rm_dev, c.my_roots
            c.w.Nb_dev.fill_(N)
        c.e_r = torch.zeros((c.Rm, D), device=dev)
        c.send = torch.zeros((2, c.Rm, D), device=dev)               # [lin_src(emb) | lin_dst(emb)] of my centres
        # Exchange between the ranks.  "peer" (default): the gathered table and the table of per-rank counts are
        # torch symmetric memory; every rank WRITES its block into every rank's table (tgn_peer_bcast: 128-bit
        # stores over NVLink / NVSwitch) and a device-side rank barrier separates writers from readers -- no
        # collective call in the captured step.  "nccl": all_gather_into_tensor + all_reduce (round-2 first design).
        c.peer = P > 1 and os.environ.get("TGN_EVAL_EXCHANGE", "peer") == "peer" and (c.Rm * D) % 2 == 0 and B % 2 == 0
        if c.peer:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            c.recv = symm_mem.empty((P, 2, c.Rm, D), dtype=torch.float32, device=dev)
            c.cnt_table = symm_mem.empty((P, 2, B), dtype=torch.int32, device=dev)
            c.recv.zero_()
            c.cnt_table.zero_()
            c.h_recv, c.h_cnt = symm_mem.rendezvous(c.recv, grp), symm_mem.rendezvous(c.cnt_table, grp)
            c.peer_recv = (ctypes.c_void_p * P)(*[int(q) for q in c.h_recv.buffer_ptrs])
            c.peer_cnt = (ctypes.c_void_p * P)(*[int(q) for q in c.h_cnt.buffer_ptrs])
            torch.cuda.synchronize()
            c.h_recv.barrier(channel=0)
        else:
            c.recv = torch.zeros((P, 2, c.Rm, D), device=dev) if P > 1 else c.send.view(1, 2, c.Rm, D)
        c.Qr = len(range(rank, Q, P))
        c.src_rows, c.dst_rows, c.neg_rows = i64(B), i64(B), i64(B, max(c.Qr, 1))
        ctxs[("dp", B, Q, rank, P)] = c
        return c

    def _eval_body_dp(self, c: SimpleNamespace, group):
        N, D, L = self.N, self.D, _L()
        B, Q, P, rank, w, mw = c.B, c.Q, c.P, c.rank, c.w, c.mw
        main, upd = torch.cuda.current_stream(), self.upd
        if not c.dense:      # global rank of every candidate in the sorted unique root list; my share of the roots
            if c.n_ids <= 8192:
                check(L.tgn_unique_mark_rank(_p(c.ids), c.n_ids, _p(self.bitmap), N, _p(c.roots_all), c.Rall,
                                             _p(self.assoc), _p(c.Rall_dev), 0, _stream()))
            else:
                check(L.tgn_unique_mark(_p(c.ids), c.n_ids, None, N, _p(self.bitmap), _stream()))
                check(L.tgn_unique_rank(_p(self.bitmap), N, _p(c.roots_all), c.Rall, _p(self.assoc), _p(c.Rall_dev), 0,
                                        _stream()))
            check(L.tgn_relabel(_p(c.ids), c.n_ids, None, _p(self.assoc), _p(c.ids_g), _stream()))
            check(L.tgn_stride_select(_p(c.roots_all), c.Rall, _p(c.Rall_dev), rank, P, _p(c.my_roots), c.Rm,
                                      _p(c.Rm_dev), _stream()))
        if c.dense:
            # every node is a root of SOME rank and (through the hubs) a neighbour of every rank's roots: the step
            # works on all N rows in place -- one ring lookup of this rank's roots, a snapshot of memory /
            # last_update (the forked state update below rewrites the touched rows), no unique / relabel launches
            check(L.tgn_nbr_lookup(_p(c.my_roots), c.Rm, _p(c.Rm_dev), self.K, N, _p(self.neighbors), _p(self.e_id),
                                   _p(self.t_ring), _p(w.nbr_l), _p(w.ctr_g), _p(w.eid), _p(w.t_e), _p(w.root_off),
                                   _p(w.E_dev), None, _p(w.lookup_ws), _stream()))
            w.z.copy_(self.memory)
            w.lu.copy_(self.last_update)
        else:
            self._sample(w, c.my_roots, c.my_l, ids_dev=c.Rm_dev)
            s = _stream()
            check(L.tgn_gather_rows(_p(self.memory), _p(w.n_id), w.Nb, _p(w.Nb_dev), D, _p(w.z), s))
            check(L.tgn_relabel(_p(w.n_id), w.Nb, _p(w.Nb_dev), _p(self.last_update), _p(w.lu), s))
        # ---- forked: the batch's state update (store first, then memory: memory_module.py:135-138; then the ring).
        # It needs the batch only, and may start once this batch's sampling and memory reads are issued.
        upd.wait_stream(main)
        with torch.cuda.stream(upd):
            src, dst = c.ids[:B], c.ids[B:2 * B]
            self.store.update(src, dst, c.t_i, c.msg, base_dev=self.log_base_dev)
            check(L.tgn_unique_mark_rank(_p(c.ids), 2 * B, _p(self.bitmap), N, _p(c.n_upd), 2 * B, None, _p(c.U_dev), 0,
                                         _stream()))
            self._memory_fwd(mw, c.n_upd, 2 * B, c.U_dev)
            self._scatter_owned(c.n_upd, mw.z, mw.lu, self.memory, self.last_update, num_dev=c.U_dev)
            check(L.tgn_nbr_insert(_p(src), _p(dst), _p(c.t_f), B, 0, _p(self.cur_e_id_dev), self.K, N,
                                   _p(self.neighbors), _p(self.e_id), _p(self.t_ring), _stream()))
        self._attention_fwd(w, w.z, w.lu, False)
        p, off = self.p, self.off
        check(L.tgn_gather_rows(_p(w.emb), _p(c.my_l), c.Rm, _p(c.Rm_dev), D, _p(c.e_r), _stream()))
        ops.gemm_batch([
            ops.gemm_desc(c.e_r, self.flat, c.send, m=c.Rm, n=D, k=D, lda=D, ldb=D, ldc=D, b_off=off["lin_src.weight"],
                          bias=p["lin_src.bias"], m_dev=c.Rm_dev),
            ops.gemm_desc(c.e_r, self.flat, c.send, m=c.Rm, n=D, k=D, lda=D, ldb=D, ldc=D, b_off=off["lin_dst.weight"],
                          bias=p["lin_dst.bias"], m_dev=c.Rm_dev, c_off=c.Rm * D),
        ], self.prec)
        if c.peer:
            # the previous batch's end-of-step barrier guarantees every rank has finished reading its table
            check(L.tgn_peer_bcast(_p(c.send), 4 * c.send.numel(), c.peer_recv, 4 * c.send.numel() * rank, P, _stream()))
            c.h_recv.barrier(channel=0)
        elif P > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(c.recv, c.send, group=group)
        # row of global root g in the gathered table: block of its rank (g % P), position g // P -- one kernel for
        # the sources, the destinations and this rank's negative columns
        g = c.ids if c.dense else c.ids_g
        check(L.tgn_dp_rows(_p(g), B, Q, rank, P, 2 * c.Rm, _p(c.src_rows), _p(c.dst_rows), _p(c.neg_rows), _stream()))
        check(L.tgn_score_negs(_p(c.recv), c.recv.data_ptr() + 4 * c.Rm * D, _p(c.src_rows), _p(c.dst_rows),
                               _p(c.neg_rows), B, c.Qr, D, _p(p["lin_final.weight"]), _p(p["lin_final.bias"]),
                               _p(c.pos), None, _p(c.gt), _p(c.ge), _stream()))
        if c.reduce:
            # the TGB rank needs the counts over ALL columns: the per-rank counts are exchanged (peer writes into
            # every rank's [P, 2, B] table + barrier + a local sum in rank order: integers, so every rank gets the
            # same totals; or one all-reduce of 2*B int32), then the batch's mean reciprocal rank is added to the
            # epoch accumulator on the device (epoch_utils.py:113,163)
            if c.peer:
                check(L.tgn_peer_bcast(_p(c.cnt), 4 * c.cnt.numel(), c.peer_cnt, 4 * c.cnt.numel() * rank, P, _stream()))
                c.h_recv.barrier(channel=1)
                torch.sum(c.cnt_table, dim=0, dtype=torch.int32, out=c.cnt)
            elif P > 1:
                import torch.distributed as dist
                dist.all_reduce(c.cnt, group=group)
            check(L.tgn_rank_accum(_p(c.gt), _p(c.ge), B, _p(self.mrr_acc), _p(c.rr), _stream()))
        elif c.peer:
            c.h_recv.barrier(channel=1)      # nobody overwrites a table that a slower rank is still scoring from
        main.wait_stream(upd)

    @torch.no_grad()
    def eval_batch_dp(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor, rank: int = 0,
                      world: int = 1, group=None, reduce: bool = False):
        """eval_batch for data-parallel evaluation over `world` model replicas, with the embedding sharded too:
        `neg` is the FULL [B, Q] negative matrix on every rank.  The roots of the batch (unique candidates) are
        dealt round-robin; each rank samples / gathers / embeds only its roots and projects them through the
        decoder's two linears, ONE all-gather assembles the projected rows on every rank, and each rank scores
        its column shard of the negatives.  Returns (pos[B], gt[B], ge[B]) with the counts of THIS rank's
        columns (sum them over the ranks: dist_eval.reduce_counts).  The state update is replicated and runs on
        a forked stream beside the embedding.
        reduce=True: the counts are all-reduced INSIDE the captured step and the batch's mean reciprocal rank is
        added to self.mrr_acc (float64 [2] on the device: sum of per-batch means, number of batches; read it once
        per epoch with epoch_mrr()); the returned counts are then the global ones."""
        if self.world != 1:
            raise _cabi.TgnError("eval_batch_dp runs on model replicas (world == 1 engines), one per rank")
        self._unprime()
        B, Q = neg.shape
        c = self._eval_ctx_dp(B, Q, rank, world, group)
        c.reduce = reduce
        if not hasattr(self, "mrr_acc"):
            self.mrr_acc = torch.zeros(2, dtype=torch.float64, device=self.dev)
        self._reserve(B)
        c.ids[:B].copy_(src, non_blocking=True)
        c.ids[B:2 * B].copy_(dst, non_blocking=True)
        if Q:
            c.ids[2 * B:].view(B, Q).copy_(neg, non_blocking=True)
        c.t_i.copy_(t, non_blocking=True)
        c.t_f.copy_(c.t_i)
        if self.De:
            c.msg.copy_(msg, non_blocking=True)
        self._run(("eval_dp", B, Q, rank, world, reduce), lambda: self._eval_body_dp(c, group))
        self._advance(B)
        return c.pos, c.gt, c.ge

    def epoch_mrr(self, reset: bool = True) -> float:
        """mean over the batches of the per-batch mean reciprocal rank accumulated by eval_batch_dp(reduce=True)
        (epoch_utils.py:163); ONE host read per epoch."""
        acc = self.mrr_acc.tolist()
        if reset:
            self.mrr_acc.zero_()
        return acc[0] / max(acc[1], 1.0)

    def eval_scores(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor):
        """(pos[B], neg[B,Q]) probabilities of one evaluation batch (see eval_batch)."""
        pos, negs, _, _ = self.eval_batch(src, dst, neg, t, msg, True)
        return pos, negs

    @torch.no_grad()
    def flush_to_eval(self):
        """TGNMemory.train(False) (memory_module.py:209-215): every node goes through the updater
        with its stored messages, then the store is cleared."""
        self._unprime()
        new_mem = torch.zeros_like(self.memory)
        new_lu = torch.zeros_like(self.last_update)
        chunk = 1 << 16
        wm = self._alloc_work(1, 1, min(self.N, chunk), 1, train=False)
        for lo in range(0, self.N, chunk):
            ids = torch.arange(lo, min(self.N, lo + chunk), device=self.dev)
            self._memory_fwd(wm, ids, ids.numel(), None)
            self._scatter_owned(ids, wm.z, wm.lu, new_mem, new_lu)
        self.memory.copy_(new_mem)
        self.last_update.copy_(new_lu)
        self.store.reset()
        self.log_base_dev.zero_()
        self.events_done = 0
        self.training = False
