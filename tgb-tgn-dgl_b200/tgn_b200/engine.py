"""Fused TGN training / evaluation step on the sm_100a kernels.

One step = the per-batch flow the reference's four hot-path pieces serve
(pyg_epoch_utils.py:106-137 commented flow == PyG/TGB tgn.py):

    roots = unique(src, dst, neg)            -> bitmap ranking           (unique.cu)
    neighbour lookup of the roots            -> ring kernel              (nbr_ring.cu)
    n_id = unique(roots + neighbours), relabel
    z, last_update = memory(n_id)            -> fused store gather + concat + time-enc
                                                + Last aggregation + GRU (msgstore.cu, dense.cu)
    z = gnn(z, last_update, edges, t, msg)   -> fused time-enc + attention (attn.cu)
    pos/neg logits, BCE loss, backward, Adam
    memory.update_state(src, dst, t, msg)    -> scatter of the rows just computed + store update
    neighbor_loader.insert(src, dst, t)      -> ring insert

All buffers are sized by upper bounds (3B roots, 3B*K edges, ...) and the true
counts stay in device memory, so the step has no host synchronisation and is
captured once into a CUDA graph; a replay needs no host work besides the launch.
State and weights use the reference's names so that state_dicts move freely
between this engine, the drop-in modules and the CPU oracle.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch
from torch import Tensor

from . import _cabi, ops
from ._cabi import check

_L = _cabi.lib
_p = ops._p
_stream = ops._stream


class TGNEngine:
    def __init__(self, num_nodes: int, raw_dim: int, hidden: int, size_k: int, batch_size: int,
                 device="cuda", lr: float = 1e-4, heads: int = 2, dropout: float = 0.1,
                 log_capacity: int = 1 << 20, seed: int = 0, use_graph: bool = True):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TGNEngine runs on CUDA only (no CPU fallback)")
        self.dev, self.N, self.De, self.D, self.K, self.B = dev, num_nodes, raw_dim, hidden, size_k, batch_size
        self.H, self.C = heads, hidden // heads
        self.HC = self.H * self.C
        self.Dt = hidden
        self.Dx = 2 * hidden + raw_dim + hidden
        self.lr, self.dropout, self.seed, self.use_graph = lr, dropout, seed, use_graph
        D, De, Dt, Dx, HC = self.D, self.De, self.Dt, self.Dx, self.HC
        # ---- parameters: one flat buffer (one Adam launch, one grad memset)
        shapes = [("time_enc.lin.weight", (Dt, 1)), ("time_enc.lin.bias", (Dt,)),
                  ("memory_updater.weight_ih", (3 * D, Dx)), ("memory_updater.weight_hh", (3 * D, D)),
                  ("memory_updater.bias_ih", (3 * D,)), ("memory_updater.bias_hh", (3 * D,)),
                  ("conv.w_node", (4 * HC, D)), ("conv.b_node", (4 * HC,)),
                  ("conv.lin_edge.weight", (HC, Dt + De)),
                  ("lin_src.weight", (D, D)), ("lin_src.bias", (D,)),
                  ("lin_dst.weight", (D, D)), ("lin_dst.bias", (D,)),
                  ("lin_final.weight", (1, D)), ("lin_final.bias", (1,))]
        total = sum(int(torch.Size(s).numel()) for _, s in shapes)
        self.flat = torch.zeros(total, device=dev)
        self.flat_grad = torch.zeros(total, device=dev)
        self.exp_avg = torch.zeros(total, device=dev)
        self.exp_avg_sq = torch.zeros(total, device=dev)
        self.adam_step_dev = torch.zeros(1, device=dev)
        self.p: Dict[str, Tensor] = {}
        o = 0
        for name, shp in shapes:
            n = int(torch.Size(shp).numel())
            v = self.flat[o:o + n].view(shp)
            v.requires_grad_()
            v.grad = self.flat_grad[o:o + n].view(shp)
            self.p[name] = v
            o += n
        # ---- state
        self.memory = torch.zeros((num_nodes, D), device=dev)
        self.last_update = torch.zeros(num_nodes, dtype=torch.long, device=dev)
        self.assoc = torch.zeros(num_nodes, dtype=torch.long, device=dev)
        self.neighbors = torch.zeros((num_nodes, size_k), dtype=torch.long, device=dev)
        self.e_id = torch.full((num_nodes, size_k), -1, dtype=torch.long, device=dev)
        self.t_ring = torch.full((num_nodes, size_k), -1.0, device=dev)
        self.store = ops.MsgStore(num_nodes, raw_dim, dev, capacity=log_capacity, t_dtype=torch.int64)
        self.bitmap = torch.zeros(_L().tgn_bitmap_bytes(num_nodes) // 4, dtype=torch.int32, device=dev)
        self.cur_e_id_dev = torch.zeros(1, dtype=torch.long, device=dev)   # ring event counter
        self.log_base_dev = torch.zeros(1, dtype=torch.long, device=dev)   # store log position
        self.pos_dev = torch.zeros(1, dtype=torch.long, device=dev)        # dataset cursor
        self.step_dev = torch.zeros(1, dtype=torch.long, device=dev)       # dropout stream (advanced after backward)
        self.events_done = 0
        self.events = None
        self._alloc_step_buffers(batch_size)
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.training = True
        self.launches_per_step = None

    # ------------------------------------------------------------------ buffers
    def _alloc_step_buffers(self, B: int):
        dev, K = self.dev, self.K
        R = 3 * B
        E = R * K
        Nb = min(self.N, R + E)
        i64 = dict(dtype=torch.long, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.in_i64 = torch.zeros(4 * B, **i64)          # [src | dst | neg | t]: one H2D copy per step
        self.in_ids3 = self.in_i64[:3 * B]
        self.in_t_i64 = self.in_i64[3 * B:]
        self.in_t_f32 = torch.zeros(B, device=dev)
        self.in_msg = torch.zeros((B, max(self.De, 1)), device=dev)
        self.roots = torch.zeros(R, **i64)
        self.R_dev = torch.zeros(1, **i32)
        self.nbr_g = torch.zeros(E, **i64)
        self.ctr_g = torch.zeros(E, **i64)
        self.eid = torch.zeros(E, **i64)
        self.t_e = torch.zeros(E, device=dev)
        self.root_off = torch.zeros(R + 1, **i32)
        self.E_dev = torch.zeros(1, **i32)
        self.lookup_ws = torch.zeros(max(_L().tgn_nbr_lookup_ws_bytes(R, K), 16) // 8, **i64)
        self.n_id = torch.zeros(Nb, **i64)
        self.Nb_dev = torch.zeros(1, **i32)
        self.nbr_l = torch.zeros(E, **i64)
        self.ctr_l = torch.zeros(R, **i64)
        self.ids3_l = torch.zeros(3 * B, **i64)
        self.loss = torch.zeros((), device=dev)
        self.bounds = (R, E, Nb)

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
            if "memory" in memory_sd:
                self.memory.copy_(memory_sd["memory"])
                self.last_update.copy_(memory_sd["last_update"])

    def export_state(self):
        HC = self.HC
        mem = {k: self.p[k].detach().clone() for k in ("time_enc.lin.weight", "time_enc.lin.bias",
               "memory_updater.weight_ih", "memory_updater.weight_hh", "memory_updater.bias_ih",
               "memory_updater.bias_hh")}
        mem.update(memory=self.memory.clone(), last_update=self.last_update.clone(), _assoc=self.assoc.clone())
        gnn = {"time_enc.lin.weight": mem["time_enc.lin.weight"], "time_enc.lin.bias": mem["time_enc.lin.bias"],
               "conv.lin_edge.weight": self.p["conv.lin_edge.weight"].detach().clone()}
        for i, n in enumerate(("query", "key", "value", "skip")):
            gnn[f"conv.lin_{n}.weight"] = self.p["conv.w_node"].detach()[i * HC:(i + 1) * HC].clone()
            gnn[f"conv.lin_{n}.bias"] = self.p["conv.b_node"].detach()[i * HC:(i + 1) * HC].clone()
        lp = {k: self.p[k].detach().clone() for k in ("lin_src.weight", "lin_src.bias", "lin_dst.weight",
              "lin_dst.bias", "lin_final.weight", "lin_final.bias")}
        return mem, gnn, lp

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
        self.store.reset()

    # ------------------------------------------------------------------ data
    def set_events(self, src: Tensor, dst: Tensor, t: Tensor, msg: Tensor, neg: Optional[Tensor] = None):
        """Device-resident event arrays (the whole split).  `t` int64, `msg` [E, De].
        e_id of the ring == row of these arrays, like data.msg[e_id] in the reference."""
        dev = self.dev
        self.events = dict(src=src.to(dev, torch.long).contiguous(), dst=dst.to(dev, torch.long).contiguous(),
                           t=t.to(dev, torch.long).contiguous(), msg=msg.to(dev, torch.float32).contiguous(),
                           neg=None if neg is None else neg.to(dev, torch.long).contiguous())
        need = self.events["src"].numel()
        if self.store.capacity < need:
            self.store._alloc_log(need)

    def stage_batch_from_device(self):
        ev = self.events
        check(_L().tgn_batch_load(_p(ev["src"]), _p(ev["dst"]), _p(ev["neg"]), _p(ev["t"]), _p(ev["msg"]),
                                  self.De, self.B, _p(self.pos_dev), _p(self.in_ids3), _p(self.in_t_i64),
                                  _p(self.in_t_f32), _p(self.in_msg), _stream()))

    def stage_packed(self, ids_t: Tensor, msg: Tensor):
        """End-to-end staging: `ids_t` = pinned host int64 [4B] = [src|dst|neg|t], `msg` pinned
        host float32 [B, De]; two H2D copies, the float timestamps are derived on the device."""
        self.in_i64.copy_(ids_t, non_blocking=True)
        if self.De:
            self.in_msg.copy_(msg, non_blocking=True)
        self.in_t_f32.copy_(self.in_t_i64)

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
        self.store.size = count

    def stage_batch(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor):
        """Copies one batch (host or device tensors) into the static step buffers."""
        B = self.B
        self.in_ids3[:B].copy_(src, non_blocking=True)
        self.in_ids3[B:2 * B].copy_(dst, non_blocking=True)
        self.in_ids3[2 * B:].copy_(neg, non_blocking=True)
        self.in_t_i64.copy_(t, non_blocking=True)
        self.in_t_f32.copy_(t, non_blocking=True)
        if self.De:
            self.in_msg.copy_(msg, non_blocking=True)

    # ------------------------------------------------------------------ the step
    def _sample(self):
        """roots -> neighbour lookup -> union -> relabel; everything bound-sized, counts on device."""
        B, N, K = self.B, self.N, self.K
        R, E, Nb = self.bounds
        L, s = _L(), _stream()
        check(L.tgn_unique_mark(_p(self.in_ids3), 3 * B, None, N, _p(self.bitmap), s))
        check(L.tgn_unique_rank(_p(self.bitmap), N, _p(self.roots), R, None, _p(self.R_dev), 1, s))
        check(L.tgn_nbr_lookup(_p(self.roots), R, _p(self.R_dev), K, N, _p(self.neighbors), _p(self.e_id),
                               _p(self.t_ring), _p(self.nbr_g), _p(self.ctr_g), _p(self.eid), _p(self.t_e),
                               _p(self.root_off), _p(self.E_dev), _p(self.bitmap), _p(self.lookup_ws), s))
        check(L.tgn_unique_rank(_p(self.bitmap), N, _p(self.n_id), Nb, _p(self.assoc), _p(self.Nb_dev), 0, s))
        check(L.tgn_relabel(_p(self.nbr_g), E, _p(self.E_dev), _p(self.assoc), _p(self.nbr_l), s))
        check(L.tgn_relabel(_p(self.roots), R, _p(self.R_dev), _p(self.assoc), _p(self.ctr_l), s))
        check(L.tgn_relabel(_p(self.in_ids3), 3 * B, None, _p(self.assoc), _p(self.ids3_l), s))

    def _embed(self, train: bool):
        p = self.p
        ev = self.events
        if train:
            z, lu = ops.memory_update(p["time_enc.lin.weight"].view(-1), p["time_enc.lin.bias"],
                                      p["memory_updater.weight_ih"], p["memory_updater.weight_hh"],
                                      p["memory_updater.bias_ih"], p["memory_updater.bias_hh"], self.store,
                                      self.n_id, self.memory, self.last_update, ops.AGG_LAST, self.Nb_dev)
        else:
            z = ops.gather_rows(self.memory, self.n_id, self.Nb_dev)
            lu = torch.empty_like(self.n_id)
            check(_L().tgn_relabel(_p(self.n_id), self.n_id.numel(), _p(self.Nb_dev), _p(self.last_update),
                                   _p(lu), _stream()))
        emb = ops.temporal_attention(
            z, p["conv.w_node"], p["conv.b_node"], p["conv.lin_edge.weight"],
            p["time_enc.lin.weight"].view(-1), p["time_enc.lin.bias"], lu, self.nbr_l, ev["t"], ev["msg"],
            self.root_off, heads=self.H, msg_rows=self.eid, centre_ids=self.ctr_l,
            dropout_p=self.dropout if train else 0.0, seed=self.seed,
            counts=(self.Nb_dev, self.E_dev, self.R_dev, self.step_dev))
        return z, lu, emb

    def _logits(self, emb: Tensor):
        p, B = self.p, self.B
        a = ops.linear(emb.index_select(0, self.ids3_l[:B]), p["lin_src.weight"], p["lin_src.bias"])
        b = ops.linear(emb.index_select(0, self.ids3_l[B:]), p["lin_dst.weight"], p["lin_dst.bias"])
        h = (a.repeat(2, 1) + b).relu()
        return ops.linear(h, p["lin_final.weight"], p["lin_final.bias"]).view(2, B)

    def _update_state(self, z: Tensor, lu: Tensor):
        B = self.B
        # memory[n] = z[assoc[n]] for n in (src, dst): the rows _update_memory would recompute
        # (memory_module.py:147-150) are the rows the forward just produced (same store, same weights)
        ops.memory_scatter(self.in_ids3[:2 * B], z.detach(), lu, self.memory, self.last_update,
                           src_rows=self.ids3_l[:2 * B])
        self.store.update(self.in_ids3[:B], self.in_ids3[B:2 * B], self.in_t_i64, self.in_msg,
                          base_dev=self.log_base_dev)
        check(_L().tgn_nbr_insert(_p(self.in_ids3), self.in_ids3[B:].data_ptr(), _p(self.in_t_f32), B, 0,
                                  _p(self.cur_e_id_dev), self.K, self.N, _p(self.neighbors), _p(self.e_id),
                                  _p(self.t_ring), _stream()))

    def _train_body(self):
        self.flat_grad.zero_()
        self._sample()
        z, lu, emb = self._embed(True)
        logits = self._logits(emb)
        loss = torch.nn.functional.softplus(-logits[0]).mean() + torch.nn.functional.softplus(logits[1]).mean()
        self._update_state(z, lu)
        loss.backward()
        ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.adam_step_dev, self.lr)
        self.loss.copy_(loss.detach())
        self.step_dev.add_(1)

    def _run(self, key: tuple, body):
        if not self.use_graph:
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            # warm-up already happened: the first calls of every configuration run eagerly
            cnt = self._graphs.get(("warm",) + key, 0)
            if cnt < 3:
                self._graphs[("warm",) + key] = cnt + 1
                body()
                return
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
        g.replay()

    def train_step(self, from_device: bool = True):
        """One training batch.  from_device=True: slice the next batch out of the resident event
        arrays (set_events); False: the caller staged it with stage_batch()."""
        if from_device:
            self._run(("train", True), lambda: (self.stage_batch_from_device(), self._train_body()))
        else:
            self._run(("train", False), self._train_body)
        self.events_done += self.B
        self.store.size = self.events_done   # host mirror of log_base_dev (graph replays skip the Python body)
        return self.loss

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def eval_scores(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor):
        """test() body for one batch (epoch_utils.py:28-157): scores of the positives and of
        the [B,Q] negatives, then eval-mode update_state (store first, then memory) and insert.
        Returns (pos[B], neg[B,Q]) probabilities."""
        dev, N = self.dev, self.N
        src, dst, neg = src.to(dev, torch.long), dst.to(dev, torch.long), neg.to(dev, torch.long)
        t_i = t.to(dev, torch.long)
        B, Q = neg.shape
        n_in = torch.cat([src, dst, neg.reshape(-1)])
        roots = ops.unique_relabel([n_in], N)
        ids, edge_index, e_id, _, root_off = ops.nbr_lookup(roots, self.neighbors, self.e_id, self.t_ring, self.assoc)
        p, ev = self.p, self.events
        z = ops.gather_rows(self.memory, ids)
        lu = self.last_update[ids]
        ctr_l = ops.relabel(roots, self.assoc)
        emb = ops.temporal_attention(z, p["conv.w_node"], p["conv.b_node"], p["conv.lin_edge.weight"],
                                     p["time_enc.lin.weight"].view(-1), p["time_enc.lin.bias"], lu,
                                     edge_index[0].contiguous(), ev["t"], ev["msg"], root_off, heads=self.H,
                                     msg_rows=e_id, centre_ids=ctr_l)
        Nb, D = emb.shape
        hs = ops.sgemm(emb, p["lin_src.weight"], p["lin_src.bias"], m=Nb, n=D, k=D, lda=D, ldb=D)
        hd = ops.sgemm(emb, p["lin_dst.weight"], p["lin_dst.bias"], m=Nb, n=D, k=D, lda=D, ldb=D)
        sl, dl, nl = self.assoc[src], self.assoc[dst], self.assoc[neg.reshape(-1)]
        wf, bf = p["lin_final.weight"].view(-1), p["lin_final.bias"]
        pos = ops.link_score(hs, hd, sl, dl, wf, bf, True)
        negs = ops.link_score(hs, hd, sl.repeat_interleave(Q), nl, wf, bf, True).view(B, Q)
        # eval ordering of update_state: store first, then memory (memory_module.py:135-138)
        self.store.update(src, dst, t_i, msg.to(dev, torch.float32))
        self.log_base_dev += B
        self.events_done += B
        n_upd = ops.unique_relabel([src, dst], N)
        m_new, lu_new = ops.memory_update(p["time_enc.lin.weight"].view(-1), p["time_enc.lin.bias"],
                                          p["memory_updater.weight_ih"], p["memory_updater.weight_hh"],
                                          p["memory_updater.bias_ih"], p["memory_updater.bias_hh"], self.store,
                                          n_upd, self.memory, self.last_update, ops.AGG_LAST, None)
        ops.memory_scatter(n_upd, m_new, lu_new, self.memory, self.last_update)
        ops.nbr_insert(src, dst, t_i.to(torch.float32), 0, self.neighbors, self.e_id, self.t_ring,
                       cur_e_id_dev=self.cur_e_id_dev)
        return pos, negs

    @torch.no_grad()
    def flush_to_eval(self):
        """TGNMemory.train(False) (memory_module.py:209-215): every node goes through the updater
        with its stored messages, then the store is cleared."""
        p = self.p
        new_mem = torch.empty_like(self.memory)
        new_lu = torch.empty_like(self.last_update)
        for lo in range(0, self.N, 1 << 16):
            ids = torch.arange(lo, min(self.N, lo + (1 << 16)), device=self.dev)
            m, lu = ops.memory_update(p["time_enc.lin.weight"].view(-1), p["time_enc.lin.bias"],
                                      p["memory_updater.weight_ih"], p["memory_updater.weight_hh"],
                                      p["memory_updater.bias_ih"], p["memory_updater.bias_hh"], self.store,
                                      ids, self.memory, self.last_update, ops.AGG_LAST, None)
            new_mem[lo:lo + ids.numel()] = m
            new_lu[lo:lo + ids.numel()] = lu
        self.memory.copy_(new_mem)
        self.last_update.copy_(new_lu)
        self.store.reset()
        self.log_base_dev.zero_()
        self.events_done = 0
        self.training = False
