"""TEST INFRASTRUCTURE -- CPU oracle of the reference's per-batch TGN hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product path (tgb-tgn-dgl_b200/)
never does and fails loudly when its CUDA library is missing.

Every class/function restates one piece of cseduashraful/tgb-tgn-dgl and cites
the file:line it follows (paths relative to the reference root).  Pinning
status (see tests/golden/make_golden.py, tests/test_oracle_golden.py):

  * LastNeighborLoader lookup/insert     PINNED against the reference class itself
                                         (neighbor_loader.py imported unmodified)
  * LastAggregator / MeanAggregator,
    TGNMemory, GraphAttentionEmbedding   PINNED against the reference files run
                                         unmodified on top of oracle/thirdparty.py
  * third-party arithmetic (scatter_max, scatter, TransformerConv, TimeEncoder)
                                         PARITY UNPINNED (packages absent; restated
                                         from their published algorithms)
  * t-CSR sampler (TGL sampler_core)     PARITY UNPINNED (source absent from the
                                         reference; restated from SURVEY.md B1,
                                         pinned only by hand-computed cases)

Known divergence, documented and tested: the reference orders a node's events
inside one batch with torch.sort, which is *unstable* for small CPU tensors
(measured in this container).  The oracle (and the CUDA path) use the stable
order.  The two agree whenever a node has at most K events in one batch
(neighbour ring) and whenever a node's events in one batch have distinct
timestamps (message store).
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import Tensor

from . import thirdparty as tp

# ---------------------------------------------------------------------------
# Philox-4x32-10 (same generator as csrc/common.cuh) so that the uniform
# sampler can be checked bit-exactly
# ---------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(seed: int, ctr_lo: int, ctr_hi: int) -> Tuple[int, int, int, int]:
    k0, k1 = seed & _MASK, (seed >> 32) & _MASK
    c0, c1 = ctr_lo & _MASK, (ctr_lo >> 32) & _MASK
    c2, c3 = ctr_hi & _MASK, (ctr_hi >> 32) & _MASK
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & _MASK, p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


# ---------------------------------------------------------------------------
# LastNeighborLoader (neighbor_loader.py:15-109)
# ---------------------------------------------------------------------------
class NeighborRing:
    """Per-node ring of the `size` most recent neighbours, restated with explicit
    per-node loops (small cases only)."""

    def __init__(self, num_nodes: int, size: int):
        self.size = size
        self.neighbors = np.zeros((num_nodes, size), dtype=np.int64)
        self.e_id = np.full((num_nodes, size), -1, dtype=np.int64)   # :108 reset_state
        self.t = np.full((num_nodes, size), -1.0, dtype=np.float32)  # :109
        self.cur_e_id = 0
        self._assoc = np.zeros(num_nodes, dtype=np.int64)

    def reset_state(self):
        self.cur_e_id = 0
        self.e_id.fill(-1)
        self.t.fill(-1.0)

    def lookup(self, n_id: np.ndarray):
        """neighbor_loader.py:26-50.  Edge order = n_id order x slot order."""
        nb, ct, ei, tt = [], [], [], []
        for n in n_id.tolist():
            for s in range(self.size):
                if self.e_id[n, s] >= 0:            # :41 mask = e_id >= 0
                    nb.append(self.neighbors[n, s]); ct.append(n)
                    ei.append(self.e_id[n, s]); tt.append(self.t[n, s])
        nb = np.asarray(nb, dtype=np.int64); ct = np.asarray(ct, dtype=np.int64)
        uniq = np.unique(np.concatenate([n_id.astype(np.int64), nb]))  # :46 sorted unique
        self._assoc[uniq] = np.arange(uniq.size)                        # :47
        edge_index = np.stack([self._assoc[nb], self._assoc[ct]]) if nb.size else np.zeros((2, 0), np.int64)
        return uniq, edge_index, np.asarray(ei, dtype=np.int64), np.asarray(tt, dtype=np.float32)

    def insert(self, src: np.ndarray, dst: np.ndarray, t: np.ndarray):
        """neighbor_loader.py:52-104 with a STABLE node sort."""
        B, K = src.size, self.size
        nodes = np.concatenate([dst, src])             # :58  (centre of entry j)
        nbrs = np.concatenate([src, dst])              # :57
        e_new = np.concatenate([np.arange(B), np.arange(B)]) + self.cur_e_id  # :59-61
        t2 = np.concatenate([t, t]).astype(np.float32)
        self.cur_e_id += B
        order = np.argsort(nodes, kind="stable")       # :68 (reference: torch.sort, unstable)
        for n in np.unique(nodes).tolist():
            run = order[nodes[order] == n]
            # slot = sorted position % K, later entries overwrite (:75-88) -> last K of the run
            kept = run[-K:]
            cand_e = np.concatenate([self.e_id[n], e_new[kept], np.full(K - kept.size, -1)])
            cand_n = np.concatenate([self.neighbors[n], nbrs[kept], np.zeros(K - kept.size, np.int64)])
            cand_t = np.concatenate([self.t[n], t2[kept], np.full(K - kept.size, -1.0, np.float32)])
            top = np.argsort(-cand_e, kind="stable")[:K]        # :99 e_id.topk
            self.e_id[n] = cand_e[top]
            self.neighbors[n] = cand_n[top]                      # :104 gather by the e_id perm
            self.t[n] = np.sort(cand_t)[::-1][:K]                # :100 t.topk, independent of e_id


# ---------------------------------------------------------------------------
# TGL sampler_core.ParallelSampler (source absent)
# ---------------------------------------------------------------------------
def tcsr_sample_ref(indptr, indices, eid, ts, roots, root_ts, k, strategy="recent", offset=0.0,
                    duration=0.0, seed=0):
    """Brute-force restatement: per root mask the row by timestamp, take the last
    k (most recent first) or k uniform draws with replacement.  Uniform draws use
    the same Philox stream as the CUDA kernel: draw j of root r = philox(seed, r, j)[0] % cand."""
    out_n, out_c, out_e, out_t, out_d, off = [], [], [], [], [], [0]
    for r, (n, t) in enumerate(zip(np.asarray(roots).tolist(), np.asarray(root_ts, dtype=np.float32))):
        s, e = int(indptr[n]), int(indptr[n + 1])
        row_ts = ts[s:e]
        t_hi = np.float32(t) + np.float32(offset)
        hi = s + int(np.searchsorted(row_ts, t_hi, side="left"))     # ts < t_hi
        lo = s + int(np.searchsorted(row_ts, t_hi - np.float32(duration), side="left")) if duration > 0 else s
        cand = hi - lo
        if strategy == "recent" or cand <= k:
            picks = list(range(hi - 1, max(lo, hi - k) - 1, -1))
        else:
            picks = [lo + philox4x32(seed, r, j)[0] % cand for j in range(k)]
        for p in picks:
            out_n.append(indices[p]); out_c.append(r); out_e.append(eid[p])
            out_t.append(ts[p]); out_d.append(np.float32(t) - np.float32(ts[p]))
        off.append(len(out_n))
    return (np.asarray(out_n, np.int32), np.asarray(out_c, np.int32), np.asarray(out_e, np.int32),
            np.asarray(out_t, np.float32), np.asarray(out_d, np.float32), np.asarray(off, np.int32))


def build_tcsr(src, dst, t, num_nodes, add_reverse=True):
    """t-CSR as TGL's gen_graph writes it (ext_full.npz keys indptr/indices/ts/eid,
    reference utils.py:73): rows sorted by (ts, eid), reverse edges share the eid."""
    src = np.asarray(src, np.int64); dst = np.asarray(dst, np.int64)
    ts = np.asarray(t, np.float32); e = np.arange(src.size, dtype=np.int64)
    if add_reverse:
        row = np.concatenate([src, dst]); col = np.concatenate([dst, src])
        tt = np.concatenate([ts, ts]); ee = np.concatenate([e, e])
    else:
        row, col, tt, ee = src, dst, ts, e
    order = np.lexsort((ee, tt, row))
    row, col, tt, ee = row[order], col[order], tt[order], ee[order]
    indptr = np.zeros(num_nodes + 1, np.int64)
    np.add.at(indptr, row + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr.astype(np.int32), col.astype(np.int32), ee.astype(np.int32), tt.astype(np.float32)


# ---------------------------------------------------------------------------
# aggregators (modules/msg_agg.py)
# ---------------------------------------------------------------------------
class LastAggregator(torch.nn.Module):
    def forward(self, msg: Tensor, index: Tensor, t: Tensor, dim_size: int):
        _, argmax = tp.scatter_max(t, index, dim=0, dim_size=dim_size)   # msg_agg.py:17
        out = msg.new_zeros((dim_size, msg.size(-1)))                      # :18
        mask = argmax < msg.size(0)                                        # :19
        out[mask] = msg[argmax[mask]]                                      # :20
        return out


class MeanAggregator(torch.nn.Module):
    def forward(self, msg: Tensor, index: Tensor, t: Tensor, dim_size: int):
        return tp.scatter(msg, index, dim=0, dim_size=dim_size, reduce="mean")  # msg_agg.py:26


class IdentityMessage(torch.nn.Module):
    """modules/msg_func.py:12-18"""

    def __init__(self, raw_msg_dim: int, memory_dim: int, time_dim: int):
        super().__init__()
        self.out_channels = raw_msg_dim + 2 * memory_dim + time_dim

    def forward(self, z_src, z_dst, raw_msg, t_enc):
        return torch.cat([z_src, z_dst, raw_msg, t_enc], dim=-1)


# ---------------------------------------------------------------------------
# TGNMemory (modules/memory_module.py:25-215) with the Python-dict message store
# ---------------------------------------------------------------------------
class TGNMemory(torch.nn.Module):
    def __init__(self, num_nodes, raw_msg_dim, memory_dim, time_dim, message_module,
                 aggregator_module, memory_updater_cell="gru"):
        super().__init__()
        self.num_nodes, self.raw_msg_dim = num_nodes, raw_msg_dim
        self.memory_dim, self.time_dim = memory_dim, time_dim
        self.msg_s_module = message_module
        self.msg_d_module = copy.deepcopy(message_module)             # :67
        self.aggr_module = aggregator_module
        self.time_enc = tp.TimeEncoder(time_dim)                      # :69
        cell = {"gru": torch.nn.GRUCell, "rnn": torch.nn.RNNCell}[memory_updater_cell]
        self.memory_updater = cell(message_module.out_channels, memory_dim)   # :72/:74
        self.register_buffer("memory", torch.zeros(num_nodes, memory_dim))
        self.register_buffer("last_update", torch.zeros(num_nodes, dtype=torch.long))
        self.register_buffer("_assoc", torch.zeros(num_nodes, dtype=torch.long))
        self.msg_s_store: Dict[int, tuple] = {}
        self.msg_d_store: Dict[int, tuple] = {}
        self.reset_state()

    def reset_state(self):                                            # :106-110
        self.memory.fill_(0)
        self.last_update.fill_(0)
        self._reset_message_store()

    def detach(self):                                                 # :112-114
        self.memory.detach_()

    def _reset_message_store(self):                                   # :140-145
        i = torch.empty((0,), dtype=torch.long)
        msg = torch.empty((0, self.raw_msg_dim))
        self.msg_s_store = {j: (i, i, i, msg) for j in range(self.num_nodes)}
        self.msg_d_store = {j: (i, i, i, msg) for j in range(self.num_nodes)}

    def forward(self, n_id: Tensor):                                  # :116-124
        if self.training:
            return self._get_updated_memory(n_id)
        return self.memory[n_id], self.last_update[n_id]

    def update_state(self, src, dst, t, raw_msg):                     # :126-138
        n_id = torch.cat([src, dst]).unique()
        if self.training:
            self._update_memory(n_id)
            self._update_msg_store(src, dst, t, raw_msg, self.msg_s_store)
            self._update_msg_store(dst, src, t, raw_msg, self.msg_d_store)
        else:
            self._update_msg_store(src, dst, t, raw_msg, self.msg_s_store)
            self._update_msg_store(dst, src, t, raw_msg, self.msg_d_store)
            self._update_memory(n_id)

    def _update_memory(self, n_id):                                   # :147-150
        memory, last_update = self._get_updated_memory(n_id)
        self.memory[n_id] = memory
        self.last_update[n_id] = last_update.to(self.last_update.dtype)

    def _get_updated_memory(self, n_id):                              # :152-178
        self._assoc[n_id] = torch.arange(n_id.size(0))
        msg_s, t_s, src_s, _ = self._compute_msg(n_id, self.msg_s_store, self.msg_s_module)
        msg_d, t_d, src_d, _ = self._compute_msg(n_id, self.msg_d_store, self.msg_d_module)
        idx = torch.cat([src_s, src_d], dim=0)
        msg = torch.cat([msg_s, msg_d], dim=0)
        t = torch.cat([t_s, t_d], dim=0)
        aggr = self.aggr_module(msg, self._assoc[idx], t, n_id.size(0))
        memory = self.memory_updater(aggr, self.memory[n_id])
        last_update = tp.scatter(t, idx, 0, self.last_update.size(0), reduce="max")[n_id]
        return memory, last_update

    def _update_msg_store(self, src, dst, t, raw_msg, msg_store):     # :180-191
        n_id, perm = src.sort(stable=True)  # reference: src.sort() (unstable on CPU)
        n_id, count = n_id.unique_consecutive(return_counts=True)
        for i, idx in zip(n_id.tolist(), perm.split(count.tolist())):
            msg_store[i] = (src[idx], dst[idx], t[idx], raw_msg[idx])

    def _compute_msg(self, n_id, msg_store, msg_module):              # :193-207
        data = [msg_store[i] for i in n_id.tolist()]
        src, dst, t, raw_msg = list(zip(*data))
        src, dst, t, raw_msg = torch.cat(src), torch.cat(dst), torch.cat(t), torch.cat(raw_msg)
        t_rel = t - self.last_update[src]
        t_enc = self.time_enc(t_rel.to(raw_msg.dtype))
        msg = msg_module(self.memory[src], self.memory[dst], raw_msg, t_enc)
        return msg, t, src, dst

    def train(self, mode: bool = True):                               # :209-215
        if self.training and not mode:
            self._update_memory(torch.arange(self.num_nodes))
            self._reset_message_store()
        super().train(mode)


class DyRepMemory(TGNMemory):
    """modules/memory_module.py:218-421: TGNMemory with a `gru` or `rnn` updater (:259-265) whose
    messages may use the current embeddings of the batch's nodes instead of their memory (:389-408)."""

    def __init__(self, num_nodes, raw_msg_dim, memory_dim, time_dim, message_module, aggregator_module,
                 memory_updater_type, use_src_emb_in_msg=False, use_dst_emb_in_msg=False):
        assert memory_updater_type in ["gru", "rnn"]                          # :258
        super().__init__(num_nodes, raw_msg_dim, memory_dim, time_dim, message_module, aggregator_module,
                         memory_updater_type)
        self.use_src_emb_in_msg, self.use_dst_emb_in_msg = use_src_emb_in_msg, use_dst_emb_in_msg
        self._emb = None

    def update_state(self, src, dst, t, raw_msg, embeddings=None, assoc=None):   # :316-329
        self._emb = (embeddings, assoc) if embeddings is not None else None
        try:
            super().update_state(src, dst, t, raw_msg)
        finally:
            self._emb = None

    def _compute_msg(self, n_id, msg_store, msg_module):                      # :375-412
        data = [msg_store[i] for i in n_id.tolist()]
        src, dst, t, raw_msg = list(zip(*data))
        src, dst, t, raw_msg = torch.cat(src), torch.cat(dst), torch.cat(t), torch.cat(raw_msg)
        t_rel = t - self.last_update[src]
        t_enc = self.time_enc(t_rel.to(raw_msg.dtype))
        z_src, z_dst = self.memory[src], self.memory[dst]
        if self._emb is not None:
            emb, assoc = self._emb
            ids = set(n_id.tolist())
            if self.use_src_emb_in_msg:                                       # :389-397 (`s in n_id` loop)
                hit = [i for i, s_ in enumerate(src.tolist()) if s_ in ids]
                if hit:
                    z_src[hit] = emb[assoc[src[hit]]]
            if self.use_dst_emb_in_msg:                                       # :399-408
                hit = [i for i, d_ in enumerate(dst.tolist()) if d_ in ids]
                if hit:
                    z_dst[hit] = emb[assoc[dst[hit]]]
        return msg_module(z_src, z_dst, raw_msg, t_enc), t, src, dst


class TimeEmbedding(torch.nn.Module):
    """modules/emb_module.py:32-52 (JODIE projection)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.embedding_layer = torch.nn.Linear(1, out_channels)
        stdv = 1.0                                                            # 1/sqrt(fan_in), fan_in = 1 (:39-43)
        self.embedding_layer.weight.data.normal_(0, stdv)
        self.embedding_layer.bias.data.normal_(0, stdv)

    def forward(self, x, last_update, t):
        rel_t = last_update - t                                               # :49
        return x * (1 + self.embedding_layer(rel_t.to(x.dtype).unsqueeze(1)))  # :50


class GraphAttentionEmbedding(torch.nn.Module):
    """modules/emb_module.py:11-29"""

    def __init__(self, in_channels, out_channels, msg_dim, time_enc):
        super().__init__()
        self.time_enc = time_enc
        edge_dim = msg_dim + time_enc.out_channels
        self.conv = tp.TransformerConv(in_channels, out_channels // 2, heads=2, dropout=0.1,
                                       edge_dim=edge_dim)

    def forward(self, x, last_update, edge_index, t, msg):
        rel_t = last_update[edge_index[0]] - t                        # :26
        rel_t_enc = self.time_enc(rel_t.to(x.dtype))                  # :27
        edge_attr = torch.cat([rel_t_enc, msg], dim=-1)               # :28
        return self.conv(x, edge_index, edge_attr)                    # :29


class LinkPredictor(torch.nn.Module):
    """modules/decoder.py:12-27 (ends in sigmoid)."""

    def __init__(self, in_channels):
        super().__init__()
        self.lin_src = torch.nn.Linear(in_channels, in_channels)
        self.lin_dst = torch.nn.Linear(in_channels, in_channels)
        self.lin_final = torch.nn.Linear(in_channels, 1)

    def logits(self, z_src, z_dst):
        return self.lin_final((self.lin_src(z_src) + self.lin_dst(z_dst)).relu())

    def forward(self, z_src, z_dst):
        return self.logits(z_src, z_dst).sigmoid()


def mrr_ref(pos: np.ndarray, neg: np.ndarray) -> np.ndarray:
    """TGB Evaluator MRR per positive (epoch_utils.py:108-113):
    rank = 1 + 0.5*(#{neg > pos} + #{neg >= pos})."""
    pos = np.asarray(pos, np.float32).reshape(-1, 1)
    neg = np.asarray(neg, np.float32).reshape(pos.shape[0], -1)
    return 1.0 / (0.5 * ((neg > pos).sum(1) + (neg >= pos).sum(1)) + 1.0)


# ---------------------------------------------------------------------------
# the per-batch training step the four kernels serve (flow reconstructed from
# pyg_epoch_utils.py:106-137 and pyg_model_utils.py:10-43)
# ---------------------------------------------------------------------------
class TorchNeighborLoader:
    """neighbor_loader.py:15-109 restated in torch (vectorised like the reference,
    with stable sorts); this is what the CPU baseline times."""

    def __init__(self, num_nodes: int, size: int):
        self.size = size
        self.neighbors = torch.zeros((num_nodes, size), dtype=torch.long)
        self.e_id = torch.full((num_nodes, size), -1, dtype=torch.long)
        self.t = torch.full((num_nodes, size), -1.0)
        self._assoc = torch.zeros(num_nodes, dtype=torch.long)
        self.cur_e_id = 0

    def reset_state(self):
        self.cur_e_id = 0
        self.e_id.fill_(-1)
        self.t.fill_(-1)

    def __call__(self, n_id):
        neighbors = self.neighbors[n_id]
        nodes = n_id.view(-1, 1).repeat(1, self.size)
        e_id, t = self.e_id[n_id], self.t[n_id]
        mask = e_id >= 0
        neighbors, nodes, e_id, t = neighbors[mask], nodes[mask], e_id[mask], t[mask]
        n_id = torch.cat([n_id, neighbors]).unique()
        self._assoc[n_id] = torch.arange(n_id.size(0))
        return n_id, torch.stack([self._assoc[neighbors], self._assoc[nodes]]), e_id, t

    def insert(self, src, dst, t):
        K = self.size
        neighbors = torch.cat([src, dst]); nodes = torch.cat([dst, src])
        e_id = torch.arange(self.cur_e_id, self.cur_e_id + src.size(0)).repeat(2)
        t = t.repeat(2)
        self.cur_e_id += src.numel()
        nodes, perm = nodes.sort(stable=True)
        neighbors, e_id, t = neighbors[perm], e_id[perm], t[perm]
        n_id = nodes.unique()
        self._assoc[n_id] = torch.arange(n_id.numel())
        dense_id = torch.arange(nodes.size(0)) % K + self._assoc[nodes] * K
        dense_e = e_id.new_full((n_id.numel() * K,), -1); dense_e[dense_id] = e_id
        dense_t = t.new_full((n_id.numel() * K,), -1); dense_t[dense_id] = t
        dense_n = e_id.new_zeros(n_id.numel() * K); dense_n[dense_id] = neighbors
        e_all = torch.cat([self.e_id[n_id], dense_e.view(-1, K)], dim=-1)
        t_all = torch.cat([self.t[n_id], dense_t.view(-1, K)], dim=-1)
        n_all = torch.cat([self.neighbors[n_id], dense_n.view(-1, K)], dim=-1)
        e_top, perm = e_all.topk(K, dim=-1)
        self.e_id[n_id] = e_top
        self.t[n_id] = t_all.topk(K, dim=-1).values
        self.neighbors[n_id] = torch.gather(n_all, 1, perm)


def build_model(raw_dim: int, hidden: int, num_nodes: int, seed: int = 1, aggregator="last"):
    """pyg_model_utils.getModel (pyg_model_utils.py:10-36) on the oracle classes."""
    torch.manual_seed(seed)
    aggr = LastAggregator() if aggregator == "last" else MeanAggregator()
    memory = TGNMemory(num_nodes, raw_dim, hidden, hidden, IdentityMessage(raw_dim, hidden, hidden), aggr)
    gnn = GraphAttentionEmbedding(hidden, hidden, raw_dim, memory.time_enc)
    link_pred = LinkPredictor(hidden)
    return {"memory": memory, "gnn": gnn, "link_pred": link_pred}


def model_parameters(model) -> List[torch.nn.Parameter]:
    seen, out = set(), []
    for m in ("memory", "gnn", "link_pred"):
        for p in model[m].parameters():
            if id(p) not in seen:
                seen.add(id(p)); out.append(p)
    return out


def train_step(model, loader: TorchNeighborLoader, optimizer, src, dst, neg, t, msg, data_t, data_msg,
               dropout: bool = True):
    """One training batch, PyG/TGB tgn.py order (commented flow at
    pyg_epoch_utils.py:106-137): sample -> memory -> embed -> decode -> BCE ->
    update_state -> insert -> backward -> step -> detach.  Returns the loss.
    data_t / data_msg are the full event arrays indexed by e_id."""
    memory, gnn, link_pred = model["memory"], model["gnn"], model["link_pred"]
    optimizer.zero_grad()
    n_id = torch.cat([src, dst, neg]).unique()
    n_id, edge_index, e_id, _ = loader(n_id)
    assoc = loader._assoc
    z, last_update = memory(n_id)
    if not dropout:
        gnn.conv.dropout = 0.0
    z = gnn(z, last_update, edge_index, data_t[e_id], data_msg[e_id])
    pos = link_pred.logits(z[assoc[src]], z[assoc[dst]])
    negs = link_pred.logits(z[assoc[src]], z[assoc[neg]])
    crit = torch.nn.BCEWithLogitsLoss()
    loss = crit(pos, torch.ones_like(pos)) + crit(negs, torch.zeros_like(negs))
    memory.update_state(src, dst, t, msg)
    loader.insert(src, dst, t.to(torch.float32))
    loss.backward()
    optimizer.step()
    memory.detach()
    return float(loss.detach())
