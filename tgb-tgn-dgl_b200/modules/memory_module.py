"""Drop-in `modules.memory_module.{TGNMemory, DyRepMemory}` (reference
modules/memory_module.py:25-215, :218-421) on the sm_100a kernels.

Same constructor arguments, buffers (`memory`, `last_update`, `_assoc`),
sub-module / parameter names (`time_enc.lin.*`, `memory_updater.weight_ih` ...),
methods and train/eval semantics as the reference:

  forward(n_id)            training: transient updated memory (with grad);
                           eval: plain gathers                       (:116-124)
  update_state(...)        training: update memory, then store;
                           eval: store, then update                  (:126-138)
  train(False)             flushes every node through the updater, then clears
                           the store                                 (:209-215)

What changes is the machinery.  The two Python dicts of per-node tuples
(:140-145,180-191) are a device-resident event log + per-node runs
(tgn_b200.ops.MsgStore, csrc/msgstore.cu).  With IdentityMessage +
Last/MeanAggregator + GRUCell -- what pyg_model_utils.getModel builds
(pyg_model_utils.py:14-21) -- `_get_updated_memory` is the fused path
gather -> concat -> time-encode -> aggregate -> GRU with no [M,472] message
matrix; any other message/aggregator module goes through the unfused path,
which materialises the stored tuples with the store kernels and calls the
user's modules exactly like memory_module.py:152-178,193-207.

Timestamps: `t` must be int64 as in the PyG/TGB loop; the reference itself
raises on float `t` at memory_module.py:150.  float32 `t` is accepted here as an
extension (stored values are truncated into `last_update` like `.long()`).
"""
import copy
from typing import Callable, Tuple

import torch
from torch import Tensor
from torch.nn import GRUCell, RNNCell

from modules.msg_agg import LastAggregator, MeanAggregator
from modules.msg_func import IdentityMessage
from modules.time_enc import TimeEncoder
from tgn_b200 import ops

_FLUSH_CHUNK = 1 << 16


class TGNMemory(torch.nn.Module):
    def __init__(self, num_nodes: int, raw_msg_dim: int, memory_dim: int, time_dim: int,
                 message_module: Callable, aggregator_module: Callable,
                 memory_updater_cell: str = "gru"):
        super().__init__()
        self.num_nodes = num_nodes
        self.raw_msg_dim = raw_msg_dim
        self.memory_dim = memory_dim
        self.time_dim = time_dim

        self.msg_s_module = message_module
        self.msg_d_module = copy.deepcopy(message_module)
        self.aggr_module = aggregator_module
        self.time_enc = TimeEncoder(time_dim)
        if memory_updater_cell == "gru":  # TGN
            self.memory_updater = GRUCell(message_module.out_channels, memory_dim)
        elif memory_updater_cell == "rnn":  # JODIE & DyRep
            self.memory_updater = RNNCell(message_module.out_channels, memory_dim)
        else:
            raise ValueError(
                "Undefined memory updater!!! Memory updater can be either 'gru' or 'rnn'.")

        self.register_buffer("memory", torch.empty(num_nodes, memory_dim))
        self.register_buffer("last_update", torch.empty(num_nodes, dtype=torch.long))
        self.register_buffer("_assoc", torch.empty(num_nodes, dtype=torch.long))
        self._store = None
        self.reset_parameters()

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self) -> torch.device:
        return self.time_enc.lin.weight.device

    @property
    def store(self) -> ops.MsgStore:
        """The tensorised msg_s_store / msg_d_store pair."""
        if self._store is None or self._store.device != self.memory.device:
            if self.memory.device.type != "cuda":
                raise RuntimeError("TGNMemory (B200 build) runs on CUDA only; move the module to the GPU first")
            self._store = ops.MsgStore(self.num_nodes, self.raw_msg_dim, self.memory.device)
        return self._store

    def _fused_mode(self):
        """agg mode for the fused kernel, or None if the modules need the unfused path."""
        if not (type(self.msg_s_module) is IdentityMessage and type(self.msg_d_module) is IdentityMessage):
            return None
        if not isinstance(self.memory_updater, GRUCell):
            return None
        if type(self.aggr_module) is LastAggregator:
            return ops.AGG_LAST
        if type(self.aggr_module) is MeanAggregator:
            # the fused backward covers LastAggregator; Mean trains through the unfused path
            return None if (self.training and torch.is_grad_enabled()) else ops.AGG_MEAN
        return None

    # ------------------------------------------------------------------ reference API
    def reset_parameters(self):
        for m in (self.msg_s_module, self.msg_d_module, self.aggr_module):
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.time_enc.reset_parameters()
        self.memory_updater.reset_parameters()
        self.reset_state()

    def reset_state(self):
        """Resets the memory to its initial state (memory_module.py:106-110)."""
        self.memory.data.fill_(0)
        self.last_update.data.fill_(0)
        self._reset_message_store()

    def detach(self):
        self.memory.detach_()

    def forward(self, n_id: Tensor) -> Tuple[Tensor, Tensor]:
        if self.training:
            return self._get_updated_memory(n_id)
        n_id = n_id.to(self.memory.device, torch.long)
        return ops.gather_rows(self.memory, n_id), self.last_update[n_id]

    def update_state(self, src: Tensor, dst: Tensor, t: Tensor, raw_msg: Tensor):
        dev = self.memory.device
        src, dst = src.to(dev, torch.long), dst.to(dev, torch.long)
        t, raw_msg = t.to(dev), raw_msg.to(dev, torch.float32)
        n_id = ops.unique_relabel([src, dst], self.num_nodes)
        if self.training:
            self._update_memory(n_id)
            self.store.update(src, dst, t, raw_msg)
        else:
            self.store.update(src, dst, t, raw_msg)
            self._update_memory(n_id)

    def _reset_message_store(self):
        if self._store is not None:
            self._store.reset()

    def _update_memory(self, n_id: Tensor):
        with torch.no_grad():
            memory, last_update = self._get_updated_memory(n_id)
            ops.memory_scatter(n_id, memory, last_update.contiguous(), self.memory, self.last_update)

    def _get_updated_memory(self, n_id: Tensor) -> Tuple[Tensor, Tensor]:
        n_id = n_id.to(self.memory.device, torch.long).contiguous()
        mode = self._fused_mode()
        if mode is not None:
            cell = self.memory_updater
            return ops.memory_update(self.time_enc.lin.weight.view(-1), self.time_enc.lin.bias,
                                     cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh,
                                     self.store, n_id, self.memory, self.last_update, mode)
        return self._get_updated_memory_unfused(n_id)

    # ------------------------------------------------------------------ unfused path
    def _compute_msg(self, n_id: Tensor, direction: int, msg_module: Callable):
        """memory_module.py:193-207 on the tensorised store."""
        src, dst, t, raw_msg = self.store.gather(n_id, direction)
        t_rel = t - self.last_update[src]
        t_enc = self.time_enc(t_rel.to(raw_msg.dtype))
        msg = msg_module(ops.gather_rows(self.memory, src), ops.gather_rows(self.memory, dst), raw_msg, t_enc)
        return msg, t, src, dst

    def _apply_updater(self, aggr: Tensor, h: Tensor) -> Tensor:
        cell = self.memory_updater
        if isinstance(cell, GRUCell):
            return ops.gru_cell(aggr, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
        if torch.is_grad_enabled() and any(p.requires_grad for p in cell.parameters()):
            # RNNCell training (JODIE/DyRep variant): gate math via torch autograd on the sgemm linears
            return torch.tanh(ops.linear(aggr, cell.weight_ih, cell.bias_ih) +
                              ops.linear(h, cell.weight_hh, cell.bias_hh))
        return ops.rnn_cell(aggr, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)

    def _get_updated_memory_unfused(self, n_id: Tensor) -> Tuple[Tensor, Tensor]:
        S = n_id.size(0)
        self._assoc[n_id] = torch.arange(S, device=n_id.device)
        msg_s, t_s, src_s, _ = self._compute_msg(n_id, 0, self.msg_s_module)
        msg_d, t_d, src_d, _ = self._compute_msg(n_id, 1, self.msg_d_module)
        idx = torch.cat([src_s, src_d], dim=0)
        msg = torch.cat([msg_s, msg_d], dim=0)
        t = torch.cat([t_s, t_d], dim=0)
        aggr = self.aggr_module(msg, self._assoc[idx], t, S)
        memory = self._apply_updater(aggr, ops.gather_rows(self.memory, n_id))
        # scatter(t, idx, dim_size=N, reduce='max')[n_id]  (memory_module.py:175-176): 0 where no message
        lu = torch.zeros(S, dtype=t.dtype, device=t.device)
        if t.numel():
            lu = lu.scatter_reduce(0, self._assoc[idx], t, "amax", include_self=False)
        return memory, lu

    # ------------------------------------------------------------------ train / eval switch
    def train(self, mode: bool = True):
        """Entering eval flushes the message store into the memory (memory_module.py:209-215)."""
        if self.training and not mode and self._store is not None:
            self._flush_all()
            self._reset_message_store()
        return super().train(mode)

    def _flush_all(self):
        # all rows are computed from the pre-flush state, then written (the reference computes
        # the whole [N, .] update before assigning it)
        with torch.no_grad():
            dev = self.memory.device
            new_mem = torch.empty_like(self.memory)
            new_lu = torch.empty_like(self.last_update)
            for lo in range(0, self.num_nodes, _FLUSH_CHUNK):
                ids = torch.arange(lo, min(self.num_nodes, lo + _FLUSH_CHUNK), device=dev)
                m, lu = self._get_updated_memory(ids)
                new_mem[lo:lo + ids.numel()] = m
                new_lu[lo:lo + ids.numel()] = lu.to(torch.long)
            self.memory.copy_(new_mem)
            self.last_update.copy_(new_lu)


class DyRepMemory(TGNMemory):
    """reference modules/memory_module.py:218-421: TGNMemory whose messages may use the
    current *embeddings* of the batch nodes instead of their memory (:389-408)."""

    def __init__(self, num_nodes: int, raw_msg_dim: int, memory_dim: int, time_dim: int,
                 message_module: Callable, aggregator_module: Callable, memory_updater_type: str,
                 use_src_emb_in_msg: bool = False, use_dst_emb_in_msg: bool = False):
        assert memory_updater_type in ["gru", "rnn"], "Memor updater can be either `rnn` or `gru`."
        super().__init__(num_nodes, raw_msg_dim, memory_dim, time_dim, message_module,
                         aggregator_module, memory_updater_type)
        self.use_src_emb_in_msg = use_src_emb_in_msg
        self.use_dst_emb_in_msg = use_dst_emb_in_msg
        self._emb = None

    def _fused_mode(self):
        if self._emb is not None and (self.use_src_emb_in_msg or self.use_dst_emb_in_msg):
            return None
        return super()._fused_mode()

    def update_state(self, src, dst, t, raw_msg, embeddings: Tensor = None, assoc: Tensor = None):
        self._emb = (embeddings, assoc) if embeddings is not None else None
        try:
            super().update_state(src, dst, t, raw_msg)
        finally:
            self._emb = None

    def _compute_msg(self, n_id: Tensor, direction: int, msg_module: Callable):
        src, dst, t, raw_msg = self.store.gather(n_id, direction)
        t_rel = t - self.last_update[src]
        t_enc = self.time_enc(t_rel.to(raw_msg.dtype))
        z_src, z_dst = ops.gather_rows(self.memory, src), ops.gather_rows(self.memory, dst)
        if self._emb is not None:
            emb, assoc = self._emb
            if self.use_src_emb_in_msg and src.numel():
                hit = torch.isin(src, n_id)
                z_src[hit] = emb[assoc[src[hit]]]
            if self.use_dst_emb_in_msg and dst.numel():
                hit = torch.isin(dst, n_id)
                z_dst[hit] = emb[assoc[dst[hit]]]
        return msg_module(z_src, z_dst, raw_msg, t_enc), t, src, dst
