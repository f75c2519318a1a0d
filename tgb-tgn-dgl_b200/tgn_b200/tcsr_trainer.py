"""TGN training step over the t-CSR graph with TGL-style multi-layer sampling (BASELINE.json configs[3]:
uniform-20 sampling, 2 attention layers; the reference's config keys `sampling.layer / neighbor / strategy`,
config/TGN.yml:1-9, and `gnn.layer`, :21 -- parsed by utils.parse_config, never consumed by the reference's own
model: its multi-layer code is commented out, model_utils.py:669-686,694-696, and it has no uniform sampler.
TGL upstream defines the semantics).

Per batch:
    roots      = [src | dst | neg] at the batch's timestamps
    blocks     = sampler_core.ParallelSampler(t-CSR).sample_device(roots)     csrc/tcsr.cu: per layer, k
                 neighbours strictly earlier than the root's time -- most recent, or uniform draws with
                 replacement (Philox) -- the sampled neighbours are the next layer's roots (with THEIR times)
    memory     = TGNMemory(unique nodes of the innermost block)               fused gather+concat+GRU op
    embedding  = L x GraphAttentionEmbedding, innermost block first: layer l runs on block l and hands the
                 rows of its centres (= the nodes of block l-1) to layer l-1
    decoder    = LinkPredictor on the root rows, BCE-with-logits, backward (autograd through the CUDA ops),
                 Adam, memory.update_state(src, dst, t, msg)

The graph is static (every event of the split is in the t-CSR, reverse edges included); causality comes from
the sampler's `ts < root time` search, so there is no neighbour ring to maintain.  Everything numeric is a
kernel behind the C-ABI (tgn_tcsr_sample, tgn_msg_build / tgn_gru_*, tgn_attn_*, tgn_gemm_*); the composition
is the drop-in module path, i.e. this is what a TGL-style script gets on this package -- the captured-graph
engine (tgn_b200.engine) covers the ring-sampled single-layer configuration of the reference.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import ops


class TCSRTrainer:
    def __init__(self, indptr: Tensor, indices: Tensor, eid: Tensor, ts: Tensor, num_nodes: int, raw_dim: int,
                 hidden: int, num_neighbors: List[int], recent: bool, edge_feats: Tensor, device="cuda",
                 lr: float = 1e-4, dropout: float = 0.1, seed: int = 0):
        import sampler_core
        from modules.decoder import LinkPredictor
        from modules.emb_module import GraphAttentionEmbedding
        from modules.memory_module import TGNMemory
        from modules.msg_agg import LastAggregator
        from modules.msg_func import IdentityMessage
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("TCSRTrainer runs on CUDA only (no CPU fallback)")
        self.dev, self.N, self.L = dev, num_nodes, len(num_neighbors)
        self.sampler = sampler_core.ParallelSampler(indptr, indices, eid, ts, 1, 1, self.L, list(num_neighbors), recent,
                                                    False, 1, 0.0, device=dev, seed=seed)
        self.memory = TGNMemory(num_nodes, raw_dim, hidden, hidden, IdentityMessage(raw_dim, hidden, hidden),
                                LastAggregator()).to(dev)
        # layers[0] is the OUTERMOST layer (runs last, produces the root embeddings)
        self.layers = [GraphAttentionEmbedding(hidden, hidden, raw_dim, self.memory.time_enc).to(dev)
                       for _ in range(self.L)]
        for g in self.layers:
            g.conv.dropout = dropout
        self.link_pred = LinkPredictor(hidden).to(dev)
        self.feats = edge_feats.to(dev, torch.float32).contiguous()       # [num_events, raw_dim], row = eid
        params, seen = [], set()
        for m in [self.memory] + self.layers + [self.link_pred]:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        self.params = params
        self.opt = torch.optim.Adam(params, lr=lr)
        self.crit = torch.nn.BCEWithLogitsLoss()
        self.assoc = torch.zeros(num_nodes, dtype=torch.long, device=dev)
        self.sampled_edges = 0

    def modules(self):
        return [self.memory] + self.layers + [self.link_pred]

    def train(self, mode: bool = True):
        for m in self.modules():
            m.train(mode)

    def reset_state(self):
        self.memory.reset_state()
        self.sampler.reset()

    # ------------------------------------------------------------------ one batch
    def embed(self, roots: Tensor, root_ts: Tensor) -> Tensor:
        """Embeddings of `roots` at times `root_ts` ([R, hidden]) through the L sampled blocks."""
        blocks = self.sampler.sample_device(roots.to(self.dev, torch.int32), root_ts.to(self.dev, torch.float32))
        counts = [int(b["count"].item()) for b in blocks]
        self.sampled_edges += sum(counts)
        inner = blocks[-1]
        n_in = counts[-1]
        nodes_in = torch.cat([inner["roots"].long(), inner["nbr"][:n_in].long()])
        n_id = ops.unique_relabel([nodes_in], self.N, self.assoc)
        z, last_update = self.memory(n_id)
        pos = self.assoc[nodes_in]
        h, lu = z[pos], last_update[pos]
        for layer in range(self.L - 1, -1, -1):
            b, n = blocks[layer], counts[layer]
            R = b["roots"].numel()
            edge_index = torch.stack([R + torch.arange(n, device=self.dev), b["col"][:n].long()])
            edge_index._tgn_csr = (b["root_off"], torch.arange(R, device=self.dev))   # edges are grouped by root
            h = self.layers[layer](h[:R + n], lu[:R + n], edge_index, b["ts"][:n], self.feats[b["eid"][:n].long()])[:R]
            # the centres of block l are the nodes of block l-1 (its roots followed by its neighbours)
        return h

    def train_step(self, src: Tensor, dst: Tensor, neg: Tensor, t: Tensor, msg: Tensor) -> Tensor:
        dev, B = self.dev, src.numel()
        src, dst, neg = src.to(dev, torch.long), dst.to(dev, torch.long), neg.to(dev, torch.long)
        t = t.to(dev)
        self.opt.zero_grad()
        roots = torch.cat([src, dst, neg])
        z = self.embed(roots, t.to(torch.float32).repeat(3))
        lp = self.link_pred
        pos = lp.logits(z[:B], z[B:2 * B])
        ngo = lp.logits(z[:B], z[2 * B:])
        loss = self.crit(pos, torch.ones_like(pos)) + self.crit(ngo, torch.zeros_like(ngo))
        self.memory.update_state(src, dst, t.long(), msg.to(dev, torch.float32))
        loss.backward()
        self.opt.step()
        self.memory.detach()
        return loss.detach()
