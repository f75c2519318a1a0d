"""Drop-in `modules.emb_module.{GraphAttentionEmbedding, TimeEmbedding}`
(reference modules/emb_module.py:11-52).

GraphAttentionEmbedding keeps the reference's parameter tree --
`conv.lin_key / lin_query / lin_value / lin_edge / lin_skip` with PyG
TransformerConv's shapes (heads=2, out_channels//2 per head, dropout=0.1,
edge_dim = msg_dim + time_dim; emb_module.py:19-23) -- so state_dicts are
interchangeable, but forward() is one node-projection GEMM plus the fused
time-encode + attention kernel (csrc/attn.cu): edge_attr = [cos(w*rel_t+b), msg]
(emb_module.py:26-28) is never written to memory.
"""
import math

import torch
from torch import Tensor

from tgn_b200 import ops


class TransformerConv(torch.nn.Module):
    """Parameter container with torch_geometric.nn.TransformerConv's attribute names for
    the configuration the reference uses (concat=True, beta=False, root_weight=True)."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, dropout: float = 0.0,
                 edge_dim: int = None):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.dropout, self.edge_dim = dropout, edge_dim
        hc = heads * out_channels
        self.lin_key = torch.nn.Linear(in_channels, hc)
        self.lin_query = torch.nn.Linear(in_channels, hc)
        self.lin_value = torch.nn.Linear(in_channels, hc)
        self.lin_edge = torch.nn.Linear(edge_dim, hc, bias=False)
        self.lin_skip = torch.nn.Linear(in_channels, hc, bias=True)

    def reset_parameters(self):
        for lin in (self.lin_key, self.lin_query, self.lin_value, self.lin_edge, self.lin_skip):
            lin.reset_parameters()

    def packed_node_weights(self):
        """[query; key; value; skip] -- the row order csrc/attn.cu expects."""
        w = torch.cat([self.lin_query.weight, self.lin_key.weight, self.lin_value.weight,
                       self.lin_skip.weight], dim=0)
        b = torch.cat([self.lin_query.bias, self.lin_key.bias, self.lin_value.bias,
                       self.lin_skip.bias], dim=0)
        return w, b


class GraphAttentionEmbedding(torch.nn.Module):
    def __init__(self, in_channels, out_channels, msg_dim, time_enc):
        super().__init__()
        self.time_enc = time_enc
        edge_dim = msg_dim + time_enc.out_channels
        self.conv = TransformerConv(in_channels, out_channels // 2, heads=2, dropout=0.1,
                                    edge_dim=edge_dim)

    def forward(self, x, last_update, edge_index, t, msg):
        conv = self.conv
        dev = x.device
        w_node, b_node = conv.packed_node_weights()
        csr = getattr(edge_index, "_tgn_csr", None)
        if csr is not None:  # edges straight from LastNeighborLoader: already grouped by centre
            row_ptr, centre_ids = csr
            edge_perm = None
        else:
            row_ptr, edge_perm = ops.group_edges_by_centre(edge_index[1], x.size(0))
            centre_ids = None
        p = conv.dropout if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        t = t.to(dev)
        if t.dtype not in (torch.int64, torch.float32):
            t = t.to(torch.float32)
        return ops.temporal_attention(
            x.to(torch.float32), w_node, b_node, conv.lin_edge.weight,
            self.time_enc.lin.weight.view(-1), self.time_enc.lin.bias, last_update,
            edge_index[0], t, msg.to(dev, torch.float32), row_ptr, heads=conv.heads,
            edge_perm=edge_perm, centre_ids=centre_ids, dropout_p=p, seed=seed)


class TimeEmbedding(torch.nn.Module):
    """JODIE-style projection (reference modules/emb_module.py:32-52)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels

        class NormalLinear(torch.nn.Linear):
            def reset_parameters(self):
                stdv = 1.0 / math.sqrt(self.weight.size(1))
                self.weight.data.normal_(0, stdv)
                if self.bias is not None:
                    self.bias.data.normal_(0, stdv)

        self.embedding_layer = NormalLinear(1, self.out_channels)

    def forward(self, x, last_update, t):
        rel_t = (last_update - t).to(x.dtype).unsqueeze(1).contiguous()
        return x * (1 + ops.linear(rel_t, self.embedding_layer.weight, self.embedding_layer.bias))
