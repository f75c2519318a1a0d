"""`modules.decoder.LinkPredictor` (reference modules/decoder.py:12-27): two
linears + relu + linear + sigmoid, with the GEMMs on tgn_sgemm.  `logits()` is
the pre-sigmoid score: pyg-mem-tgn.py trains with BCEWithLogitsLoss
(pyg-mem-tgn.py:51), so the drop-in model hands it logits."""
import torch
from torch.nn import Linear

from tgn_b200 import ops


class LinkPredictor(torch.nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.lin_src = Linear(in_channels, in_channels)
        self.lin_dst = Linear(in_channels, in_channels)
        self.lin_final = Linear(in_channels, 1)

    def logits(self, z_src, z_dst):
        h = ops.linear(z_src, self.lin_src.weight, self.lin_src.bias)
        h = h + ops.linear(z_dst, self.lin_dst.weight, self.lin_dst.bias)
        return ops.linear(h.relu(), self.lin_final.weight, self.lin_final.bias)

    def forward(self, z_src, z_dst):
        return self.logits(z_src, z_dst).sigmoid()


class NodePredictor(torch.nn.Module):
    """reference modules/decoder.py:30-41"""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.lin_node = Linear(in_dim, in_dim)
        self.out = Linear(in_dim, out_dim)

    def forward(self, node_embed):
        h = ops.linear(node_embed, self.lin_node.weight, self.lin_node.bias).relu()
        return ops.linear(h, self.out.weight, self.out.bias)
