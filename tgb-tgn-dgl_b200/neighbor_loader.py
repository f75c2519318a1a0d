"""Drop-in for the reference's `neighbor_loader.LastNeighborLoader`
(neighbor_loader.py:15-109) backed by the sm_100a ring kernels
(csrc/nbr_ring.cu, csrc/unique.cu) through the C-ABI.

Same constructor, call signature, return tuple, public attributes
(`size, neighbors, e_id, t, _assoc, cur_e_id`) and dtypes as the reference, so
`pyg-mem-tgn.py` and `epoch_utils.py` (which reads `_assoc` directly,
epoch_utils.py:99,262) run on it unchanged.  CUDA only: there is no CPU path.
"""
from typing import Tuple

import torch
from torch import Tensor

from tgn_b200 import ops


class LastNeighborLoader:
    def __init__(self, num_nodes: int, size: int, device=None):
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise RuntimeError("LastNeighborLoader (B200 build) needs a CUDA device; there is no CPU fallback")
        self.size = size
        self.num_nodes = num_nodes
        print("Total number of nodes: ", num_nodes)  # the reference prints this (neighbor_loader.py:18)
        self.neighbors = torch.zeros((num_nodes, size), dtype=torch.long, device=device)
        self.e_id = torch.empty((num_nodes, size), dtype=torch.long, device=device)
        self.t = torch.empty((num_nodes, size), dtype=torch.float, device=device)
        self._assoc = torch.zeros(num_nodes, dtype=torch.long, device=device)
        self.reset_state()

    def __call__(self, n_id: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        n_id = n_id.to(self.neighbors.device, torch.long)
        ids, edge_index, e_id, t, root_off = ops.nbr_lookup(n_id, self.neighbors, self.e_id, self.t,
                                                            self._assoc)
        # CSR of the edges by centre (edges leave the kernel grouped by root, in
        # root order); GraphAttentionEmbedding picks this up to skip regrouping.
        # Only valid when every root occurs once (all reference callers pass `.unique()` output).
        if n_id.numel() and bool((n_id[1:] > n_id[:-1]).all()):
            edge_index._tgn_csr = (root_off, ops.relabel(n_id, self._assoc))
        return ids, edge_index, e_id, t

    def insert(self, src: Tensor, dst: Tensor, t: Tensor = None):
        if t is None:
            raise TypeError("insert() needs the event timestamps t (the reference dereferences t unconditionally, neighbor_loader.py:63)")
        dev = self.neighbors.device
        src, dst = src.to(dev, torch.long), dst.to(dev, torch.long)
        ops.nbr_insert(src, dst, t.to(dev, torch.float), self.cur_e_id, self.neighbors, self.e_id, self.t)
        self.cur_e_id += src.numel()

    def reset_state(self):
        self.cur_e_id = 0
        self.e_id.fill_(-1)
        self.t.fill_(-1)
