"""A minimal edge-list graph object for the DGL-flavoured twins (dgl_model_utils.py).

The reference's live stack passes DGLGraph objects around (model_utils.py, dgl_utils.py); DGL is not part of
this image.  The twins only need what is below -- `edges()`, `num_nodes()`, `num_edges()`, `in_degrees()`,
the `ndata` / `edata` frames, `local_var()` / `local_scope()` -- and they duck-type their argument, so a real
`dgl.DGLGraph` works in place of this class where DGL is installed."""
from contextlib import contextmanager

import torch

NID = "_ID"


class Graph:
    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int):
        self._src, self._dst, self._n = src.long(), dst.long(), int(num_nodes)
        self.ndata, self.edata = {}, {}

    def edges(self):
        return self._src, self._dst

    def num_nodes(self) -> int:
        return self._n

    def num_edges(self) -> int:
        return int(self._src.numel())

    def in_degrees(self) -> torch.Tensor:
        return torch.bincount(self._dst, minlength=self._n)

    def add_edges(self, u, v):
        self._src = torch.cat([self._src, u.long()])
        self._dst = torch.cat([self._dst, v.long()])

    def local_var(self) -> "Graph":
        g = Graph(self._src, self._dst, self._n)
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    @contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed


def graph(data, num_nodes=None) -> Graph:
    """`dgl.graph((src, dst), num_nodes=...)` (dgl_utils.py:4)."""
    src, dst = data
    if num_nodes is None:
        num_nodes = int(max(int(src.max()), int(dst.max()))) + 1 if src.numel() else 0
    return Graph(src, dst, num_nodes)
