"""TEST INFRASTRUCTURE -- not part of the product path.

Restatements of the third-party arithmetic the reference *calls* but does not
ship (none of these packages is installed here, no version is pinned by the
reference, and there is no network):

  torch_scatter.scatter_max          used at reference modules/msg_agg.py:12,17
  torch_geometric.utils.scatter      used at modules/msg_agg.py:11,26, memory_module.py:17,176
  torch_geometric.nn.inits.zeros     used at modules/memory_module.py:16,108-109
  torch_geometric.nn.TransformerConv used at modules/emb_module.py:7,21-23,29
  modules.time_enc.TimeEncoder       imported at modules/memory_module.py:19 but the file is
                                     ABSENT from the reference tree (TGB/PyG upstream:
                                     Linear(1, out) then cos)

PARITY UNPINNED for these five: they follow the published upstream algorithms
(SURVEY.md Appendix B3-B5), with no golden vectors from the real packages.  The
reference's OWN logic built on top of them (memory_module.py, msg_agg.py,
emb_module.py, neighbor_loader.py) is pinned: tests/golden/make_golden.py runs
the real reference files on top of these functions and freezes the outputs.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def scatter_max(src: Tensor, index: Tensor, dim: int = 0, dim_size: int | None = None):
    """torch_scatter.scatter_max, CPU semantics: sequential scan with strict `>`,
    so the FIRST maximal element of a segment wins; empty segments give value 0
    and argmax == src.size(0) (the sentinel modules/msg_agg.py:19 tests for).

    Vectorised (the real package runs a C++ loop here, so a per-element Python loop
    would handicap the CPU baseline): two stable sorts -- by value descending, then by
    segment -- leave, at the head of every segment, its largest value with the smallest
    original position among equals, i.e. exactly the winner of the strict-`>` scan."""
    assert dim == 0 and src.dim() == 1
    n = src.size(0)
    size = int(dim_size if dim_size is not None else (int(index.max()) + 1 if n else 0))
    out = torch.zeros(size, dtype=src.dtype)
    arg = torch.full((size,), n, dtype=torch.long)
    if n:
        order = torch.argsort(src, descending=True, stable=True)
        order = order[torch.argsort(index[order], stable=True)]
        seg = index[order]
        head = torch.ones(n, dtype=torch.bool)
        head[1:] = seg[1:] != seg[:-1]
        win = order[head]
        out[seg[head]] = src[win]
        arg[seg[head]] = win
    return out, arg


def scatter_max_loop(src: Tensor, index: Tensor, dim: int = 0, dim_size: int | None = None):
    """The same rule as an explicit sequential scan (small cases; pins scatter_max in tests)."""
    n = src.size(0)
    size = int(dim_size if dim_size is not None else (int(index.max()) + 1 if n else 0))
    out = torch.zeros(size, dtype=src.dtype)
    arg = torch.full((size,), n, dtype=torch.long)
    best = [None] * size
    s, ix = src.tolist(), index.tolist()
    for i in range(n):
        k = ix[i]
        if best[k] is None or s[i] > best[k]:
            best[k] = s[i]
            arg[k] = i
    for k in range(size):
        if best[k] is not None:
            out[k] = best[k]
    return out, arg


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: int | None = None,
            reduce: str = "sum") -> Tensor:
    """torch_geometric.utils.scatter for dim=0: 'mean' = index_add then divide by
    count.clamp(min=1); 'max' = zeros.scatter_reduce_(amax, include_self=False)."""
    assert dim == 0
    size = int(dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0))
    shape = (size,) + tuple(src.shape[1:])
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype).index_add_(0, index, src)
    if reduce == "mean":
        out = torch.zeros(shape, dtype=src.dtype).index_add_(0, index, src)
        count = torch.zeros(size, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        count = count.clamp(min=1).view((size,) + (1,) * (src.dim() - 1))
        return out / count
    if reduce == "max":
        idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
        return torch.zeros(shape, dtype=src.dtype).scatter_reduce_(0, idx, src, "amax", include_self=False)
    raise ValueError(reduce)


def zeros(t: Tensor) -> None:
    if t is not None:
        t.data.fill_(0)


class TimeEncoder(torch.nn.Module):
    """cos(Linear(1, out)(t)) -- contract recovered from the reference's use
    sites (memory_module.py:69,92,102,203; emb_module.py:19-20,27)."""

    def __init__(self, out_channels: int):
        super().__init__()
        self.out_channels = out_channels
        self.lin = torch.nn.Linear(1, out_channels)

    def reset_parameters(self):
        self.lin.reset_parameters()

    def forward(self, t: Tensor) -> Tensor:
        return self.lin(t.view(-1, 1)).cos()


def _softmax_by_target(score: Tensor, target: Tensor, num_nodes: int) -> Tensor:
    """softmax over the edges that share a target node (PyG `softmax(src, index)`)."""
    H = score.size(1)
    idx = target.view(-1, 1).expand(-1, H)
    mx = torch.full((num_nodes, H), float("-inf"), dtype=score.dtype)
    mx = mx.scatter_reduce_(0, idx, score, "amax", include_self=True)
    ex = (score - mx[target]).exp()
    den = torch.zeros((num_nodes, H), dtype=score.dtype).index_add_(0, target, ex)
    return ex / (den[target] + 1e-16)


class TransformerConv(torch.nn.Module):
    """torch_geometric.nn.TransformerConv restricted to the configuration the
    reference instantiates (concat=True, beta=False, root_weight=True, bias=True,
    edge_dim given) -- SURVEY.md B5."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, dropout: float = 0.0,
                 edge_dim: int | None = None):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.dropout, self.edge_dim = dropout, edge_dim
        hc = heads * out_channels
        self.lin_key = torch.nn.Linear(in_channels, hc)
        self.lin_query = torch.nn.Linear(in_channels, hc)
        self.lin_value = torch.nn.Linear(in_channels, hc)
        self.lin_edge = torch.nn.Linear(edge_dim, hc, bias=False)
        self.lin_skip = torch.nn.Linear(in_channels, hc, bias=True)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor) -> Tensor:
        H, C = self.heads, self.out_channels
        src, dst = edge_index[0], edge_index[1]  # messages flow j=src -> i=dst
        q = self.lin_query(x).view(-1, H, C)
        k = self.lin_key(x).view(-1, H, C)
        v = self.lin_value(x).view(-1, H, C)
        e = self.lin_edge(edge_attr).view(-1, H, C)
        kj = k[src] + e
        score = (q[dst] * kj).sum(-1) / math.sqrt(C)
        alpha = _softmax_by_target(score, dst, x.size(0))
        alpha = torch.nn.functional.dropout(alpha, p=self.dropout, training=self.training)
        msg = (v[src] + e) * alpha.unsqueeze(-1)
        out = torch.zeros((x.size(0), H, C), dtype=x.dtype).index_add_(0, dst, msg)
        return out.reshape(-1, H * C) + self.lin_skip(x)
