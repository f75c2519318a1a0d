"""TEST INFRASTRUCTURE -- a minimal stand-in for the `dgl` package (not installed in this image, no
version pinned by the reference, no network), just large enough to import the reference's
model_utils.py / dgl_utils.py UNMODIFIED and run its DGL-flavoured hot-path classes
(TimeEncode, MemoryOperation, TemporalEdgePreprocess, EdgeGATConv, TemporalTransformerConv, TGNN) on the
CPU so that tests/golden/make_golden.py can freeze their outputs (tests/golden/dgl_twins.npz).

PARITY UNPINNED for this layer: the message-passing semantics below are restated from DGL's published
behaviour -- `apply_edges` / `update_all` with UDFs (degree bucketing, mailbox ordered by
edge id, zero-in-degree nodes reduce to zeros and still run the apply function), the `u_add_e`, `e_add_v`,
`sum`, `copy_e` built-ins, `edge_softmax` by destination, `in_subgraph` (all nodes kept, only the in-edges
of the given nodes, `edata['_ID']` = parent edge ids), `add_edges`, `add_self_loop` (loops appended after
the existing edges).  Everything the REFERENCE implements on top of them runs from its own source.
Never imported by the product path.
"""
from contextlib import contextmanager

import torch

from . import base, function, ops  # noqa: F401

NID = "_ID"
EID = "_ID"


class _Frame(dict):
    pass


class _EdgeBatch:
    def __init__(self, g, eids=None):
        sel = slice(None) if eids is None else eids
        self.src = {k: v[g._src[sel]] for k, v in g.ndata.items()}
        self.dst = {k: v[g._dst[sel]] for k, v in g.ndata.items()}
        self.data = {k: v[sel] for k, v in g.edata.items()}
        self._n = g._src[sel].numel()

    def __len__(self):
        return self._n


class _NodeBatch:
    def __init__(self, data, mailbox=None):
        self.data, self.mailbox = data, mailbox


class DGLGraph:
    def __init__(self, src, dst, num_nodes):
        self._src, self._dst = src.long(), dst.long()
        self._n = int(num_nodes)
        self.ndata, self.edata = _Frame(), _Frame()

    # ---- structure
    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def add_edges(self, u, v):
        self._src = torch.cat([self._src, u.long()])
        self._dst = torch.cat([self._dst, v.long()])
        for k, val in list(self.edata.items()):
            pad = torch.zeros((u.numel(),) + tuple(val.shape[1:]), dtype=val.dtype)
            self.edata[k] = torch.cat([val, pad])

    @contextmanager
    def local_scope(self):
        nd, ed = _Frame(self.ndata), _Frame(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    def local_var(self):
        g = DGLGraph(self._src, self._dst, self._n)
        g.ndata, g.edata = _Frame(self.ndata), _Frame(self.edata)
        return g

    # ---- message passing
    def apply_edges(self, func):
        out = func(self) if isinstance(func, function.BuiltinEdge) else func(_EdgeBatch(self))
        self.edata.update(out)

    def update_all(self, message_func, reduce_func, apply_node_func=None):
        msgs = message_func(self) if isinstance(message_func, function.BuiltinEdge) else message_func(_EdgeBatch(self))
        if isinstance(reduce_func, function.BuiltinReduce):
            red = reduce_func(self, msgs)
        else:
            red = self._bucketed_reduce(msgs, reduce_func)
        self.ndata.update(red)
        if apply_node_func is not None:
            self.ndata.update(apply_node_func(_NodeBatch(self.ndata)))

    def _bucketed_reduce(self, msgs, reduce_func):
        deg = self.in_degrees()
        order = torch.argsort(self._dst, stable=True)            # per destination: edges in edge-id order
        start = torch.cumsum(deg, 0) - deg
        out = {}
        for d in sorted(set(deg.tolist()) - {0}):
            nodes = (deg == d).nonzero(as_tuple=True)[0]
            eids = order[(start[nodes].view(-1, 1) + torch.arange(d).view(1, -1)).view(-1)]
            mailbox = {k: v[eids].view((nodes.numel(), d) + tuple(v.shape[1:])) for k, v in msgs.items()}
            data = {k: v[nodes] for k, v in self.ndata.items()}
            res = reduce_func(_NodeBatch(data, mailbox))
            for k, v in res.items():
                if k not in out:
                    out[k] = torch.zeros((self._n,) + tuple(v.shape[1:]), dtype=v.dtype)
                out[k] = out[k].index_put((nodes,), v)
        return out


def graph(data, num_nodes=None):
    src, dst = data
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
    return DGLGraph(src, dst, num_nodes)


def add_self_loop(g):
    out = DGLGraph(torch.cat([g._src, torch.arange(g._n)]), torch.cat([g._dst, torch.arange(g._n)]), g._n)
    out.ndata = _Frame(g.ndata)
    for k, v in g.edata.items():
        out.edata[k] = torch.cat([v, torch.zeros((g._n,) + tuple(v.shape[1:]), dtype=v.dtype)])
    return out


def in_subgraph(g, nodes):
    keep = torch.zeros(g._n, dtype=torch.bool)
    keep[torch.as_tensor(nodes).long()] = True
    eids = keep[g._dst].nonzero(as_tuple=True)[0]
    sub = DGLGraph(g._src[eids], g._dst[eids], g._n)
    sub.ndata = _Frame(g.ndata)
    sub.edata = _Frame({k: v[eids] for k, v in g.edata.items()})
    sub.edata["_ID"] = eids
    return sub
