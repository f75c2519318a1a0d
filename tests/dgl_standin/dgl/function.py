"""dgl.function built-ins used by the reference (model_utils.py:594-599): u_add_e, e_add_v, sum, copy_e."""
import torch


class BuiltinEdge:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, g):
        return self.fn(g)


class BuiltinReduce:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, g, msgs):
        return self.fn(g, msgs)


def u_add_e(u, e, out):
    return BuiltinEdge(lambda g: {out: g.ndata[u][g._src] + g.edata[e]})


def e_add_v(e, v, out):
    return BuiltinEdge(lambda g: {out: g.edata[e] + g.ndata[v][g._dst]})


def copy_e(e, out):
    return BuiltinEdge(lambda g: {out: g.edata[e]})


def sum(msg, out):  # noqa: A001
    def red(g, msgs):
        m = msgs[msg]
        return {out: torch.zeros((g._n,) + tuple(m.shape[1:]), dtype=m.dtype).index_add(0, g._dst, m)}
    return BuiltinReduce(red)
