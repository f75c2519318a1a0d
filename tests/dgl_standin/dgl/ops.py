"""dgl.ops.edge_softmax: softmax of edge logits over the incoming edges of each destination node."""
import torch


def edge_softmax(graph, logits):
    dst, n = graph._dst, graph._n
    shape = (n,) + tuple(logits.shape[1:])
    idx = dst.view((-1,) + (1,) * (logits.dim() - 1)).expand_as(logits)
    mx = torch.full(shape, float("-inf"), dtype=logits.dtype).scatter_reduce(0, idx, logits, "amax", include_self=True)
    ex = (logits - mx[dst]).exp()
    den = torch.zeros(shape, dtype=logits.dtype).index_add(0, dst, ex)
    return ex / den[dst]
