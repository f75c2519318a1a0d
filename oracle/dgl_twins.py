"""TEST INFRASTRUCTURE -- CPU oracle of the DGL-flavoured twins of the hot path:
the classes the reference's LIVE driver wiring uses (pyg-mem-tgn.py:24 -> model_utils.py).

Each function restates one class of /root/reference/model_utils.py on plain edge lists (no graph object)
and cites the lines it follows.  Pinning: tests/golden/make_golden.py runs the UNMODIFIED model_utils.py
on top of tests/dgl_standin (a minimal `dgl`) and freezes inputs, parameters, outputs and gradients in
tests/golden/dgl_twins.npz; tests/test_oracle_golden.py checks these functions against them.  The DGL
message-passing primitives themselves (edge_softmax, update_all bucketing, zero-in-degree handling) are
PARITY UNPINNED (package absent; restated from its published behaviour).
Only tests/ may import this module.
"""
from __future__ import annotations

import torch
from torch import Tensor


def time_encode(t: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """TimeEncode.forward (model_utils.py:232-237): cos(Linear(1, dim)(t)), t [M] or [M,1] -> [M, dim].
    Initial weights 1 / 10**linspace(0, 9, dim), zero bias (:227-230)."""
    return torch.cos(t.reshape(-1, 1) * w.reshape(1, -1) + b.reshape(1, -1))


def time_encode_init(dim: int):
    import numpy as np
    w = torch.from_numpy(1 / 10 ** np.linspace(0, 9, dim)).float().reshape(dim, 1)
    return w, torch.zeros(dim)


def edge_preprocess(feats: Tensor, edge_ts: Tensor, node_ts: Tensor, src: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """TemporalEdgePreprocess.edge_fn (model_utils.py:440-449): time_diff = edge timestamp - SOURCE node
    timestamp (:442), edge feature = [feats, time_encode] (:448 -- feats first, the opposite of the PyG
    stack's [time, msg], emb_module.py:28)."""
    dt = edge_ts.reshape(-1) - node_ts.reshape(-1)[src]
    return torch.cat([feats, time_encode(dt, w, b)], dim=1)


def _softmax_by_dst(z: Tensor, dst: Tensor, n: int) -> Tensor:
    idx = dst.view((-1,) + (1,) * (z.dim() - 1)).expand_as(z)
    mx = torch.full((n,) + tuple(z.shape[1:]), float("-inf"), dtype=z.dtype).scatter_reduce(0, idx, z, "amax")
    ex = (z - mx[dst]).exp()
    den = torch.zeros((n,) + tuple(z.shape[1:]), dtype=z.dtype).index_add(0, dst, ex)
    return ex / den[dst]


def edge_gat_conv(x: Tensor, efeat: Tensor, src: Tensor, dst: Tensor, p: dict, heads: int, out_feats: int,
                  negative_slope: float = 0.2) -> Tensor:
    """EdgeGATConv.forward in eval mode (model_utils.py:565-612; dropouts are identities), residual=True.
    p: fc_node.weight/bias, fc_edge.weight/bias, attn_l, attn_r, attn_e (+ res_fc.weight when
    node_feats != out_feats).  Returns rst [N, H, out_feats]."""
    n, H, F = x.size(0), heads, out_feats
    node_feat = (x @ p["fc_node.weight"].t() + p["fc_node.bias"]).view(-1, H, F)             # :582-583
    edge_feat = (efeat @ p["fc_edge.weight"].t() + p["fc_edge.bias"]).view(-1, H, F)         # :584-585
    el = (node_feat * p["attn_l"]).sum(-1, keepdim=True)                                     # :587
    er = (node_feat * p["attn_r"]).sum(-1, keepdim=True)                                     # :588
    ee = (edge_feat * p["attn_e"]).sum(-1, keepdim=True)                                     # :589
    el_prime = el[src] + ee                                                                  # :594 u_add_e
    e = torch.nn.functional.leaky_relu(el_prime + er[dst], negative_slope)                   # :595-596
    a = _softmax_by_dst(e, dst, n)                                                           # :597
    m = a.view(-1, H, 1) * el_prime                                                          # :560-563 msg_fn
    ft = torch.zeros((n, H, 1), dtype=x.dtype).index_add(0, dst, m)                          # :599 fn.sum
    if "res_fc.weight" in p:                                                                 # :601-604
        resval = (x @ p["res_fc.weight"].t()).view(n, -1, F)
    else:
        resval = x.view(n, -1, F)
    return ft + resval


def temporal_transformer_conv(x: Tensor, feats: Tensor, edge_ts: Tensor, node_ts: Tensor, src: Tensor, dst: Tensor,
                              time_w: Tensor, time_b: Tensor, p: dict, heads: int, out_feats: int) -> Tensor:
    """TemporalTransformerConv.forward (model_utils.py:688-697): preprocess, one EdgeGATConv, mean over heads."""
    efeat = edge_preprocess(feats, edge_ts, node_ts, src, time_w, time_b).float()            # :691
    return edge_gat_conv(x, efeat, src, dst, p, heads, out_feats).mean(1)                    # :693


def memory_operation(memory: Tensor, last_update_t: Tensor, feats: Tensor, edge_ts: Tensor, src: Tensor, dst: Tensor,
                     time_w: Tensor, time_b: Tensor, cell: torch.nn.Module, reference_tiled_gather: bool = False):
    """MemoryOperation.forward (model_utils.py:393-416) on a positive-pair graph with every node of the
    graph: message = [memory[src], memory[dst], feats, time_encode(edge ts - last_update_t[src])]
    (:394-398), per destination the message with the LATEST timestamp (torch.max over the mailbox: the first
    of equal maxima, edges in edge-id order, :402), then the GRU/RNN cell on (message_bar, memory) (:409-410).
    Nodes without in-edges reduce to a zero message and still run the cell (DGL applies the node function to
    all nodes; SURVEY.md B7 -- unpinned).  Returns (new memory [N, D], timestamp [N]).

    reference_tiled_gather=True reproduces a latent defect of the reference's agg_last (:403-404): the gather
    index is built with `latest_idx.repeat(message_dim).view(-1, 1, message_dim)`, which TILES the bucket's
    argmax vector instead of repeating each node's index -- column c of node r (r-th node of its in-degree
    bucket, n nodes) is taken from mailbox slot latest_idx[(r*message_dim + c) % n].  The class is never
    instantiated by the reference, so the defect is dead code; the documented behaviour
    ("last(m_i(t_1),...,m_i(t_b))", :344-346) is what the default path and the CUDA path implement.  The
    two agree on every node whose bucket has a single node or identical argmax indices (e.g. in-degree 1)."""
    n = memory.size(0)
    dt = edge_ts.reshape(-1) - last_update_t.reshape(-1)[src]
    msg = torch.cat([memory[src], memory[dst], feats, time_encode(dt, time_w, time_b)], dim=1)
    md = msg.size(1)
    ts = edge_ts.reshape(-1)
    bar = torch.zeros((n, md), dtype=msg.dtype)
    ts_out = torch.zeros(n, dtype=ts.dtype)
    order = torch.argsort(dst, stable=True)
    deg = torch.bincount(dst, minlength=n)
    start = torch.cumsum(deg, 0) - deg
    for d in sorted(set(deg.tolist()) - {0}):
        nodes = (deg == d).nonzero(as_tuple=True)[0]                       # bucket, ascending node id
        eids = order[(start[nodes].view(-1, 1) + torch.arange(d).view(1, -1))]   # [n_b, d] mailbox, edge-id order
        latest = torch.argmax(ts[eids], dim=1)                             # first maximal element
        ts_out[nodes] = ts[eids].gather(1, latest.view(-1, 1)).view(-1)
        if reference_tiled_gather:
            idx = latest.repeat(md).view(-1, md)                           # :403 (tiled, not interleaved)
        else:
            idx = latest.view(-1, 1).expand(-1, md)
        win = eids.gather(1, idx)                                          # [n_b, md] edge id per column
        bar[nodes] = msg[win, torch.arange(md).view(1, -1).expand_as(win)]
    return cell(bar.float(), memory.float()), ts_out


def edge_predictor(h_src, h_pos, h_neg, p: dict, neg_samples: int = 1):
    """EdgePredictor.forward (model_utils.py:186-195): logits, no sigmoid; the negatives are paired with
    h_src.tile(neg_samples, 1) (:192 -- row r of the tiled sources is source r % B)."""
    lin = lambda x, k: x @ p[k + ".weight"].t() + p[k + ".bias"]
    hs, hp, hn = lin(h_src, "src_fc"), lin(h_pos, "dst_fc"), lin(h_neg, "dst_fc")
    pos = torch.relu(hs + hp)
    neg = torch.relu(hs.tile(neg_samples, 1) + hn)
    return lin(pos, "out_fc"), lin(neg, "out_fc")
