"""CPU: oracle/dgl_twins.py (restatement of the reference's DGL-flavoured hot-path classes) against
tests/golden/dgl_twins.npz, which tests/golden/make_golden.py produced by running the UNMODIFIED
/root/reference/model_utils.py on the minimal dgl stand-in: outputs and (via autograd) gradients."""
import os

import numpy as np
import torch

from oracle import dgl_twins as tw

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dgl_twins.npz"))
T = lambda k: torch.from_numpy(G[k])


def _case(c):
    pre = f"c{c}_"
    N, E, De, D, H = (int(v) for v in G[pre + "cfg"])
    p = {k[len(pre) + 2 + len("edge_gatconv."):]: T(k).clone().requires_grad_() for k in G.files
         if k.startswith(pre + "p.edge_gatconv.")}
    tw_ = T(pre + "p.temporal_encoder.w.weight").clone().requires_grad_()
    tb_ = T(pre + "p.temporal_encoder.w.bias").clone().requires_grad_()
    return pre, (N, E, De, D, H), p, tw_, tb_


def test_time_encode_init_matches_reference():
    pre, (N, E, De, D, H), p, w, b = _case(0)
    w0, b0 = tw.time_encode_init(D)
    assert torch.equal(w0, w.detach()) and torch.equal(b0, b.detach())


def test_temporal_transformer_conv_matches_reference():
    for c in range(int(G["num_cases"])):
        pre, (N, E, De, D, H), p, w, b = _case(c)
        src, dst = T(pre + "src"), T(pre + "dst")
        efeat = tw.edge_preprocess(T(pre + "feats"), T(pre + "edge_ts"), T(pre + "node_ts"), src, w, b)
        torch.testing.assert_close(efeat.detach(), T(pre + "efeat"), rtol=1e-6, atol=1e-6)
        out = tw.temporal_transformer_conv(T(pre + "mem"), T(pre + "feats"), T(pre + "edge_ts"), T(pre + "node_ts"),
                                           src, dst, w, b, p, H, D)
        torch.testing.assert_close(out.detach(), T(pre + "out"), rtol=1e-5, atol=1e-5)
        (out * T(pre + "out_w")).sum().backward()
        if E:
            for k, v in p.items():
                torch.testing.assert_close(v.grad, T(pre + "g.edge_gatconv." + k), rtol=1e-4, atol=1e-6, msg=lambda m: f"{c} {k}: {m}")
            torch.testing.assert_close(w.grad, T(pre + "g.temporal_encoder.w.weight"), rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(b.grad, T(pre + "g.temporal_encoder.w.bias"), rtol=1e-4, atol=1e-6)


def test_memory_operation_matches_reference():
    """Exactly equal to the reference when its tiled-gather defect (model_utils.py:403) is reproduced; the
    documented `last` semantics agree with it on every node whose in-degree bucket cannot mix indices."""
    for c in range(int(G["num_cases"])):
        pre, (N, E, De, D, H), p, w, b = _case(c)
        if not E:
            continue
        src, dst, ets = T(pre + "src"), T(pre + "dst"), T(pre + "edge_ts").view(-1)
        for kind, cls in (("gru", torch.nn.GRUCell), ("rnn", torch.nn.RNNCell)):
            cell = cls(2 * D + De + D, D)
            cell.load_state_dict({k.split(".", 1)[1]: T(k) for k in G.files if k.startswith(pre + f"mo_{kind}.")})
            args = (T(pre + "mem"), T(pre + "node_ts").view(-1), T(pre + "feats"), ets, src, dst, w.detach(), b.detach(), cell)
            mem_ref, ts = tw.memory_operation(*args, reference_tiled_gather=True)
            torch.testing.assert_close(mem_ref.detach(), T(pre + f"mo_{kind}_memory"), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(ts, T(pre + f"mo_{kind}_ts"), rtol=0, atol=0)
            mem, ts2 = tw.memory_operation(*args)
            assert torch.equal(ts, ts2)
            deg = torch.bincount(dst, minlength=N)
            safe = torch.zeros(N, dtype=torch.bool)
            for d in set(deg.tolist()):
                nodes = (deg == d).nonzero(as_tuple=True)[0]
                if d <= 1 or nodes.numel() == 1:
                    safe[nodes] = True
            assert int(safe.sum()) > 0
            torch.testing.assert_close(mem.detach()[safe], T(pre + f"mo_{kind}_memory")[safe], rtol=1e-5, atol=1e-6)


def test_edge_predictor_matches_reference():
    p = {k[len("ep_p."):]: T(k) for k in G.files if k.startswith("ep_p.")}
    pos, neg = tw.edge_predictor(T("ep_hs"), T("ep_hp"), T("ep_hn"), p, neg_samples=3)
    torch.testing.assert_close(pos, T("ep_pos"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(neg, T("ep_neg"), rtol=1e-6, atol=1e-6)
