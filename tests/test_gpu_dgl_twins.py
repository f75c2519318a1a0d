"""GPU: the DGL-flavoured twins (dgl_model_utils.py on csrc/egat.cu + the shared kernels) against
tests/golden/dgl_twins.npz -- outputs and gradients of the UNMODIFIED reference model_utils.py classes run on
the minimal dgl stand-in -- and against oracle/dgl_twins.py.  fp32 bar: 1e-5 relative (+ small atol; the
GEMMs run as 3xTF32), gradients 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import dgl_twins as tw

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dgl_twins.npz"))
T = lambda k: torch.from_numpy(G[k])


def _conv(c):
    import dgl_model_utils as dm
    from tgn_b200.graph import graph
    pre = f"c{c}_"
    N, E, De, D, H = (int(v) for v in G[pre + "cfg"])
    te = dm.TimeEncode(D)
    conv = dm.TemporalTransformerConv(De, D, te, D, H, allow_zero_in_degree=True)
    sd = {k[len(pre) + 2:]: T(k) for k in G.files if k.startswith(pre + "p.")}
    assert set(sd) == set(conv.state_dict().keys())          # same parameter tree as the reference class
    conv.load_state_dict(sd)
    conv = conv.to(DEV).eval()
    g = graph((T(pre + "src").to(DEV), T(pre + "dst").to(DEV)), num_nodes=N)
    g.ndata["timestamp"] = T(pre + "node_ts").to(DEV)
    g.edata["timestamp"], g.edata["feats"] = T(pre + "edge_ts").to(DEV), T(pre + "feats").to(DEV)
    return pre, (N, E, De, D, H), conv, g


def test_time_encode_has_the_reference_init():
    import dgl_model_utils as dm
    te = dm.TimeEncode(100)
    w, b = tw.time_encode_init(100)
    assert torch.equal(te.w.weight.detach(), w) and torch.equal(te.w.bias.detach(), b)
    te = te.to(DEV)
    t = torch.tensor([[0.0], [1.0], [12345.0], [2.6e6]], device=DEV)
    torch.testing.assert_close(te(t).cpu(), tw.time_encode(t.cpu(), w, b), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("c", range(int(G["num_cases"])))
def test_temporal_transformer_conv_matches_reference(c):
    pre, (N, E, De, D, H), conv, g = _conv(c)
    efeat = conv.preprocessor(g.local_var())
    torch.testing.assert_close(efeat.cpu(), T(pre + "efeat"), rtol=1e-5, atol=2e-6)
    out = conv(g, T(pre + "mem").to(DEV))
    torch.testing.assert_close(out.detach().cpu(), T(pre + "out"), rtol=1e-5, atol=2e-5)
    (out * T(pre + "out_w").to(DEV)).sum().backward()
    if E:
        for k, v in conv.named_parameters():
            want = T(pre + "g." + k)
            torch.testing.assert_close(v.grad.cpu(), want, rtol=2e-4, atol=2e-5 * max(1.0, float(want.abs().max())),
                                       msg=lambda m: f"case {c} {k}: {m}")


def test_edge_gat_conv_attention_weights_and_zero_in_degree():
    pre, (N, E, De, D, H), conv, g = _conv(0)
    gat = conv.edge_gatconv
    efeat = conv.preprocessor(g.local_var())
    rst, a = gat(g, T(pre + "mem").to(DEV), efeat, get_attention=True)
    assert tuple(rst.shape) == (N, H, D) and tuple(a.shape) == (E, H, 1)
    dst = T(pre + "dst")
    sums = torch.zeros(N, H).index_add(0, dst, a.view(E, H).cpu())
    deg = torch.bincount(dst, minlength=N)
    torch.testing.assert_close(sums[deg > 0], torch.ones_like(sums[deg > 0]), rtol=1e-5, atol=1e-5)
    # nodes without in-edges: only the residual (allow_zero_in_degree=True, model_utils.py:34,601-604)
    zero = (deg == 0).nonzero(as_tuple=True)[0]
    assert zero.numel() > 0
    torch.testing.assert_close(rst[zero.to(DEV)].mean(1).cpu(), T(pre + "mem")[zero], rtol=1e-6, atol=1e-6)
    strict = type(gat)(D, De + D, D, H, residual=True, allow_zero_in_degree=False).to(DEV)
    with pytest.raises(RuntimeError, match="0-in-degree"):
        strict(g, T(pre + "mem").to(DEV), efeat)


def test_attention_dropout_is_unbiased_and_differentiable():
    from tgn_b200 import ops
    torch.manual_seed(0)
    N, E, H = 6, 40, 8
    el, er, ee = torch.randn(N, H, device=DEV), torch.randn(N, H, device=DEV), torch.randn(E, H, device=DEV)
    src, dst = torch.randint(0, N, (E,), device=DEV), torch.randint(0, N, (E,), device=DEV)
    s0, _ = ops.egat_attention(el, er, ee, src, dst, 0.2, 0.0, 0)
    acc = torch.zeros_like(s0)
    n = 600
    for seed in range(n):
        acc += ops.egat_attention(el, er, ee, src, dst, 0.2, 0.6, 1000 + seed)[0]
    assert float((acc / n - s0).abs().max()) < 0.25 * float(s0.abs().max())
    # backward under a (seed-fixed) dropout mask: central differences of the same forward
    el.requires_grad_(); er.requires_grad_(); ee.requires_grad_()
    s, alpha = ops.egat_attention(el, er, ee, src, dst, 0.2, 0.6, 77)
    w = torch.randn_like(s)
    (s * w).sum().backward()
    f = lambda a_, b_, c_: float((ops.egat_attention(a_, b_, c_, src, dst, 0.2, 0.6, 77)[0] * w).sum())
    eps = 1e-2
    for name, t, idxs in (("el", el, [(0, 0), (3, 5)]), ("er", er, [(1, 2), (5, 7)]), ("ee", ee, [(0, 0), (17, 3), (39, 7)])):
        for ij in idxs:
            base = [x.detach().clone() for x in (el, er, ee)]
            k = ("el", "er", "ee").index(name)
            base[k][ij] += eps
            up = f(*base)
            base[k][ij] -= 2 * eps
            dn = f(*base)
            num = (up - dn) / (2 * eps)
            assert abs(num - float(t.grad[ij])) < 2e-2 * max(1.0, abs(num)), (name, ij, num, float(t.grad[ij]))


def test_memory_operation_last_then_cell():
    import dgl_model_utils as dm
    from tgn_b200.graph import NID, graph
    for c in range(int(G["num_cases"])):
        pre = f"c{c}_"
        N, E, De, D, H = (int(v) for v in G[pre + "cfg"])
        if not E:
            continue
        src, dst, ets = T(pre + "src"), T(pre + "dst"), T(pre + "edge_ts").view(-1)
        for kind, cls in (("gru", torch.nn.GRUCell), ("rnn", torch.nn.RNNCell)):
            te = dm.TimeEncode(D)
            mm = dm.MemoryModule(N, D)
            mm.memory.data.copy_(T(pre + "mem")); mm.last_update_t.data.copy_(T(pre + "node_ts").view(-1))
            mo = dm.MemoryOperation(kind, mm, De, te)
            mo.updater.load_state_dict({k.split(".", 1)[1]: T(k) for k in G.files if k.startswith(pre + f"mo_{kind}.")})
            mo = mo.to(DEV)
            g = graph((src.to(DEV), dst.to(DEV)), num_nodes=N)
            g.ndata[NID] = torch.arange(N, device=DEV)
            g.edata["timestamp"], g.edata["feats"] = ets.to(DEV), T(pre + "feats").to(DEV)
            with torch.no_grad():
                res = mo(g)
            cell = cls(2 * D + De + D, D)
            cell.load_state_dict(mo.updater.cpu().state_dict())
            w, b = tw.time_encode_init(D)
            want_mem, want_ts = tw.memory_operation(T(pre + "mem"), T(pre + "node_ts").view(-1), T(pre + "feats"), ets,
                                                    src, dst, w, b, cell)
            torch.testing.assert_close(res.ndata["memory"].cpu(), want_mem.detach(), rtol=1e-4, atol=2e-5)
            assert torch.equal(res.ndata["timestamp"].cpu(), want_ts)
            assert torch.equal(res.ndata["timestamp"].cpu(), T(pre + f"mo_{kind}_ts"))      # pinned by the reference
            deg = torch.bincount(dst, minlength=N)
            safe = torch.zeros(N, dtype=torch.bool)
            for d in set(deg.tolist()):
                nodes = (deg == d).nonzero(as_tuple=True)[0]
                if d <= 1 or nodes.numel() == 1:
                    safe[nodes] = True
            torch.testing.assert_close(res.ndata["memory"].cpu()[safe], T(pre + f"mo_{kind}_memory")[safe], rtol=1e-4, atol=2e-5)


def test_edge_predictor_matches_reference():
    import dgl_model_utils as dm
    ep = dm.EdgePredictor(16, 16)
    ep.load_state_dict({k[len("ep_p."):]: T(k) for k in G.files if k.startswith("ep_p.")})
    ep = ep.to(DEV)
    pos, neg = ep(T("ep_hs").to(DEV), T("ep_hp").to(DEV), T("ep_hn").to(DEV), neg_samples=3)
    torch.testing.assert_close(pos.detach().cpu(), T("ep_pos"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(neg.detach().cpu(), T("ep_neg"), rtol=1e-5, atol=1e-5)
