"""GPU parity tests of the drop-in modules (reference class names and
signatures) against the golden vectors of the unmodified reference and against
the CPU oracle, including gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def load_sd(module, z, prefix):
    sd = {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}
    module.load_state_dict(sd)


def make_memory(N, De, D, aggr, cell="gru"):
    from modules.memory_module import TGNMemory
    from modules.msg_agg import LastAggregator, MeanAggregator
    from modules.msg_func import IdentityMessage
    return TGNMemory(N, De, D, D, IdentityMessage(De, D, D),
                     LastAggregator() if aggr == 0 else MeanAggregator(), cell)


# ------------------------------------------------------------------ TGNMemory
@pytest.mark.parametrize("force_unfused", [False, True])
def test_memory_golden(force_unfused):
    z = np.load(os.path.join(G, "memory.npz"))
    for c in range(int(z["num_cases"])):
        N, De, D, B, steps, aggr, _ = z[f"m{c}_meta"].tolist()
        mem = make_memory(N, De, D, aggr)
        load_sd(mem, z, f"m{c}_sd")
        mem = mem.to(DEV)
        if force_unfused:
            mem._fused_mode = lambda: None
        mem.train()
        for s in range(steps):
            p = f"m{c}_s{s}_"
            if not int(z[p + "training"]) and mem.training:
                mem.eval()
            with torch.no_grad():
                zz, lu = mem(cu(z[p + "q"]))
            torch.testing.assert_close(zz.cpu(), torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=5e-6)
            assert np.array_equal(lu.cpu().numpy(), z[p + "lu"])            # bit-exact last_update
            mem.update_state(cu(z[p + "src"]), cu(z[p + "dst"]), cu(z[p + "t"]), cu(z[p + "raw"]))
            mem.detach()
            torch.testing.assert_close(mem.memory.cpu(), torch.from_numpy(z[p + "memory"]), rtol=1e-5, atol=5e-6)
            assert np.array_equal(mem.last_update.cpu().numpy(), z[p + "last_update"])


def test_dyrep_memory_and_time_embedding_golden():
    """DyRepMemory (rnn / gru updater, embeddings substituted into the messages, memory_module.py:218-421)
    and TimeEmbedding (emb_module.py:32-52) against vectors from the unmodified reference."""
    from modules.memory_module import DyRepMemory
    from modules.emb_module import TimeEmbedding
    from modules.msg_agg import LastAggregator, MeanAggregator
    from modules.msg_func import IdentityMessage
    z = np.load(os.path.join(G, "variants.npz"))
    for c in range(int(z["num_cases"])):
        N, De, D, B, steps, aggr, upd, use_s, use_d = z[f"d{c}_meta"].tolist()
        mem = DyRepMemory(N, De, D, D, IdentityMessage(De, D, D), LastAggregator() if aggr == 0 else MeanAggregator(),
                          "gru" if upd == 0 else "rnn", bool(use_s), bool(use_d))
        load_sd(mem, z, f"d{c}_sd")
        mem = mem.to(DEV)
        mem.train()
        for s in range(steps):
            p = f"d{c}_s{s}_"
            if not int(z[p + "training"]) and mem.training:
                mem.eval()
            with torch.no_grad():
                zz, lu = mem(cu(z[p + "q"]))
            torch.testing.assert_close(zz.cpu(), torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=5e-6)
            assert np.array_equal(lu.cpu().numpy(), z[p + "lu"])
            with torch.no_grad():
                mem.update_state(*(cu(z[p + k]) for k in ("src", "dst", "t", "raw", "emb", "assoc")))
            mem.detach()
            torch.testing.assert_close(mem.memory.cpu(), torch.from_numpy(z[p + "memory"]), rtol=1e-5, atol=5e-6)
            assert np.array_equal(mem.last_update.cpu().numpy(), z[p + "last_update"])
    te = TimeEmbedding(16, 16)
    load_sd(te, z, "te_sd")
    te = te.to(DEV)
    out = te(cu(z["te_x"]), cu(z["te_lu"]), cu(z["te_t"]))
    torch.testing.assert_close(out.detach().cpu(), torch.from_numpy(z["te_out"]), rtol=1e-5, atol=1e-5)


def _run_memory_pair(N, De, D, B, steps, aggr, seed, ties):
    torch.manual_seed(seed)
    ref = orc.TGNMemory(N, De, D, D, orc.IdentityMessage(De, D, D),
                        orc.LastAggregator() if aggr == 0 else orc.MeanAggregator())
    mem = make_memory(N, De, D, aggr)
    mem.load_state_dict(ref.state_dict())
    mem = mem.to(DEV)
    ref.train(); mem.train()
    rng = np.random.default_rng(seed)
    tcur = 0
    return ref, mem, rng, tcur


@pytest.mark.parametrize("aggr", [0, 1])
def test_memory_vs_oracle_with_ties_and_grads(aggr):
    """coarse timestamps (many equal t inside a batch and across s/d stores), hubs, nodes
    with empty stores; gradients of a random projection of the output."""
    N, De, D, B, steps = 120, 5, 16, 64, 6
    ref, mem, rng, tcur = _run_memory_pair(N, De, D, B, steps, aggr, 3, True)
    for s in range(steps):
        src = (rng.random(B) ** 2 * N).astype(np.int64); dst = (rng.random(B) ** 2 * N).astype(np.int64)
        t = np.sort(rng.integers(0, 4, B) + tcur).astype(np.int64); tcur = int(t[-1])
        raw = rng.standard_normal((B, De)).astype(np.float32)
        q = np.unique(np.concatenate([src, dst, rng.integers(0, N, 20)]))
        zr, lur = ref(torch.from_numpy(q))
        zg, lug = mem(cu(q))
        torch.testing.assert_close(zg.detach().cpu(), zr.detach(), rtol=1e-5, atol=1e-5)
        assert np.array_equal(lug.cpu().numpy(), lur.numpy())
        wgt = torch.from_numpy(rng.standard_normal(zr.shape).astype(np.float32))
        ref.zero_grad(); mem.zero_grad()
        (zr * wgt).sum().backward()
        (zg * wgt.to(DEV)).sum().backward()
        for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mem.named_parameters()):
            assert n1 == n2
            if p1.grad is None:
                assert p2.grad is None or float(p2.grad.abs().max()) == 0.0
                continue
            torch.testing.assert_close(p2.grad.cpu(), p1.grad, rtol=1e-4, atol=1e-4, msg=lambda m: f"{n1}: {m}")
        args = (src, dst, t, raw)
        ref.update_state(*(torch.from_numpy(a) for a in args)); ref.detach()
        mem.update_state(*(cu(a) for a in args)); mem.detach()
        torch.testing.assert_close(mem.memory.cpu(), ref.memory.detach(), rtol=1e-5, atol=1e-5)
        assert np.array_equal(mem.last_update.cpu().numpy(), ref.last_update.numpy())
    ref.eval(); mem.eval()      # flush of all N nodes, then an eval-mode update
    torch.testing.assert_close(mem.memory.cpu(), ref.memory.detach(), rtol=1e-5, atol=1e-5)
    assert np.array_equal(mem.last_update.cpu().numpy(), ref.last_update.numpy())


def test_memory_wiki_dims_state_dict_keys():
    mem = make_memory(1000, 172, 100, 0)
    assert sorted(mem.state_dict().keys()) == sorted([
        "memory", "last_update", "_assoc", "time_enc.lin.weight", "time_enc.lin.bias",
        "memory_updater.weight_ih", "memory_updater.weight_hh", "memory_updater.bias_ih",
        "memory_updater.bias_hh"])
    assert mem.memory_updater.weight_ih.shape == (300, 472)
    assert mem.msg_s_module.out_channels == 472


# ------------------------------------------------------------------ embedding + decoder
def test_embedding_decoder_golden():
    from modules.decoder import LinkPredictor
    from modules.emb_module import GraphAttentionEmbedding
    from modules.time_enc import TimeEncoder
    z = np.load(os.path.join(G, "embedding.npz"))
    for c in range(int(z["num_cases"])):
        Nb, E, D, De = z[f"e{c}_meta"].tolist()
        gnn = GraphAttentionEmbedding(D, D, De, TimeEncoder(D)).eval()
        lp = LinkPredictor(D)
        load_sd(gnn, z, f"e{c}_gnn"); load_sd(lp, z, f"e{c}_lp")
        gnn, lp = gnn.to(DEV), lp.to(DEV)
        p = f"e{c}_"
        with torch.no_grad():
            out = gnn(*(cu(z[p + k]) for k in ("x", "lu", "edge_index", "t", "msg")))
            prob = lp(out[cu(z[p + "a"])], out[cu(z[p + "b"])])
        torch.testing.assert_close(out.cpu(), torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(prob.cpu(), torch.from_numpy(z[p + "prob"]), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("grouped", [True, False])
def test_embedding_grads_vs_oracle(grouped):
    from modules.decoder import LinkPredictor
    from modules.emb_module import GraphAttentionEmbedding
    from modules.time_enc import TimeEncoder
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    Nb, E, D, De = 90, 400, 100, 172
    te_r = orc.tp.TimeEncoder(D)
    with torch.no_grad():
        te_r.lin.weight.mul_(0.01)
    ref = orc.GraphAttentionEmbedding(D, D, De, te_r).eval()
    lp_r = orc.LinkPredictor(D)
    gnn = GraphAttentionEmbedding(D, D, De, TimeEncoder(D)).eval()
    gnn.load_state_dict(ref.state_dict())
    lp = LinkPredictor(D); lp.load_state_dict(lp_r.state_dict())
    gnn, lp = gnn.to(DEV), lp.to(DEV)
    x = torch.randn(Nb, D, requires_grad=True)
    lu = torch.from_numpy(rng.integers(0, 3000, Nb))
    centres = rng.integers(0, Nb // 3, E); centres = np.sort(centres) if grouped else centres
    ei = torch.from_numpy(np.stack([rng.integers(0, Nb, E), centres]))
    t = torch.from_numpy(rng.integers(0, 3000, E).astype(np.float32))
    msg = torch.randn(E, De)
    a = torch.from_numpy(rng.integers(0, Nb, 64)); b = torch.from_numpy(rng.integers(0, Nb, 64))
    zr = ref(x, lu, ei, t, msg)
    loss_r = torch.nn.functional.binary_cross_entropy_with_logits(lp_r.logits(zr[a], zr[b]), torch.ones(64, 1))
    loss_r.backward()
    xg = x.detach().to(DEV).requires_grad_()
    zg = gnn(xg, lu.to(DEV), ei.to(DEV), t.to(DEV), msg.to(DEV))
    loss_g = torch.nn.functional.binary_cross_entropy_with_logits(lp.logits(zg[a.to(DEV)], zg[b.to(DEV)]),
                                                                  torch.ones(64, 1, device=DEV))
    loss_g.backward()
    torch.testing.assert_close(zg.detach().cpu(), zr.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(loss_g.detach().cpu(), loss_r.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(xg.grad.cpu(), x.grad, rtol=1e-4, atol=1e-6)
    for (n1, p1), (n2, p2) in zip(list(ref.named_parameters()) + list(lp_r.named_parameters()),
                                  list(gnn.named_parameters()) + list(lp.named_parameters())):
        assert n1 == n2
        torch.testing.assert_close(p2.grad.cpu(), p1.grad, rtol=1e-4, atol=1e-6, msg=lambda m: f"{n1}: {m}")


def test_attention_dropout_statistics():
    """training-mode dropout (p=0.1, emb_module.py:22) keeps E[out] and is reproducible per seed."""
    from tgn_b200 import ops
    torch.manual_seed(0)
    Nb, K, HC, De, Dt = 2000, 10, 100, 16, 100
    x = torch.randn(Nb, 100, device=DEV)
    wn = torch.randn(4 * HC, 100, device=DEV) * 0.1; bn = torch.zeros(4 * HC, device=DEV)
    we = torch.randn(HC, De + Dt, device=DEV) * 0.1
    tw = torch.rand(Dt, device=DEV) * 0.01; tb = torch.zeros(Dt, device=DEV)
    lu = torch.randint(0, 1000, (Nb,), device=DEV)
    nbr = torch.randint(0, Nb, (Nb * K,), device=DEV)
    t = torch.rand(Nb * K, device=DEV) * 1000
    msg = torch.randn(Nb * K, De, device=DEV)
    row_ptr = torch.arange(0, Nb * K + 1, K, dtype=torch.int32, device=DEV)
    kw = dict(heads=2)
    base = ops.temporal_attention(x, wn, bn, we, tw, tb, lu, nbr, t, msg, row_ptr, **kw)
    a = ops.temporal_attention(x, wn, bn, we, tw, tb, lu, nbr, t, msg, row_ptr, dropout_p=0.1, seed=1, **kw)
    a2 = ops.temporal_attention(x, wn, bn, we, tw, tb, lu, nbr, t, msg, row_ptr, dropout_p=0.1, seed=1, **kw)
    b = ops.temporal_attention(x, wn, bn, we, tw, tb, lu, nbr, t, msg, row_ptr, dropout_p=0.1, seed=2, **kw)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    mean_of_many = torch.stack([ops.temporal_attention(x, wn, bn, we, tw, tb, lu, nbr, t, msg, row_ptr,
                                                       dropout_p=0.1, seed=s, **kw) for s in range(3, 43)]).mean(0)
    err = (mean_of_many - base).abs().mean() / base.abs().mean()
    assert float(err) < 0.03, float(err)


# ------------------------------------------------------------------ sampler_core API
def test_sampler_core_api_two_layers():
    import sampler_core
    rng = np.random.default_rng(2)
    N, E = 300, 5000
    src = rng.integers(0, N, E); dst = rng.integers(0, N, E); t = np.sort(rng.integers(0, 10000, E)).astype(np.float32)
    g = orc.build_tcsr(src, dst, t, N)
    s = sampler_core.ParallelSampler(*g, 8, 1, 2, [10, 5], True, False, 1, 0.0)
    roots = rng.integers(0, N, 64).astype(np.int32); rts = rng.integers(5000, 10000, 64).astype(np.float32)
    s.sample(roots, rts)
    ret = s.get_ret()
    assert len(ret) == 2
    r0 = orc.tcsr_sample_ref(*g, roots, rts, 10)
    b0 = ret[0]
    assert b0.dim_out() == 64 and b0.dim_in() == 64 + r0[0].size
    assert np.array_equal(b0.nodes(), np.concatenate([roots, r0[0]])) and np.array_equal(b0.col(), r0[1])
    assert np.array_equal(b0.eid(), r0[2]) and np.array_equal(b0.row(), np.arange(64, b0.dim_in()))
    assert np.array_equal(b0.ts(), np.concatenate([rts, r0[3]]))
    assert np.array_equal(b0.dts(), np.concatenate([np.zeros(64, np.float32), r0[4]]))
    r1 = orc.tcsr_sample_ref(*g, b0.nodes(), b0.ts(), 5)     # layer 2 roots = layer 1 nodes with their ts
    b1 = ret[1]
    assert b1.dim_out() == b0.dim_in() and np.array_equal(b1.eid(), r1[2]) and np.array_equal(b1.col(), r1[1])
    s.reset()
    assert s.get_ret() == []


def test_two_layer_uniform_embedding_matches_oracle():
    """BASELINE configs[3] in miniature (comment shape: uniform-20 sampling, 2 attention layers): the
    two TGL blocks come from sampler_core (uniform draws are Philox draws the oracle restates, so the
    blocks are compared exactly), then the layers run innermost-first on the drop-in
    GraphAttentionEmbedding and are compared with the oracle's modules on the same blocks."""
    import sampler_core
    from modules.emb_module import GraphAttentionEmbedding
    from modules.time_enc import TimeEncoder
    rng = np.random.default_rng(4)
    N, E, De, D, R = 400, 9000, 2, 32, 48
    src = rng.integers(0, N, E); dst = rng.integers(0, N, E); t = np.sort(rng.integers(0, 20000, E)).astype(np.float32)
    feat = torch.from_numpy(rng.standard_normal((E, De)).astype(np.float32))
    g = orc.build_tcsr(src, dst, t, N)
    smp = sampler_core.ParallelSampler(*g, 8, 1, 2, [20, 20], False, False, 1, 0.0, seed=9)
    roots = rng.integers(0, N, R).astype(np.int32); rts = rng.integers(10000, 20000, R).astype(np.float32)
    smp.sample(roots, rts)
    b1, b2 = smp.get_ret()
    assert b2.dim_out() == b1.dim_in() and b1.dim_out() == R
    # uniform-k semantics: every root draws k entries with replacement (or all when it has <= k), all earlier
    for blk, k in ((b1, 20), (b2, 20)):
        cnt = np.bincount(blk.col(), minlength=blk.dim_out())
        assert cnt.max() <= k and bool((blk.ts()[blk.row()] < blk.ts()[blk.col()]).all())
    memory = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32))
    last_update = torch.from_numpy(rng.integers(0, 10000, N).astype(np.int64))
    torch.manual_seed(3)
    te = orc.tp.TimeEncoder(D)
    with torch.no_grad():
        te.lin.weight.mul_(0.01)
    ref = [orc.GraphAttentionEmbedding(D, D, De, te) for _ in range(2)]
    for r in ref:
        r.eval()
    te_g = TimeEncoder(D); te_g.load_state_dict(te.state_dict())
    gpu = [GraphAttentionEmbedding(D, D, De, te_g) for _ in range(2)]
    for r, m in zip(ref, gpu):
        m.load_state_dict(r.state_dict()); m.to(DEV).eval()

    def run(layers, dev):
        mv = lambda a: a.to(dev)
        n2 = torch.from_numpy(b2.nodes()).long()
        ei2 = torch.from_numpy(np.stack([b2.row(), b2.col()])).long()
        h = layers[0](mv(memory[n2]), mv(last_update[n2]), mv(ei2), mv(torch.from_numpy(b2.ts()[b2.row()]).long()),
                      mv(feat[torch.from_numpy(b2.eid()).long()]))
        n1 = torch.from_numpy(b1.nodes()).long()
        ei1 = torch.from_numpy(np.stack([b1.row(), b1.col()])).long()
        out = layers[1](h[:b1.dim_in()], mv(last_update[n1]), mv(ei1), mv(torch.from_numpy(b1.ts()[b1.row()]).long()),
                        mv(feat[torch.from_numpy(b1.eid()).long()]))
        return out[:R]
    with torch.no_grad():
        want, got = run(ref, "cpu"), run(gpu, DEV)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-4, atol=1e-5)
