"""GPU: BASELINE configs[3] as a real training step -- t-CSR graph, uniform-k (and recent-k) sampling over TWO
layers, two attention layers, memory update -- tgn_b200.tcsr_trainer.TCSRTrainer against the CPU oracle composed
the same way (oracle t-CSR sampler incl. the restated Philox draws, oracle TGNMemory / GraphAttentionEmbedding /
LinkPredictor): identical sampled blocks, loss per step within 1e-4, memory and last_update after the steps."""
import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle_blocks(g, roots, rts, ks, recent, seed, call):
    """sampler_core.ParallelSampler.sample_device restated on the oracle sampler (same per-layer seeds)."""
    blocks, rn, rt = [], np.asarray(roots, np.int32), np.asarray(rts, np.float32)
    for layer, k in enumerate(ks):
        s = (seed + 1315423911 * call + 2654435761 * layer) & 0xFFFFFFFFFFFFFFFF
        nbr, col, eid, ts, dts, off = orc.tcsr_sample_ref(*g, rn, rt, k, "recent" if recent else "uniform", seed=s)
        blocks.append(dict(roots=rn, nbr=nbr, col=col, eid=eid, ts=ts))
        rn, rt = np.concatenate([rn, nbr]), np.concatenate([rt, ts])
    return blocks


@pytest.mark.parametrize("recent", [False, True])
def test_two_layer_tcsr_training_matches_oracle(recent):
    from tgn_b200.tcsr_trainer import TCSRTrainer
    rng = np.random.default_rng(12)
    N, E, De, D, B, steps, ks = 300, 4000, 2, 32, 20, 3, [6, 5]
    ns = N // 2
    src = rng.integers(0, ns, E); dst = rng.integers(ns, N, E)
    t = np.sort(rng.integers(0, 40000, E)).astype(np.int64)
    feat = rng.standard_normal((E, De)).astype(np.float32)
    g = orc.build_tcsr(src, dst, t, N)
    tr = TCSRTrainer(*(torch.from_numpy(a) for a in g), N, De, D, ks, recent, torch.from_numpy(feat), device=DEV,
                     lr=1e-3, dropout=0.0, seed=5)
    with torch.no_grad():
        tr.memory.time_enc.lin.weight.mul_(0.002)
    tr.train()
    # oracle twin with the same weights
    ref = orc.build_model(De, D, N, seed=1)
    ref["memory"].load_state_dict({k: v.cpu() for k, v in tr.memory.state_dict().items()})
    layers = [orc.GraphAttentionEmbedding(D, D, De, ref["memory"].time_enc) for _ in ks]
    for lr_, lg in zip(layers, tr.layers):
        lr_.load_state_dict({k: v.cpu() for k, v in lg.state_dict().items()})
        lr_.time_enc = ref["memory"].time_enc
        lr_.conv.dropout = 0.0
    lp = ref["link_pred"]
    lp.load_state_dict({k: v.cpu() for k, v in tr.link_pred.state_dict().items()})
    mods = [ref["memory"]] + layers + [lp]
    for m in mods:
        m.train()
    params = list({id(p): p for m in mods for p in m.parameters()}.values())
    opt = torch.optim.Adam(params, lr=1e-3)
    crit = torch.nn.BCEWithLogitsLoss()
    featc = torch.from_numpy(feat)
    e0 = E - steps * B                                # the last batches: their roots have history in the graph
    for s in range(steps):
        sl = slice(e0 + s * B, e0 + (s + 1) * B)
        bs, bd = torch.from_numpy(src[sl]), torch.from_numpy(dst[sl])
        bn = torch.from_numpy(rng.integers(ns, N, B))
        bt, bm = torch.from_numpy(t[sl]), torch.from_numpy(feat[sl])
        loss = float(tr.train_step(bs, bd, bn, bt, bm))
        # ---- oracle
        opt.zero_grad()
        roots = torch.cat([bs, bd, bn]).numpy()
        blocks = _oracle_blocks(g, roots, np.tile(t[sl].astype(np.float32), 3), ks, recent, 5, s)
        inner = blocks[-1]
        nodes_in = torch.from_numpy(np.concatenate([inner["roots"], inner["nbr"]])).long()
        n_id = nodes_in.unique()
        assoc = torch.zeros(N, dtype=torch.long); assoc[n_id] = torch.arange(n_id.numel())
        z, lu = ref["memory"](n_id)
        h, lus = z[assoc[nodes_in]], lu[assoc[nodes_in]]
        for layer in range(len(ks) - 1, -1, -1):
            b = blocks[layer]
            R, n = b["roots"].size, b["nbr"].size
            ei = torch.stack([R + torch.arange(n), torch.from_numpy(b["col"]).long()])
            h = layers[layer](h[:R + n], lus[:R + n], ei, torch.from_numpy(b["ts"]), featc[torch.from_numpy(b["eid"]).long()])[:R]
        pos, ngo = lp.logits(h[:B], h[B:2 * B]), lp.logits(h[:B], h[2 * B:])
        loss_ref = crit(pos, torch.ones_like(pos)) + crit(ngo, torch.zeros_like(ngo))
        ref["memory"].update_state(bs, bd, bt, bm)
        loss_ref.backward()
        opt.step()
        ref["memory"].detach()
        assert abs(loss - float(loss_ref)) < 1e-4 * max(1.0, abs(float(loss_ref))), (s, loss, float(loss_ref))
        # keep the two weight sets identical (Adam amplifies rounding-level gradient noise)
        tr.memory.load_state_dict({k: v.to(DEV) for k, v in ref["memory"].state_dict().items() if k not in ("memory", "last_update", "_assoc")}, strict=False)
        for lr_, lg in zip(layers, tr.layers):
            lg.load_state_dict({k: v.to(DEV) for k, v in lr_.state_dict().items()})
        tr.link_pred.load_state_dict({k: v.to(DEV) for k, v in lp.state_dict().items()})
    assert tr.sampled_edges > 0
    torch.testing.assert_close(tr.memory.memory.cpu(), ref["memory"].memory.detach(), rtol=1e-4, atol=1e-5)
    assert torch.equal(tr.memory.last_update.cpu(), ref["memory"].last_update)
