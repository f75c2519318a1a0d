"""CPU: pins the oracle (oracle/tgn_oracle.py) against vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  The GPU tests then compare
the CUDA path against this oracle and against the same vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def load_sd(module, z, prefix):
    sd = {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}
    module.load_state_dict(sd)


# ---------------------------------------------------------------- neighbour ring
@pytest.mark.parametrize("impl", ["loops", "torch"])
def test_neighbor_ring_matches_reference(impl):
    z = _load("neighbor_loader.npz")
    for c in range(int(z["num_cases"])):
        N, K, B, steps = z[f"c{c}_meta"].tolist()
        ring = orc.NeighborRing(N, K) if impl == "loops" else orc.TorchNeighborLoader(N, K)
        for s in range(steps):
            p = f"c{c}_s{s}_"
            roots = z[p + "roots"]
            if impl == "loops":
                n_id, ei, e_id, t = ring.lookup(roots)
            else:
                n_id, ei, e_id, t = [x.numpy() for x in ring(torch.from_numpy(roots))]
            assert np.array_equal(n_id, z[p + "n_id"])
            assert np.array_equal(ei, z[p + "edge_index"])
            assert np.array_equal(e_id, z[p + "e_id"])
            assert np.array_equal(t, z[p + "t"])
            if impl == "loops":
                ring.insert(z[p + "src"], z[p + "dst"], z[p + "tin"])
                e_state, n_state, t_state = ring.e_id, ring.neighbors, ring.t
            else:
                ring.insert(torch.from_numpy(z[p + "src"]), torch.from_numpy(z[p + "dst"]),
                            torch.from_numpy(z[p + "tin"]))
                e_state, n_state, t_state = ring.e_id.numpy(), ring.neighbors.numpy(), ring.t.numpy()
            assert np.array_equal(e_state, z[p + "state_e"])
            assert np.array_equal(np.where(e_state >= 0, n_state, 0), z[p + "state_n"])
            assert np.array_equal(t_state, z[p + "state_t"])


def test_ring_more_than_k_events_keeps_last_k_of_stable_run():
    """> K events of one node in one batch: slots collide (neighbor_loader.py:75-88);
    with the stable order the survivors are the last K entries of cat[dst-side, src-side]."""
    ring = orc.NeighborRing(6, 2)
    src = np.array([0, 0, 0, 5]); dst = np.array([1, 2, 3, 0]); t = np.array([1, 2, 3, 4], np.float32)
    ring.insert(src, dst, t)
    # node 0: dst-side entry (e3, nbr 5) comes first in the run, then src-side e0,e1,e2 -> keep e1,e2
    assert ring.e_id[0].tolist() == [2, 1] and ring.neighbors[0].tolist() == [3, 2]
    tl = orc.TorchNeighborLoader(6, 2)
    tl.insert(torch.from_numpy(src), torch.from_numpy(dst), torch.from_numpy(t))
    assert tl.e_id[0].tolist() == [2, 1] and tl.neighbors[0].tolist() == [3, 2]
    # the independent t.topk (neighbor_loader.py:100) keeps the two largest t of the kept slots
    assert ring.t[0].tolist() == [3.0, 2.0]


# ---------------------------------------------------------------- aggregators
def test_aggregators_match_reference():
    z = _load("aggregators.npz")
    for i in range(int(z["num_cases"])):
        msg, idx, t = (torch.from_numpy(z[f"a{i}_{k}"]) for k in ("msg", "index", "t"))
        S = int(z[f"a{i}_S"])
        assert torch.equal(orc.LastAggregator()(msg, idx, t, S), torch.from_numpy(z[f"a{i}_last"]))
        torch.testing.assert_close(orc.MeanAggregator()(msg, idx, t, S), torch.from_numpy(z[f"a{i}_mean"]),
                                   rtol=1e-6, atol=1e-6)


def test_last_aggregator_first_wins_on_ties():
    msg = torch.arange(12.0).view(6, 2)
    idx = torch.tensor([1, 1, 0, 1, 0, 3])
    t = torch.tensor([5, 7, 2, 7, 2, 0])
    out = orc.LastAggregator()(msg, idx, t, 4)
    assert out.tolist() == [[4.0, 5.0], [2.0, 3.0], [0.0, 0.0], [10.0, 11.0]]


# ---------------------------------------------------------------- memory
def test_memory_matches_reference():
    z = _load("memory.npz")
    for c in range(int(z["num_cases"])):
        N, De, D, B, steps, aggr, _ = z[f"m{c}_meta"].tolist()
        mem = orc.TGNMemory(N, De, D, D, orc.IdentityMessage(De, D, D),
                            orc.LastAggregator() if aggr == 0 else orc.MeanAggregator())
        load_sd(mem, z, f"m{c}_sd")
        mem.train()
        for s in range(steps):
            p = f"m{c}_s{s}_"
            if not int(z[p + "training"]) and mem.training:
                mem.eval()
            zz, lu = mem(torch.from_numpy(z[p + "q"]))
            torch.testing.assert_close(zz.detach(), torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=1e-6)
            assert np.array_equal(lu.numpy(), z[p + "lu"])
            mem.update_state(*(torch.from_numpy(z[p + k]) for k in ("src", "dst", "t", "raw")))
            mem.detach()
            torch.testing.assert_close(mem.memory.detach(), torch.from_numpy(z[p + "memory"]), rtol=1e-5, atol=1e-6)
            assert np.array_equal(mem.last_update.numpy(), z[p + "last_update"])


def test_dyrep_memory_and_time_embedding_match_reference():
    z = _load("variants.npz")
    for c in range(int(z["num_cases"])):
        N, De, D, B, steps, aggr, upd, use_s, use_d = z[f"d{c}_meta"].tolist()
        mem = orc.DyRepMemory(N, De, D, D, orc.IdentityMessage(De, D, D),
                              orc.LastAggregator() if aggr == 0 else orc.MeanAggregator(),
                              "gru" if upd == 0 else "rnn", bool(use_s), bool(use_d))
        load_sd(mem, z, f"d{c}_sd")
        mem.train()
        for s in range(steps):
            p = f"d{c}_s{s}_"
            if not int(z[p + "training"]) and mem.training:
                mem.eval()
            zz, lu = mem(torch.from_numpy(z[p + "q"]))
            torch.testing.assert_close(zz.detach(), torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=1e-6)
            assert np.array_equal(lu.numpy(), z[p + "lu"])
            mem.update_state(*(torch.from_numpy(z[p + k]) for k in ("src", "dst", "t", "raw", "emb", "assoc")))
            mem.detach()
            torch.testing.assert_close(mem.memory.detach(), torch.from_numpy(z[p + "memory"]), rtol=1e-5, atol=1e-6)
            assert np.array_equal(mem.last_update.numpy(), z[p + "last_update"])
    te = orc.TimeEmbedding(16, 16)
    load_sd(te, z, "te_sd")
    out = te(torch.from_numpy(z["te_x"]), torch.from_numpy(z["te_lu"]), torch.from_numpy(z["te_t"]))
    torch.testing.assert_close(out.detach(), torch.from_numpy(z["te_out"]), rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- embedding + decoder
def test_embedding_and_decoder_match_reference():
    z = _load("embedding.npz")
    for c in range(int(z["num_cases"])):
        Nb, E, D, De = z[f"e{c}_meta"].tolist()
        te = orc.tp.TimeEncoder(D)
        gnn = orc.GraphAttentionEmbedding(D, D, De, te).eval()
        lp = orc.LinkPredictor(D)
        load_sd(gnn, z, f"e{c}_gnn"); load_sd(lp, z, f"e{c}_lp")
        p = f"e{c}_"
        out = gnn(*(torch.from_numpy(z[p + k]) for k in ("x", "lu", "edge_index", "t", "msg")))
        torch.testing.assert_close(out, torch.from_numpy(z[p + "z"]), rtol=1e-5, atol=1e-6)
        prob = lp(out[torch.from_numpy(z[p + "a"])], out[torch.from_numpy(z[p + "b"])])
        torch.testing.assert_close(prob, torch.from_numpy(z[p + "prob"]), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- t-CSR sampler (hand cases)
def test_tcsr_recent_hand_case():
    #   node 0: events at t=1(e0,->1) 2(e1,->2) 2(e2,->3) 5(e3,->1);  node 1..3 get the reverse edges
    src = [0, 0, 0, 0]; dst = [1, 2, 3, 1]; t = [1, 2, 2, 5]
    indptr, indices, eid, ts = orc.build_tcsr(src, dst, t, 4)
    assert indptr.tolist() == [0, 4, 6, 7, 8]
    assert indices[:4].tolist() == [1, 2, 3, 1] and eid[:4].tolist() == [0, 1, 2, 3]
    n, c, e, tt, dt, off = orc.tcsr_sample_ref(indptr, indices, eid, ts, [0, 0, 1, 2], [2.0, 6.0, 5.0, 0.5], k=2)
    # root (0, t=2): strictly earlier -> only e0;  root (0, t=6): two most recent = e3, e2
    assert e.tolist() == [0, 3, 2, 0] and n.tolist() == [1, 1, 3, 0]
    assert c.tolist() == [0, 1, 1, 2] and off.tolist() == [0, 1, 3, 4, 4]
    assert dt.tolist() == [1.0, 1.0, 4.0, 4.0]


def test_tcsr_uniform_is_within_window_and_reproducible():
    rng = np.random.default_rng(0)
    src = rng.integers(0, 20, 400); dst = rng.integers(0, 20, 400); t = np.sort(rng.integers(0, 1000, 400))
    g = orc.build_tcsr(src, dst, t, 20)
    roots = rng.integers(0, 20, 50); rts = rng.integers(0, 1000, 50).astype(np.float32)
    a = orc.tcsr_sample_ref(*g, roots, rts, k=5, strategy="uniform", seed=7)
    b = orc.tcsr_sample_ref(*g, roots, rts, k=5, strategy="uniform", seed=7)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert (a[3] < rts[a[1]]).all()
    w = orc.tcsr_sample_ref(*g, roots, rts, k=5, strategy="recent", duration=100.0)
    assert ((w[3] < rts[w[1]]) & (w[3] >= rts[w[1]] - 100.0)).all()


def test_philox_known_answer():
    # Random123 known-answer test for philox4x32-10: counter = key = 0
    assert orc.philox4x32(0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    # counter = key = all ones
    assert orc.philox4x32(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF) == \
        (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)


def test_mrr_convention():
    pos = np.array([0.9, 0.5, 0.5]); neg = np.array([[0.1, 0.2], [0.5, 0.7], [0.6, 0.7]])
    # ranks: 1, 1+0.5*(1+2)=2.5, 3
    np.testing.assert_allclose(orc.mrr_ref(pos, neg), [1.0, 1 / 2.5, 1 / 3.0])


def test_vendored_driver_is_the_reference_file(golden_dir):
    """tests/golden/ref_driver_pyg-mem-tgn.py.txt (run unchanged by tests/test_gpu_unchanged_driver.py) is byte-identical
    to the reference's pyg-mem-tgn.py wherever the reference tree exists, and imports the drop-in names."""
    import os
    txt = open(os.path.join(golden_dir, "ref_driver_pyg-mem-tgn.py.txt"), "rb").read()
    ref = "/root/reference/pyg-mem-tgn.py"
    if os.path.exists(ref):
        assert txt == open(ref, "rb").read()
    for name in (b"from utils import parse_config, getDataWithDependecyBlock", b"from epoch_utils import train, test",
                 b"from model_utils import getModel, getOptimizer", b"from neighbor_loader import LastNeighborLoader"):
        assert name in txt


def test_scatter_max_vectorised_equals_sequential_scan():
    """oracle/thirdparty.scatter_max (two stable sorts) == the strict-`>` sequential scan it restates
    (first maximal element wins; empty segments -> value 0, argmax == len)."""
    import torch
    from oracle import thirdparty as tp
    g = torch.Generator().manual_seed(0)
    for dt in (torch.long, torch.float32):
        for n, S in ((0, 3), (1, 1), (50, 7), (1000, 40), (3000, 3000)):
            src = torch.randint(0, 6, (n,), generator=g).to(dt)
            idx = torch.randint(0, S, (n,), generator=g)
            a, b = tp.scatter_max(src, idx, dim_size=S + 2), tp.scatter_max_loop(src, idx, dim_size=S + 2)
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), (dt, n, S)
