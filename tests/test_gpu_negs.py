"""GPU: device-side negatives and epoch metric: NegLinkSamplerDest on the device
(reference neg_sampler.py:8-23), resident evaluation negative tables with the reference's
batch-min truncation (epoch_utils.py:48-56), synthetic [B,Q] negatives, and the epoch MRR
accumulated in device memory (epoch_utils.py:108-113,163)."""
import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_neg_dest_sampler_distribution_and_collisions():
    from tgn_b200.neg_table import DeviceNegSamplerDest
    dst_nodes = torch.tensor([3, 7, 11, 19, 23, 42])
    s = DeviceNegSamplerDest(dst_nodes, seed=5)
    pos = torch.tensor([7, 42, 100, 3] * 30000, device=DEV)
    neg = s.sample(pos)
    assert neg.dtype == pos.dtype and bool((neg != pos).all())
    assert bool(torch.isin(neg, dst_nodes.to(DEV)).all())
    # per positive value: uniform over the destination set without that value
    for p, k in ((7, 5), (42, 5), (100, 6), (3, 5)):
        vals = neg[pos == p].cpu().numpy()
        hist = np.array([(vals == d).sum() for d in dst_nodes.tolist() if d != p], dtype=np.float64)
        assert hist.size == k and hist.sum() == 30000
        chi2 = ((hist - 30000 / k) ** 2 / (30000 / k)).sum()
        assert chi2 < 30, (p, chi2)                       # dof <= 5: 30 is beyond 1e-5
    # the reference sampler (CPU) has the same support and marginal; successive calls differ, a
    # sampler with the same seed replays them
    again = DeviceNegSamplerDest(dst_nodes, seed=5)
    assert torch.equal(again.sample(pos), neg) and not torch.equal(s.sample(pos), neg)
    one = DeviceNegSamplerDest(torch.tensor([9]))
    assert one.sample(torch.tensor([9, 1], device=DEV)).tolist() == [9, 9]        # single destination: kept


def test_neg_fill_uniform_without_positive():
    from tgn_b200.neg_table import SyntheticNegatives
    gen = SyntheticNegatives(num_neg=999, lo=100, hi=150, seed=2)
    pos = torch.tensor([100, 149, 125, 7], device=DEV).repeat(500)
    neg = gen.batch(pos, call=3)
    assert neg.shape == (2000, 999) and int(neg.min()) >= 100 and int(neg.max()) <= 149
    assert not bool((neg == pos[:, None]).any())
    for p, k in ((100, 49), (149, 49), (125, 49), (7, 50)):
        vals = neg[pos == p].reshape(-1).cpu().numpy()
        hist = np.bincount(vals - 100, minlength=50).astype(np.float64)
        hist = hist[hist > 0] if p != 7 else hist
        assert hist.size == k
        chi2 = ((hist - vals.size / k) ** 2 / (vals.size / k)).sum()
        assert chi2 < 120, (p, chi2)                      # dof ~ 49
    assert torch.equal(gen.batch(pos, call=3), neg) and not torch.equal(gen.batch(pos, call=4), neg)
    odd = SyntheticNegatives(num_neg=5, lo=0, hi=3).batch(torch.tensor([1], device=DEV), call=0)
    assert set(odd.reshape(-1).tolist()) <= {0, 2}


def test_negative_table_truncates_like_the_reference():
    from tgn_b200 import dist_eval
    from tgn_b200.neg_table import DeviceNegativeTable
    rng = np.random.default_rng(0)
    rows = [rng.integers(0, 1000, rng.integers(3, 9)).tolist() for _ in range(37)]
    tab = DeviceNegativeTable.from_lists(rows, DEV)
    for i0 in range(0, 37, 10):
        i1 = min(i0 + 10, 37)
        assert torch.equal(tab.batch(i0, i1).cpu(), dist_eval.truncate_negatives(rows[i0:i1]))
    assert len(tab) == 37


def test_rank_accum_matches_oracle_mrr():
    from tgn_b200.neg_table import rank_accum
    g = torch.Generator().manual_seed(1)
    acc = torch.zeros(2, dtype=torch.float64, device=DEV)
    per_batch = []
    for B, Q in ((200, 999), (37, 20), (1, 5)):
        pos, neg = torch.rand(B, generator=g), torch.rand(B, Q, generator=g)
        neg[:, 0] = pos                                                # ties
        gt, ge = (neg > pos[:, None]).sum(1).int().to(DEV), (neg >= pos[:, None]).sum(1).int().to(DEV)
        rr = torch.empty(B, device=DEV)
        rank_accum(gt, ge, acc, rr)
        want = orc.mrr_ref(pos.numpy(), neg.numpy())
        np.testing.assert_allclose(rr.cpu().numpy(), want, rtol=1e-6)
        per_batch.append(float(want.mean()))
    a = acc.cpu()
    assert float(a[1]) == 3.0 and abs(float(a[0] / a[1]) - float(np.mean(per_batch))) < 1e-6


def test_evaluate_table_equals_host_list_evaluation():
    """a whole evaluation split through the resident table == the per-batch host-list path of
    dist_eval.evaluate_dp (identical MRR, identical final state)."""
    from test_gpu_engine import _setup
    from tgn_b200 import dist_eval, synth
    from tgn_b200.neg_table import DeviceNegativeTable, SyntheticNegatives, evaluate_table
    N, De, D, K, B, nb, Q = 300, 8, 16, 5, 40, 5, 23
    lists = None
    out = []
    for mode in ("host", "table"):
        ref, eng, ev = _setup(N, De, D, K, B, B * (nb + 3), 5, False)
        for _ in range(3):
            eng.train_step(from_device=True)
        eng.flush_to_eval()
        sl = slice(3 * B, (3 + nb) * B)
        src, dst, t, msg = ev["src"][sl], ev["dst"][sl], ev["t"][sl], ev["msg"][sl]
        if lists is None:
            neg = synth.eval_negatives(src.numpy(), dst.numpy(), N, Q, seed=3, dst_lo=N // 2)
            lists = [r[:Q - (i % 3)].tolist() for i, r in enumerate(neg)]         # ragged: the truncation matters
        if mode == "host":
            batches = [(src[i:i + B], dst[i:i + B], dist_eval.truncate_negatives(lists[i:i + B]), t[i:i + B],
                        msg[i:i + B]) for i in range(0, nb * B, B)]
            mrr = dist_eval.evaluate_dp(eng, batches)
        else:
            mrr = evaluate_table(eng, src, dst, t, msg, DeviceNegativeTable.from_lists(lists, DEV), B)
        out.append((mrr, eng.memory.clone(), eng.last_update.clone()))
    assert abs(out[0][0] - out[1][0]) < 1e-3
    # (the two engines were trained separately: split-K gradient atomics make their weights differ at
    # rounding level, so the memories agree to tolerance, the integer state exactly)
    torch.testing.assert_close(out[0][1], out[1][1], rtol=1e-4, atol=1e-5)
    assert torch.equal(out[0][2], out[1][2])
    # synthetic generator path runs end to end and gives a valid MRR
    ref, eng, ev = _setup(N, De, D, K, B, B * (nb + 3), 5, False)
    eng.flush_to_eval()
    m = evaluate_table(eng, ev["src"][:nb * B], ev["dst"][:nb * B], ev["t"][:nb * B], ev["msg"][:nb * B],
                       SyntheticNegatives(Q, N // 2, N), B)
    assert 0.0 < m <= 1.0


def test_ap_auc_accum_matches_sklearn():
    """tgn_ap_auc_accum == sklearn's average_precision_score / roc_auc_score on sigmoid(logits), the per-batch
    metrics the reference computes on the host (epoch_utils.py:312-315), including tied scores."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    from tgn_b200 import ops
    g = torch.Generator().manual_seed(0)
    acc = torch.zeros(3, dtype=torch.float64, device=DEV)
    want_ap, want_auc = 0.0, 0.0
    cases = [(200, 200), (37, 37), (5, 11), (2000, 2000)]
    for i, (P, Nn) in enumerate(cases):
        pos = torch.randn(P, generator=g) + 0.5
        neg = torch.randn(Nn, generator=g)
        if i == 1:                       # heavy ties, saturated sigmoids
            pos, neg = (pos * 2).round() / 2, (neg * 2).round() / 2
            pos[:3], neg[:3] = 40.0, 40.0
        y_pred = torch.cat([pos, neg]).sigmoid().numpy()
        y_true = np.r_[np.ones(P), np.zeros(Nn)]
        want_ap += average_precision_score(y_true, y_pred)
        want_auc += roc_auc_score(y_true, y_pred)
        ops.ap_auc_accum(pos.to(DEV), neg.to(DEV), acc)
    got = acc.tolist()
    assert got[2] == len(cases)
    # (the kernel evaluates the sigmoid itself: a last-bit difference can split or merge a tie)
    assert abs(got[0] - want_ap) < 2e-4 * len(cases) and abs(got[1] - want_auc) < 2e-4 * len(cases), (got, want_ap, want_auc)
