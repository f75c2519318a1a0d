"""GPU parity tests of the individual kernels, through the C-ABI (ctypes), against
the oracle and the committed golden vectors.  Bit-exact for ids / indices /
argmax / last_update; 1e-5 for fp32 math."""
import os

import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


@pytest.fixture(scope="module")
def ops():
    from tgn_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------ unique / relabel
@pytest.mark.parametrize("N,count", [(10, 0), (10, 7), (1024, 300), (1025, 2000), (9227, 600),
                                     (352637, 6600), (994790, 200200), (3_000_000, 50_000)])
def test_unique_relabel(ops, N, count):
    rng = np.random.default_rng(N + count)
    ids = rng.integers(0, N, count).astype(np.int64)
    if count:
        ids[0], ids[-1] = 0, N - 1
    assoc = torch.full((N,), -7, dtype=torch.int64, device=DEV)
    parts = [cu(ids[: count // 2]), cu(ids[count // 2:])] if count else [cu(ids)]
    out = ops.unique_relabel(parts, N, assoc)
    ref = np.unique(ids)
    assert np.array_equal(out.cpu().numpy(), ref)
    if count:
        assert np.array_equal(assoc[cu(ref)].cpu().numpy(), np.arange(ref.size))
    # the bitmap is left clean: a second call on other ids is unaffected
    ids2 = rng.integers(0, N, 13).astype(np.int64)
    assert np.array_equal(ops.unique_relabel([cu(ids2)], N).cpu().numpy(), np.unique(ids2))


# ------------------------------------------------------------------ neighbour ring
def test_ring_golden(ops):
    from neighbor_loader import LastNeighborLoader
    z = np.load(os.path.join(G, "neighbor_loader.npz"))
    for c in range(int(z["num_cases"])):
        N, K, B, steps = z[f"c{c}_meta"].tolist()
        ld = LastNeighborLoader(N, K, device=DEV)
        for s in range(steps):
            p = f"c{c}_s{s}_"
            n_id, ei, e_id, t = ld(cu(z[p + "roots"]))
            assert np.array_equal(n_id.cpu().numpy(), z[p + "n_id"])
            assert np.array_equal(ei.cpu().numpy(), z[p + "edge_index"])
            assert np.array_equal(e_id.cpu().numpy(), z[p + "e_id"])
            assert np.array_equal(t.cpu().numpy(), z[p + "t"])
            assert np.array_equal(ld._assoc[n_id].cpu().numpy(), np.arange(n_id.numel()))
            ld.insert(cu(z[p + "src"]), cu(z[p + "dst"]), cu(z[p + "tin"]))
            e_state = ld.e_id.cpu().numpy()
            assert np.array_equal(e_state, z[p + "state_e"])
            assert np.array_equal(np.where(e_state >= 0, ld.neighbors.cpu().numpy(), 0), z[p + "state_n"])
            assert np.array_equal(ld.t.cpu().numpy(), z[p + "state_t"])
        assert ld.cur_e_id == B * steps


@pytest.mark.parametrize("N,K,B,steps,bip", [(9227, 10, 200, 30, True), (500, 20, 600, 10, False),
                                             (64, 3, 50, 12, False), (2000, 10, 2000, 4, True),
                                             (300, 64, 100, 5, False)])
def test_ring_vs_oracle_random(ops, N, K, B, steps, bip):
    """wiki-like popularity (hubs get > K events per batch), self loops, non-bipartite."""
    from neighbor_loader import LastNeighborLoader
    rng = np.random.default_rng(N * K + B)
    ld = LastNeighborLoader(N, K, device=DEV)
    ring = orc.NeighborRing(N, K)
    tcur = 0
    for s in range(steps):
        u = rng.random(B)
        if bip:
            src = (u ** 3 * (N // 2)).astype(np.int64); dst = N // 2 + (rng.random(B) ** 3 * (N - N // 2)).astype(np.int64)
        else:
            src = (u ** 2 * N).astype(np.int64); dst = (rng.random(B) ** 2 * N).astype(np.int64)
        t = np.sort(rng.integers(0, 30, B) + tcur).astype(np.float32); tcur = int(t[-1])
        roots = np.unique(np.concatenate([src, dst, rng.integers(0, N, B)]))
        if s % 3 == 2:
            roots = rng.permutation(roots)      # lookup order follows the caller's order
        got = ld(cu(roots))
        ref = ring.lookup(roots)
        for a, b in zip(got, ref):
            assert np.array_equal(a.cpu().numpy(), b)
        ld.insert(cu(src), cu(dst), cu(t))
        ring.insert(src, dst, t)
        e_state = ld.e_id.cpu().numpy()
        assert np.array_equal(e_state, ring.e_id)
        assert np.array_equal(np.where(e_state >= 0, ld.neighbors.cpu().numpy(), 0),
                              np.where(ring.e_id >= 0, ring.neighbors, 0))
        assert np.array_equal(ld.t.cpu().numpy(), ring.t)
    ld.reset_state()
    assert ld.cur_e_id == 0 and int((ld.e_id != -1).sum()) == 0
    out = ld(cu(np.array([1, 2, 3])))
    assert out[1].shape == (2, 0) and out[0].tolist() == [1, 2, 3]


@pytest.mark.parametrize("N,K,R,sort_rows", [(994790, 10, 200_200, True), (100_000, 7, 60_000, False),
                                             (50_000, 3, 120_000, False), (30_000, 64, 9_000, True)])
def test_ring_large_lookup_properties(ops, N, K, R, sort_rows):
    """eval-sized lookup (200 x 1001 roots worth of nodes on the comment shape): size-independent
    properties -- edge order follows roots, every emitted slot is valid, counts add up.  The large-launch
    path (count pass: one thread per root, 128-bit row loads for even K; emit pass) with odd K, tiles of more
    roots than a count CTA has threads (K = 3), K = 64, and rows whose empty slots are not at the end."""
    g = torch.Generator(device=DEV).manual_seed(K)
    e_id = torch.randint(-1, 10_000_000, (N, K), device=DEV, generator=g)
    e_id[torch.rand((N, K), device=DEV, generator=g) < 0.3] = -1
    if sort_rows:
        e_id = torch.sort(e_id, dim=1, descending=True).values
    nbrs = torch.randint(0, N, (N, K), device=DEV, generator=g)
    t = torch.rand((N, K), device=DEV, generator=g)
    roots = torch.unique(torch.randint(0, N, (R,), device=DEV, generator=g))
    assoc = torch.zeros(N, dtype=torch.int64, device=DEV)
    ids, ei, eo, to, off = ops.nbr_lookup(roots, nbrs, e_id, t, assoc)
    valid = e_id[roots] >= 0
    assert eo.numel() == int(valid.sum())
    assert torch.equal(eo, e_id[roots][valid]) and torch.equal(to, t[roots][valid])
    assert torch.equal(ids[ei[0]], nbrs[roots][valid])
    assert torch.equal(ids[ei[1]], roots.view(-1, 1).expand(-1, K)[valid])
    assert torch.equal(off[1:].long() - off[:-1].long(), valid.sum(1))
    assert torch.equal(ids, torch.unique(torch.cat([roots, nbrs[roots][valid]])))


# ------------------------------------------------------------------ t-CSR sampler
def _graph(rng, N, E, tmax):
    src = (rng.random(E) ** 2 * N).astype(np.int64); dst = rng.integers(0, N, E)
    t = np.sort(rng.integers(0, tmax, E)).astype(np.float32)
    return orc.build_tcsr(src, dst, t, N)


@pytest.mark.parametrize("skip_index", [False, True])
@pytest.mark.parametrize("strategy", ["recent", "uniform"])
@pytest.mark.parametrize("N,E,R,k,dur", [(50, 600, 300, 5, 0.0), (2000, 40000, 3000, 10, 0.0),
                                         (2000, 40000, 1000, 20, 300.0), (10, 3, 40, 4, 0.0),
                                         (3, 5000, 500, 10, 40.0)])
def test_tcsr_vs_oracle(ops, strategy, N, E, R, k, dur, skip_index):
    rng = np.random.default_rng(E + k)
    indptr, indices, eid, ts = _graph(rng, N, E, 5000 if N > 3 else 300)     # N=3: long rows, many duplicate stamps
    roots = rng.integers(0, N, R).astype(np.int32)
    rts = rng.integers(0, 5200 if N > 3 else 320, R).astype(np.float32)
    ref = orc.tcsr_sample_ref(indptr, indices, eid, ts, roots, rts, k, strategy, 0.0, dur, seed=11)
    ts_d = cu(ts)
    coarse = ops.tcsr_build_index(ts_d) if skip_index else None
    (n, c, e, t, d), off, cnt = ops.tcsr_sample(cu(indptr), cu(indices), cu(eid), ts_d, cu(roots), cu(rts), k,
                                                0 if strategy == "recent" else 1, 0.0, dur, 11, coarse=coarse)
    m = int(cnt.item())
    assert m == ref[0].size
    for got, want in zip((n, c, e, t, d), ref[:5]):
        assert np.array_equal(got[:m].cpu().numpy(), want)
    assert np.array_equal(off.cpu().numpy(), ref[5])


@pytest.mark.parametrize("N,E,tmax,sorted_t,rev", [(50, 600, 40, True, True), (50, 600, 40, False, True),
                                                    (3000, 70000, 500, True, True), (3000, 70000, 500, False, False),
                                                    (1, 5, 3, True, True), (70000, 9000, 10**9, False, True)])
def test_tcsr_build_matches_oracle(ops, N, E, tmax, sorted_t, rev):
    """GPU t-CSR builder (TGL gen_graph stand-in) == oracle lexsort by (row, ts, eid), bit for bit;
    duplicate stamps, self-loops, isolated nodes, unsorted streams, float32 collapse of large stamps."""
    rng = np.random.default_rng(N + E)
    src = (rng.random(E) ** 2 * N).astype(np.int64); dst = rng.integers(0, N, E)
    dst[::7] = src[::7]                                            # self-loops
    t = rng.integers(0, tmax, E).astype(np.int64)
    if sorted_t:
        t = np.sort(t)
    ref = orc.build_tcsr(src, dst, t, N, add_reverse=rev)
    got = ops.tcsr_build(cu(src), cu(dst), cu(t), N, add_reverse=rev)
    for g, w, name in zip(got, ref, ("indptr", "indices", "eid", "ts")):
        assert np.array_equal(g.cpu().numpy(), w), name
    gotf = ops.tcsr_build(cu(src), cu(dst), cu(t.astype(np.float32)), N, add_reverse=rev)
    assert all(torch.equal(a, b) for a, b in zip(got, gotf))


def test_tcsr_build_edge_cases(ops):
    e = ops.tcsr_build(cu(np.zeros(0, np.int64)), cu(np.zeros(0, np.int64)), cu(np.zeros(0, np.int64)), 5)
    assert e[0].cpu().tolist() == [0] * 6 and e[1].numel() == 0
    from tgn_b200 import _cabi
    with pytest.raises(_cabi.TgnError):
        ops.tcsr_build(cu(np.array([0, 9])), cu(np.array([1, 2])), cu(np.array([0, 1])), 5)


def test_tcsr_build_large_feeds_sampler(ops):
    """4M events (8M entries) of the review shape: rows sorted by (ts, eid), degrees match a bincount,
    and the sampler on the built graph agrees with the sampler on the oracle-built graph."""
    rng = np.random.default_rng(9)
    N, E = 352637, 4_000_000
    src = (rng.random(E) ** 3 * N).astype(np.int64); dst = rng.integers(0, N, E)
    t = np.sort(rng.integers(0, 600000, E)).astype(np.int64)
    indptr, indices, eid, ts = ops.tcsr_build(cu(src), cu(dst), cu(t), N)
    deg = torch.bincount(torch.cat([cu(src), cu(dst)]), minlength=N)
    assert torch.equal((indptr[1:] - indptr[:-1]).long(), deg) and int(indptr[-1]) == 2 * E
    row = torch.repeat_interleave(torch.arange(N, device=DEV), deg)
    same = row[1:] == row[:-1]
    assert bool(((ts[1:] > ts[:-1]) | ((ts[1:] == ts[:-1]) & (eid[1:] >= eid[:-1])))[same].all())
    e64 = eid.long()
    assert torch.equal(ts, cu(t)[e64].float())
    assert bool(((indices.long() == cu(dst)[e64]) & (row == cu(src)[e64]) |
                 (indices.long() == cu(src)[e64]) & (row == cu(dst)[e64])).all())
    assert torch.equal(torch.bincount(e64, minlength=E), torch.full((E,), 2, device=DEV))


def test_tcsr_empty_and_ragged(ops):
    indptr, indices, eid, ts = orc.build_tcsr([0, 0], [1, 1], [3.0, 3.0], 4)   # nodes 2,3 isolated; tie at t=3
    (n, c, e, t, d), off, cnt = ops.tcsr_sample(cu(indptr), cu(indices), cu(eid), cu(ts),
                                                cu(np.array([2, 0, 0, 1, 3], np.int32)),
                                                cu(np.array([9, 3, 3.5, 100, 0], np.float32)), 3)
    assert off.cpu().tolist() == [0, 0, 0, 2, 4, 4] and int(cnt) == 4    # ts < t is strict
    assert e[:4].cpu().tolist() == [1, 0, 1, 0]
    empty = ops.tcsr_sample(cu(indptr), cu(indices), cu(eid), cu(ts), cu(np.zeros(0, np.int32)),
                            cu(np.zeros(0, np.float32)), 3)
    assert int(empty[2]) == 0


def test_tcsr_uniform_statistics(ops):
    """one hub with 1000 earlier events, 20000 roots x k=20 draws: chi-square against uniform."""
    E = 1000
    indptr, indices, eid, ts = orc.build_tcsr(np.zeros(E, np.int64), np.arange(1, E + 1), np.arange(E), E + 1)
    R, k = 20000, 20
    (n, c, e, t, d), off, cnt = ops.tcsr_sample(cu(indptr), cu(indices), cu(eid), cu(ts),
                                                cu(np.zeros(R, np.int32)), cu(np.full(R, 2000.0, np.float32)),
                                                k, 1, 0.0, 0.0, 1234)
    assert int(cnt) == R * k
    hist = np.bincount(e.cpu().numpy(), minlength=E).astype(np.float64)
    chi2 = ((hist - R * k / E) ** 2 / (R * k / E)).sum()
    assert 800 < chi2 < 1250, chi2       # dof = 999, +-5.6 sigma


def test_tcsr_large_properties(ops):
    """review-sized graph: recent-k output is sorted by root, most-recent-first, strictly earlier."""
    rng = np.random.default_rng(5)
    N, E = 352637, 2_000_000
    indptr, indices, eid, ts = _graph(rng, N, E, 100000)
    R, k = 500_000, 10
    roots = rng.integers(0, N, R).astype(np.int32); rts = rng.integers(0, 100000, R).astype(np.float32)
    g_d = [cu(indptr), cu(indices), cu(eid), cu(ts)]
    (n, c, e, t, d), off, cnt = ops.tcsr_sample(*g_d, cu(roots), cu(rts), k)
    m = int(cnt)
    # the skip-index search returns the identical sample (recent and windowed)
    coarse = ops.tcsr_build_index(g_d[3])
    for dur in (0.0, 5000.0):
        a = ops.tcsr_sample(*g_d, cu(roots), cu(rts), k, 0, 0.0, dur)
        b = ops.tcsr_sample(*g_d, cu(roots), cu(rts), k, 0, 0.0, dur, coarse=coarse)
        ma = int(a[2])
        assert ma == int(b[2]) and torch.equal(a[1], b[1])
        assert all(torch.equal(x[:ma], y[:ma]) for x, y in zip(a[0], b[0]))
    c, t, d, off = c[:m].long(), t[:m], d[:m], off.long()
    assert m == int(off[-1]) and bool((c[1:] >= c[:-1]).all())
    rt = cu(rts)
    assert bool((t < rt[c]).all()) and torch.equal(d, rt[c] - t)
    same = c[1:] == c[:-1]
    assert bool((t[1:][same] <= t[:-1][same]).all())
    assert bool(((off[1:] - off[:-1]) <= k).all())
    # count = min(k, #earlier in row): recompute on the host for a sample of roots
    for r in rng.integers(0, R, 200):
        row = ts[indptr[roots[r]]:indptr[roots[r] + 1]]
        assert int(off[r + 1] - off[r]) == min(k, int((row < rts[r]).sum()))


# ------------------------------------------------------------------ aggregators
def test_aggregators_golden_and_ties(ops):
    z = np.load(os.path.join(G, "aggregators.npz"))
    for i in range(int(z["num_cases"])):
        msg, idx, t, S = cu(z[f"a{i}_msg"]), cu(z[f"a{i}_index"]), cu(z[f"a{i}_t"]), int(z[f"a{i}_S"])
        out, arg = ops.agg_last(msg, idx, t, S)
        assert np.array_equal(out.cpu().numpy(), z[f"a{i}_last"])
        np.testing.assert_allclose(ops.agg_mean(msg, idx, S).cpu().numpy(), z[f"a{i}_mean"], rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(3)
    for M, S, W, dt in [(5000, 700, 472, np.int64), (10000, 50, 33, np.float32), (7, 3, 1, np.int64)]:
        msg = rng.standard_normal((M, W)).astype(np.float32)
        idx = rng.integers(0, S, M); t = rng.integers(-3, 4, M).astype(dt)   # heavy ties, negative t
        ref = orc.LastAggregator()(torch.from_numpy(msg), torch.from_numpy(idx), torch.from_numpy(t), S)
        _, ref_arg = orc.tp.scatter_max(torch.from_numpy(t), torch.from_numpy(idx), 0, S)
        out, arg = ops.agg_last(cu(msg), cu(idx), cu(t), S)
        assert np.array_equal(arg.cpu().numpy(), ref_arg.numpy())
        assert np.array_equal(out.cpu().numpy(), ref.numpy())
        refm = orc.MeanAggregator()(torch.from_numpy(msg), torch.from_numpy(idx), None, S)
        np.testing.assert_allclose(ops.agg_mean(cu(msg), cu(idx), S).cpu().numpy(), refm.numpy(), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ dense
@pytest.mark.parametrize("prec", [0, 3, 1])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
@pytest.mark.parametrize("m,n,k,split", [(1, 1, 1, 1), (65, 300, 472, 1), (600, 400, 100, 1), (300, 472, 4000, 8),
                                         (129, 63, 17, 2), (5023, 300, 472, 1), (50_000, 300, 472, 1)])
def test_sgemm(ops, prec, ta, tb, m, n, k, split):
    """prec 0: fp32 CUDA cores; 3: tcgen05 3xTF32 (same tolerance); 1: tcgen05 tf32 (2e-2 bar)."""
    g = torch.Generator(device="cpu").manual_seed(m * n + k)
    A = torch.randn((k, m) if ta else (m, k), generator=g)
    B = torch.randn((k, n) if tb else (n, k), generator=g)
    bias = torch.randn(n, generator=g)
    ref = (A.t() if ta else A).double() @ (B if tb else B.t()).double() + bias.double()
    out = ops.sgemm(A.to(DEV), B.to(DEV), bias.to(DEV), m=m, n=n, k=k, lda=A.shape[1], ldb=B.shape[1],
                    trans_a=ta, trans_b=tb, split_k=split, prec=prec)
    if prec == 1:
        scale = float(ref.abs().max())
        torch.testing.assert_close(out.cpu().double(), ref, rtol=2e-2, atol=2e-3 * max(scale, 1.0))
        return
    # fp32 accumulation over k terms of N(0,1) products: error grows ~ sqrt(k) * eps * |sum|
    torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-5 * max(4.0, k ** 0.5))


@pytest.mark.parametrize("prec", [0, 3])
def test_sgemm_gather_and_device_counts(ops, prec):
    g = torch.Generator(device="cpu").manual_seed(0)
    table = torch.randn(50, 24, generator=g); W = torch.randn(30, 24, generator=g)
    rows = torch.randint(0, 50, (40,), generator=g)
    cnt = torch.tensor([33], dtype=torch.int32, device=DEV)
    out = torch.full((40, 30), 7.0, device=DEV)
    ops.sgemm(table.to(DEV), W.to(DEV), m=40, n=30, k=24, lda=24, ldb=24, a_rows=rows.to(DEV), out=out, m_dev=cnt,
              prec=prec)
    torch.testing.assert_close(out[:33].cpu(), table[rows[:33]] @ W.t(), rtol=1e-5, atol=1e-5)
    assert bool((out[33:] == 7.0).all())
    # reduction length from the device: C = A^T B over the first 33 rows only
    A = torch.randn(40, 12, generator=g); B = torch.randn(40, 9, generator=g)
    got = ops.sgemm(A.to(DEV), B.to(DEV), m=12, n=9, k=40, lda=12, ldb=9, trans_a=True, trans_b=True, k_dev=cnt,
                    prec=prec)
    torch.testing.assert_close(got.cpu(), A[:33].t() @ B[:33], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("prec", [3, 1])
@pytest.mark.parametrize("m,n,k,ta,tb,split,live", [(300, 472, 5000, True, True, 8, 4321), (400, 100, 5000, True, True, 8, 4999),
                                                   (100, 104, 6000, True, True, 11, 5555), (260, 100, 400, False, True, 1, 333),
                                                   (200, 64, 96, False, False, 1, 70)])
def test_tma_gemm_device_reduction_length(ops, prec, m, n, k, ta, tb, split, live):
    """TMA path: the reduction stops at *k_dev exactly (tail of the last k-block masked in shared
    memory) in both precisions; rows past the live length hold garbage that must not leak in."""
    g = torch.Generator(device="cpu").manual_seed(k + live)
    A = torch.randn((k, m) if ta else (m, k), generator=g)
    B = torch.randn((k, n) if tb else (n, k), generator=g)
    Am, Bm = A.clone(), B.clone()
    if ta: A[live:] = 1e30
    else: A[:, live:] = 1e30
    if tb: B[live:] = -1e30
    else: B[:, live:] = -1e30
    for X in (Am, Bm):
        pass
    ref = ((Am[:live].t() if ta else Am[:, :live]).double() @ (Bm[:live] if tb else Bm[:, :live].t()).double())
    kdev = torch.tensor([live], dtype=torch.int32, device=DEV)
    out = torch.zeros(m, n, device=DEV)
    ops.sgemm(A.to(DEV), B.to(DEV), m=m, n=n, k=k, lda=A.shape[1], ldb=B.shape[1], trans_a=ta, trans_b=tb, out=out,
              split_k=split, k_dev=kdev, prec=prec)
    scale = float(ref.abs().max())
    if prec == 1:
        torch.testing.assert_close(out.cpu().double(), ref, rtol=2e-2, atol=2e-3 * scale)
    else:
        torch.testing.assert_close(out.cpu().double(), ref, rtol=1e-5, atol=1e-5 * max(4.0, live ** 0.5) * 2)


@pytest.mark.parametrize("prec", [3, 1])
def test_tma_gemm_persistent_multiwave(ops, prec):
    """Launches of several waves of tiles take the persistent kernel (one CTA per SM walking tiles, TMEM
    double-buffered accumulators): a batch of three problems in ONE launch -- a store problem with a device row
    count and a bias (rows past the count untouched, column tail of a 300-wide output), a transposed split-K
    problem with a device reduction length that ends inside a k-block (atomic accumulation into a pre-filled
    C), and a short-K problem -- against float64."""
    g = torch.Generator(device="cpu").manual_seed(7)
    m0, n0, k0, live0 = 30_000, 300, 116, 25_001
    A0 = torch.randn(m0, k0, generator=g); B0 = torch.randn(n0, k0, generator=g); bias0 = torch.randn(n0, generator=g)
    C0 = torch.full((m0, n0), 7.0, device=DEV)
    m1, n1, k1, live1, split1 = 300, 472, 40_000, 39_987, 48
    A1 = torch.randn(k1, m1, generator=g); B1 = torch.randn(k1, n1, generator=g)
    A1[live1:] = 1e30; B1[live1:] = -1e30
    C1 = torch.ones(m1, n1, device=DEV)
    m2, n2, k2 = 20_000, 100, 100
    A2 = torch.randn(m2, k2, generator=g); B2 = torch.randn(n2, k2, generator=g)
    C2 = torch.zeros(m2, n2, device=DEV)
    d = lambda t: t.to(DEV)
    A0d, B0d, b0d, A1d, B1d, A2d, B2d = map(d, (A0, B0, bias0, A1, B1, A2, B2))
    mdev = torch.tensor([live0], dtype=torch.int32, device=DEV)
    kdev = torch.tensor([live1], dtype=torch.int32, device=DEV)
    ops.gemm_batch([
        ops.gemm_desc(A0d, B0d, C0, m=m0, n=n0, k=k0, lda=k0, ldb=k0, ldc=n0, bias=b0d, m_dev=mdev),
        ops.gemm_desc(A1d, B1d, C1, m=m1, n=n1, k=k1, lda=m1, ldb=n1, ldc=n1, trans_a=True, trans_b=True, mode=2,
                      split_k=split1, k_dev=kdev),
        ops.gemm_desc(A2d, B2d, C2, m=m2, n=n2, k=k2, lda=k2, ldb=k2, ldc=n2, mode=1),
    ], prec)
    tol = (lambda k, ref: dict(rtol=2e-2, atol=2e-3 * float(ref.abs().max()))) if prec == 1 else \
          (lambda k, ref: dict(rtol=1e-5, atol=2e-5 * max(4.0, k ** 0.5)))
    ref0 = A0[:live0].double() @ B0.double().t() + bias0.double()
    torch.testing.assert_close(C0[:live0].cpu().double(), ref0, **tol(k0, ref0))
    assert bool((C0[live0:] == 7.0).all())
    ref1 = A1[:live1].double().t() @ B1[:live1].double() + 1.0
    torch.testing.assert_close(C1.cpu().double(), ref1, **tol(live1, ref1))
    ref2 = A2.double() @ B2.double().t()
    torch.testing.assert_close(C2.cpu().double(), ref2, **tol(k2, ref2))


@pytest.mark.parametrize("S,Dx,D", [(1, 5, 3), (333, 472, 100), (4000, 274, 100)])
def test_gru_cell_fwd_bwd(ops, S, Dx, D):
    torch.manual_seed(S)
    cell = torch.nn.GRUCell(Dx, D)
    x = torch.randn(S, Dx, requires_grad=True); h = torch.randn(S, D, requires_grad=True)
    ref = cell(x, h)
    wgt = torch.randn(S, D)
    (ref * wgt).sum().backward()
    p = [q.detach().to(DEV).requires_grad_() for q in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)]
    xg = x.detach().to(DEV).requires_grad_(); hg = h.detach().to(DEV).requires_grad_()
    out = ops.gru_cell(xg, hg, *p)
    torch.testing.assert_close(out.cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    (out * wgt.to(DEV)).sum().backward()
    tol = dict(rtol=1e-4, atol=1e-4 * max(1.0, S / 300))
    torch.testing.assert_close(xg.grad.cpu(), x.grad, **tol)
    torch.testing.assert_close(hg.grad.cpu(), h.grad, **tol)
    for a, b in zip(p, (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)):
        torch.testing.assert_close(a.grad.cpu(), b.grad, **tol)


@pytest.mark.parametrize("prec", [3, 1])
@pytest.mark.parametrize("S,Dx,D,live", [(1, 5, 4, 1), (333, 472, 100, 333), (4785, 301, 100, 4700), (130, 64, 40, 129),
                                          (700, 100, 172, 700)])
def test_gru_fused_fwd_matches_grucell(ops, S, Dx, D, live, prec):
    """tgn_gru_fused_fwd (both gate GEMMs in TMEM, gate math in the epilogue) against torch.nn.GRUCell
    (memory_module.py:72,172) and against the unfused path (tgn_gemm_batch + tgn_gru_gates_fwd)."""
    torch.manual_seed(S + D)
    cell = torch.nn.GRUCell(Dx, D)
    x, h = torch.randn(S, Dx), torch.randn(S, D)
    ref = cell(x, h).detach()
    ldx = (Dx + 3) // 4 * 4
    xp = torch.zeros(S, ldx); xp[:, :Dx] = x
    wp = torch.zeros(3 * D, ldx); wp[:, :Dx] = cell.weight_ih.detach()
    xd, hd, wd = xp.to(DEV), h.to(DEV), wp.to(DEV)
    whh, bih, bhh = (q.detach().to(DEV).contiguous() for q in (cell.weight_hh, cell.bias_ih, cell.bias_hh))
    out = torch.full((S, D), 7.0, device=DEV); gates = torch.full((S, 4 * D), 7.0, device=DEV)
    ndev = torch.tensor([live], dtype=torch.int32, device=DEV)
    ops.gru_fused_fwd(xd, hd, wd, whh, bih, bhh, dx=Dx, ldx=ldx, ldw=ldx, num=S, num_dev=ndev, prec=prec,
                      out=out, gates=gates)
    tol = dict(rtol=1e-5, atol=1e-5) if prec == 3 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(out[:live].cpu(), ref[:live], **tol)
    assert bool((out[live:] == 7.0).all()) and bool((gates[live:] == 7.0).all())      # dead rows untouched
    # unfused path at the same precision
    gi, gh = torch.empty(S, 3 * D, device=DEV), torch.empty(S, 3 * D, device=DEV)
    ops.gemm_batch([ops.gemm_desc(xd, wd, gi, m=S, n=3 * D, k=Dx, lda=ldx, ldb=ldx, ldc=3 * D, bias=bih),
                    ops.gemm_desc(hd, whh, gh, m=S, n=3 * D, k=D, lda=D, ldb=D, ldc=3 * D, bias=bhh)], prec)
    from tgn_b200 import _cabi
    out2, gates2 = torch.empty(S, D, device=DEV), torch.empty(S, 4 * D, device=DEV)
    _cabi.check(_cabi.lib().tgn_gru_gates_fwd(gi.data_ptr(), gh.data_ptr(), hd.data_ptr(), None, S, None, D,
                                              out2.data_ptr(), gates2.data_ptr(), 0))
    torch.testing.assert_close(out[:live], out2[:live], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(gates[:live], gates2[:live], rtol=1e-6, atol=1e-6)


def test_time_encode_fwd_bwd(ops):
    torch.manual_seed(0)
    te = orc.tp.TimeEncoder(100)
    t = (torch.rand(777) * 2.6e6).float()        # wiki-scale deltas: full-range cosf, fma rounding
    ref = te(t)
    (ref * torch.linspace(-1, 1, 100)).sum().backward()
    w = te.lin.weight.detach().view(-1).to(DEV).requires_grad_(); b = te.lin.bias.detach().to(DEV).requires_grad_()
    out = ops.time_encode_autograd(t.to(DEV), w, b)
    torch.testing.assert_close(out.cpu(), ref.detach(), rtol=0, atol=2e-6)
    (out * torch.linspace(-1, 1, 100, device=DEV)).sum().backward()
    torch.testing.assert_close(w.grad.cpu(), te.lin.weight.grad.view(-1), rtol=1e-4, atol=1.0)  # grads ~ 1e6-scale
    torch.testing.assert_close(b.grad.cpu(), te.lin.bias.grad, rtol=1e-4, atol=1e-3)


def test_link_score_and_mrr(ops):
    torch.manual_seed(1)
    lp = orc.LinkPredictor(100)
    z = torch.randn(300, 100)
    a = torch.randint(0, 300, (500,)); b = torch.randint(0, 300, (500,))
    ref_logit = lp.logits(z[a], z[b]).view(-1).detach(); ref_prob = lp(z[a], z[b]).view(-1).detach()
    zd = z.to(DEV)
    hs = ops.sgemm(zd, lp.lin_src.weight.detach().to(DEV), lp.lin_src.bias.detach().to(DEV), m=300, n=100, k=100, lda=100, ldb=100)
    hd = ops.sgemm(zd, lp.lin_dst.weight.detach().to(DEV), lp.lin_dst.bias.detach().to(DEV), m=300, n=100, k=100, lda=100, ldb=100)
    wf = lp.lin_final.weight.detach().view(-1).to(DEV); bf = lp.lin_final.bias.detach().to(DEV)
    torch.testing.assert_close(ops.link_score(hs, hd, a.to(DEV), b.to(DEV), wf, bf, False).cpu(), ref_logit, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ops.link_score(hs, hd, a.to(DEV), b.to(DEV), wf, bf, True).cpu(), ref_prob, rtol=1e-5, atol=1e-5)
    rng = np.random.default_rng(0)
    pos = rng.random(200).astype(np.float32); neg = rng.random((200, 999)).astype(np.float32)
    neg[:, :5] = pos[:, None]     # ties
    np.testing.assert_allclose(ops.mrr(cu(pos), cu(neg)).cpu().numpy(), orc.mrr_ref(pos, neg), rtol=1e-6)


def test_adam_matches_torch(ops):
    torch.manual_seed(0)
    p = torch.randn(1000, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-3)
    pg = p.detach().clone().to(DEV); m = torch.zeros_like(pg); v = torch.zeros_like(pg)
    step = torch.zeros(1, device=DEV)
    for i in range(5):
        g = torch.randn(1000)
        p.grad = g.clone(); opt.step()
        ops.adam_step(pg, g.to(DEV), m, v, step, 1e-3)
    torch.testing.assert_close(pg.cpu(), p.detach(), rtol=1e-5, atol=1e-6)
    assert float(step) == 5.0


# ------------------------------------------------------------------ decoder kernels
@pytest.mark.parametrize("B,D,Nb", [(1, 4, 3), (37, 32, 60), (200, 100, 500)])
def test_dec_fused_matches_autograd(ops, B, D, Nb):
    """tgn_dec_fused (forward + loss + all decoder gradients, one launch) against torch autograd on the
    reference decoder (modules/decoder.py:24-27) + BCEWithLogitsLoss (pyg-mem-tgn.py:51)."""
    from tgn_b200 import _cabi
    L = _cabi.lib()
    p = lambda t: None if t is None else t.data_ptr()
    g = torch.Generator(device="cpu").manual_seed(B * D)
    emb = torch.randn(Nb, D, generator=g, requires_grad=True)
    lp = orc.LinkPredictor(D)
    ids = torch.randint(0, Nb, (3 * B,), generator=g)
    zs, zd, zn = emb[ids[:B]], emb[ids[B:2 * B]], emb[ids[2 * B:]]
    logit = lambda a, b: lp.lin_final((lp.lin_src(a) + lp.lin_dst(b)).relu()).view(-1)
    pos, neg = logit(zs, zd), logit(zs, zn)
    crit = torch.nn.BCEWithLogitsLoss()
    loss = crit(pos, torch.ones_like(pos)) + crit(neg, torch.zeros_like(neg))
    loss.backward()
    cu = lambda t: t.detach().to(DEV).contiguous()
    w = {k: cu(v) for k, v in lp.state_dict().items()}
    out_loss = torch.zeros(1, device=DEV); logits = torch.zeros(2 * B, device=DEV)
    d_emb = torch.zeros(Nb, D, device=DEV)
    gr = {k: torch.zeros_like(v) for k, v in w.items()}
    emb_d, ids_d = cu(emb), cu(ids)      # keep the device copies alive across the asynchronous launch
    _cabi.check(L.tgn_dec_fused(p(emb_d), p(ids_d), B, D, p(w["lin_src.weight"]), p(w["lin_src.bias"]),
                                p(w["lin_dst.weight"]), p(w["lin_dst.bias"]), p(w["lin_final.weight"]),
                                p(w["lin_final.bias"]), p(out_loss), p(logits), p(d_emb), p(gr["lin_src.weight"]),
                                p(gr["lin_src.bias"]), p(gr["lin_dst.weight"]), p(gr["lin_dst.bias"]),
                                p(gr["lin_final.weight"]), p(gr["lin_final.bias"]), None, None, 0))
    # deferred weight gradients: same launch without the [D,D] accumulators; dW formed from the emitted rows
    out2 = torch.zeros(1, device=DEV); d_emb2 = torch.zeros(Nb, D, device=DEV)
    gr2 = {k: torch.zeros_like(v) for k, v in w.items()}
    z_rows, g_rows = torch.zeros(3 * B, D, device=DEV), torch.zeros(3 * B, D, device=DEV)
    _cabi.check(L.tgn_dec_fused(p(emb_d), p(ids_d), B, D, p(w["lin_src.weight"]), p(w["lin_src.bias"]),
                                p(w["lin_dst.weight"]), p(w["lin_dst.bias"]), p(w["lin_final.weight"]),
                                p(w["lin_final.bias"]), p(out2), None, p(d_emb2), None,
                                p(gr2["lin_src.bias"]), None, p(gr2["lin_dst.bias"]),
                                p(gr2["lin_final.weight"]), p(gr2["lin_final.bias"]), p(z_rows), p(g_rows), 0))
    torch.testing.assert_close(out2, out_loss, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(d_emb2, d_emb, rtol=1e-5, atol=1e-7)
    assert torch.equal(z_rows, emb_d[ids_d])
    torch.testing.assert_close(g_rows[:B].t() @ z_rows[:B], gr["lin_src.weight"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(g_rows[B:].t() @ z_rows[B:], gr["lin_dst.weight"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out_loss.cpu()[0], loss.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits.cpu(), torch.cat([pos, neg]).detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d_emb.cpu(), emb.grad, rtol=1e-4, atol=1e-7)
    for k, v in lp.named_parameters():
        torch.testing.assert_close(gr[k].cpu(), v.grad, rtol=1e-4, atol=1e-7, msg=lambda m: f"{k}: {m}")


def test_score_negs_counts(ops):
    """tgn_score_negs: sigmoid scores of the positives / negatives and the two TGB rank counts."""
    from tgn_b200 import _cabi
    L = _cabi.lib()
    p = lambda t: None if t is None else t.data_ptr()
    g = torch.Generator(device="cpu").manual_seed(3)
    Nb, D, B, Q = 300, 100, 23, 57
    hs, hd = torch.randn(Nb, D, generator=g), torch.randn(Nb, D, generator=g)
    wf, bf = torch.randn(D, generator=g) * 0.1, torch.randn(1, generator=g)
    src, dst = torch.randint(0, Nb, (B,), generator=g), torch.randint(0, Nb, (B,), generator=g)
    neg = torch.randint(0, Nb, (B, Q), generator=g)
    neg[:, 5] = dst                                    # exact ties with the positive
    score = lambda a, b: torch.sigmoid((hs[a] + hd[b]).relu() @ wf + bf)
    pos_r = score(src, dst)
    neg_r = score(src.repeat_interleave(Q), neg.reshape(-1)).view(B, Q)
    cu = lambda t: t.to(DEV).contiguous()
    pos, negs = torch.zeros(B, device=DEV), torch.zeros(B, Q, device=DEV)
    gt, ge = torch.zeros(B, dtype=torch.int32, device=DEV), torch.zeros(B, dtype=torch.int32, device=DEV)
    keep = [cu(x) for x in (hs, hd, src, dst, neg, wf, bf)]   # alive across the asynchronous launch
    _cabi.check(L.tgn_score_negs(*(p(x) for x in keep[:5]), B, Q, D, p(keep[5]), p(keep[6]),
                                 p(pos), p(negs), p(gt), p(ge), 0))
    torch.testing.assert_close(pos.cpu(), pos_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(negs.cpu(), neg_r, rtol=1e-5, atol=1e-6)
    # counts are computed from the kernel's own scores: consistent with them, and ties count as >= only
    assert torch.equal(gt.cpu().long(), (negs > pos[:, None]).sum(1).cpu()) and torch.equal(ge.cpu().long(), (negs >= pos[:, None]).sum(1).cpu())
    assert bool((ge > gt).all())
    assert torch.equal(ops.mrr(pos, negs).cpu(), 1.0 / (0.5 * (gt + ge).float().cpu() + 1.0))


# ------------------------------------------------------------------ owner-partitioned memory
@pytest.mark.parametrize("world", [2, 3, 8])
def test_partitioned_memory_rows_assemble_exactly(ops, world):
    """Emulates `world` ranks on one GPU: each rank's tgn_part_gather contribution is summed (what
    the NCCL all-reduce does) and fed to tgn_msg_build_gathered; the result must equal
    tgn_msg_build_ld on the unpartitioned table bit for bit.  Same for the owner-side write-back."""
    import ctypes
    from tgn_b200 import _cabi
    L = _cabi.lib()
    p = lambda t: None if t is None else t.data_ptr()
    N, De, D, Dt, B = 257, 5, 16, 12, 64
    rng = np.random.default_rng(world)
    store = ops.MsgStore(N, De, DEV, capacity=4 * B, t_dtype=torch.int64)
    for b in range(3):   # three batches: later batches overwrite the per-node runs of earlier ones
        src = torch.from_numpy(rng.integers(0, N // 2, B)).to(DEV)
        dst = torch.from_numpy(rng.integers(N // 2, N, B)).to(DEV)
        t = torch.from_numpy(np.sort(rng.integers(100 * b, 100 * b + 50, B))).to(DEV)
        store.update(src, dst, t, torch.from_numpy(rng.standard_normal((B, De)).astype(np.float32)).to(DEV))
    memory = torch.from_numpy(rng.standard_normal((N, D)).astype(np.float32)).to(DEV)
    last_update = torch.from_numpy(rng.integers(0, 90, N)).to(DEV)
    tw = torch.from_numpy(rng.standard_normal(Dt).astype(np.float32) * 0.01).to(DEV)
    tb = torch.from_numpy(rng.standard_normal(Dt).astype(np.float32)).to(DEV)
    n_id = torch.from_numpy(np.unique(rng.integers(0, N, 150))).to(DEV)
    S = n_id.numel()
    cnt = torch.tensor([S - 7], dtype=torch.int32, device=DEV)      # live rows < bound
    ldx = 2 * D + De + Dt + 3
    st = ctypes.byref(store.struct())

    def outputs():
        return (torch.zeros(S, ldx, device=DEV), torch.zeros(S, D, device=DEV), torch.zeros(S, Dt, device=DEV),
                torch.zeros(S, dtype=torch.long, device=DEV), torch.zeros(S, dtype=torch.int32, device=DEV),
                torch.zeros(S, device=DEV))
    ref = outputs()
    _cabi.check(L.tgn_msg_build_ld(st, p(n_id), S, p(cnt), 0, p(memory), p(last_update), D, p(tw), p(tb), Dt,
                                   p(ref[0]), ldx, p(ref[1]), p(ref[2]), p(ref[3]), p(ref[4]), p(ref[5]), 0))
    g_sum = torch.zeros(2, S, D, device=DEV)
    lu_sum = torch.zeros(S, dtype=torch.long, device=DEV)
    Nloc = (N + world - 1) // world
    for r in range(world):
        mem_loc = torch.zeros(Nloc, D, device=DEV)
        lu_loc = torch.zeros(Nloc, dtype=torch.long, device=DEV)
        own = memory[r::world]
        mem_loc[:own.shape[0]] = own
        lu_loc[:own.shape[0]] = last_update[r::world]
        g = torch.full((2, S, D), 7.0, device=DEV)
        lu = torch.full((S,), 7, dtype=torch.long, device=DEV)
        _cabi.check(L.tgn_part_gather(st, p(n_id), S, p(cnt), p(mem_loc), p(lu_loc), D, r, world, p(g),
                                      g.data_ptr() + 4 * S * D, p(lu), None, 0))
        g_sum += g
        lu_sum += lu
    got = outputs()
    _cabi.check(L.tgn_msg_build_gathered(st, p(n_id), S, p(cnt), p(g_sum), g_sum.data_ptr() + 4 * S * D, p(lu_sum),
                                         D, p(tw), p(tb), Dt, p(got[0]), ldx, p(got[1]), p(got[2]), p(got[3]),
                                         p(got[4]), p(got[5]), 0))
    live = S - 7
    for a, b in zip(ref, got):
        assert torch.equal(a[:live], b[:live])
    # owner-side write-back == tgn_memory_scatter on the full table
    new_mem = torch.from_numpy(rng.standard_normal((S, D)).astype(np.float32)).to(DEV)
    new_lu = torch.from_numpy(rng.integers(0, 1000, S)).to(DEV)
    full_m, full_l = memory.clone(), last_update.clone()
    ops.memory_scatter(n_id, new_mem, new_lu, full_m, full_l)
    for r in range(world):
        own = memory[r::world]
        mem_loc = torch.zeros(Nloc, D, device=DEV); lu_loc = torch.zeros(Nloc, dtype=torch.long, device=DEV)
        mem_loc[:own.shape[0]] = own; lu_loc[:own.shape[0]] = last_update[r::world]
        _cabi.check(L.tgn_memory_scatter_owned(p(n_id), S, None, p(new_mem), p(new_lu), 0, None, D, r, world,
                                               p(mem_loc), p(lu_loc), 0))
        assert torch.equal(mem_loc[:own.shape[0]], full_m[r::world]) and torch.equal(lu_loc[:own.shape[0]], full_l[r::world])


@pytest.mark.parametrize("H,C", [(2, 50), (1, 32), (4, 16)])
def test_attn_core_small_path_equals_general_path_with_dropout(ops, H, C):
    """The register-resident low-degree attention kernels (lane-parallel Philox dropout masks) against the
    general kernels (one draw per (edge, head)): same seed -> same mask, so outputs, attention weights and
    all gradients agree; also checks that forward and backward apply the SAME mask."""
    from tgn_b200 import _cabi
    L = _cabi.lib()
    p = lambda t: None if t is None else t.data_ptr()
    g = torch.Generator(device="cpu").manual_seed(H * C)
    R, Nb, K, HC = 77, 300, 10, H * C
    deg = torch.randint(0, K + 1, (R,), generator=g)
    row_ptr = torch.zeros(R + 1, dtype=torch.int32); row_ptr[1:] = deg.cumsum(0).int()
    E = int(row_ptr[-1])
    nbr = torch.randint(0, Nb, (E,), generator=g)
    centres = torch.randperm(Nb, generator=g)[:R]
    proj, ee = torch.randn(Nb, 4 * HC, generator=g), torch.randn(E, HC, generator=g)
    d_out = torch.randn(Nb, HC, generator=g)
    dev = [x.to(DEV).contiguous() for x in (proj, nbr, row_ptr, centres, ee, d_out)]
    res = []
    for max_deg in (K, 0):                       # K: small path, 0: general path
        out, alpha = torch.zeros(Nb, HC, device=DEV), torch.zeros(E, H, device=DEV)
        d_proj, d_ee = torch.zeros(Nb, 4 * HC, device=DEV), torch.zeros(E, HC, device=DEV)
        _cabi.check(L.tgn_attn_core_fwd(p(dev[0]), p(dev[1]), p(dev[2]), p(dev[3]), R, None, H, C, p(dev[4]), 0.3,
                                        99, None, max_deg, p(out), p(alpha), 0))
        _cabi.check(L.tgn_attn_core_bwd(p(dev[0]), p(dev[1]), p(dev[2]), p(dev[3]), R, None, H, C, p(dev[4]),
                                        p(alpha), p(dev[5]), 0.3, 99, None, max_deg, Nb, p(d_proj), p(d_ee), 0))
        res.append((out, alpha, d_proj, d_ee))
    for a, b, name in zip(res[0], res[1], ("out", "alpha", "d_proj", "d_ee")):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5, msg=lambda m: f"{name}: {m}")
    # dropout really happened (30 % of the (edge, head) pairs carry no weight into `out`)
    out_nodrop = torch.zeros(Nb, HC, device=DEV); al2 = torch.zeros(E, H, device=DEV)
    _cabi.check(L.tgn_attn_core_fwd(p(dev[0]), p(dev[1]), p(dev[2]), p(dev[3]), R, None, H, C, p(dev[4]), 0.0,
                                    99, None, K, p(out_nodrop), p(al2), 0))
    assert not torch.allclose(out_nodrop, res[0][0])


@pytest.mark.parametrize("K,R", [(10, 150_000), (7, 40_000), (10, 500)])
def test_ring_lookup_writes_stay_inside_their_buffers(ops, K, R):
    """Guard bands around every output of tgn_nbr_lookup (and its workspace) on the large-launch path (count +
    emit passes) and the single-launch path: nothing outside [0, count) / [0, R] / the declared workspace is
    written (compute-sanitizer is not available on the pool)."""
    import ctypes
    from tgn_b200 import _cabi
    L = _cabi.lib()
    N, G = 20_000, 64
    g = torch.Generator(device=DEV).manual_seed(R)
    e_id = torch.randint(-1, 1_000_000, (N, K), device=DEV, generator=g)
    nbrs = torch.randint(0, N, (N, K), device=DEV, generator=g)
    t = torch.rand((N, K), device=DEV, generator=g)
    roots = torch.randint(0, N, (R,), device=DEV, generator=g)
    count = int((e_id[roots] >= 0).sum())
    cap = R * K
    mk = lambda n, dt, v: torch.full((n + 2 * G,), v, dtype=dt, device=DEV)
    o_n, o_c, o_e = (mk(cap, torch.int64, -777) for _ in range(3))
    o_t = mk(cap, torch.float32, -777.0)
    off = mk(R + 1, torch.int32, -777)
    cnt = mk(1, torch.int32, -777)
    wsn = L.tgn_nbr_lookup_ws_bytes(R, K) // 8
    ws = mk(wsn, torch.int64, -777)
    p = lambda x: x.data_ptr() + G * x.element_size()
    _cabi.check(L.tgn_nbr_lookup(roots.data_ptr(), R, None, K, N, nbrs.data_ptr(), e_id.data_ptr(), t.data_ptr(),
                                 p(o_n), p(o_c), p(o_e), p(o_t), p(off), p(cnt), None, p(ws),
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(cnt[G]) == count
    for buf, n in ((o_n, cap), (o_c, cap), (o_e, cap), (o_t, cap), (off, R + 1), (cnt, 1), (ws, wsn)):
        assert bool((buf[:G] == -777).all()) and bool((buf[G + n:] == -777).all())
    for buf in (o_n, o_c, o_e, o_t):          # the compacted outputs end at `count`
        assert bool((buf[G + count:G + cap] == -777).all())
    valid = e_id[roots] >= 0
    assert torch.equal(o_e[G:G + count], e_id[roots][valid])
