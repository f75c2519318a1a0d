"""GPU: dependency-aware block ids (csrc/depblock.cu, tgn_dep_blocks) against the golden vectors produced by
the UNMODIFIED reference get_block (dependencyGraph.py:8-28; tests/golden/callers.npz) and, on larger and
nastier streams, against the host restatement that is itself pinned to the same golden vectors
(tests/test_callers_cpu.py): bit-exact integer work."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers.npz"))


def _host(src, dst, B):
    from dependencyGraph import get_block
    out = []
    for lo in range(0, len(src), B):
        out += get_block(None, src[lo:lo + B], dst[lo:lo + B])
    return np.asarray(out, np.int32)


def test_dep_blocks_golden():
    from tgn_b200 import ops
    for c in range(int(G["num_block_cases"])):
        src, dst = G[f"b{c}_src"], G[f"b{c}_dst"]
        got = ops.dep_blocks(torch.from_numpy(src).to(DEV), torch.from_numpy(dst).to(DEV), len(src))
        assert got.cpu().tolist() == G[f"b{c}_blocks"].tolist(), c


@pytest.mark.parametrize("E,B,N", [(1, 1, 3), (999, 200, 50), (10_000, 200, 9_227), (30_000, 2000, 1000),
                                   (4096 * 3 + 5, 4096, 300), (5000, 600, 2)])
def test_dep_blocks_random_streams(E, B, N):
    """ragged last batch, hubs (Zipf-like popularity), self loops (src == dst), two-node worst-case chains,
    the largest batch the shared-memory sort takes (2B = TGN_SORT_MAX)"""
    from tgn_b200 import ops
    rng = np.random.default_rng(E + B)
    src = np.floor(rng.random(E) ** 3 * N).astype(np.int64)
    dst = np.floor(rng.random(E) ** 3 * N).astype(np.int64)      # same range: self loops occur
    ids, cnt = ops.dep_blocks(torch.from_numpy(src).to(DEV), torch.from_numpy(dst).to(DEV), B, want_counts=True)
    want = _host(src, dst, B)
    assert np.array_equal(ids.cpu().numpy(), want)
    per_batch = [int(want[lo:lo + B].max()) + 1 for lo in range(0, E, B)]
    assert cnt.cpu().tolist() == per_batch
    # the defining property (size independent): events of one block inside a batch touch disjoint nodes
    lo = (E // B // 2) * B
    b = want[lo:lo + B]
    for k in range(int(b.max()) + 1):
        s, d = src[lo:lo + B][b == k], dst[lo:lo + B][b == k]
        assert len(np.unique(np.concatenate([s, d]))) == 2 * len(s) - int((s == d).sum())


def test_dependency_aware_batch_uses_the_kernel_and_equals_the_host_walk():
    import dependencyGraph as dg
    from temporal_dataset import TemporalGraphDataset, TensorBatchLoader
    rng = np.random.default_rng(3)
    E, B = 2345, 200
    src = torch.from_numpy(rng.integers(0, 80, E)); dst = torch.from_numpy(rng.integers(60, 160, E))
    t = torch.arange(E); msg = torch.zeros(E, 1)
    loader = TensorBatchLoader(TemporalGraphDataset(src, dst, t, msg), B)
    assert dg._device_blocks(loader) is not None
    flat = dg.dependecyAwareBatch(loader)
    host = []
    for batch in loader:
        host += dg.get_block(batch["t"], batch["src"], batch["dst"])
    assert flat == host
    nested = dg.dependecyAwareBatch(loader, flat=False)
    assert sum(nested, []) == host and len(nested) == len(loader)


def test_dep_blocks_rejects_oversized_batches():
    from tgn_b200 import _cabi, ops
    x = torch.zeros(10, dtype=torch.long, device=DEV)
    with pytest.raises(_cabi.TgnError):
        ops.dep_blocks(x, x, 4097)
