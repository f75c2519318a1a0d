"""GPU: the torch C++ extension (torch.ops.tgn.*, csrc_ext/tgn_torch.cpp, built by setup.py build_ext) returns
exactly what the ctypes route to the same C-ABI returns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_extension_ops_equal_ctypes_route():
    from tgn_b200 import _cabi, ops, torch_ext
    assert torch_ext.available(), "run `python setup.py build_ext --inplace` (or __graft_entry__.build())"
    g = torch.Generator().manual_seed(0)
    N, K, R = 500, 6, 200
    nb = torch.randint(0, N, (N, K), generator=g).to(DEV)
    ei = torch.randint(-1, 1000, (N, K), generator=g).to(DEV)
    tt = torch.rand(N, K, generator=g).to(DEV)
    n_id = torch.randperm(N, generator=g)[:R].to(DEV)
    via_ext = torch.ops.tgn.nbr_lookup(n_id, nb, ei, tt, None)
    L = _cabi.lib()
    p = ops._p
    cap = R * K
    o = [torch.empty(cap, dtype=torch.long, device=DEV) for _ in range(3)] + [torch.empty(cap, device=DEV)]
    off, cnt = torch.empty(R + 1, dtype=torch.int32, device=DEV), torch.empty(1, dtype=torch.int32, device=DEV)
    ws = torch.empty(max(L.tgn_nbr_lookup_ws_bytes(R, K), 16) // 8, dtype=torch.long, device=DEV)
    _cabi.check(L.tgn_nbr_lookup(p(n_id), R, None, K, N, p(nb), p(ei), p(tt), p(o[0]), p(o[1]), p(o[2]), p(o[3]), p(off),
                                 p(cnt), None, p(ws), ops._stream()))
    n = int(cnt)
    assert int(via_ext[5]) == n and torch.equal(via_ext[4], off)
    for a, b in zip(via_ext[:4], o):
        assert torch.equal(a[:n], b[:n])
    # dependency blocks and the aggregators
    src, dst = torch.randint(0, 50, (1000,), generator=g).to(DEV), torch.randint(0, 50, (1000,), generator=g).to(DEV)
    assert torch.equal(torch.ops.tgn.dep_blocks(src, dst, 200), ops.dep_blocks(src, dst, 200))
    msg = torch.randn(300, 12, generator=g).to(DEV); idx = torch.randint(0, 40, (300,), generator=g).to(DEV)
    tm = torch.randint(0, 9, (300,), generator=g).to(DEV)
    a, b = torch.ops.tgn.agg_last(msg, idx, tm, 40), ops.agg_last(msg, idx, tm, 40)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(torch.ops.tgn.agg_mean(msg, idx, 40), ops.agg_mean(msg, idx, 40))
    with pytest.raises(NotImplementedError, match="CPU"):      # registered for the CUDA backend only: no CPU fallback
        torch.ops.tgn.agg_mean(msg.cpu(), idx.cpu(), 40)


def test_neighbor_loader_runs_on_the_extension():
    """LastNeighborLoader / sampler_core go through ops.*, which prefers the extension when it is built: golden
    parity of those classes (tests/test_gpu_kernels.py) therefore covers it; here: the route is actually taken."""
    from neighbor_loader import LastNeighborLoader
    from tgn_b200 import ops
    assert ops._ext()
    calls = []
    orig = torch.ops.tgn.nbr_insert
    loader = LastNeighborLoader(100, 4, device=DEV)
    src, dst = torch.arange(0, 10, device=DEV), torch.arange(50, 60, device=DEV)
    loader.insert(src, dst, torch.arange(10, device=DEV).float())
    ids, ei, e_id, t = loader(torch.arange(0, 60, device=DEV))
    assert e_id.numel() == 20 and int(loader.e_id.max()) == 9
