"""CPU checks of the drop-in boundary: the shared library loads, exports every
symbol include/tgn_b200.h declares, and argument validation answers without a GPU."""
import ctypes
import os
import subprocess

import pytest

from tgn_b200 import _cabi


def test_header_declares_the_hot_path():
    syms = _cabi.declared_symbols()
    for must in ["tgn_nbr_lookup", "tgn_nbr_insert", "tgn_tcsr_sample", "tgn_agg_last", "tgn_agg_mean",
                 "tgn_msgstore_update", "tgn_msg_build", "tgn_sgemm", "tgn_gru_gates_fwd",
                 "tgn_gru_gates_bwd", "tgn_attn_fwd", "tgn_attn_bwd", "tgn_unique_rank", "tgn_mrr"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "libtgn_b200.so not built (run __graft_entry__.build())"
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in _cabi.declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = [s for s in exported if s.startswith("tgn_") and s not in _cabi.declared_symbols()]
    assert not extra, f"exported but not declared: {extra}"


def test_library_loads_and_reports_version():
    lib = _cabi.lib()
    assert lib.tgn_abi_version() == 1
    assert lib.tgn_bitmap_bytes(9227) == (10 * 32 + 1) * 4   # 10 groups of 1024 nodes + 1 summary word
    # ticket word + one word per 2048-slot tile (3) + one word per group of 8 tiles (count pass of large launches)
    assert lib.tgn_nbr_lookup_ws_bytes(600, 10) == (1 + 3 + 1) * 8


def test_argument_validation_needs_no_gpu():
    lib = _cabi.lib()
    rc = lib.tgn_nbr_insert(None, None, None, 5000, 0, None, 10, 100, None, None, None, None)
    assert rc == _cabi.TGN_EINVAL and b"TGN_SORT_MAX" in lib.tgn_last_error()
    rc = lib.tgn_nbr_lookup(None, 4, None, 0, 10, None, None, None, None, None, None, None, None, None,
                            None, None, None)
    assert rc == _cabi.TGN_EINVAL and b"size_k" in lib.tgn_last_error()
    with pytest.raises(_cabi.TgnError):
        _cabi.check(lib.tgn_tcsr_sample(None, None, None, None, None, 0, 10, None, None, 4, 10, 7, 0.0, 0.0, 0,
                                        None, None, None, None, None, None, None, None, None))


def test_round2_entry_points_validate_without_a_gpu():
    """The entry points added in round 2 answer bad arguments with TGN_EINVAL before any CUDA call."""
    lib = _cabi.lib()
    buf = (ctypes.c_void_p * 2)(1024, 2048)
    # tgn_peer_bcast: sizes / offsets are multiples of 16 bytes, world <= 16
    assert lib.tgn_peer_bcast(ctypes.c_void_p(4096), 60, buf, 0, 2, None) == _cabi.TGN_EINVAL
    assert b"16" in lib.tgn_last_error()
    assert lib.tgn_peer_bcast(ctypes.c_void_p(4096), 64, buf, 0, 17, None) == _cabi.TGN_EINVAL
    assert lib.tgn_peer_bcast(None, 0, None, 0, 1, None) == 0          # an empty block is a no-op
    # tgn_dec_attn_fused: two heads, at most 10 edges per centre
    args = [None] * 4 + [4, 4, 25] + [None, 0.0, 0, None, 10] + [None] * 3 + [8] + [None] * 16
    assert lib.tgn_dec_attn_fused(*args) == _cabi.TGN_EINVAL and b"two heads" in lib.tgn_last_error()
    args[5], args[11] = 2, 12
    assert lib.tgn_dec_attn_fused(*args) == _cabi.TGN_EINVAL and b"10 edges" in lib.tgn_last_error()
    # tgn_msg_build_p2p: needs the peer tables
    assert lib.tgn_msg_build_p2p(None, None, 4, None, None, None, 2, 100, None, None, 100, None, 304, None, None,
                                 None, None, None, None) == _cabi.TGN_EINVAL
    # workspace of the ring lookup: ticket + tiles + groups of eight tiles
    assert lib.tgn_nbr_lookup_ws_bytes(1_000_000, 10) == (1 + 4902 + 613) * 8


def test_msgstore_struct_matches_header():
    src = open(_cabi.HEADER).read()
    body = src[src.index("typedef struct tgn_msgstore {"):src.index("} tgn_msgstore;")]
    import re
    names = re.findall(r"(\w+);", _cabi._strip_comments(body))
    assert names == [f[0] for f in _cabi.MsgStoreStruct._fields_]


def test_product_modules_refuse_cpu():
    import torch
    from neighbor_loader import LastNeighborLoader
    with pytest.raises(RuntimeError):
        LastNeighborLoader(10, 2, device="cpu")
    from tgn_b200 import ops
    with pytest.raises(_cabi.TgnError):
        ops.agg_last(torch.zeros(2, 2), torch.zeros(2, dtype=torch.long), torch.zeros(2, dtype=torch.long), 2)
