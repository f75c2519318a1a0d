"""Tensor-level wrappers over the C-ABI (include/tgn_b200.h).

torch is used here for what the task calls plumbing: device memory, streams and
autograd bookkeeping.  Every numeric operation below is a kernel of
libtgn_b200.so; there is no eager/torch fallback -- a missing library or a CPU
tensor raises.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import AGG_LAST, AGG_MEAN, SAMPLE_RECENT, SAMPLE_UNIFORM, SORT_MAX, check

__all__ = [
    "unique_relabel", "nbr_lookup", "nbr_insert", "tcsr_sample", "tcsr_build_index", "tcsr_build", "gru_fused_fwd", "agg_last", "agg_mean",
    "MsgStore", "sgemm", "gru_cell", "time_encode", "temporal_attention", "link_score", "mrr",
    "memory_scatter", "gather_rows", "adam_step",
]


def _L():
    return _cabi.lib()


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need(t: Tensor, dtype, name: str) -> Tensor:
    if not t.is_cuda:
        raise _cabi.TgnError(f"{name}: expected a CUDA tensor (the sm_100a path has no CPU fallback)")
    if t.dtype != dtype:
        raise _cabi.TgnError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t.contiguous()


_bitmaps = {}


def _ext() -> bool:
    """True when the torch C++ extension (torch.ops.tgn.*, setup.py build_ext) is built: the sampler and
    aggregator entry points then go through it instead of ctypes (same C-ABI underneath)."""
    from . import torch_ext
    return torch_ext.available()


def _bitmap(num_nodes: int, device) -> Tensor:
    key = (int(num_nodes), str(device))
    bm = _bitmaps.get(key)
    if bm is None:
        nbytes = _L().tgn_bitmap_bytes(num_nodes)
        bm = torch.zeros(nbytes // 4, dtype=torch.int32, device=device)
        _bitmaps[key] = bm
    return bm


# ---------------------------------------------------------------------------
# unique / relabel
# ---------------------------------------------------------------------------
def unique_mark(ids: Tensor, num_nodes: int, bitmap: Tensor, count_dev: Optional[Tensor] = None):
    ids = _need(ids, torch.int64, "ids")
    check(_L().tgn_unique_mark(_p(ids), ids.numel(), _p(count_dev), num_nodes, _p(bitmap), _stream()))


def unique_rank(num_nodes: int, bitmap: Tensor, out_ids: Tensor, assoc: Optional[Tensor],
                out_count: Tensor, keep_marks: bool = False):
    check(_L().tgn_unique_rank(_p(bitmap), num_nodes, _p(out_ids), out_ids.numel(), _p(assoc),
                               _p(out_count), int(keep_marks), _stream()))


def unique_relabel(id_lists: Sequence[Tensor], num_nodes: int,
                   assoc: Optional[Tensor] = None) -> Tensor:
    """sorted-unique of the concatenation of id_lists; assoc[id] = rank (if given).
    Replaces torch.cat(...).unique() (+ `_assoc[n_id] = arange`)."""
    dev = id_lists[0].device
    bm = _bitmap(num_nodes, dev)
    total = 0
    for ids in id_lists:
        if ids.numel():
            unique_mark(ids.reshape(-1), num_nodes, bm)
            total += ids.numel()
    cap = min(total, num_nodes)
    out = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    unique_rank(num_nodes, bm, out, assoc, cnt)
    return out[: int(cnt.item())]


def relabel(ids: Tensor, assoc: Tensor, count_dev: Optional[Tensor] = None) -> Tensor:
    ids = _need(ids, torch.int64, "ids")
    out = torch.empty_like(ids)
    check(_L().tgn_relabel(_p(ids), ids.numel(), _p(count_dev), _p(assoc), _p(out), _stream()))
    return out


# ---------------------------------------------------------------------------
# LastNeighborLoader ring
# ---------------------------------------------------------------------------
def nbr_lookup_raw(n_id: Tensor, neighbors: Tensor, e_id: Tensor, t: Tensor, bitmap: Optional[Tensor],
                   num_roots_dev: Optional[Tensor] = None):
    """Bound-sized outputs + device count (no host sync).  Returns
    (nbr_global, centre_global, e_id, t, root_off, count_dev)."""
    n_id = _need(n_id, torch.int64, "n_id")
    if num_roots_dev is None and _ext():
        return tuple(torch.ops.tgn.nbr_lookup(n_id, neighbors, e_id, t, bitmap))
    N, K = neighbors.shape
    R = n_id.numel()
    dev = n_id.device
    cap = max(R * K, 1)
    o_n = torch.empty(cap, dtype=torch.int64, device=dev)
    o_c = torch.empty(cap, dtype=torch.int64, device=dev)
    o_e = torch.empty(cap, dtype=torch.int64, device=dev)
    o_t = torch.empty(cap, dtype=torch.float32, device=dev)
    off = torch.empty(R + 1, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(_L().tgn_nbr_lookup_ws_bytes(R, K), 16) // 8, dtype=torch.int64, device=dev)
    check(_L().tgn_nbr_lookup(_p(n_id), R, _p(num_roots_dev), K, N, _p(neighbors), _p(e_id), _p(t),
                              _p(o_n), _p(o_c), _p(o_e), _p(o_t), _p(off), _p(cnt), _p(bitmap),
                              _p(ws), _stream()))
    return o_n, o_c, o_e, o_t, off, cnt


def nbr_lookup(n_id: Tensor, neighbors: Tensor, e_id: Tensor, t: Tensor, assoc: Tensor):
    """LastNeighborLoader.__call__ (neighbor_loader.py:26-50).
    Returns (n_id_sorted_unique, edge_index[2,E], e_id[E], t[E], root_off[R+1])."""
    n_id = _need(n_id, torch.int64, "n_id")
    N = neighbors.shape[0]
    dev = n_id.device
    bm = _bitmap(N, dev)
    if n_id.numel():
        unique_mark(n_id, N, bm)
    o_n, o_c, o_e, o_t, off, cnt = nbr_lookup_raw(n_id, neighbors, e_id, t, bm)
    cap = min(n_id.numel() * (neighbors.shape[1] + 1), N)
    ids = torch.empty(max(cap, 1), dtype=torch.int64, device=dev)
    ucnt = torch.empty(1, dtype=torch.int32, device=dev)
    unique_rank(N, bm, ids, assoc, ucnt)
    E, U = int(cnt.item()), int(ucnt.item())  # the reference syncs here too (boolean-mask indexing)
    edge_index = torch.empty((2, E), dtype=torch.int64, device=dev)
    if E:
        check(_L().tgn_relabel(_p(o_n), E, None, _p(assoc), edge_index[0].data_ptr(), _stream()))
        check(_L().tgn_relabel(_p(o_c), E, None, _p(assoc), edge_index[1].data_ptr(), _stream()))
    return ids[:U], edge_index, o_e[:E], o_t[:E], off


def nbr_insert(src: Tensor, dst: Tensor, t: Tensor, cur_e_id: int, neighbors: Tensor, e_id: Tensor,
               t_state: Tensor, cur_e_id_dev: Optional[Tensor] = None):
    src = _need(src, torch.int64, "src")
    dst = _need(dst, torch.int64, "dst")
    t = _need(t, torch.float32, "t")
    if cur_e_id_dev is None and _ext() and src.numel():
        torch.ops.tgn.nbr_insert(src, dst, t, int(cur_e_id), neighbors, e_id, t_state)
        return
    N, K = neighbors.shape
    check(_L().tgn_nbr_insert(_p(src), _p(dst), _p(t), src.numel(), cur_e_id, _p(cur_e_id_dev), K, N,
                              _p(neighbors), _p(e_id), _p(t_state), _stream()))


# ---------------------------------------------------------------------------
# t-CSR sampler
# ---------------------------------------------------------------------------
def tcsr_build(src: Tensor, dst: Tensor, t: Tensor, num_nodes: int, add_reverse: bool = True,
               t_sorted: Optional[bool] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(indptr int32[N+1], indices int32, eid int32, ts float32) of the t-CSR graph TGL's gen_graph
    writes to ext_full.npz, built on the device.  `t_sorted=None` checks on the device whether the
    event stream is chronological (one reduction + host read; building a graph is not a per-batch op)."""
    src = _need(src, torch.int64, "src")
    dst = _need(dst, torch.int64, "dst")
    t = t.contiguous()
    if t.dtype not in (torch.int64, torch.float32) or not t.is_cuda:
        raise _cabi.TgnError(f"t must be a CUDA int64 or float32 tensor, got {t.dtype} on {t.device}")
    E, dev = src.numel(), src.device
    if dst.numel() != E or t.numel() != E:
        raise _cabi.TgnError("src, dst, t must have the same length")
    if t_sorted is None:
        t_sorted = bool((t[1:] >= t[:-1]).all()) if E > 1 else True
    n = (2 if add_reverse else 1) * E
    indptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    eid = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ts = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(_L().tgn_tcsr_build_ws_bytes(E, int(add_reverse)), 256), dtype=torch.uint8, device=dev)
    check(_L().tgn_tcsr_build(_p(src), _p(dst), _p(t), 1 if t.dtype == torch.float32 else 0, E, num_nodes,
                              int(add_reverse), int(t_sorted), _p(indptr), _p(indices), _p(eid), _p(ts),
                              _p(bad), _p(ws), _stream()))
    if int(bad.item()):
        raise _cabi.TgnError(f"tcsr_build: an endpoint is outside [0, {num_nodes})")
    return indptr, indices[:n], eid[:n], ts[:n]


def tcsr_build_index(ts: Tensor) -> Tensor:
    """Skip index of a t-CSR timestamp array (every 16th entry); build once per graph."""
    _need(ts, torch.float32, "ts")
    coarse = torch.empty(max(_L().tgn_tcsr_index_len(ts.numel()), 1), dtype=torch.float32, device=ts.device)
    check(_L().tgn_tcsr_build_index(_p(ts), ts.numel(), _p(coarse), _stream()))
    return coarse


def tcsr_sample(indptr: Tensor, indices: Tensor, eid: Tensor, ts: Tensor, roots: Tensor,
                root_ts: Tensor, k: int, strategy: int = SAMPLE_RECENT, offset: float = 0.0,
                duration: float = 0.0, seed: int = 0, coarse: Optional[Tensor] = None):
    """Returns bound-sized (nbr, col, eid, ts, dts), root_off[R+1], count_dev.
    `coarse` = tcsr_build_index(ts) switches the per-root search to the skip index."""
    if coarse is not None:
        _need(coarse, torch.float32, "coarse")
    for name, x in (("indptr", indptr), ("indices", indices), ("eid", eid), ("roots", roots)):
        _need(x, torch.int32, name)
    _need(ts, torch.float32, "ts")
    root_ts = _need(root_ts, torch.float32, "root_ts")
    roots = roots.contiguous()
    if _ext():
        s64 = int(seed) & 0xFFFFFFFFFFFFFFFF          # same 64 bits as the ctypes route, as a signed schema int
        s64 = s64 - (1 << 64) if s64 >= (1 << 63) else s64
        o = torch.ops.tgn.tcsr_sample(indptr, indices, eid, ts, coarse, roots, root_ts, int(k), int(strategy),
                                      float(offset), float(duration), s64)
        return tuple(o[:5]), o[5], o[6]
    R = roots.numel()
    dev = roots.device
    cap = max(R * k, 1)
    o_n = torch.empty(cap, dtype=torch.int32, device=dev)
    o_c = torch.empty(cap, dtype=torch.int32, device=dev)
    o_e = torch.empty(cap, dtype=torch.int32, device=dev)
    o_t = torch.empty(cap, dtype=torch.float32, device=dev)
    o_d = torch.empty(cap, dtype=torch.float32, device=dev)
    off = torch.empty(R + 1, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    ws = torch.empty(max(_L().tgn_tcsr_sample_ws_bytes(R), 16) // 8, dtype=torch.int64, device=dev)
    check(_L().tgn_tcsr_sample(_p(indptr), _p(indices), _p(eid), _p(ts), _p(coarse), ts.numel(), indptr.numel() - 1,
                               _p(roots), _p(root_ts), R, k, strategy, offset, duration,
                               int(seed) & 0xFFFFFFFFFFFFFFFF,
                               _p(o_n), _p(o_c), _p(o_e), _p(o_t), _p(o_d), _p(off), _p(cnt),
                               _p(ws), _stream()))
    return (o_n, o_c, o_e, o_t, o_d), off, cnt


# ---------------------------------------------------------------------------
# aggregators on materialised messages
# ---------------------------------------------------------------------------
def _t_flag(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return 1
    if t.dtype == torch.int64:
        return 0
    raise _cabi.TgnError(f"timestamps must be int64 or float32, got {t.dtype}")


def agg_last(msg: Tensor, index: Tensor, t: Tensor, dim_size: int) -> Tuple[Tensor, Tensor]:
    msg = _need(msg, torch.float32, "msg")
    index = _need(index, torch.int64, "index")
    t = t.contiguous()
    M, W = msg.shape
    out = torch.empty((dim_size, W), dtype=torch.float32, device=msg.device)
    argmax = torch.empty(dim_size, dtype=torch.int64, device=msg.device)
    ws = torch.empty(max(2 * dim_size, 1), dtype=torch.int64, device=msg.device)
    check(_L().tgn_agg_last(_p(msg), _p(index), _p(t), _t_flag(t), M, dim_size, W, _p(out),
                            _p(argmax), _p(ws), _stream()))
    return out, argmax


def agg_mean(msg: Tensor, index: Tensor, dim_size: int) -> Tensor:
    msg = _need(msg, torch.float32, "msg")
    index = _need(index, torch.int64, "index")
    M, W = msg.shape
    out = torch.empty((dim_size, W), dtype=torch.float32, device=msg.device)
    ws = torch.empty(max(_L().tgn_agg_mean_ws_bytes(M, dim_size), 16) // 8, dtype=torch.int64, device=msg.device)
    check(_L().tgn_agg_mean(_p(msg), _p(index), M, dim_size, W, _p(out), _p(ws), _stream()))
    return out


def ap_auc_accum(pos_logits: Tensor, neg_logits: Tensor, acc: Tensor):
    """acc (float64 [3] on the device) += (AP, AUC, 1) of one training batch (epoch_utils.py:312-315)."""
    logits = torch.cat([pos_logits.detach().reshape(-1), neg_logits.detach().reshape(-1)]).float().contiguous()
    check(_L().tgn_ap_auc_accum(_p(logits), pos_logits.numel(), neg_logits.numel(), _p(acc), _stream()))


def dep_blocks(src: Tensor, dst: Tensor, batch: int, want_counts: bool = False):
    """dependencyGraph.get_block for every consecutive batch of `batch` events of the stream (src, dst) in
    one launch; returns int32 block ids [E] (and the number of blocks per batch)."""
    src, dst = _need(src, torch.int64, "src"), _need(dst, torch.int64, "dst")
    E = src.numel()
    out = torch.empty(E, dtype=torch.int32, device=src.device)
    nb = (E + batch - 1) // batch
    cnt = torch.empty(nb, dtype=torch.int32, device=src.device) if want_counts else None
    check(_L().tgn_dep_blocks(_p(src), _p(dst), E, int(batch), _p(out), _p(cnt), _stream()))
    return (out, cnt) if want_counts else out


# ---------------------------------------------------------------------------
# message store
# ---------------------------------------------------------------------------
class MsgStore:
    """Device-resident replacement of TGNMemory's msg_s_store / msg_d_store dicts
    (modules/memory_module.py:140-145,180-191)."""

    def __init__(self, num_nodes: int, raw_dim: int, device, capacity: int = 1 << 16,
                 t_dtype=torch.int64):
        self.num_nodes, self.raw_dim, self.device = int(num_nodes), int(raw_dim), device
        self.t_dtype = t_dtype
        self.size = 0
        i32 = dict(dtype=torch.int32, device=device)
        self.s_start = torch.zeros(num_nodes, **i32)
        self.s_cnt = torch.zeros(num_nodes, **i32)
        self.s_last = torch.full((num_nodes,), -1, **i32)
        self.d_start = torch.zeros(num_nodes, **i32)
        self.d_cnt = torch.zeros(num_nodes, **i32)
        self.d_last = torch.full((num_nodes,), -1, **i32)
        self._alloc_log(int(capacity))

    def _alloc_log(self, capacity: int):
        dev = self.device
        old = getattr(self, "ev_src", None)
        n = self.size
        ev_src = torch.empty(capacity, dtype=torch.int64, device=dev)
        ev_dst = torch.empty(capacity, dtype=torch.int64, device=dev)
        ev_t = torch.empty(capacity, dtype=self.t_dtype, device=dev)
        ev_msg = torch.empty((capacity, max(self.raw_dim, 1)), dtype=torch.float32, device=dev)
        s_perm = torch.empty(capacity, dtype=torch.int32, device=dev)
        d_perm = torch.empty(capacity, dtype=torch.int32, device=dev)
        if old is not None and n:
            ev_src[:n], ev_dst[:n], ev_t[:n] = self.ev_src[:n], self.ev_dst[:n], self.ev_t[:n]
            ev_msg[:n], s_perm[:n], d_perm[:n] = self.ev_msg[:n], self.s_perm[:n], self.d_perm[:n]
        self.ev_src, self.ev_dst, self.ev_t, self.ev_msg = ev_src, ev_dst, ev_t, ev_msg
        self.s_perm, self.d_perm = s_perm, d_perm
        self.capacity = capacity
        self._struct = None
        # bumped whenever the log arrays move: whoever baked their addresses into a captured CUDA graph
        # (tgn_b200.engine) compares it before a replay
        self.generation = getattr(self, "generation", 0) + 1

    def struct(self) -> _cabi.MsgStoreStruct:
        if self._struct is None:
            s = _cabi.MsgStoreStruct()
            s.num_nodes, s.capacity, s.raw_dim = self.num_nodes, self.capacity, self.raw_dim
            s.t_is_float = 1 if self.t_dtype == torch.float32 else 0
            for f in ("ev_src", "ev_dst", "ev_t", "ev_msg", "s_perm", "d_perm", "s_start", "s_cnt",
                      "s_last", "d_start", "d_cnt", "d_last"):
                setattr(s, f, getattr(self, f).data_ptr())
            self._struct = s
        return self._struct

    def _set_t_dtype(self, t: Tensor):
        if t.dtype != self.t_dtype:
            if self.size:
                raise _cabi.TgnError("message store: timestamp dtype changed mid-epoch")
            _t_flag(t)
            self.t_dtype = t.dtype
            self.ev_t = torch.empty(self.capacity, dtype=t.dtype, device=self.device)
            self._struct = None
            self.generation += 1

    def reset(self):
        self.size = 0
        check(_L().tgn_msgstore_reset(ctypes.byref(self.struct()), _stream()))

    def update(self, src: Tensor, dst: Tensor, t: Tensor, raw_msg: Tensor,
               base_dev: Optional[Tensor] = None):
        B = src.numel()
        if B == 0:
            return
        self._set_t_dtype(t)
        if self.size + B > self.capacity:
            self._alloc_log(max(2 * self.capacity, self.size + B))
        src = _need(src, torch.int64, "src")
        dst = _need(dst, torch.int64, "dst")
        raw = _need(raw_msg, torch.float32, "raw_msg") if self.raw_dim else None
        t = t.contiguous()
        if B > SORT_MAX:
            raise _cabi.TgnError(f"update_state batch {B} exceeds TGN_SORT_MAX={SORT_MAX}")
        check(_L().tgn_msgstore_update(ctypes.byref(self.struct()), _p(src), _p(dst), _p(t), _p(raw),
                                       B, self.size, _p(base_dev), _stream()))
        self.size += B

    def gather(self, n_id: Tensor, direction: int):
        """_compute_msg's tuple gather (memory_module.py:196-201) for one store.
        Returns (src, dst, t, raw_msg) in store order."""
        n_id = _need(n_id, torch.int64, "n_id")
        S = n_id.numel()
        dev = n_id.device
        off = torch.empty(S + 1, dtype=torch.int32, device=dev)
        ws = torch.empty(max(_L().tgn_msgstore_count_ws_bytes(S), 16) // 8, dtype=torch.int64, device=dev)
        st = ctypes.byref(self.struct())
        check(_L().tgn_msgstore_count(st, _p(n_id), S, direction, _p(off), _p(ws), _stream()))
        M = int(off[S].item())
        o_s = torch.empty(M, dtype=torch.int64, device=dev)
        o_d = torch.empty(M, dtype=torch.int64, device=dev)
        o_t = torch.empty(M, dtype=self.t_dtype, device=dev)
        o_r = torch.empty((M, self.raw_dim), dtype=torch.float32, device=dev)
        if M:
            check(_L().tgn_msgstore_gather(st, _p(n_id), S, direction, _p(off), _p(o_s), _p(o_d),
                                           _p(o_t), _p(o_r) if self.raw_dim else None, _stream()))
        return o_s, o_d, o_t, o_r

    def build(self, n_id: Tensor, agg_mode: int, memory: Tensor, last_update: Tensor,
              time_w: Tensor, time_b: Tensor, num_dev: Optional[Tensor] = None):
        """Fused gather + IdentityMessage + TimeEncoder + aggregation.
        Returns x [S, 2*Dm+De+Dt], lu [S] (dtype of stored t), sel_ev [S], sel_dt [S]."""
        n_id = _need(n_id, torch.int64, "n_id")
        S = n_id.numel()
        dev = n_id.device
        Dm, Dt = memory.shape[1], time_w.numel()
        W = 2 * Dm + self.raw_dim + Dt
        x = torch.empty((S, W), dtype=torch.float32, device=dev)
        lu = torch.empty(S, dtype=self.t_dtype, device=dev)
        sel_ev = torch.empty(S, dtype=torch.int32, device=dev)
        sel_dt = torch.empty(S, dtype=torch.float32, device=dev)
        check(_L().tgn_msg_build(ctypes.byref(self.struct()), _p(n_id), S, _p(num_dev), agg_mode,
                                 _p(memory), _p(last_update), Dm, _p(time_w), _p(time_b), Dt, _p(x),
                                 _p(lu), _p(sel_ev), _p(sel_dt), _stream()))
        return x, lu, sel_ev, sel_dt


# ---------------------------------------------------------------------------
# dense
# ---------------------------------------------------------------------------
# GEMM engine of the dense path: 0 = fp32 FMA on CUDA cores (tgn_sgemm, the 1e-5 parity path),
# 1 = tcgen05 kind::tf32 (fastest, ~1e-3), 3 = tcgen05 3xTF32 (tensor cores, fp32-level accuracy).
GEMM_PRECISION = 3


def set_gemm_precision(mode: int) -> int:
    global GEMM_PRECISION
    if mode not in (0, 1, 3):
        raise ValueError("gemm precision must be 0 (fp32 CUDA cores), 1 (tf32) or 3 (3xtf32)")
    old, GEMM_PRECISION = GEMM_PRECISION, mode
    return old


def sgemm(a: Tensor, b: Tensor, bias: Optional[Tensor] = None, *, m: int, n: int, k: int,
          lda: int, ldb: int, trans_a: bool = False, trans_b: bool = False,
          out: Optional[Tensor] = None, ldc: Optional[int] = None, accumulate: bool = False,
          split_k: int = 1, a_rows: Optional[Tensor] = None, m_dev: Optional[Tensor] = None,
          k_dev: Optional[Tensor] = None, a_off: int = 0, b_off: int = 0,
          prec: Optional[int] = None) -> Tensor:
    """C[m,n] (=|+=) op(A) op(B) (+bias); engine chosen by `prec` / GEMM_PRECISION.
    a_off / b_off are element offsets into a / b (for column slices)."""
    prec = GEMM_PRECISION if prec is None else prec
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
        if split_k > 1:
            out.zero_()
    ldc = ldc if ldc is not None else out.stride(0)
    if prec == 0:
        check(_L().tgn_sgemm(a.data_ptr() + 4 * a_off, _p(a_rows), b.data_ptr() + 4 * b_off, _p(bias),
                             _p(out), m, _p(m_dev), n, k, _p(k_dev), lda, ldb, ldc, int(trans_a),
                             int(trans_b), int(accumulate), split_k, _stream()))
    else:
        check(_L().tgn_tc_gemm(a.data_ptr() + 4 * a_off, _p(a_rows), b.data_ptr() + 4 * b_off, _p(bias),
                               _p(out), m, _p(m_dev), n, k, _p(k_dev), lda, ldb, ldc, int(trans_a),
                               int(trans_b), int(accumulate), split_k, prec, _stream()))
    return out


def gemm_desc(a, b, c, *, m: int, n: int, k: int, lda: int, ldb: int, ldc: int, bias=None,
              trans_a: bool = False, trans_b: bool = False, mode: int = 0, split_k: int = 1,
              m_dev=None, k_dev=None, a_off: int = 0, b_off: int = 0, c_off: int = 0) -> _cabi.GemmDesc:
    """One problem of a tgn_gemm_batch launch: C[m,n] (=, +=, atomic +=) op(A) op(B) (+bias).
    mode 0 store / 1 accumulate / 2 atomic (needed for split_k > 1).  *_off are element offsets."""
    d = _cabi.GemmDesc()
    d.a, d.b, d.c = a.data_ptr() + 4 * a_off, b.data_ptr() + 4 * b_off, c.data_ptr() + 4 * c_off
    d.bias = _p(bias)
    d.m_dev, d.k_dev = _p(m_dev), _p(k_dev)
    d.m, d.n, d.k, d.lda, d.ldb, d.ldc = m, n, k, lda, ldb, ldc
    d.trans_a, d.trans_b, d.mode, d.split_k = int(trans_a), int(trans_b), mode, split_k
    return d


def gemm_batch(descs: Sequence[_cabi.GemmDesc], prec: Optional[int] = None):
    """Up to 4 GEMMs in one TMA + tcgen05 launch (tgn_gemm_batch)."""
    prec = GEMM_PRECISION if prec is None else prec
    if prec not in (1, 3):
        raise _cabi.TgnError("gemm_batch runs on the tensor cores: precision must be 1 (tf32) or 3 (3xtf32)")
    arr = (_cabi.GemmDesc * len(descs))(*descs)
    check(_L().tgn_gemm_batch(ctypes.byref(arr), len(descs), prec, _stream()))


def gather_rows(table: Tensor, rows: Tensor, num_dev: Optional[Tensor] = None) -> Tensor:
    rows = _need(rows, torch.int64, "rows")
    out = torch.empty((rows.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    check(_L().tgn_gather_rows(_p(table), _p(rows), rows.numel(), _p(num_dev), table.shape[1], _p(out),
                               _stream()))
    return out


def colsum(x: Tensor, rows: int, cols: int, ld: int, out: Tensor, accumulate: bool,
           rows_dev: Optional[Tensor] = None, x_off: int = 0):
    check(_L().tgn_colsum(x.data_ptr() + 4 * x_off, rows, _p(rows_dev), cols, ld, _p(out),
                          int(accumulate), _stream()))


def time_encode(t: Tensor, w: Tensor, b: Tensor) -> Tensor:
    t = _need(t, torch.float32, "t")
    out = torch.empty((t.numel(), w.numel()), dtype=torch.float32, device=t.device)
    check(_L().tgn_time_encode(_p(t), t.numel(), _p(w), _p(b), w.numel(), _p(out), _stream()))
    return out


def time_encode_bwd(t: Tensor, grad: Tensor, ld_grad: int, w: Tensor, b: Tensor, d_w: Tensor,
                    d_b: Tensor, row_mask: Optional[Tensor] = None, num_dev: Optional[Tensor] = None,
                    grad_off: int = 0):
    check(_L().tgn_time_encode_bwd(_p(t), _p(row_mask), t.numel(), _p(num_dev), _p(w), _p(b),
                                   w.numel(), grad.data_ptr() + 4 * grad_off, ld_grad, _p(d_w),
                                   _p(d_b), _stream()))


def memory_scatter(n_id: Tensor, new_mem: Tensor, new_lu: Optional[Tensor], memory: Tensor,
                   last_update: Tensor, src_rows: Optional[Tensor] = None,
                   num_dev: Optional[Tensor] = None):
    n_id = _need(n_id, torch.int64, "n_id")
    lu_flag = 0 if new_lu is None else _t_flag(new_lu)
    check(_L().tgn_memory_scatter(_p(n_id), n_id.numel(), _p(num_dev), _p(new_mem), _p(new_lu),
                                  lu_flag, _p(src_rows), memory.shape[1], _p(memory),
                                  _p(last_update), _stream()))


class _TimeEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, w, b):
        ctx.save_for_backward(t, w, b)
        return time_encode(t, w, b)

    @staticmethod
    def backward(ctx, g):
        t, w, b = ctx.saved_tensors
        g = g.contiguous()
        d_w = torch.zeros_like(w)
        d_b = torch.zeros_like(b)
        time_encode_bwd(t, g, g.shape[1], w, b, d_w, d_b)
        return None, d_w, d_b


def time_encode_autograd(t: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """w: flat [D] view of lin.weight, b: lin.bias"""
    return _TimeEncodeFn.apply(t, w, b)


class _GRUFn(torch.autograd.Function):
    """torch.nn.GRUCell(x, h) with the gate GEMMs and gate math on the C-ABI kernels.
    Optional extras let the fused memory path return the TimeEncoder gradient
    through the time columns of x."""

    @staticmethod
    def forward(ctx, x, h, w_ih, w_hh, b_ih, b_hh, num_dev):
        S, Dx = x.shape
        D = h.shape[1]
        gi = sgemm(x, w_ih, b_ih, m=S, n=3 * D, k=Dx, lda=Dx, ldb=Dx, m_dev=num_dev)
        gh = sgemm(h, w_hh, b_hh, m=S, n=3 * D, k=D, lda=D, ldb=D, m_dev=num_dev)
        out = torch.empty((S, D), dtype=torch.float32, device=x.device)
        gates = torch.empty((S, 4 * D), dtype=torch.float32, device=x.device)
        check(_L().tgn_gru_gates_fwd(_p(gi), _p(gh), _p(h), None, S, _p(num_dev), D, _p(out),
                                     _p(gates), _stream()))
        ctx.save_for_backward(x, h, w_ih, w_hh, gates)
        ctx.num_dev = num_dev
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, h, w_ih, w_hh, gates = ctx.saved_tensors
        num_dev = ctx.num_dev
        S, Dx = x.shape
        D = h.shape[1]
        dev = x.device
        d_out = d_out.contiguous()
        d_gi = torch.empty((S, 3 * D), dtype=torch.float32, device=dev)
        d_gh = torch.empty((S, 3 * D), dtype=torch.float32, device=dev)
        need_h = ctx.needs_input_grad[1]
        d_h = torch.empty((S, D), dtype=torch.float32, device=dev) if need_h else None
        check(_L().tgn_gru_gates_bwd(_p(d_out), _p(gates), _p(h), None, S, _p(num_dev), D, _p(d_gi),
                                     _p(d_gh), _p(d_h), _stream()))
        split = max(1, min(16, S // 256))
        d_wih = sgemm(d_gi, x, m=3 * D, n=Dx, k=S, lda=3 * D, ldb=Dx, trans_a=True, trans_b=True,
                      split_k=split, k_dev=num_dev)
        d_whh = sgemm(d_gh, h, m=3 * D, n=D, k=S, lda=3 * D, ldb=D, trans_a=True, trans_b=True,
                      split_k=split, k_dev=num_dev)
        d_bih = torch.empty(3 * D, dtype=torch.float32, device=dev)
        d_bhh = torch.empty(3 * D, dtype=torch.float32, device=dev)
        colsum(d_gi, S, 3 * D, 3 * D, d_bih, False, rows_dev=num_dev)
        colsum(d_gh, S, 3 * D, 3 * D, d_bhh, False, rows_dev=num_dev)
        d_x = None
        if ctx.needs_input_grad[0]:
            d_x = sgemm(d_gi, w_ih, m=S, n=Dx, k=3 * D, lda=3 * D, ldb=Dx, trans_b=True)
        if need_h:
            sgemm(d_gh, w_hh, m=S, n=D, k=3 * D, lda=3 * D, ldb=D, trans_b=True, out=d_h,
                  accumulate=True)
        return d_x, d_h, d_wih, d_whh, d_bih, d_bhh, None


def gru_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor,
             num_dev: Optional[Tensor] = None) -> Tensor:
    return _GRUFn.apply(x.contiguous(), h.contiguous(), w_ih, w_hh, b_ih, b_hh, num_dev)


def gru_fused_fwd(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, *, dx: int, ldx: int,
                  ldw: int, num: int, num_dev: Optional[Tensor] = None, prec: int = 3, out: Optional[Tensor] = None,
                  gates: Optional[Tensor] = None) -> Tensor:
    """Fused GRUCell forward (tgn_gru_fused_fwd): gate GEMMs on tcgen05 + gate math in the epilogue.
    x [num, ldx] with dx live columns, w_ih [3D, ldw]; all operands 16-byte aligned."""
    D = h.shape[-1]
    if out is None:
        out = torch.empty((num, D), dtype=torch.float32, device=h.device)
    check(_L().tgn_gru_fused_fwd(_p(x), ldx, dx, _p(h), D, _p(w_ih), ldw, _p(w_hh), _p(b_ih), _p(b_hh), num,
                                 _p(num_dev), prec, _p(out), _p(gates), _stream()))
    return out


def rnn_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """torch.nn.RNNCell (tanh) forward -- inference only (JODIE/DyRep updater)."""
    S, Dx = x.shape
    D = h.shape[1]
    gi = sgemm(x, w_ih, b_ih, m=S, n=D, k=Dx, lda=Dx, ldb=Dx)
    gh = sgemm(h, w_hh, b_hh, m=S, n=D, k=D, lda=D, ldb=D)
    out = torch.empty((S, D), dtype=torch.float32, device=x.device)
    check(_L().tgn_rnn_gates_fwd(_p(gi), _p(gh), S, D, _p(out), _stream()))
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on tgn_sgemm (forward and both gradients)."""

    @staticmethod
    def forward(ctx, x, w, b):
        M, K = x.shape
        N = w.shape[0]
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return sgemm(x, w, b, m=M, n=N, k=K, lda=K, ldb=K)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        M, K = x.shape
        N = w.shape[0]
        d_x = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            d_x = sgemm(g, w, m=M, n=K, k=N, lda=N, ldb=K, trans_b=True)
        if ctx.needs_input_grad[1]:
            d_w = sgemm(g, x, m=N, n=K, k=M, lda=N, ldb=K, trans_a=True, trans_b=True,
                        split_k=max(1, min(16, M // 256)))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = torch.empty(N, dtype=torch.float32, device=x.device)
            colsum(g, M, N, N, d_b, False)
        return d_x, d_w, d_b


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    return _LinearFn.apply(x.contiguous(), w, b)


class _MemoryUpdateFn(torch.autograd.Function):
    """TGNMemory._get_updated_memory (modules/memory_module.py:152-178) for
    IdentityMessage + LastAggregator|MeanAggregator + GRUCell, fused:
    store gather -> message concat -> time encoding -> aggregation -> GRU.
    Gradients reach the GRU weights and (last mode) the TimeEncoder."""

    @staticmethod
    def forward(ctx, time_w, time_b, w_ih, w_hh, b_ih, b_hh, store, n_id, memory, last_update,
                agg_mode, num_dev):
        S = n_id.numel()
        D = memory.shape[1]
        dev = memory.device
        x, lu, sel_ev, sel_dt = store.build(n_id, agg_mode, memory, last_update, time_w, time_b,
                                            num_dev=num_dev)
        h = gather_rows(memory, n_id, num_dev=num_dev)
        Dx = x.shape[1]
        gi = sgemm(x, w_ih, b_ih, m=S, n=3 * D, k=Dx, lda=Dx, ldb=Dx, m_dev=num_dev)
        gh = sgemm(h, w_hh, b_hh, m=S, n=3 * D, k=D, lda=D, ldb=D, m_dev=num_dev)
        out = torch.empty((S, D), dtype=torch.float32, device=dev)
        gates = torch.empty((S, 4 * D), dtype=torch.float32, device=dev)
        check(_L().tgn_gru_gates_fwd(_p(gi), _p(gh), _p(h), None, S, _p(num_dev), D, _p(out),
                                     _p(gates), _stream()))
        ctx.save_for_backward(x, h, w_ih, w_hh, gates, sel_ev, sel_dt, time_w, time_b)
        ctx.meta = (agg_mode, num_dev, store.raw_dim)
        ctx.mark_non_differentiable(lu)
        return out, lu

    @staticmethod
    def backward(ctx, d_out, _d_lu):
        x, h, w_ih, w_hh, gates, sel_ev, sel_dt, time_w, time_b = ctx.saved_tensors
        agg_mode, num_dev, De = ctx.meta
        S, Dx = x.shape
        D = h.shape[1]
        Dt = time_w.numel()
        dev = x.device
        d_out = d_out.contiguous()
        d_gi = torch.empty((S, 3 * D), dtype=torch.float32, device=dev)
        d_gh = torch.empty((S, 3 * D), dtype=torch.float32, device=dev)
        check(_L().tgn_gru_gates_bwd(_p(d_out), _p(gates), _p(h), None, S, _p(num_dev), D, _p(d_gi),
                                     _p(d_gh), None, _stream()))
        split = max(1, min(16, S // 256))
        d_wih = sgemm(d_gi, x, m=3 * D, n=Dx, k=S, lda=3 * D, ldb=Dx, trans_a=True, trans_b=True,
                      split_k=split, k_dev=num_dev)
        d_whh = sgemm(d_gh, h, m=3 * D, n=D, k=S, lda=3 * D, ldb=D, trans_a=True, trans_b=True,
                      split_k=split, k_dev=num_dev)
        d_bih = torch.empty(3 * D, dtype=torch.float32, device=dev)
        d_bhh = torch.empty(3 * D, dtype=torch.float32, device=dev)
        colsum(d_gi, S, 3 * D, 3 * D, d_bih, False, rows_dev=num_dev)
        colsum(d_gh, S, 3 * D, 3 * D, d_bhh, False, rows_dev=num_dev)
        d_tw = d_tb = None
        if Dt and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            if agg_mode != AGG_LAST:
                raise _cabi.TgnError("fused memory backward supports LastAggregator only; "
                                     "MeanAggregator trains through the unfused module path")
            off = 2 * D + De
            # d x[:, time columns] = d_gi @ W_ih[:, off:off+Dt]
            d_xt = sgemm(d_gi, w_ih, m=S, n=Dt, k=3 * D, lda=3 * D, ldb=Dx, trans_b=True, b_off=off,
                         m_dev=num_dev)
            d_tw = torch.zeros_like(time_w)
            d_tb = torch.zeros_like(time_b)
            time_encode_bwd(sel_dt, d_xt, Dt, time_w, time_b, d_tw, d_tb, row_mask=sel_ev,
                            num_dev=num_dev)
        return d_tw, d_tb, d_wih, d_whh, d_bih, d_bhh, None, None, None, None, None, None


def memory_update(time_w, time_b, w_ih, w_hh, b_ih, b_hh, store: "MsgStore", n_id: Tensor,
                  memory: Tensor, last_update: Tensor, agg_mode: int = AGG_LAST,
                  num_dev: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    return _MemoryUpdateFn.apply(time_w, time_b, w_ih, w_hh, b_ih, b_hh, store, n_id, memory,
                                 last_update, agg_mode, num_dev)


# ---------------------------------------------------------------------------
# temporal attention
# ---------------------------------------------------------------------------
class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_node, b_node, w_edge, time_w, time_b, lu, nbr, t_edge, msg, msg_rows,
                row_ptr, edge_perm, centre_ids, heads, dropout_p, seed, counts):
        # counts = (rows_dev, edges_dev, centres_dev, seed_dev) or None
        rows_dev, edges_dev, centres_dev, seed_dev = counts if counts is not None else (None,) * 4
        Nb, Din = x.shape
        HC = w_edge.shape[0]
        C = HC // heads
        De = msg.shape[1] if msg is not None and msg.dim() == 2 else 0
        Dt = time_w.numel()
        E = nbr.numel()
        nC = row_ptr.numel() - 1
        dev = x.device
        proj = sgemm(x, w_node, b_node, m=Nb, n=4 * HC, k=Din, lda=Din, ldb=Din, m_dev=rows_dev)
        out = torch.empty((Nb, HC), dtype=torch.float32, device=dev)
        check(_L().tgn_attn_fill_skip(_p(proj), Nb, _p(rows_dev), HC, _p(out), _stream()))
        alpha = torch.empty((max(E, 1), heads), dtype=torch.float32, device=dev)
        ee = torch.empty((max(E, 1), HC), dtype=torch.float32, device=dev)
        lu_f, t_f = _t_flag(lu), _t_flag(t_edge)
        if nC and E:  # no edges: every row is its skip projection (already written)
            check(_L().tgn_attn_fwd(_p(proj), _p(lu), lu_f, _p(nbr), _p(t_edge), t_f, _p(msg),
                                    _p(msg_rows), _p(row_ptr), _p(edge_perm), _p(centre_ids), nC,
                                    _p(centres_dev), heads, C, De, Dt, _p(w_edge), _p(time_w),
                                    _p(time_b), float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _p(seed_dev), _p(out), _p(alpha),
                                    _p(ee), _stream()))
        ctx.save_for_backward(x, w_node, w_edge, time_w, time_b, lu, nbr, t_edge, msg, msg_rows,
                              row_ptr, edge_perm, centre_ids, proj, alpha, ee)
        ctx.meta = (heads, C, De, Dt, float(dropout_p), int(seed), rows_dev, edges_dev, centres_dev,
                    seed_dev)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (x, w_node, w_edge, time_w, time_b, lu, nbr, t_edge, msg, msg_rows, row_ptr, edge_perm,
         centre_ids, proj, alpha, ee) = ctx.saved_tensors
        heads, C, De, Dt, p, seed, rows_dev, edges_dev, centres_dev, seed_dev = ctx.meta
        Nb, Din = x.shape
        HC = heads * C
        E = nbr.numel()
        nC = row_ptr.numel() - 1
        dev = x.device
        d_out = d_out.contiguous()
        d_proj = torch.empty((Nb, 4 * HC), dtype=torch.float32, device=dev)
        check(_L().tgn_attn_bwd_init(_p(d_out), Nb, _p(rows_dev), HC, _p(d_proj), _stream()))
        d_ee = torch.empty((max(E, 1), HC), dtype=torch.float32, device=dev)
        De_in = Dt + De
        d_we = torch.zeros((HC, De_in), dtype=torch.float32, device=dev)
        d_tw = torch.zeros_like(time_w)
        d_tb = torch.zeros_like(time_b)
        if nC and E:
            check(_L().tgn_attn_bwd(_p(proj), _p(nbr), _p(row_ptr), _p(edge_perm), _p(centre_ids), nC,
                                    _p(centres_dev), heads, C, _p(alpha), _p(ee), _p(d_out), p, seed & 0xFFFFFFFFFFFFFFFF, _p(seed_dev),
                                    _p(d_proj), _p(d_ee), _stream()))
            ea = torch.empty((E, De_in), dtype=torch.float32, device=dev)
            rel = torch.empty(E, dtype=torch.float32, device=dev)
            check(_L().tgn_attn_edge_attr(_p(lu), _t_flag(lu), _p(nbr), _p(t_edge), _t_flag(t_edge),
                                          _p(msg), _p(msg_rows), E, _p(edges_dev), De, Dt,
                                          _p(time_w), _p(time_b), _p(ea), _p(rel), _stream()))
            split = max(1, min(16, E // 256))
            sgemm(d_ee, ea, m=HC, n=De_in, k=E, lda=HC, ldb=De_in, trans_a=True, trans_b=True,
                  out=d_we, split_k=split, accumulate=True, k_dev=edges_dev)
            if Dt:
                # d edge_attr[:, :Dt] = d_ee @ W_edge[:, :Dt]
                d_eat = sgemm(d_ee, w_edge, m=E, n=Dt, k=HC, lda=HC, ldb=De_in, trans_b=True,
                              m_dev=edges_dev)
                time_encode_bwd(rel, d_eat, Dt, time_w, time_b, d_tw, d_tb, num_dev=edges_dev)
        split = max(1, min(16, Nb // 256))
        d_wn = sgemm(d_proj, x, m=4 * HC, n=Din, k=Nb, lda=4 * HC, ldb=Din, trans_a=True,
                     trans_b=True, split_k=split, k_dev=rows_dev)
        d_bn = torch.empty(4 * HC, dtype=torch.float32, device=dev)
        colsum(d_proj, Nb, 4 * HC, 4 * HC, d_bn, False, rows_dev=rows_dev)
        d_x = None
        if ctx.needs_input_grad[0]:
            d_x = sgemm(d_proj, w_node, m=Nb, n=Din, k=4 * HC, lda=4 * HC, ldb=Din, trans_b=True)
        return (d_x, d_wn, d_bn, d_we, d_tw, d_tb) + (None,) * 12


def temporal_attention(x: Tensor, w_node: Tensor, b_node: Tensor, w_edge: Tensor, time_w: Tensor,
                       time_b: Tensor, last_update: Tensor, nbr_local: Tensor, t_edge: Tensor,
                       msg: Tensor, row_ptr: Tensor, *, heads: int, msg_rows: Optional[Tensor] = None,
                       edge_perm: Optional[Tensor] = None, centre_ids: Optional[Tensor] = None,
                       dropout_p: float = 0.0, seed: int = 0, counts=None) -> Tensor:
    """w_node = cat[lin_query, lin_key, lin_value, lin_skip].weight  [4*H*C, in],
    b_node likewise [4*H*C]; w_edge = lin_edge.weight [H*C, time_dim+raw_dim]."""
    return _AttnFn.apply(x.contiguous(), w_node, b_node, w_edge, time_w, time_b,
                         last_update.contiguous(), nbr_local.contiguous(), t_edge.contiguous(),
                         msg.contiguous() if msg is not None else None, msg_rows, row_ptr, edge_perm,
                         centre_ids, heads, dropout_p, seed, counts)


class _EgatFn(torch.autograd.Function):
    """EdgeGATConv attention core (model_utils.py:594-599) on csrc/egat.cu: s[v,h] = sum_e a_e * (el[src]+ee)."""

    @staticmethod
    def forward(ctx, el, er, ee, src, row_ptr, perm, slope, p, seed):
        N, H = el.shape
        E = ee.shape[0]
        s = torch.empty((N, H), dtype=torch.float32, device=el.device)
        alpha = torch.empty((E, H), dtype=torch.float32, device=el.device)
        check(_L().tgn_egat_attn_fwd(_p(el), _p(er), _p(ee), _p(row_ptr), _p(perm), _p(src), N, E, H, float(slope),
                                     float(p), int(seed), _p(s), _p(alpha), _stream()))
        ctx.save_for_backward(el, er, ee, src, row_ptr, perm, alpha)
        ctx.cfg = (float(slope), float(p), int(seed))
        ctx.mark_non_differentiable(alpha)
        return s, alpha

    @staticmethod
    def backward(ctx, d_s, _d_alpha):
        el, er, ee, src, row_ptr, perm, alpha = ctx.saved_tensors
        slope, p, seed = ctx.cfg
        N, H = el.shape
        E = ee.shape[0]
        d_el = torch.zeros_like(el)
        d_er, d_ee = torch.empty_like(er), torch.empty_like(ee)
        check(_L().tgn_egat_attn_bwd(_p(el), _p(er), _p(ee), _p(row_ptr), _p(perm), _p(src), N, E, H, slope, p, seed,
                                     _p(alpha), _p(d_s.contiguous()), _p(d_el), _p(d_er), _p(d_ee), _stream()))
        return d_el, d_er, d_ee, None, None, None, None, None, None


def egat_attention(el: Tensor, er: Tensor, ee: Tensor, src: Tensor, dst: Tensor, negative_slope: float = 0.2,
                   dropout_p: float = 0.0, seed: int = 0):
    """(s [N,H], alpha [E,H]) of the reference's EdgeGATConv for logits el/er [N,H], ee [E,H] and the edge
    list (src, dst) [E]; alpha is the softmax weight before dropout (what get_attention returns in eval)."""
    N = el.shape[0]
    src, dst = _need(src, torch.int64, "src"), _need(dst, torch.int64, "dst")
    row_ptr, perm = group_edges_by_centre(dst, N)
    return _EgatFn.apply(el.contiguous().float(), er.contiguous().float(), ee.contiguous().float(), src, row_ptr,
                         perm.contiguous(), negative_slope, dropout_p, seed)


def group_edges_by_centre(centre_local: Tensor, num_rows: int):
    """CSR over centres for an arbitrary edge list: returns (row_ptr[num_rows+1], edge_perm[E]).
    A stable counting sort -- plumbing for callers that hand GraphAttentionEmbedding an
    edge_index that is not already grouped by centre."""
    order = torch.sort(centre_local, stable=True).indices.to(torch.int32)
    counts = torch.bincount(centre_local, minlength=num_rows)
    row_ptr = torch.zeros(num_rows + 1, dtype=torch.int32, device=centre_local.device)
    row_ptr[1:] = counts.cumsum(0).to(torch.int32)
    return row_ptr, order


# ---------------------------------------------------------------------------
# decoder / metric / optimiser
# ---------------------------------------------------------------------------
def link_score(hs: Tensor, hd: Tensor, a_rows: Tensor, b_rows: Tensor, w_final: Tensor,
               b_final: Tensor, apply_sigmoid: bool) -> Tensor:
    out = torch.empty(a_rows.numel(), dtype=torch.float32, device=hs.device)
    check(_L().tgn_link_score(_p(hs), _p(hd), _p(a_rows), _p(b_rows), a_rows.numel(), hs.shape[1],
                              _p(w_final), _p(b_final), int(apply_sigmoid), _p(out), _stream()))
    return out


def mrr(pos: Tensor, neg: Tensor) -> Tensor:
    """per-positive reciprocal rank, TGB convention (epoch_utils.py:108-113)."""
    pos = _need(pos.reshape(-1), torch.float32, "pos")
    neg = _need(neg.reshape(pos.numel(), -1), torch.float32, "neg")
    out = torch.empty(pos.numel(), dtype=torch.float32, device=pos.device)
    check(_L().tgn_mrr(_p(pos), _p(neg), pos.numel(), neg.shape[1], _p(out), _stream()))
    return out


def adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step_dev: Tensor,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    check(_L().tgn_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), params.numel(), lr,
                             beta1, beta2, eps, _p(step_dev), _stream()))
