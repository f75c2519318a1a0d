"""Negatives and the evaluation metric on the device.

The reference's evaluation loop builds, per batch, a Python list of lists from
`neg_sampler.query_batch`, truncates it on the host to the shortest list and uploads it
(epoch_utils.py:43-56), and its training loop draws one negative per positive by indexing a Python
list (neg_sampler.py:8-23).  Here

* `DeviceNegativeTable` holds a split's pre-generated negatives as ONE ragged-packed device tensor
  (built once from the sampler's own `query_batch`, so the negatives are exactly the reference's);
  `batch(i0, i1)` is a device view truncated to the batch's shortest list -- no per-batch upload;
* `SyntheticNegatives` generates `[B, Q]` uniform negatives on the device where real TGB tables cannot
  be downloaded (`tgn_neg_fill`);
* `DeviceNegSamplerDest` is `NegLinkSamplerDest` on the device (`tgn_neg_dest_sample`);
* `evaluate_table` runs a whole evaluation split on a `TGNEngine` with the epoch MRR accumulated in
  device memory (`tgn_rank_accum`): one host read per epoch.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _cabi
from ._cabi import check

_L = _cabi.lib
_p = lambda t: None if t is None else t.data_ptr()
_stream = lambda: torch.cuda.current_stream().cuda_stream


class DeviceNegativeTable:
    """neg[E, Qmax] int64 on the device (+ per-row lengths on device and host)."""

    def __init__(self, neg: Tensor, lens: Tensor):
        if not neg.is_cuda:
            raise _cabi.TgnError("DeviceNegativeTable lives on a CUDA device; there is no CPU fallback")
        self.neg, self.lens = neg, lens.to(neg.device, torch.int32)
        self._lens_host = lens.cpu().numpy().astype(np.int64)

    @classmethod
    def from_lists(cls, rows: Sequence[Sequence[int]], device="cuda") -> "DeviceNegativeTable":
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
        qmax = int(lens.max()) if len(rows) else 0
        host = np.zeros((len(rows), max(qmax, 1)), dtype=np.int64)
        for i, r in enumerate(rows):
            host[i, :len(r)] = r
        return cls(torch.from_numpy(host).to(device), torch.from_numpy(lens))

    @classmethod
    def from_sampler(cls, neg_sampler, src: Tensor, dst: Tensor, t: Tensor, split_mode: str, batch_size: int,
                     device="cuda") -> "DeviceNegativeTable":
        """One host pass over the split with the reference's own call (epoch_utils.py:43), batch by batch
        (TGB keys its tables by the positive edge, so the batching does not change the negatives)."""
        rows = []
        for i0 in range(0, src.numel(), batch_size):
            sl = slice(i0, i0 + batch_size)
            rows.extend(neg_sampler.query_batch(src[sl], dst[sl], t[sl], split_mode=split_mode))
        return cls.from_lists(rows, device)

    def __len__(self) -> int:
        return self.neg.shape[0]

    def batch(self, i0: int, i1: int) -> Tensor:
        """[i1-i0, q] view, q = shortest list of the batch (epoch_utils.py:48-56)."""
        q = int(self._lens_host[i0:i1].min()) if i1 > i0 else 0
        return self.neg[i0:i1, :q]


class SyntheticNegatives:
    """[B, Q] negatives uniform over [lo, hi) without the positive, drawn on the device.  `call` keys the
    Philox stream: the same (seed, call) reproduces the batch (one call id per evaluation batch)."""

    def __init__(self, num_neg: int, lo: int, hi: int, seed: int = 2):
        self.Q, self.lo, self.hi, self.seed = int(num_neg), int(lo), int(hi), int(seed)

    def batch(self, pos_dst: Tensor, call: int, out: Optional[Tensor] = None) -> Tensor:
        pos_dst = pos_dst.contiguous()
        if not pos_dst.is_cuda or pos_dst.dtype != torch.int64:
            raise _cabi.TgnError("pos_dst must be a CUDA int64 tensor")
        B = pos_dst.numel()
        if out is None:
            out = torch.empty((B, self.Q), dtype=torch.int64, device=pos_dst.device)
        check(_L().tgn_neg_fill(_p(pos_dst), B, self.Q, self.lo, self.hi, self.seed & 0xFFFFFFFFFFFFFFFF, int(call),
                                _p(out), _stream()))
        return out


class DeviceNegSamplerDest:
    """neg_sampler.NegLinkSamplerDest (reference neg_sampler.py:3-23) with device-side draws."""

    def __init__(self, dst_nodes, device="cuda", seed: int = 0):
        self.dst_nodes = torch.as_tensor(dst_nodes).reshape(-1).to(device, torch.int64).contiguous()
        if not self.dst_nodes.is_cuda:
            raise _cabi.TgnError("DeviceNegSamplerDest needs a CUDA device; there is no CPU fallback")
        self.seed, self.calls = int(seed), 0

    def sample(self, pos_dst: Tensor) -> Tensor:
        pos = pos_dst.to(self.dst_nodes.device, torch.int64).contiguous()
        out = torch.empty_like(pos)
        check(_L().tgn_neg_dest_sample(_p(self.dst_nodes), self.dst_nodes.numel(), _p(pos), pos.numel(),
                                       self.seed & 0xFFFFFFFFFFFFFFFF, self.calls, _p(out), _stream()))
        self.calls += 1
        return out.to(pos_dst.dtype)


def rank_accum(gt: Tensor, ge: Tensor, acc: Tensor, rr_out: Optional[Tensor] = None) -> None:
    """acc (float64[2], device) += (per-batch mean reciprocal rank, 1)."""
    check(_L().tgn_rank_accum(_p(gt), _p(ge), gt.numel(), _p(acc), _p(rr_out), _stream()))


def evaluate_table(engine, src: Tensor, dst: Tensor, t: Tensor, msg: Tensor, negatives, batch_size: int,
                   rank: int = 0, world: int = 1, group=None) -> float:
    """test() (epoch_utils.py:28-165) for one split on a TGNEngine: `negatives` is a
    DeviceNegativeTable (rows aligned with the split) or a SyntheticNegatives generator.  The split's
    arrays are uploaded once; per batch there is no host<->device traffic and no synchronisation except
    the shape of a ragged table batch.  With world > 1 the negative columns are sharded as in dist_eval."""
    from . import dist_eval
    dev = engine.dev
    src, dst, t, msg = (x.to(dev) for x in (src, dst, t, msg))
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    for b, i0 in enumerate(range(0, src.numel(), batch_size)):
        i1 = min(i0 + batch_size, src.numel())
        neg = negatives.batch(i0, i1) if isinstance(negatives, DeviceNegativeTable) else \
            negatives.batch(dst[i0:i1], call=b)
        if world > 1:
            neg = dist_eval.shard_columns(neg, rank, world)
        _, _, gt, ge = engine.eval_batch(src[i0:i1], dst[i0:i1], neg, t[i0:i1], msg[i0:i1], want_neg_scores=False)
        gt, ge = dist_eval.reduce_counts(gt, ge, group)
        rank_accum(gt, ge, acc)
    a = acc.cpu()
    return float(a[0] / a[1]) if float(a[1]) else 0.0
