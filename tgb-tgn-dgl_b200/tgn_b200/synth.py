"""Synthetic temporal graphs with the shapes of the TGB link-prediction datasets
BASELINE.json names (no network access: real TGB data cannot be downloaded).

The generator follows SURVEY.md 8(d): bipartite where the real dataset is,
Zipf-like node popularity floor(u^3 * n), sorted timestamps (integers; coarse on
the large shapes so that equal timestamps occur), N(0,1) edge features, 70/15/15
chronological split (utils.py:38-40), one uniform destination negative per training
event drawn like neg_sampler.NegLinkSamplerDest (neg_sampler.py:8-23: uniform over the
observed destinations, resampled while it collides with the positive).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

SHAPES = {
    # name: nodes, events, raw_dim, (n_src, n_dst) or None for non-bipartite, time span, default batch, K
    "tgbl-wiki": dict(N=9_227, E=157_474, De=172, bip=(8_227, 1_000), tmax=2_678_373, B=200, K=10),
    "tgbl-review": dict(N=352_637, E=4_873_540, De=1, bip=(298_349, 54_288), tmax=600_000, B=200, K=10),
    "tgbl-coin": dict(N=638_486, E=22_809_486, De=1, bip=None, tmax=1_000_000, B=600, K=10),
    "tgbl-comment": dict(N=994_790, E=44_314_507, De=2, bip=None, tmax=2_000_000, B=200, K=20),
    "tgbl-flight": dict(N=18_143, E=67_169_570, De=16, bip=None, tmax=1_500, B=200, K=10),
}


def synth_events(name: str, seed: int = 0, max_events: Optional[int] = None, batch: Optional[int] = None,
                 extend: bool = False) -> Dict[str, np.ndarray]:
    """Returns numpy arrays: src, dst (int64), t (int64, sorted), msg (float32 [E, De]),
    neg (int64, one negative destination per event), plus num_nodes / raw_dim / split sizes.
    `batch` overrides the shape's default batch size (config/TGN.yml:27 says 2000 where BASELINE.json
    says 200).  extend=True lets `max_events` exceed the dataset's event count (same nodes, same
    event rate: the time span grows with it) -- throughput runs on the small wiki shape need more
    batches than its 157k events hold."""
    cfg = SHAPES[name]
    rng = np.random.default_rng(seed)
    E = cfg["E"] if max_events is None else (int(max_events) if extend else min(cfg["E"], int(max_events)))
    N = cfg["N"]
    if cfg["bip"] is not None:
        ns, nd = cfg["bip"]
        src = np.floor(rng.random(E) ** 3 * ns).astype(np.int64)
        dst = ns + np.floor(rng.random(E) ** 3 * nd).astype(np.int64)
    else:
        src = np.floor(rng.random(E) ** 3 * N).astype(np.int64)
        dst = np.floor(rng.random(E) ** 2 * N).astype(np.int64)
    # spread timestamps over the dataset's span proportionally to the number of events kept
    span = max(1, int(cfg["tmax"] * E / cfg["E"]))
    t = np.sort(rng.integers(0, span + 1, E)).astype(np.int64)
    msg = rng.standard_normal((E, cfg["De"])).astype(np.float32)
    neg = sample_negatives(dst, np.unique(dst), np.random.default_rng(seed + 2))
    n_train = int(E * 0.70)
    n_val = int(E * 0.15)
    return dict(src=src, dst=dst, t=t, msg=msg, neg=neg, num_nodes=N, raw_dim=cfg["De"],
                n_train=n_train, n_val=n_val, n_test=E - n_train - n_val, batch=cfg["B"] if batch is None else int(batch), K=cfg["K"])


def sample_negatives(pos_dst: np.ndarray, dst_nodes: np.ndarray, rng) -> np.ndarray:
    """neg_sampler.NegLinkSamplerDest.sample (neg_sampler.py:8-23), vectorised."""
    neg = dst_nodes[rng.integers(0, dst_nodes.size, pos_dst.size)]
    if dst_nodes.size > 1:
        bad = neg == pos_dst
        while bad.any():
            neg[bad] = dst_nodes[rng.integers(0, dst_nodes.size, int(bad.sum()))]
            bad = neg == pos_dst
    return neg.astype(np.int64)


def eval_negatives(src: np.ndarray, dst: np.ndarray, num_nodes: int, q: int, seed: int = 2,
                   dst_lo: int = 0) -> np.ndarray:
    """[B, q] negative destinations per positive, uniform over [dst_lo, num_nodes) without the
    positive (stand-in for TGB's pre-generated negative_sampler.query_batch, epoch_utils.py:43)."""
    rng = np.random.default_rng(seed)
    neg = rng.integers(dst_lo, num_nodes - 1, (src.size, q)).astype(np.int64)
    neg += (neg >= dst[:, None])
    return neg
