"""`python tgb_gen_graph.py --data tgbl-wiki` (reference README.md:4-5): writes the t-CSR file
`DATA/<name>/ext_full.npz` (keys indptr / indices / ts / eid, read at reference utils.py:73) that
TGL's `sampler_core.ParallelSampler` consumes.  The reference tree does not contain the script
(SURVEY.md 0.1); this one builds the graph on the GPU (`tgn_tcsr_build`: one stable radix sort of
the 2E directed entries by (node, timestamp) + one emit pass) instead of upstream's Python row loop.

The events come from the `tgb` package when it is importable (the offline stand-in `tgb_synth`
ships with this repo: synthetic graphs with the TGB shapes, no network access)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tgn_b200 import ops  # noqa: E402


def build_ext_full(src, dst, t, num_nodes: int, add_reverse: bool = True, device="cuda"):
    """Device t-CSR (indptr, indices, eid, ts) from an event list (any array-likes)."""
    dev = torch.device(device)
    to = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if not torch.is_tensor(a) else a).to(dev, dt)
    t_t = torch.as_tensor(np.ascontiguousarray(t) if not torch.is_tensor(t) else t)
    t_d = t_t.to(dev, torch.float32 if t_t.dtype.is_floating_point else torch.int64)
    return ops.tcsr_build(to(src, torch.int64), to(dst, torch.int64), t_d, int(num_nodes), add_reverse)


def save_ext_full(path: str, indptr, indices, eid, ts) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, indptr=indptr.cpu().numpy(), indices=indices.cpu().numpy(), ts=ts.cpu().numpy(),
             eid=eid.cpu().numpy())


def load_ext_full(path: str):
    g = np.load(path)
    return g["indptr"], g["indices"], g["eid"], g["ts"]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--data", required=True)
    ap.add_argument("--root", default="datasets")
    ap.add_argument("--out", default=None, help="default DATA/<data>/ext_full.npz")
    ap.add_argument("--no_reverse", action="store_true", help="do not add the dst->src entries")
    ap.add_argument("--max_events", type=int, default=None, help="offline stand-in only: cap the synthetic stream")
    a = ap.parse_args(argv)
    name = a.data
    try:
        from tgb.linkproppred.dataset_pyg import PyGLinkPropPredDataset
    except ImportError:                                   # offline image: synthetic TGB shapes
        from tgb_synth import PyGLinkPropPredDataset
        if a.max_events:
            name = f"{a.data}@{a.max_events}"
    data = PyGLinkPropPredDataset(name=name, root=a.root).get_TemporalData()
    g = build_ext_full(data.src, data.dst, data.t, int(data.num_nodes), not a.no_reverse)
    out = a.out or os.path.join("DATA", a.data, "ext_full.npz")
    save_ext_full(out, *g)
    print(f"{out}: {g[0].numel() - 1} nodes, {g[1].numel()} entries")


if __name__ == "__main__":
    main()
