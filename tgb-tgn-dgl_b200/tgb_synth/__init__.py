"""Offline stand-in for the parts of the `tgb` package the reference imports (utils.py:9-10,
epoch_utils.py:43,113): a TGB-shaped link-prediction dataset generated locally (the image has no
network, so the real TGB datasets cannot be downloaded), its pre-generated evaluation negatives
and the TGB MRR evaluator.  `utils.getDataWithDependecyBlock` uses the real `tgb` package when it
is importable and this module otherwise.

Dataset names: the TGB names of `tgn_b200.synth.SHAPES` ("tgbl-wiki", "tgbl-review", ...), with an
optional size cap "name@events" (e.g. "tgbl-wiki@20000") for quick runs.
"""
from types import SimpleNamespace
from typing import List

import zlib

import numpy as np
import torch

from tgn_b200 import synth


class TemporalData:
    """The subset of torch_geometric.data.TemporalData the reference touches."""

    def __init__(self, src, dst, t, msg, num_nodes=None):
        self.src, self.dst, self.t, self.msg = src, dst, t, msg
        self._num_nodes = num_nodes

    @property
    def num_nodes(self):
        if self._num_nodes is not None:
            return self._num_nodes
        return int(max(int(self.src.max()), int(self.dst.max()))) + 1

    def __getitem__(self, mask):
        return TemporalData(self.src[mask], self.dst[mask], self.t[mask], self.msg[mask], self._num_nodes)

    def __len__(self):
        return self.src.numel()


class NegativeSampler:
    """Pre-generated evaluation negatives, `query_batch` -> list of lists (tgb NegativeEdgeSampler)."""

    def __init__(self, num_nodes: int, num_neg: int, dst_lo: int, seed: int):
        self.num_nodes, self.num_neg, self.dst_lo, self.seed = num_nodes, num_neg, dst_lo, seed

    def query_batch(self, pos_src, pos_dst, pos_t, split_mode: str = "val") -> List[List[int]]:
        src = torch.as_tensor(pos_src).cpu().numpy()
        dst = torch.as_tensor(pos_dst).cpu().numpy()
        t = torch.as_tensor(pos_t).cpu().numpy()
        # deterministic per (split, first event of the batch): the same batch always gets the same negatives
        # (zlib.crc32, not hash(): str hashes are randomised per process, and every rank / run must draw the same)
        key = (zlib.crc32(str(split_mode).encode()) & 0xFFFF) * 1_000_003 + int(src[0]) * 7919 + int(dst[0]) * 104_729 + int(t[0])
        neg = synth.eval_negatives(src, dst, self.num_nodes, self.num_neg, seed=(self.seed + key) & 0x7FFFFFFF,
                                   dst_lo=self.dst_lo)
        return neg.tolist()


class PyGLinkPropPredDataset:
    def __init__(self, name: str, root: str = "datasets", num_neg: int = 20, seed: int = 0):
        base, _, cap = name.partition("@")
        data = synth.synth_events(base, seed=seed, max_events=int(cap) if cap else None)
        E = data["src"].size
        self.name, self.eval_metric = name, "mrr"
        self._data = TemporalData(torch.from_numpy(data["src"]), torch.from_numpy(data["dst"]),
                                  torch.from_numpy(data["t"]), torch.from_numpy(data["msg"]), data["num_nodes"])
        idx = torch.arange(E)
        self.train_mask = idx < data["n_train"]
        self.val_mask = (idx >= data["n_train"]) & (idx < data["n_train"] + data["n_val"])
        self.test_mask = idx >= data["n_train"] + data["n_val"]
        bip = synth.SHAPES[base]["bip"]
        self.negative_sampler = NegativeSampler(data["num_nodes"], num_neg, bip[0] if bip else 0, seed + 17)

    def get_TemporalData(self):
        return self._data

    def load_val_ns(self):
        pass

    def load_test_ns(self):
        pass


class Evaluator:
    """TGB link-prediction evaluator: MRR with the optimistic/pessimistic tie average."""

    def __init__(self, name: str):
        self.name = name

    def eval(self, input_dict):
        pos = np.asarray(input_dict["y_pred_pos"], dtype=np.float32).reshape(-1, 1)
        neg = np.asarray(input_dict["y_pred_neg"], dtype=np.float32).reshape(pos.shape[0], -1)
        rank = 0.5 * ((neg > pos).sum(1) + (neg >= pos).sum(1)) + 1.0
        return {m: float((1.0 / rank).mean()) for m in input_dict.get("eval_metric", ["mrr"])}


linkproppred = SimpleNamespace(evaluate=SimpleNamespace(Evaluator=Evaluator),
                               dataset_pyg=SimpleNamespace(PyGLinkPropPredDataset=PyGLinkPropPredDataset))
