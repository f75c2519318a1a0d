"""Drop-in `temporal_dataset.TemporalGraphDataset` (reference temporal_dataset.py:4-57): a
map-style dataset of events whose items are dicts {'src','dst','t' (float32),'msg','idx'} plus
'b' (dependency block id) when block ids were given.  `TensorBatchLoader` yields the same
collated dict batches as `torch.utils.data.DataLoader(dataset, batch_size, shuffle=False)` by
slicing the tensors -- no per-item collation (the reference spends most of an epoch's host time
there)."""
import torch
from torch.utils.data import Dataset


class TemporalGraphDataset(Dataset):
    def __init__(self, src, dst, t, msg, batch=None):
        self.src, self.dst, self.t, self.msg = src, dst, t, msg
        self.batch = batch

    def __len__(self):
        return len(self.src)

    def __getitem__(self, idx):
        item = {"src": self.src[idx], "dst": self.dst[idx], "t": self.t[idx].float(), "msg": self.msg[idx]}
        if self.batch is not None:
            item["b"] = self.batch[idx]
        item["idx"] = idx
        return item


class TensorBatchLoader:
    """Sequential fixed-size batches of a TemporalGraphDataset as collated dicts."""

    def __init__(self, dataset: TemporalGraphDataset, batch_size: int, drop_last: bool = False):
        self.dataset, self.batch_size, self.drop_last = dataset, int(batch_size), drop_last
        self._b = None if dataset.batch is None else torch.as_tensor(dataset.batch)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        d, n = self.dataset, len(self.dataset)
        for lo in range(0, n, self.batch_size):
            hi = min(n, lo + self.batch_size)
            if self.drop_last and hi - lo < self.batch_size:
                return
            out = {"src": d.src[lo:hi], "dst": d.dst[lo:hi], "t": d.t[lo:hi].float(), "msg": d.msg[lo:hi]}
            if self._b is not None:
                out["b"] = self._b[lo:hi]
            out["idx"] = torch.arange(lo, hi)
            yield out
