"""Drop-in `dependencyGraph` (reference dependencyGraph.py:8-49): dependency-aware block ids.

Inside one DataLoader batch every event gets a block id such that the events of one block touch
pairwise disjoint nodes: walking the batch in order, an event's block is one more than the
highest block already assigned to either of its endpoints (0 if both are new in this batch).
The reference does this with Python dicts and `.item()` per node (hours on the large TGB
shapes).  `dependecyAwareBatch` over one of this package's TensorBatchLoaders computes the ids of ALL
batches in one launch of the sm_100a kernel (csrc/depblock.cu: one CTA per batch);
`get_block` on host lists / CPU tensors -- the reference's own per-batch entry point -- is a single
numpy pass over a last-block table (host-side preprocessing, no device involved).
"""
from typing import List

import numpy as np
import torch

try:  # the reference wraps its loop in tqdm; keep that when it is available
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = lambda x: x


def _np(x) -> np.ndarray:
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def get_block(tss, src_b, dst_b) -> List[int]:
    """Block ids of one batch (dependencyGraph.py:8-28).  `tss` only fixes the length/order."""
    src, dst = _np(src_b).astype(np.int64), _np(dst_b).astype(np.int64)
    n = len(src)
    if n == 0:
        return []
    nodes, inv = np.unique(np.concatenate([src, dst]), return_inverse=True)
    s_loc, d_loc = inv[:n], inv[n:]
    last = np.full(nodes.size, -1, dtype=np.int64)   # last block id that touched each node
    out = np.empty(n, dtype=np.int64)
    for i in range(n):                                # inherently sequential: a chain per node
        b = max(last[s_loc[i]], last[d_loc[i]]) + 1
        last[s_loc[i]] = last[d_loc[i]] = b
        out[i] = b
    return out.tolist()


def _device_blocks(loader):
    """All batches of a TensorBatchLoader on the GPU (None when the loader is something else, the batch is
    above the kernel's shared-memory sort capacity, or no CUDA device is present)."""
    try:
        from temporal_dataset import TensorBatchLoader
        from tgn_b200 import _cabi, ops
    except Exception:       # pragma: no cover
        return None
    if type(loader) is not TensorBatchLoader or loader.drop_last or not torch.cuda.is_available():
        return None
    if 2 * loader.batch_size > _cabi.SORT_MAX or len(loader.dataset) == 0:
        return None
    ds = loader.dataset
    dev = torch.device("cuda")
    ids = ops.dep_blocks(torch.as_tensor(ds.src).to(dev, torch.long), torch.as_tensor(ds.dst).to(dev, torch.long),
                         loader.batch_size)
    return ids.cpu().tolist()


def dependecyAwareBatch(loader, flat: bool = True):
    """Block ids for every batch of `loader` (dependencyGraph.py:33-49); batches are dicts with
    'src', 'dst', 't', 'msg'."""
    fast = _device_blocks(loader)
    if fast is not None:
        B = loader.batch_size
        return fast if flat else [fast[lo:lo + B] for lo in range(0, len(fast), B)]
    block_ids = []
    for pos_batch in tqdm(loader):
        ids = get_block(pos_batch["t"], pos_batch["src"], pos_batch["dst"])
        if flat:
            block_ids.extend(ids)
        else:
            block_ids.append(ids)
    return block_ids
