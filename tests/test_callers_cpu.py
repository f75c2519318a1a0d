"""CPU: the host-side callers of the hot path against vectors produced by the
reference's own files (tests/golden/callers.npz, written by tests/golden/make_golden.py):
dependency-aware block ids, dataset items / collated batches, plus the properties of the negative
sampler, the synthetic TGB stand-in and the config parser."""
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tgb-tgn-dgl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
G = np.load(os.path.join(REPO, "tests", "golden", "callers.npz"))


def test_block_ids_match_reference():
    from dependencyGraph import dependecyAwareBatch, get_block
    for c in range(int(G["num_block_cases"])):
        src, dst, t = (torch.from_numpy(G[f"b{c}_{k}"]) for k in ("src", "dst", "t"))
        assert get_block(t, src, dst) == G[f"b{c}_blocks"].tolist()
    # events of one block touch disjoint nodes; flat / nested output of the loader walk
    src, dst = torch.from_numpy(G["b3_src"]), torch.from_numpy(G["b3_dst"])
    blocks = np.asarray(get_block(torch.from_numpy(G["b3_t"]), src, dst))
    for k in range(blocks.max() + 1):
        nodes = torch.cat([src[blocks == k], dst[blocks == k]]).numpy()
        pairs = np.stack([src[blocks == k].numpy(), dst[blocks == k].numpy()], 1)
        loops = int((pairs[:, 0] == pairs[:, 1]).sum())
        assert len(np.unique(nodes)) == len(nodes) - loops
    loader = [dict(src=src[:50], dst=dst[:50], t=src[:50], msg=None), dict(src=src[50:], dst=dst[50:], t=src[50:], msg=None)]
    flat = dependecyAwareBatch(loader)
    nested = dependecyAwareBatch(loader, flat=False)
    assert flat == nested[0] + nested[1] and len(flat) == 200


@pytest.mark.parametrize("tag", ["plain", "blk"])
def test_dataset_and_batch_loader_match_reference_dataloader(tag):
    from temporal_dataset import TemporalGraphDataset, TensorBatchLoader
    src, dst, t, msg = (torch.from_numpy(G[f"dl_{k}"]) for k in ("src", "dst", "t", "msg"))
    ds = TemporalGraphDataset(src, dst, t, msg, batch=None if tag == "plain" else list(range(len(src))))
    item = ds[3]
    assert item["t"].dtype == torch.float32 and int(item["idx"]) == 3 and ("b" in item) == (tag == "blk")
    batches = list(TensorBatchLoader(ds, int(G["dl_bs"])))
    assert len(batches) == int(G[f"dl_{tag}_batches"])
    for i, b in enumerate(batches):
        keys = {k.split("_", 3)[3] for k in G.files if k.startswith(f"dl_{tag}_{i}_")}
        assert keys == set(b.keys())
        for k in keys:
            ref = G[f"dl_{tag}_{i}_{k}"]
            got = b[k].numpy()
            assert got.dtype == ref.dtype and np.array_equal(got, ref), (i, k)


def test_negative_sampler_properties():
    from neg_sampler import NegLinkSamplerDest
    torch.manual_seed(0)
    dst_nodes = torch.tensor([5, 9, 11, 40])
    s = NegLinkSamplerDest(dst_nodes)
    pos = dst_nodes[torch.randint(0, 4, (5000,))]
    neg = s.sample(pos)
    assert neg.dtype == pos.dtype and neg.shape == pos.shape
    assert bool((neg != pos).all()) and set(neg.tolist()) <= set(dst_nodes.tolist())
    # uniform over the three remaining destinations
    for p in dst_nodes.tolist():
        cnt = torch.bincount(neg[pos == p], minlength=41)[dst_nodes]
        cnt = cnt[cnt > 0].float()
        assert cnt.numel() == 3 and float(cnt.max() / cnt.min()) < 1.25
    torch.manual_seed(1); a = s.sample(pos)
    torch.manual_seed(1); b = s.sample(pos)
    assert torch.equal(a, b)


def test_synthetic_tgb_stand_in_and_utils(tmp_path):
    import tgb_synth
    import utils
    from oracle import tgn_oracle as orc
    ds = tgb_synth.PyGLinkPropPredDataset("tgbl-wiki@3000")
    data = ds.get_TemporalData()
    assert data.num_nodes == 9227 and data.msg.shape == (3000, 172) and ds.eval_metric == "mrr"
    assert int(ds.train_mask.sum()) == 2100 and int(ds.val_mask.sum()) == 450 and int(ds.test_mask.sum()) == 450
    assert bool((data.t[1:] >= data.t[:-1]).all())
    val = data[ds.val_mask]
    neg = ds.negative_sampler.query_batch(val.src[:7], val.dst[:7], val.t[:7], split_mode="val")
    assert len(neg) == 7 and all(len(r) == 20 for r in neg)
    assert all(int(val.dst[i]) not in neg[i] for i in range(7))
    assert neg == ds.negative_sampler.query_batch(val.src[:7], val.dst[:7], val.t[:7], split_mode="val")
    rng = np.random.default_rng(0)
    pos, ng = rng.random(9).astype(np.float32), rng.random((9, 20)).astype(np.float32)
    ng[:, 3] = pos
    got = tgb_synth.Evaluator("x").eval({"y_pred_pos": pos, "y_pred_neg": ng, "eval_metric": ["mrr"]})["mrr"]
    assert abs(got - float(orc.mrr_ref(pos, ng).mean())) < 1e-7
    cfg = tmp_path / "TGN.yml"
    cfg.write_text("sampling:\n  - layer: 1\n    neighbor: [10]\nmemory:\n  - type: node\ngnn:\n  - dim_out: 100\n"
                   "train:\n  - epoch: 1\n    batch_size: 200\n    lr: 0.0001\n")
    sample_param, memory_param, gnn_param, train_param = utils.parse_config(str(cfg))
    assert sample_param["neighbor"][0] == 10 and gnn_param["dim_out"] == 100 and train_param["batch_size"] == 200
    data, tr, va, te, ns, ev, metric = utils.getDataWithDependecyBlock("tgbl-wiki@3000", train_param)
    assert metric == "mrr" and len(tr) == 11 and len(va) == 3 and len(te) == 3
    first = next(iter(tr))
    assert set(first.keys()) == {"src", "dst", "t", "msg", "b", "idx"} and first["src"].numel() == 200
    assert int(first["b"].min()) == 0
