"""GPU: the reference driver's call sequence (pyg-mem-tgn.py:36-61) on the drop-in callers --
utils.getDataWithDependecyBlock, model_utils.getModel/getOptimizer, epoch_utils.train/test,
neighbor_loader, neg_sampler -- for one epoch on a synthetic TGB-shaped dataset, against the CPU
oracle running the same loop on the same data, weights and negative draws: training loss and
validation MRR (BASELINE.json: MRR parity within 0.005)."""
import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle_epoch(ref, data, train_loader, val_loader, neg_dest_sampler, ns, evaluator, K, lr):
    loader = orc.TorchNeighborLoader(data.num_nodes, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=lr)
    for m in ref.values():
        m.train()
    total = 0.0
    for b in train_loader:
        neg = neg_dest_sampler.sample(b["dst"])
        loss = orc.train_step(ref, loader, opt, b["src"], b["dst"], neg, b["t"].long(), b["msg"], data.t, data.msg,
                              dropout=False)
        total += loss * b["src"].numel()
    for m in ref.values():
        m.eval()
    perf = []
    with torch.no_grad():
        for b in val_loader:
            src, dst, t, msg = b["src"], b["dst"], b["t"].long(), b["msg"]
            negs = ns.query_batch(src, dst, b["t"], split_mode="val")
            q = min(len(r) for r in negs)
            neg = torch.tensor([r[:q] for r in negs])
            n_id = torch.cat([src, dst, neg.view(-1)]).unique()
            n_id, ei, e_id, _ = loader(n_id)
            z, lu = ref["memory"](n_id)
            z = ref["gnn"](z, lu, ei, data.t[e_id], data.msg[e_id])
            a = loader._assoc
            pos_o = ref["link_pred"](z[a[src]], z[a[dst]]).view(-1)
            neg_o = ref["link_pred"](z[a[src]].repeat_interleave(q, 0), z[a[neg.view(-1)]]).view(src.numel(), q)
            perf.append(evaluator.eval({"y_pred_pos": pos_o.numpy(), "y_pred_neg": neg_o.numpy(), "eval_metric": ["mrr"]})["mrr"])
            ref["memory"].update_state(src, dst, t, msg)
            loader.insert(src, dst, t.float())
    return total, float(np.mean(perf))


@pytest.mark.parametrize("use_engine", [None, False])
def test_driver_sequence_matches_oracle(use_engine):
    """use_engine=None: what the unchanged script gets -- train() recognises the standard model and runs the
    epoch on TGNEngine attached to the modules' state; False: the module-by-module loop.  Both against the
    oracle's epoch."""
    use_blocks = False
    import utils
    from epoch_utils import test as run_test, train as run_train
    from model_utils import getModel, getOptimizer
    from neg_sampler import NegLinkSamplerDest
    from neighbor_loader import LastNeighborLoader
    train_param = {"batch_size": 200, "lr": 1e-4, "epoch": 1}
    K, hidden = 10, 100
    data, tr, va, te, ns, evaluator, metric = utils.getDataWithDependecyBlock("tgbl-wiki@3000", train_param)
    neg_dest_sampler = NegLinkSamplerDest(torch.unique(data.dst))
    ref = orc.build_model(data.msg.shape[1], hidden, data.num_nodes, seed=3)
    with torch.no_grad():
        ref["memory"].time_enc.lin.weight.mul_(0.002)
    ref["gnn"].conv.dropout = 0.0
    sds = {k: {n: v.clone() for n, v in m.state_dict().items()} for k, m in ref.items()}
    torch.manual_seed(5)
    loss_ref, mrr_ref = _oracle_epoch(ref, data, tr, va, neg_dest_sampler, ns, evaluator, K, train_param["lr"])

    device = torch.device(DEV)
    model = getModel(data.msg.shape[1], hidden, data.num_nodes, device, gnn_param={"dim_out": hidden})
    for k in ("memory", "gnn", "link_pred"):
        model[k].load_state_dict(sds[k])
        model[k].to(device)
    model["gnn"].time_enc = model["memory"].time_enc
    model["gnn"].conv.dropout = 0.0
    optimizer = getOptimizer(model, train_param["lr"])
    criterion = torch.nn.BCEWithLogitsLoss()
    assoc = torch.empty(data.num_nodes, dtype=torch.long, device=device)
    neighbor_loader = LastNeighborLoader(data.num_nodes, size=K, device=device)
    torch.manual_seed(5)
    loss = run_train(model, data.msg, tr, neighbor_loader, neg_dest_sampler, assoc, device, optimizer, criterion,
                     use_blocks=use_blocks, use_engine=use_engine)
    if use_engine is None:
        import epoch_utils
        assert len(epoch_utils._ENGINES) == 1, "the standard model must take the engine path"
        assert neighbor_loader.cur_e_id == len(tr.dataset) and model["memory"].store.size == len(tr.dataset)
    mrr = run_test(model, data.msg, va, neighbor_loader, ns, assoc, device, optimizer, criterion, evaluator, metric, "val")
    assert abs(loss - loss_ref) < 2e-3 * abs(loss_ref), (loss, loss_ref)
    assert abs(mrr - mrr_ref) < 0.005, (mrr, mrr_ref)


def test_engine_epochs_equal_module_epochs_and_keep_the_optimizer():
    """Two epochs (train + validation each) through epoch_utils on the engine path and on the module path, same
    seeds: per-epoch loss sums and validation MRR agree, and the torch optimizer carries Adam's moments and
    step count across the epochs on both paths (the engine hands them back after every epoch)."""
    import utils
    from epoch_utils import test as run_test, train as run_train
    from model_utils import getModel, getOptimizer
    from neg_sampler import NegLinkSamplerDest
    from neighbor_loader import LastNeighborLoader
    train_param = {"batch_size": 100, "lr": 1e-4, "epoch": 2}
    data, tr, va, te, ns, evaluator, metric = utils.getDataWithDependecyBlock("tgbl-wiki@2500", train_param)
    device = torch.device(DEV)
    res = {}
    for use_engine in (None, False):
        torch.manual_seed(0)
        model = getModel(data.msg.shape[1], 100, data.num_nodes, device)
        with torch.no_grad():
            model["memory"].time_enc.lin.weight.mul_(0.002)    # free-running Adam: keep cos() well-conditioned
        model["gnn"].conv.dropout = 0.0
        opt = getOptimizer(model, 1e-4)
        nl = LastNeighborLoader(data.num_nodes, size=10, device=device)
        nds = NegLinkSamplerDest(torch.unique(data.dst))
        out = []
        torch.manual_seed(1)
        for _ in range(2):
            loss = run_train(model, data.msg, tr, nl, nds, None, device, opt, torch.nn.BCEWithLogitsLoss(),
                             use_engine=use_engine)
            mrr = run_test(model, data.msg, va, nl, ns, None, device, opt, None, evaluator, metric, "val")
            out.append((loss, mrr))
        p0 = next(iter(model["link_pred"].parameters()))
        res[use_engine] = (out, float(opt.state[p0]["step"]), opt.state[p0]["exp_avg"].clone())
    n_steps = 2 * len(tr)
    assert res[None][1] == n_steps and res[False][1] == n_steps, (res[None][1], res[False][1], n_steps)
    for (la, ma), (lb, mb) in zip(res[None][0], res[False][0]):
        assert abs(la - lb) < 2e-3 * abs(lb), (la, lb)
        assert abs(ma - mb) < 0.005, (ma, mb)
    # Adam's first moment of one weight matrix after two free-running epochs on either path: the same
    # quantity up to the accumulated rounding-level differences of two different reduction orders
    a, b = res[None][2], res[False][2]
    assert float((a - b).norm() / b.norm()) < 0.3, float((a - b).norm() / b.norm())
    import epoch_utils
    eng = next(iter(epoch_utils._ENGINES.values()))[0]
    assert float(eng.adam_step_dev) == n_steps


def test_blockwise_training_runs_and_differs_only_by_ordering():
    """use_blocks=True processes the dependency blocks of a batch one after the other; with one event per
    block it degenerates to batch size 1 semantics, with a single block it equals the plain loop."""
    import utils
    from epoch_utils import train as run_train
    from model_utils import getModel, getOptimizer
    from neg_sampler import NegLinkSamplerDest
    from neighbor_loader import LastNeighborLoader
    train_param = {"batch_size": 100, "lr": 1e-4, "epoch": 1}
    data, tr, *_ = utils.getDataWithDependecyBlock("tgbl-wiki@600", train_param)
    device = torch.device(DEV)
    losses = []
    for use_blocks in (False, True):
        torch.manual_seed(0)
        model = getModel(data.msg.shape[1], 32, data.num_nodes, device)
        model["gnn"].conv.dropout = 0.0
        opt = getOptimizer(model, 1e-4)
        nl = LastNeighborLoader(data.num_nodes, size=5, device=device)
        torch.manual_seed(1)
        losses.append(run_train(model, data.msg, tr, nl, NegLinkSamplerDest(torch.unique(data.dst)), None, device, opt,
                                torch.nn.BCEWithLogitsLoss(), use_blocks=use_blocks))
    assert all(np.isfinite(l) and l > 0 for l in losses)
    assert abs(losses[0] - losses[1]) < 0.2 * losses[0]


def test_tgb_gen_graph_cli_writes_ext_full(tmp_path):
    """`python tgb_gen_graph.py --data <name>` (reference README.md:4-5): the file has the keys the reference
    reads (utils.py:73), equals the oracle's builder, and feeds sampler_core."""
    import numpy as np
    import tgb_gen_graph
    import sampler_core
    from tgn_b200 import synth
    from oracle import tgn_oracle as orc
    out = str(tmp_path / "DATA" / "tgbl-wiki" / "ext_full.npz")
    tgb_gen_graph.main(["--data", "tgbl-wiki", "--max_events", "6000", "--out", out])
    g = np.load(out)
    assert sorted(g.files) == ["eid", "indices", "indptr", "ts"]
    d = synth.synth_events("tgbl-wiki", seed=0, max_events=6000)
    ref = orc.build_tcsr(d["src"], d["dst"], d["t"], d["num_nodes"])
    for k, w in zip(("indptr", "indices", "eid", "ts"), ref):
        assert np.array_equal(g[k], w), k
    s = sampler_core.ParallelSampler(g["indptr"], g["indices"], g["eid"], g["ts"], 8, 1, 1, [10], True, False, 1, 0.0)
    roots = d["src"][-64:].astype(np.int32); rts = d["t"][-64:].astype(np.float32)
    s.sample(roots, rts)
    blk = s.get_ret()[0]
    want = orc.tcsr_sample_ref(*ref, roots, rts, 10)
    assert np.array_equal(blk.eid(), want[2]) and np.array_equal(blk.col(), want[1])
