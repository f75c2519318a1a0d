#!/usr/bin/env python
"""End-to-end driver on the drop-in modules: the call sequence of the reference's pyg-mem-tgn.py
(:36-67) -- config, data + dependency blocks, neighbour loader, model, optimiser, then per epoch
train() and validation test() -- written against this package's modules.

    python run_tgn.py --data tgbl-wiki@20000 --config config/TGN.yml [--engine]

`--data` takes a TGB name (the real `tgb` package is used when installed, otherwise the offline
synthetic stand-in; "name@events" caps the size).  `--engine` trains with the fused CUDA-graph step
(tgn_b200.engine.TGNEngine) instead of the module-by-module loop and reports events/s.
"""
import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from epoch_utils import test, train  # noqa: E402
from model_utils import getModel, getOptimizer  # noqa: E402
from neg_sampler import NegLinkSamplerDest  # noqa: E402
from neighbor_loader import LastNeighborLoader  # noqa: E402
from utils import getDataWithDependecyBlock, parse_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", type=str, default="tgbl-wiki@20000", help="dataset name")
    ap.add_argument("--config", type=str, default=os.path.join(HERE, "config", "TGN_b200.yml"))
    ap.add_argument("--epochs", type=int, default=None, help="overrides train.epoch of the config")
    ap.add_argument("--engine", action="store_true", help="fused CUDA-graph training step")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        sys.exit("run_tgn.py needs a CUDA device: the B200 hot path has no CPU fallback")
    device = torch.device("cuda")
    sample_param, memory_param, gnn_param, train_param = parse_config(args.config)
    data, train_loader, val_loader, test_loader, neg_sampler, evaluator, metric = \
        getDataWithDependecyBlock(args.data, train_param)
    if args.epochs is not None:
        train_param["epoch"] = args.epochs
    K, hidden = sample_param["neighbor"][0], gnn_param["dim_out"]
    neg_dest_sampler = NegLinkSamplerDest(torch.unique(data.dst))
    assoc = torch.empty(data.num_nodes, dtype=torch.long, device=device)
    neighbor_loader = LastNeighborLoader(data.num_nodes, size=K, device=device)
    model = getModel(data.msg.shape[1], hidden, data.num_nodes, device, gnn_param=gnn_param)
    optimizer = getOptimizer(model, train_param["lr"])
    criterion = torch.nn.BCEWithLogitsLoss()
    t_start = time.time()
    for epoch in range(train_param["epoch"]):
        t0 = time.time()
        if args.engine:
            loss, n_ev = train_epoch_engine(model, data, train_loader, neg_dest_sampler, neighbor_loader, K, hidden,
                                            train_param, device)
        else:
            loss, n_ev = train(model, data.msg, train_loader, neighbor_loader, neg_dest_sampler, assoc, device,
                               optimizer, criterion), len(train_loader.dataset)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"Epoch: {epoch + 1:02d}, Loss: {loss:.4f}, Training elapsed Time (s): {dt:.4f} ({n_ev / dt:,.0f} events/s)")
        t0 = time.time()
        val = test(model, data.msg, val_loader, neighbor_loader, neg_sampler, assoc, device, optimizer, criterion,
                   evaluator, metric, "val")
        print(f"Validation {metric}: {val:.4f}, elapsed Time (s): {time.time() - t0:.4f}")
    print(f"Execution Time: {time.time() - t_start:.6f} seconds")


_ENGINES = {}


def train_epoch_engine(model, data, train_loader, neg_dest_sampler, neighbor_loader, K, hidden, train_param, device):
    """One training epoch on TGNEngine; weights and state are handed back to the modules afterwards so that
    validation runs on the module path.  ONE engine (weights, Adam moments and step count, captured graphs)
    lives across the epochs, as the reference's optimizer does (pyg-mem-tgn.py:50,54-57).  The events that do
    not fill a whole batch are trained by a second step geometry on the SAME state (TGNEngine(share=...)), so
    memory, ring and e_id numbering see every training event, like the reference's last short DataLoader batch."""
    from tgn_b200.engine import TGNEngine
    ds = train_loader.dataset
    B = train_param["batch_size"]
    n_all = len(ds)
    n, tail = (n_all // B) * B, n_all % B
    key = id(model["memory"])
    if key not in _ENGINES:
        eng = TGNEngine(data.num_nodes, data.msg.shape[1], hidden, K, B, device=device, lr=train_param["lr"],
                        dropout=model["gnn"].conv.dropout, log_capacity=max(n_all, 1))
        eng.load_state(model["memory"].state_dict(), model["gnn"].state_dict(), model["link_pred"].state_dict())
        tail_eng = TGNEngine(data.num_nodes, data.msg.shape[1], hidden, K, tail, device=device, lr=train_param["lr"],
                             dropout=model["gnn"].conv.dropout, share=eng, use_graph=False) if tail else None
        _ENGINES[key] = (eng, tail_eng)
    eng, tail_eng = _ENGINES[key]
    eng.reset_state()
    eng.set_events(ds.src, ds.dst, ds.t.long(), ds.msg, neg_dest_sampler.sample(ds.dst))   # fresh negatives per epoch
    total = 0.0
    for _ in range(n // B):                      # every step's loss is logged, read back one step late
        prev = eng.train_step_logged()
        if prev is not None:
            total += prev * B
    if n // B:
        total += eng.flush_loss() * B
    if tail_eng is not None:
        eng.handover()
        total += float(tail_eng.train_step(from_device=True)) * tail
        tail_eng.handover()
    eng.check_device_errors()
    eng.flush_to_eval()   # what memory.eval() does on the module side: pending messages -> memory
    mem_sd, gnn_sd, lp_sd = eng.export_state()
    model["memory"].eval()
    model["memory"].reset_state()          # the module-side message store starts validation empty (it was flushed)
    model["memory"].load_state_dict(mem_sd, strict=False)
    model["gnn"].load_state_dict(gnn_sd, strict=False)
    model["link_pred"].load_state_dict(lp_sd)
    # validation continues from the flushed memory; the neighbour ring is copied over
    neighbor_loader.neighbors.copy_(eng.neighbors)
    neighbor_loader.e_id.copy_(eng.e_id)
    neighbor_loader.t.copy_(eng.t_ring)
    neighbor_loader.cur_e_id = n_all
    return total, n_all


if __name__ == "__main__":
    main()
