"""Drop-in `epoch_utils.{train, test}` (reference epoch_utils.py:168-318, :15-165 signatures and
return values) running the TGN loop the reference keeps in comments (pyg_epoch_utils.py:106-137,
epoch_utils.py:129-157) on the sm_100a modules:

    n_id = unique(src, dst, neg) -> neighbor_loader(n_id) -> memory(n_id) -> gnn -> link_pred
    -> BCE-with-logits -> memory.update_state -> neighbor_loader.insert -> backward -> step

Kept from the reference: `train` returns the SUM of loss * batch_size (:310,318); the criterion is
BCEWithLogitsLoss, so the decoder hands over logits (LinkPredictor.logits; the reference's own
PyG decoder ends in a sigmoid); `test` truncates every negative list to the
shortest one of the batch (:48-56) and returns the mean of the per-batch evaluator outputs (:163);
edge timestamps come from the neighbour loader's fourth output (:220); edge features are
`feats[e_id]` (:224) -- gathered on the device from a copy uploaded once, not through the host.
`use_blocks=True` additionally honours the dependency-aware block ids `b` of each batch
(dependencyGraph.py): the events of a batch are processed block by block, each block seeing the
memory and neighbour state left by the previous one.

Fast path.  When `train` is handed exactly what the driver script builds (pyg-mem-tgn.py:44-51 on this
package's model_utils: TGNMemory with IdentityMessage + LastAggregator + GRUCell, GraphAttentionEmbedding,
LinkPredictor, torch.optim.Adam with default betas, BCEWithLogitsLoss, a TensorBatchLoader) the epoch runs
on tgn_b200.engine.TGNEngine -- the same step as one captured CUDA graph -- ATTACHED to the modules' own
state tensors (memory, last_update, message store, neighbour ring): after the epoch the modules, the
neighbour loader and the optimizer are in the state the module-by-module loop leaves them in, so `test`
(and a later module-path `train`) continue from it.  The unchanged script therefore gets the engine's
throughput.  `use_engine=False` forces the module-by-module loop; anything non-standard falls back to it.
"""
import numpy as np
import torch

_feat_cache = {}


def _device_feats(feats: torch.Tensor, device) -> torch.Tensor:
    key = (feats.data_ptr(), tuple(feats.shape), str(device))
    if key not in _feat_cache:
        _feat_cache.clear()
        _feat_cache[key] = feats.to(device, torch.float32).contiguous()
    return _feat_cache[key]


def _embed(model, neighbor_loader, feats_dev, ids, device):
    n_id = torch.cat(ids).unique().to(device)
    n_id, edge_index, e_id, edge_t = neighbor_loader(n_id)
    z, last_update = model["memory"](n_id)
    z = model["gnn"](z, last_update, edge_index, edge_t, feats_dev[e_id])
    return z, neighbor_loader._assoc


def _logits(link_pred, z_src, z_dst):
    return link_pred.logits(z_src, z_dst) if hasattr(link_pred, "logits") else link_pred(z_src, z_dst)


_ENGINES = {}


def _engine_eligible(model, feats, train_loader, neighbor_loader, optimizer, criterion, use_blocks) -> bool:
    from modules.decoder import LinkPredictor
    from modules.emb_module import GraphAttentionEmbedding
    from modules.memory_module import TGNMemory
    from modules.msg_agg import LastAggregator
    from modules.msg_func import IdentityMessage
    from temporal_dataset import TensorBatchLoader
    if use_blocks or set(model.keys()) != {"memory", "gnn", "link_pred"}:
        return False
    mem, gnn, lp = model["memory"], model["gnn"], model["link_pred"]
    if not (type(mem) is TGNMemory and type(gnn) is GraphAttentionEmbedding and type(lp) is LinkPredictor):
        return False
    if not (type(mem.msg_s_module) is IdentityMessage and type(mem.aggr_module) is LastAggregator and
            isinstance(mem.memory_updater, torch.nn.GRUCell) and gnn.time_enc is mem.time_enc):
        return False
    if not (mem.memory_dim == mem.time_dim and mem.memory_dim % 4 == 0 and mem.memory_dim % gnn.conv.heads == 0 and
            gnn.conv.heads * gnn.conv.out_channels == mem.memory_dim and mem.memory.is_cuda):
        return False
    if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
        return False
    g = optimizer.param_groups[0]
    if not (tuple(g["betas"]) == (0.9, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 0 and
            not g.get("amsgrad") and not g.get("maximize")):
        return False
    want = {id(p) for k in ("memory", "gnn", "link_pred") for p in model[k].parameters()}
    if {id(p) for p in g["params"]} != want:
        return False
    if type(criterion) is not torch.nn.BCEWithLogitsLoss or criterion.reduction != "mean" or \
            criterion.weight is not None or criterion.pos_weight is not None:
        return False
    if type(train_loader) is not TensorBatchLoader or train_loader.drop_last:
        return False
    ds = train_loader.dataset
    n = len(ds)
    # edge features are feats[e_id] with e_id = position in the stream (epoch_utils.py:224): the training
    # events must be the first rows of `feats`
    if n < 1 or feats.shape[0] < n or feats.shape[1] != ds.msg.shape[1] or n // train_loader.batch_size < 1:
        return False
    probe = torch.linspace(0, n - 1, min(n, 64)).long()
    return bool(torch.equal(feats[probe].to(ds.msg.dtype).cpu(), ds.msg[probe].cpu()))


def _train_engine(model, train_loader, neighbor_loader, neg_dest_sampler, device, optimizer):
    from tgn_b200.engine import TGNEngine
    mem, gnn = model["memory"], model["gnn"]
    ds, B = train_loader.dataset, train_loader.batch_size
    n_all = len(ds)
    n, tail = (n_all // B) * B, n_all % B
    lr = optimizer.param_groups[0]["lr"]
    key = (id(mem), id(neighbor_loader), B, tail, float(gnn.conv.dropout), mem.memory.data_ptr())
    if key not in _ENGINES:
        _ENGINES.clear()            # one model at a time: drop the graphs / workspaces of the previous one
        kw = dict(device=mem.memory.device, lr=lr, heads=gnn.conv.heads, dropout=float(gnn.conv.dropout),
                  track_metrics=True)
        eng = TGNEngine(mem.num_nodes, mem.raw_msg_dim, mem.memory_dim, neighbor_loader.size, B, log_capacity=16, **kw)
        eng.attach_modules(mem, neighbor_loader)
        tail_eng = TGNEngine(mem.num_nodes, mem.raw_msg_dim, mem.memory_dim, neighbor_loader.size, tail, share=eng,
                             use_graph=False, **kw) if tail else None
        _ENGINES[key] = (eng, tail_eng)
    eng, tail_eng = _ENGINES[key]
    if eng.lr != lr:                # the captured graphs hold the learning rate by value
        eng.lr = lr
        eng._drop_graphs()
        if tail_eng is not None:
            tail_eng.lr = lr
    eng.sync_from_modules(model, optimizer)
    eng.begin_epoch_on_modules()
    eng.metric_acc.zero_()
    # one negative per positive, drawn batch by batch exactly as the module loop draws them (same use of
    # torch's generator, epoch_utils.py:198), then the whole epoch is resident
    neg = torch.cat([neg_dest_sampler.sample(ds.dst[lo:lo + B]) for lo in range(0, n_all, B)])
    # the loader hands out float32 timestamps (temporal_dataset.py:42) and update_state gets t.long()
    eng.set_events(ds.src, ds.dst, ds.t.float().long(), ds.msg, neg)
    # Every step's loss enters the epoch total (epoch_utils.py:305), summed ON THE DEVICE in float64: the steps run as
    # captured graphs of `group_size` batches (no per-step host round trip), the group's per-slot loss words are
    # added to the accumulator behind each replay, and the host reads the total once per epoch.
    steps, G = n // B, eng.group_size
    acc = torch.zeros((), dtype=torch.float64, device=mem.memory.device)
    idx_cache = eng.__dict__.setdefault("_loss_idx", {})
    done = 0
    if steps:
        acc += eng.train_step(from_device=True)           # primes the sampling pipeline
        done = 1
    while eng.use_graph and steps - done >= G:
        c0 = eng.cur
        if c0 not in idx_cache:
            idx_cache[c0] = torch.tensor([(c0 + k) % eng.nslots for k in range(G)], device=mem.memory.device)
        eng.train_steps(G)
        acc += eng.loss_slots.index_select(0, idx_cache[c0]).sum()
        done += G
    while done < steps:
        acc += eng.train_step(from_device=True, _capture=False)
        done += 1
    total = float(acc) * B
    if tail_eng is not None:
        eng.handover()
        total += float(tail_eng.train_step(from_device=True)) * tail
        tail_eng.handover()
    torch.cuda.synchronize()
    eng.check_device_errors()
    eng.end_epoch_on_modules()
    eng.sync_to_modules(model, optimizer)
    mem.detach()
    ap, auc = eng.epoch_metrics()
    print("ap and auc: ", ap, auc)                       # epoch_utils.py:317
    return total


def train(model, feats, train_loader, neighbor_loader, neg_dest_sampler, assoc, device, optimizer, criterion,
          use_blocks: bool = False, use_engine=None):
    for m in model.values():
        m.train()
    model["memory"].reset_state()
    neighbor_loader.reset_state()
    if use_engine is None:
        use_engine = _engine_eligible(model, feats, train_loader, neighbor_loader, optimizer, criterion, use_blocks)
    if use_engine:
        return _train_engine(model, train_loader, neighbor_loader, neg_dest_sampler, device, optimizer)
    feats_dev = _device_feats(feats, device)
    total_loss = 0.0
    from tgn_b200 import ops
    metric_acc = torch.zeros(3, dtype=torch.float64, device=device)   # per-batch AP / AUC, summed on the device
    for batch in train_loader:
        optimizer.zero_grad()
        src, pos_dst, t, msg = batch["src"], batch["dst"], batch["t"], batch["msg"]
        neg_dst = neg_dest_sampler.sample(pos_dst)
        b = batch.get("b") if use_blocks else None
        groups = [torch.arange(src.numel())] if b is None else \
            [(torch.as_tensor(b) == k).nonzero(as_tuple=True)[0] for k in range(int(torch.as_tensor(b).max()) + 1)]
        loss = 0.0
        pos_all, neg_all = [], []
        for g in groups:
            s, d, n = src[g].to(device), pos_dst[g].to(device), neg_dst[g].to(device)
            tt, mm = t[g].to(device), msg[g].to(device, torch.float32)
            z, a = _embed(model, neighbor_loader, feats_dev, [s, d, n], device)
            pos_out = _logits(model["link_pred"], z[a[s]], z[a[d]])
            neg_out = _logits(model["link_pred"], z[a[s]], z[a[n]])
            pos_all.append(pos_out.detach().reshape(-1))
            neg_all.append(neg_out.detach().reshape(-1))
            part = criterion(pos_out, torch.ones_like(pos_out)) + criterion(neg_out, torch.zeros_like(neg_out))
            loss = loss + part * (g.numel() / src.numel())
            model["memory"].update_state(s, d, tt.long(), mm)
            neighbor_loader.insert(s, d, tt)
        loss.backward()
        optimizer.step()
        model["memory"].detach()
        total_loss += float(loss.detach()) * src.shape[0]
        # epoch_utils.py:312-315 (sklearn on the host, every batch) -> one kernel, no host copy
        ops.ap_auc_accum(torch.cat(pos_all), torch.cat(neg_all), metric_acc)
    acc = metric_acc.tolist()
    print("ap and auc: ", acc[0] / max(acc[2], 1.0), acc[1] / max(acc[2], 1.0))   # epoch_utils.py:317
    return total_loss


@torch.no_grad()
def test(model, feats, loader, neighbor_loader, neg_sampler, assoc, device, optimizer, criterion, evaluator, metric,
         split_mode):
    for m in model.values():
        m.eval()
    feats_dev = _device_feats(feats, device)
    perf_list = []
    for batch in loader:
        src, pos_dst, t, msg = batch["src"], batch["dst"], batch["t"], batch["msg"]
        neg_lists = neg_sampler.query_batch(src, pos_dst, t, split_mode=split_mode)
        min_size = min(len(r) for r in neg_lists)                      # epoch_utils.py:48-56
        neg_dst = torch.tensor([r[:min_size] for r in neg_lists], dtype=torch.long)
        s, d, n = src.to(device), pos_dst.to(device), neg_dst.to(device)
        z, a = _embed(model, neighbor_loader, feats_dev, [s, d, n.reshape(-1)], device)
        pos_out = model["link_pred"](z[a[s]], z[a[d]])
        neg_out = model["link_pred"](z[a[s]].repeat_interleave(min_size, 0), z[a[n.reshape(-1)]])
        input_dict = {"y_pred_pos": pos_out.reshape(-1).cpu().numpy(),
                      "y_pred_neg": neg_out.reshape(s.numel(), -1).cpu().numpy(),
                      "eval_metric": [metric]}
        perf_list.append(evaluator.eval(input_dict)[metric])
        model["memory"].update_state(s, d, t.to(device).long(), msg.to(device, torch.float32))
        neighbor_loader.insert(s, d, t.to(device))
    return float(torch.tensor(perf_list).mean())
