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


def train(model, feats, train_loader, neighbor_loader, neg_dest_sampler, assoc, device, optimizer, criterion,
          use_blocks: bool = False):
    for m in model.values():
        m.train()
    model["memory"].reset_state()
    neighbor_loader.reset_state()
    feats_dev = _device_feats(feats, device)
    total_loss = 0.0
    for batch in train_loader:
        optimizer.zero_grad()
        src, pos_dst, t, msg = batch["src"], batch["dst"], batch["t"], batch["msg"]
        neg_dst = neg_dest_sampler.sample(pos_dst)
        b = batch.get("b") if use_blocks else None
        groups = [torch.arange(src.numel())] if b is None else \
            [(torch.as_tensor(b) == k).nonzero(as_tuple=True)[0] for k in range(int(torch.as_tensor(b).max()) + 1)]
        loss = 0.0
        for g in groups:
            s, d, n = src[g].to(device), pos_dst[g].to(device), neg_dst[g].to(device)
            tt, mm = t[g].to(device), msg[g].to(device, torch.float32)
            z, a = _embed(model, neighbor_loader, feats_dev, [s, d, n], device)
            pos_out = _logits(model["link_pred"], z[a[s]], z[a[d]])
            neg_out = _logits(model["link_pred"], z[a[s]], z[a[n]])
            part = criterion(pos_out, torch.ones_like(pos_out)) + criterion(neg_out, torch.zeros_like(neg_out))
            loss = loss + part * (g.numel() / src.numel())
            model["memory"].update_state(s, d, tt.long(), mm)
            neighbor_loader.insert(s, d, tt)
        loss.backward()
        optimizer.step()
        model["memory"].detach()
        total_loss += float(loss.detach()) * src.shape[0]
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
