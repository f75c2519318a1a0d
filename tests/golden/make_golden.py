"""Generates tests/golden/*.npz by running the UNMODIFIED reference files from
/root/reference (this container only -- the GPU box has no reference tree, so
the vectors are committed).

  python tests/golden/make_golden.py

What runs from the reference itself:
  neighbor_loader.LastNeighborLoader            (imports as-is)
  modules.msg_agg.{LastAggregator,MeanAggregator}
  modules.msg_func.IdentityMessage
  modules.memory_module.TGNMemory
  modules.emb_module.GraphAttentionEmbedding
  modules.decoder.LinkPredictor
The last five import torch_geometric / torch_scatter / modules.time_enc, none of
which exists here; they are satisfied with stand-in modules that re-export
oracle/thirdparty.py (restated third-party arithmetic -- "parity unpinned" for
that layer, see its header).  Everything the reference itself implements on top
(message store, train/eval ordering, flush on eval, time deltas, concat order,
edge_attr order, ...) is exercised from its own source.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

from oracle import thirdparty as tp  # noqa: E402


def install_stubs():
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_inits = types.ModuleType("torch_geometric.nn.inits")
    tg_utils = types.ModuleType("torch_geometric.utils")
    ts = types.ModuleType("torch_scatter")
    tg_nn.TransformerConv = tp.TransformerConv
    tg_inits.zeros = tp.zeros
    tg_utils.scatter = tp.scatter
    ts.scatter_max = tp.scatter_max
    tg.nn, tg.utils, tg_nn.inits = tg_nn, tg_utils, tg_inits
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                        "torch_geometric.nn.inits": tg_inits, "torch_geometric.utils": tg_utils,
                        "torch_scatter": ts})
    sys.path.insert(0, REF)
    te = types.ModuleType("modules.time_enc")   # file absent from the reference tree
    te.TimeEncoder = tp.TimeEncoder
    import modules  # the reference's package
    sys.modules["modules.time_enc"] = te
    modules.time_enc = te


def sd_np(module, prefix):
    return {f"{prefix}.{k}": v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def gen_neighbor_loader(path):
    import contextlib
    import io
    from neighbor_loader import LastNeighborLoader
    rng = np.random.default_rng(0)
    out = {}
    case = 0
    for (N, K, B, steps, bip) in [(50, 4, 8, 12, True), (200, 10, 40, 10, True), (30, 3, 6, 15, False),
                                  (64, 20, 16, 6, False)]:
        with contextlib.redirect_stdout(io.StringIO()):
            ld = LastNeighborLoader(N, K, device="cpu")
        tcur = 0.0
        for s in range(steps):
            # keep every node's multiplicity per batch <= K so the reference's
            # unstable sort cannot change the result (see oracle header)
            while True:
                if bip:
                    src = rng.integers(0, N // 2, B); dst = rng.integers(N // 2, N, B)
                else:
                    src = rng.integers(0, N, B); dst = rng.integers(0, N, B)
                cnt = np.bincount(np.concatenate([src, dst]), minlength=N)
                if cnt.max() <= K:
                    break
            t = np.sort(rng.integers(0, 50, B) + tcur).astype(np.float32); tcur = float(t[-1])
            roots = np.unique(np.concatenate([src, dst, rng.integers(0, N, B)]))
            n_id, ei, e_id, tt = ld(torch.from_numpy(roots))
            pre = f"c{case}_s{s}_"
            out[pre + "roots"] = roots
            out[pre + "n_id"] = n_id.numpy(); out[pre + "edge_index"] = ei.numpy()
            out[pre + "e_id"] = e_id.numpy(); out[pre + "t"] = tt.numpy()
            out[pre + "src"] = src; out[pre + "dst"] = dst; out[pre + "tin"] = t
            ld.insert(torch.from_numpy(src), torch.from_numpy(dst), torch.from_numpy(t))
            # state after the insert; slots with e_id < 0 hold uninitialised neighbours in the
            # reference (torch.empty) -> mask them
            eid_state = ld.e_id.numpy().copy()
            nb_state = np.where(eid_state >= 0, ld.neighbors.numpy(), 0)
            out[pre + "state_e"] = eid_state; out[pre + "state_n"] = nb_state
            out[pre + "state_t"] = ld.t.numpy().copy()
        out[f"c{case}_meta"] = np.array([N, K, B, steps])
        case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(path, **out)


def gen_aggregators(path):
    from modules.msg_agg import LastAggregator, MeanAggregator
    rng = np.random.default_rng(1)
    out = {}
    for i, (M, S, W, tmax, dt) in enumerate([(40, 12, 8, 5, "i64"), (300, 50, 472, 1000, "i64"),
                                              (64, 70, 5, 3, "f32"), (0, 4, 6, 1, "i64")]):
        msg = rng.standard_normal((M, W)).astype(np.float32)
        index = rng.integers(0, S, M).astype(np.int64)
        if i == 1:
            index = np.sort(index)
        t = rng.integers(0, tmax, M)
        t = t.astype(np.int64) if dt == "i64" else t.astype(np.float32)
        tm, ti, tt = torch.from_numpy(msg), torch.from_numpy(index), torch.from_numpy(t)
        out[f"a{i}_msg"], out[f"a{i}_index"], out[f"a{i}_t"] = msg, index, t
        out[f"a{i}_S"] = np.array(S)
        out[f"a{i}_last"] = LastAggregator()(tm, ti, tt, S).numpy()
        out[f"a{i}_mean"] = MeanAggregator()(tm, ti, tt, S).numpy()
    out["num_cases"] = np.array(4)
    np.savez_compressed(path, **out)


def gen_memory(path):
    from modules.memory_module import TGNMemory
    from modules.msg_agg import LastAggregator, MeanAggregator
    from modules.msg_func import IdentityMessage
    out = {}
    case = 0
    for (N, De, D, B, steps, aggr, tdt) in [(40, 6, 8, 10, 6, "last", "i64"), (60, 4, 12, 16, 5, "mean", "i64"),
                                            (300, 172, 100, 40, 5, "last", "i64")]:
        # NB: float timestamps cannot be pinned -- the reference itself raises at
        # memory_module.py:150 ("Index put requires the source and destination dtypes match")
        torch.manual_seed(10 + case)
        rng = np.random.default_rng(20 + case)
        mem = TGNMemory(N, De, D, D, IdentityMessage(De, D, D),
                        LastAggregator() if aggr == "last" else MeanAggregator())
        # small time weights keep cos() well-conditioned (wiki-scale deltas are exercised on the GPU side)
        with torch.no_grad():
            mem.time_enc.lin.weight.mul_(0.05)
        out.update(sd_np(mem, f"m{case}_sd"))
        mem.train()
        tcur = 0
        for s in range(steps):
            src = rng.integers(0, N // 2, B).astype(np.int64); dst = rng.integers(N // 2, N, B).astype(np.int64)
            # distinct timestamps inside a batch (see the oracle header on sort stability)
            t = np.sort(rng.choice(np.arange(tcur + 1, tcur + 200), B, replace=False)); tcur = int(t[-1])
            t = t.astype(np.int64) if tdt == "i64" else t.astype(np.float32)
            raw = rng.standard_normal((B, De)).astype(np.float32)
            q = np.unique(np.concatenate([src, dst, rng.integers(0, N, 6)])).astype(np.int64)
            if s == steps - 2:
                mem.eval()       # exercises the flush (memory_module.py:209-215) and eval ordering
            z, lu = mem(torch.from_numpy(q))
            pre = f"m{case}_s{s}_"
            out[pre + "q"], out[pre + "z"], out[pre + "lu"] = q, z.detach().numpy(), lu.detach().numpy()
            out[pre + "src"], out[pre + "dst"], out[pre + "t"], out[pre + "raw"] = src, dst, t, raw
            out[pre + "training"] = np.array(int(mem.training))
            mem.update_state(torch.from_numpy(src), torch.from_numpy(dst), torch.from_numpy(t),
                             torch.from_numpy(raw))
            mem.detach()
            out[pre + "memory"] = mem.memory.detach().numpy().copy()
            out[pre + "last_update"] = mem.last_update.numpy().copy()
        out[f"m{case}_meta"] = np.array([N, De, D, B, steps, 0 if aggr == "last" else 1, 0 if tdt == "i64" else 1])
        case += 1
    out["num_cases"] = np.array(case)
    np.savez_compressed(path, **out)


def gen_embedding(path):
    from modules.emb_module import GraphAttentionEmbedding
    from modules.decoder import LinkPredictor
    out = {}
    for case, (Nb, E, D, De, ludt) in enumerate([(30, 80, 8, 6, "i64"), (50, 0, 12, 4, "i64"),
                                                 (40, 150, 100, 172, "f32")]):
        torch.manual_seed(30 + case)
        rng = np.random.default_rng(40 + case)
        te = tp.TimeEncoder(D)
        with torch.no_grad():
            te.lin.weight.mul_(0.05)
        gnn = GraphAttentionEmbedding(D, D, De, te).eval()
        lp = LinkPredictor(D)
        x = rng.standard_normal((Nb, D)).astype(np.float32)
        lu = rng.integers(0, 500, Nb)
        lu = lu.astype(np.int64) if ludt == "i64" else lu.astype(np.float32)
        # edges grouped by centre the way the neighbour loader emits them; some nodes get none
        centres = np.sort(rng.integers(0, Nb // 2, E)).astype(np.int64)
        nbrs = rng.integers(0, Nb, E).astype(np.int64)
        t = rng.integers(0, 500, E).astype(np.float32)
        msg = rng.standard_normal((E, De)).astype(np.float32)
        ei = torch.from_numpy(np.stack([nbrs, centres]))
        z = gnn(torch.from_numpy(x), torch.from_numpy(lu), ei, torch.from_numpy(t), torch.from_numpy(msg))
        a = rng.integers(0, Nb, 20).astype(np.int64); b = rng.integers(0, Nb, 20).astype(np.int64)
        prob = lp(z[torch.from_numpy(a)], z[torch.from_numpy(b)])
        pre = f"e{case}_"
        out.update(sd_np(gnn, pre + "gnn")); out.update(sd_np(lp, pre + "lp"))
        out[pre + "x"], out[pre + "lu"], out[pre + "edge_index"] = x, lu, ei.numpy()
        out[pre + "t"], out[pre + "msg"], out[pre + "z"] = t, msg, z.detach().numpy()
        out[pre + "a"], out[pre + "b"], out[pre + "prob"] = a, b, prob.detach().numpy()
        out[pre + "meta"] = np.array([Nb, E, D, De])
    out["num_cases"] = np.array(3)
    np.savez_compressed(path, **out)


def gen_variants(path):
    """DyRepMemory (memory_module.py:218-421: rnn / gru updater, embeddings substituted for memory in the
    messages, :389-408) and TimeEmbedding (emb_module.py:32-52), both from the unmodified reference."""
    from modules.memory_module import DyRepMemory
    from modules.msg_agg import LastAggregator, MeanAggregator
    from modules.msg_func import IdentityMessage
    from modules.emb_module import TimeEmbedding
    out = {}
    case = 0
    for (N, De, D, B, steps, aggr, upd, use_s, use_d) in [(40, 6, 8, 10, 6, "last", "rnn", True, True),
                                                          (60, 4, 12, 16, 5, "mean", "gru", False, True),
                                                          (50, 3, 8, 12, 5, "last", "rnn", True, False)]:
        torch.manual_seed(40 + case)
        rng = np.random.default_rng(50 + case)
        mem = DyRepMemory(N, De, D, D, IdentityMessage(De, D, D),
                          LastAggregator() if aggr == "last" else MeanAggregator(), upd,
                          use_src_emb_in_msg=use_s, use_dst_emb_in_msg=use_d)
        with torch.no_grad():
            mem.time_enc.lin.weight.mul_(0.05)
        out.update(sd_np(mem, f"d{case}_sd"))
        mem.train()
        tcur = 0
        for s in range(steps):
            src = rng.integers(0, N // 2, B).astype(np.int64); dst = rng.integers(N // 2, N, B).astype(np.int64)
            t = np.sort(rng.choice(np.arange(tcur + 1, tcur + 200), B, replace=False)).astype(np.int64); tcur = int(t[-1])
            raw = rng.standard_normal((B, De)).astype(np.float32)
            q = np.unique(np.concatenate([src, dst, rng.integers(0, N, 6)])).astype(np.int64)
            if s == steps - 2:
                mem.eval()
            z, lu = mem(torch.from_numpy(q))
            # "current embeddings" of the batch's nodes, addressed through assoc as in the callers
            emb = rng.standard_normal((q.size, D)).astype(np.float32)
            assoc = np.zeros(N, np.int64); assoc[q] = np.arange(q.size)
            pre = f"d{case}_s{s}_"
            out[pre + "q"], out[pre + "z"], out[pre + "lu"] = q, z.detach().numpy(), lu.detach().numpy()
            out[pre + "src"], out[pre + "dst"], out[pre + "t"], out[pre + "raw"] = src, dst, t, raw
            out[pre + "emb"], out[pre + "assoc"] = emb, assoc
            out[pre + "training"] = np.array(int(mem.training))
            mem.update_state(torch.from_numpy(src), torch.from_numpy(dst), torch.from_numpy(t),
                             torch.from_numpy(raw), torch.from_numpy(emb), torch.from_numpy(assoc))
            mem.detach()
            out[pre + "memory"] = mem.memory.detach().numpy().copy()
            out[pre + "last_update"] = mem.last_update.numpy().copy()
        out[f"d{case}_meta"] = np.array([N, De, D, B, steps, 0 if aggr == "last" else 1, 0 if upd == "gru" else 1,
                                         int(use_s), int(use_d)])
        case += 1
    out["num_cases"] = np.array(case)
    torch.manual_seed(60)
    te = TimeEmbedding(16, 16)
    out.update(sd_np(te, "te_sd"))
    x = torch.randn(23, 16); lu = torch.randint(0, 500, (23,)); t = torch.randint(0, 500, (23,))
    out["te_x"], out["te_lu"], out["te_t"] = x.numpy(), lu.numpy(), t.numpy()
    out["te_out"] = te(x, lu, t).detach().numpy()
    np.savez_compressed(path, **out)


def gen_callers(path):
    """Host-side callers that import as-is from the reference: dependencyGraph.get_block
    (dependencyGraph.py:8-28), temporal_dataset.TemporalGraphDataset items (temporal_dataset.py:34-57)
    collated by torch's DataLoader exactly as utils.py:52-54 does."""
    from dependencyGraph import get_block
    from temporal_dataset import TemporalGraphDataset
    from torch.utils.data import DataLoader
    rng = np.random.default_rng(11)
    out = {}
    cases = [(1, 5), (7, 3), (200, 40), (200, 400), (64, 2)]
    for c, (B, N) in enumerate(cases):
        src = rng.integers(0, N, B); dst = rng.integers(0, N, B)
        t = np.sort(rng.integers(0, 1000, B)).astype(np.float32)
        blocks = get_block(torch.from_numpy(t), torch.from_numpy(src), torch.from_numpy(dst))
        out[f"b{c}_src"], out[f"b{c}_dst"], out[f"b{c}_t"] = src, dst, t
        out[f"b{c}_blocks"] = np.asarray(blocks, dtype=np.int64)
    out["num_block_cases"] = np.int64(len(cases))
    E, De, bs = 23, 3, 5
    src = torch.from_numpy(rng.integers(0, 9, E)); dst = torch.from_numpy(rng.integers(9, 18, E))
    t = torch.from_numpy(np.sort(rng.integers(0, 500, E))); msg = torch.from_numpy(rng.standard_normal((E, De)).astype(np.float32))
    blk = list(range(E))
    for tag, ds in (("plain", TemporalGraphDataset(src, dst, t, msg)), ("blk", TemporalGraphDataset(src, dst, t, msg, batch=blk))):
        for i, batch in enumerate(DataLoader(ds, batch_size=bs, shuffle=False)):
            for k, v in batch.items():
                out[f"dl_{tag}_{i}_{k}"] = v.numpy()
        out[f"dl_{tag}_batches"] = np.int64(i + 1)
    out["dl_src"], out["dl_dst"], out["dl_t"], out["dl_msg"], out["dl_bs"] = src.numpy(), dst.numpy(), t.numpy(), msg.numpy(), np.int64(bs)
    np.savez_compressed(path, **out)


def gen_dgl_twins(path):
    """The DGL-flavoured twins from the UNMODIFIED reference model_utils.py, imported on top of
    tests/dgl_standin (minimal `dgl`): TimeEncode, TemporalEdgePreprocess, EdgeGATConv, TemporalTransformerConv
    (eval mode: its 0.6 dropouts are identities; gradients by autograd), MemoryOperation (gru and rnn),
    EdgePredictor.  Inputs, parameters, outputs and gradients are frozen."""
    sys.path.insert(0, os.path.join(REPO, "tests", "dgl_standin"))
    import warnings
    warnings.simplefilter("ignore")
    import dgl
    mu = importlib.import_module("model_utils")
    out = {}
    g_ = torch.Generator().manual_seed(41)
    cases = [dict(N=12, E=30, De=5, D=8, H=4), dict(N=40, E=300, De=172, D=100, H=8), dict(N=9, E=0, De=3, D=8, H=2),
             dict(N=30, E=90, De=1, D=16, H=8)]
    out["num_cases"] = np.int64(len(cases))
    for c, cfg in enumerate(cases):
        N, E, De, D, H = cfg["N"], cfg["E"], cfg["De"], cfg["D"], cfg["H"]
        torch.manual_seed(100 + c)
        te = mu.TimeEncode(D)
        conv = mu.TemporalTransformerConv(De, D, te, D, H, allow_zero_in_degree=True).eval()
        src = torch.randint(0, N, (E,), generator=g_); dst = torch.randint(0, max(N - 3, 1), (E,), generator=g_)
        node_ts = (torch.rand(N, 1, generator=g_) * 1000).floor()
        edge_ts = (torch.rand(E, 1, generator=g_) * 1000).floor() + 1000
        feats = torch.randn(E, De, generator=g_)
        mem = torch.randn(N, D, generator=g_)
        g = dgl.graph((src, dst), num_nodes=N)
        g.ndata["timestamp"], g.edata["timestamp"], g.edata["feats"] = node_ts, edge_ts, feats
        efeat = conv.preprocessor(g.local_var())
        rst = conv(g, mem)
        w = torch.randn_like(rst)
        (rst * w).sum().backward()
        pre = f"c{c}_"
        for k, v in dict(src=src, dst=dst, node_ts=node_ts, edge_ts=edge_ts, feats=feats, mem=mem, efeat=efeat,
                         out=rst, out_w=w).items():
            out[pre + k] = v.detach().numpy().copy()
        out[pre + "cfg"] = np.asarray([N, E, De, D, H], np.int64)
        for k, v in conv.state_dict().items():
            out[pre + "p." + k] = v.detach().numpy().copy()
        for k, v in conv.named_parameters():
            out[pre + "g." + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy().copy()
        # MemoryOperation on the same graph (it is defined but never instantiated by the reference)
        for cell in ("gru", "rnn"):
            torch.manual_seed(200 + c)
            mm = mu.MemoryModule(N, D)
            mm.memory.data.copy_(mem)
            mm.last_update_t.data.copy_(node_ts.view(-1))
            mo = mu.MemoryOperation(cell, mm, De, te)
            g2 = dgl.graph((src, dst), num_nodes=N)
            g2.ndata[dgl.NID] = torch.arange(N)
            g2.edata["timestamp"], g2.edata["feats"] = edge_ts.view(-1), feats
            if E:
                res = mo(g2)
                out[pre + f"mo_{cell}_memory"] = res.ndata["memory"].detach().numpy().copy()
                out[pre + f"mo_{cell}_ts"] = res.ndata["timestamp"].detach().numpy().copy()
            for k, v in mo.updater.state_dict().items():
                out[pre + f"mo_{cell}.{k}"] = v.detach().numpy().copy()
    # EdgePredictor (model_utils.py:165-195), incl. the tile() pairing of several negatives per positive
    torch.manual_seed(7)
    ep = mu.EdgePredictor(16, 16)
    hs, hp, hn = torch.randn(5, 16, generator=g_), torch.randn(5, 16, generator=g_), torch.randn(15, 16, generator=g_)
    pos, neg = ep(hs, hp, hn, neg_samples=3)
    for k, v in dict(hs=hs, hp=hp, hn=hn, pos=pos, neg=neg).items():
        out["ep_" + k] = v.detach().numpy().copy()
    for k, v in ep.state_dict().items():
        out["ep_p." + k] = v.detach().numpy().copy()
    np.savez_compressed(path, **out)


def vendor_driver(path):
    """The reference's CLI driver, byte for byte, as a DATA fixture (`.txt`, never imported from here):
    tests/test_gpu_unchanged_driver.py copies it to a scratch directory as pyg-mem-tgn.py and runs it
    unchanged on top of this package's drop-in modules; tests/test_oracle_golden.py checks it is still
    identical to /root/reference/pyg-mem-tgn.py wherever the reference tree is present."""
    import shutil
    shutil.copyfile(os.path.join(REF, "pyg-mem-tgn.py"), path)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: golden vectors can only be regenerated where /root/reference exists")
    install_stubs()
    gen_neighbor_loader(os.path.join(HERE, "neighbor_loader.npz"))
    gen_aggregators(os.path.join(HERE, "aggregators.npz"))
    gen_memory(os.path.join(HERE, "memory.npz"))
    gen_embedding(os.path.join(HERE, "embedding.npz"))
    gen_callers(os.path.join(HERE, "callers.npz"))
    gen_variants(os.path.join(HERE, "variants.npz"))
    vendor_driver(os.path.join(HERE, "ref_driver_pyg-mem-tgn.py.txt"))
    gen_dgl_twins(os.path.join(HERE, "dgl_twins.npz"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
