"""GPU: the fused training step (tgn_b200.engine.TGNEngine, eager and CUDA-graph
replay) against the CPU oracle's training step on the same weights and events:
loss per step, memory, last_update (bit-exact), neighbour ring (bit-exact), weights
after Adam; and the evaluation step's MRR."""
import numpy as np
import pytest
import torch

from oracle import tgn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(N, De, D, K, B, E, seed, use_graph, lr=1e-3):
    from tgn_b200.engine import TGNEngine
    from tgn_b200 import synth
    rng = np.random.default_rng(seed)
    ns = N // 2
    src = np.floor(rng.random(E) ** 2 * ns).astype(np.int64)
    dst = ns + np.floor(rng.random(E) ** 2 * (N - ns)).astype(np.int64)
    t = np.sort(rng.integers(0, 40 * E, E)).astype(np.int64)
    msg = rng.standard_normal((E, De)).astype(np.float32)
    neg = synth.sample_negatives(dst, np.unique(dst), np.random.default_rng(seed + 1))
    ref = orc.build_model(De, D, N, seed=seed)
    with torch.no_grad():
        ref["memory"].time_enc.lin.weight.mul_(0.002)   # keep cos() well-conditioned at these time deltas
    ref["gnn"].conv.dropout = 0.0
    for m in ref.values():
        m.train()
    eng = TGNEngine(N, De, D, K, B, device=DEV, lr=lr, dropout=0.0, use_graph=use_graph, log_capacity=E)
    eng.load_state(ref["memory"].state_dict(), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    ev = dict(src=torch.from_numpy(src), dst=torch.from_numpy(dst), t=torch.from_numpy(t),
              msg=torch.from_numpy(msg), neg=torch.from_numpy(neg))
    eng.set_events(**ev)
    return ref, eng, ev


def _oracle_grads(ref):
    g = {k: v.grad for k, v in ref["memory"].named_parameters()}
    gg = dict(ref["gnn"].named_parameters())
    g["conv.w_node"] = torch.cat([gg[f"conv.lin_{n}.weight"].grad for n in ("query", "key", "value", "skip")])
    g["conv.b_node"] = torch.cat([gg[f"conv.lin_{n}.bias"].grad for n in ("query", "key", "value", "skip")])
    g["conv.lin_edge.weight"] = gg["conv.lin_edge.weight"].grad
    g.update({k: v.grad for k, v in ref["link_pred"].named_parameters()})
    return g


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_steps_match_oracle(use_graph):
    """Every step: same loss and same gradients from the same weights and state.  Adam turns
    rounding-level gradient noise on near-zero gradients into lr-sized weight differences, so
    the oracle's post-step weights are copied into the engine after each step (Adam itself is
    checked in test_adam_matches_torch and by the free-running test below)."""
    N, De, D, K, B, steps = 400, 12, 32, 5, 50, 10
    ref, eng, ev = _setup(N, De, D, K, B, B * steps, 7, use_graph)
    loader = orc.TorchNeighborLoader(N, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=1e-3)
    strip = lambda sd: {k: v for k, v in sd.items() if k not in ("memory", "last_update", "_assoc")}
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        loss = float(eng.train_step(from_device=True))
        loss_ref = orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl],
                                  ev["msg"][sl], ev["t"], ev["msg"], dropout=False)
        assert abs(loss - loss_ref) < 1e-4 * max(1.0, abs(loss_ref)), (s, loss, loss_ref)
        for k, g in _oracle_grads(ref).items():
            torch.testing.assert_close(eng.p[k].grad.cpu(), g, rtol=2e-3, atol=2e-6, msg=lambda m: f"step {s} {k}: {m}")
        torch.testing.assert_close(eng.memory.cpu(), ref["memory"].memory.detach(), rtol=1e-4, atol=1e-5)
        assert torch.equal(eng.last_update.cpu(), ref["memory"].last_update)
        eng.load_state(strip(ref["memory"].state_dict()), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    assert torch.equal(eng.e_id.cpu(), loader.e_id)
    valid = loader.e_id >= 0
    assert torch.equal(eng.neighbors.cpu()[valid], loader.neighbors[valid])
    assert torch.equal(eng.t_ring.cpu(), loader.t)


def test_gemm_decoder_path_matches_oracle():
    """the decoder on batched GEMMs (used when the weights do not fit the fused kernel's shared memory);
    K = 14 > 12 also takes the general (any-degree) attention kernels instead of the register-resident ones"""
    N, De, D, K, B, steps = 400, 12, 32, 14, 50, 6
    ref, eng, ev = _setup(N, De, D, K, B, B * steps, 13, False)
    eng.fused_decoder = False
    loader = orc.TorchNeighborLoader(N, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=1e-3)
    strip = lambda sd: {k: v for k, v in sd.items() if k not in ("memory", "last_update", "_assoc")}
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        loss = float(eng.train_step(from_device=True))
        loss_ref = orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl],
                                  ev["msg"][sl], ev["t"], ev["msg"], dropout=False)
        assert abs(loss - loss_ref) < 1e-4 * max(1.0, abs(loss_ref)), (s, loss, loss_ref)
        for k, g in _oracle_grads(ref).items():
            torch.testing.assert_close(eng.p[k].grad.cpu(), g, rtol=2e-3, atol=2e-6, msg=lambda m: f"step {s} {k}: {m}")
        eng.load_state(strip(ref["memory"].state_dict()), ref["gnn"].state_dict(), ref["link_pred"].state_dict())


def test_free_running_training_tracks_oracle():
    """no weight syncing: 12 Adam steps stay close (loose tolerance, see above)."""
    N, De, D, K, B, steps = 400, 12, 32, 5, 50, 12
    ref, eng, ev = _setup(N, De, D, K, B, B * steps, 9, True)
    loader = orc.TorchNeighborLoader(N, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=1e-3)
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        loss = float(eng.train_step(from_device=True))
        loss_ref = orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl],
                                  ev["msg"][sl], ev["t"], ev["msg"], dropout=False)
        assert abs(loss - loss_ref) < 1e-2, (s, loss, loss_ref)
    assert torch.equal(eng.last_update.cpu(), ref["memory"].last_update)
    assert torch.equal(eng.e_id.cpu(), loader.e_id)
    assert float(eng.adam_step_dev) == steps


def test_host_staged_batches_equal_device_staged():
    """three ways to feed the same batches: resident arrays (pipelined), host staging, host staging with
    one batch of lookahead (the next batch is sampled on the forked stream while the current one trains)"""
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 8
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 11, True)
    _, eng_b, _ = _setup(N, De, D, K, B, B * steps, 11, True)
    _, eng_c, _ = _setup(N, De, D, K, B, B * steps, 11, True)
    pin = {k: v.pin_memory() for k, v in ev.items()}
    batch = lambda s: tuple(pin[k][s * B:(s + 1) * B] for k in ("src", "dst", "neg", "t", "msg"))
    eng_c.stage_batch(*batch(0))
    for s in range(steps):
        la = float(eng_a.train_step(from_device=True))
        eng_b.stage_batch(*batch(s))
        lb = float(eng_b.train_step(from_device=False))
        eng_c.stage_batch(*batch(min(s + 1, steps - 1)), ahead=True)
        lc = float(eng_c.train_step(from_device=False, lookahead=True))
        assert abs(la - lb) < 1e-4 and abs(la - lc) < 1e-4, (s, la, lb, lc)   # atomics in the gradient reductions: not bitwise
    torch.testing.assert_close(eng_a.memory, eng_b.memory, rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(eng_a.memory, eng_c.memory, rtol=1e-2, atol=1e-3)
    assert torch.equal(eng_a.e_id, eng_b.e_id) and torch.equal(eng_a.e_id, eng_c.e_id)
    assert torch.equal(eng_a.last_update, eng_c.last_update)
    assert torch.equal(eng_a.t_ring, eng_b.t_ring) and torch.equal(eng_a.t_ring, eng_c.t_ring)


def test_single_copy_staging_and_lagged_loss():
    """stage_packed1 (one pinned H2D copy per batch, float ring timestamps derived inside the step) with
    train_step_logged (loss read back one step late) == the resident-array path."""
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 8
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 11, True)
    _, eng_d, _ = _setup(N, De, D, K, B, B * steps, 11, True)
    host = torch.empty((steps, eng_d.packed_nbytes()), dtype=torch.uint8).pin_memory()
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        eng_d.pack_host_batch(host[s], ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], ev["msg"][sl])
    ref_losses = [float(eng_a.train_step(from_device=True)) for _ in range(steps)]
    got = []
    eng_d.stage_packed1(host[0])
    for s in range(steps):
        eng_d.stage_packed1(host[min(s + 1, steps - 1)], ahead=True)
        got.append(eng_d.train_step_logged(from_device=False, lookahead=True))
    got = got[1:] + [eng_d.flush_loss()]
    assert got[0] is not None and len(got) == steps
    for s, (a, b) in enumerate(zip(ref_losses, got)):
        assert abs(a - b) < 1e-4, (s, a, b)
    torch.testing.assert_close(eng_a.memory, eng_d.memory, rtol=1e-2, atol=1e-3)
    assert torch.equal(eng_a.e_id, eng_d.e_id) and torch.equal(eng_a.t_ring, eng_d.t_ring)
    assert torch.equal(eng_a.last_update, eng_d.last_update)


def test_fused_zero_grad_is_equivalent():
    """Adam clearing the gradients / loss / embedding-gradient rows itself (no per-step memset)
    trains exactly like the memset variant."""
    from tgn_b200.engine import TGNEngine
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 8
    ref, eng_a, ev = _setup(N, De, D, K, B, B * steps, 17, True)
    eng_b = TGNEngine(N, De, D, K, B, device=DEV, lr=1e-3, dropout=0.0, use_graph=True, log_capacity=B * steps,
                      fused_zero_grad=True)
    eng_b.load_state(ref["memory"].state_dict(), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    eng_b.set_events(**ev)
    for s in range(steps):
        la, lb = float(eng_a.train_step()), float(eng_b.train_step())
        assert abs(la - lb) < 1e-4, (s, la, lb)
    assert float(eng_b.flat_grad.abs().sum()) == 0.0 and float(eng_b.d_emb.abs().sum()) == 0.0
    assert float(eng_b.adam_step_dev) == steps and int(eng_b.step_dev) == steps
    torch.testing.assert_close(eng_a.memory, eng_b.memory, rtol=1e-2, atol=1e-3)


def test_eval_mrr_matches_oracle():
    """train a few steps, switch to eval (flush), score one batch against Q negatives."""
    from tgn_b200 import ops, synth
    N, De, D, K, B, steps, Q = 300, 8, 16, 5, 40, 6, 20
    ref, eng, ev = _setup(N, De, D, K, B, B * (steps + 2), 5, False)
    loader = orc.TorchNeighborLoader(N, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=1e-3)
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], ev["msg"][sl],
                       ev["t"], ev["msg"], dropout=False)
        eng.train_step(from_device=True)
    for m in ref.values():
        m.eval()
    eng.flush_to_eval()
    torch.testing.assert_close(eng.memory.cpu(), ref["memory"].memory.detach(), rtol=1e-3, atol=1e-4)
    assert torch.equal(eng.last_update.cpu(), ref["memory"].last_update)
    for s in range(steps, steps + 2):
        sl = slice(s * B, (s + 1) * B)
        src, dst, t, msg = ev["src"][sl], ev["dst"][sl], ev["t"][sl], ev["msg"][sl]
        neg = torch.from_numpy(synth.eval_negatives(src.numpy(), dst.numpy(), N, Q, seed=s, dst_lo=N // 2))
        with torch.no_grad():
            n_id = torch.cat([src, dst, neg.view(-1)]).unique()
            n_id, ei, e_id, _ = loader(n_id)
            z, lu = ref["memory"](n_id)
            z = ref["gnn"](z, lu, ei, ev["t"][e_id], ev["msg"][e_id])
            a = loader._assoc
            pos_r = ref["link_pred"](z[a[src]], z[a[dst]]).view(-1)
            neg_r = ref["link_pred"](z[a[src]].repeat_interleave(Q, 0), z[a[neg.view(-1)]]).view(B, Q)
            ref["memory"].update_state(src, dst, t, msg)
            loader.insert(src, dst, t.float())
        pos_g, neg_g = eng.eval_scores(src, dst, neg, t, msg)
        torch.testing.assert_close(pos_g.cpu(), pos_r, rtol=1e-3, atol=1e-5)
        torch.testing.assert_close(neg_g.cpu(), neg_r, rtol=1e-3, atol=1e-5)
        mrr_g = float(ops.mrr(pos_g, neg_g).mean())
        mrr_r = float(orc.mrr_ref(pos_r.numpy(), neg_r.numpy()).mean())
        assert abs(mrr_g - mrr_r) < 0.005
    torch.testing.assert_close(eng.memory.cpu(), ref["memory"].memory.detach(), rtol=1e-3, atol=1e-4)
    assert torch.equal(eng.last_update.cpu(), ref["memory"].last_update)


def test_eval_counts_are_additive_over_negative_shards():
    """data-parallel evaluation: the TGB rank counts of column shards of the negatives sum to the
    counts of the full matrix (bit-exact), the state after the batch does not depend on the shard,
    and the MRR equals the one from the materialised scores."""
    from tgn_b200 import dist_eval, ops, synth
    N, De, D, K, B, steps, Q = 300, 8, 16, 5, 40, 6, 37
    engs = []
    for _ in range(3):   # identical replicas; the eval path has no atomics, so they stay bit-identical
        _, eng, ev = _setup(N, De, D, K, B, B * steps, 21, False)
        eng.flush_to_eval()
        engs.append(eng)
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        src, dst, t, msg = ev["src"][sl], ev["dst"][sl], ev["t"][sl], ev["msg"][sl]
        neg = torch.from_numpy(synth.eval_negatives(src.numpy(), dst.numpy(), N, Q, seed=s, dst_lo=N // 2))
        pos, negs, gt, ge = engs[0].eval_batch(src, dst, neg, t, msg)
        parts = [engs[1 + r].eval_batch(src, dst, dist_eval.shard_columns(neg, r, 2), t, msg, want_neg_scores=False)
                 for r in range(2)]
        assert torch.equal(gt, parts[0][2] + parts[1][2]) and torch.equal(ge, parts[0][3] + parts[1][3])
        torch.testing.assert_close(pos, parts[0][0], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(pos, parts[1][0], rtol=1e-6, atol=1e-7)
        rr = dist_eval.reciprocal_ranks(gt, ge)
        assert torch.equal(rr, ops.mrr(pos, negs))
        for e in engs[1:]:
            assert torch.equal(e.memory, engs[0].memory) and torch.equal(e.last_update, engs[0].last_update)
            assert torch.equal(e.e_id, engs[0].e_id)
    assert int(engs[0].e_id.ge(0).sum()) > 0 and float(engs[0].memory.abs().sum()) > 0


def test_multi_step_graph_equals_single_steps():
    """train_steps(n): three steps per captured graph (+ remainder) == n single-step launches."""
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 26
    # (small lr: two identical engines drift apart after ~15 steps at lr 1e-3 -- split-K atomics give
    # rounding-level gradient noise that Adam amplifies; the comparison is about the step plumbing)
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 13, True, lr=1e-6)
    _, eng_b, _ = _setup(N, De, D, K, B, B * steps, 13, True, lr=1e-6)
    la = [float(eng_a.train_step(from_device=True)) for _ in range(steps)]
    done = 0
    for n in (1, 9, 2, 14):            # eager warm-up calls, capture, replays, remainders
        lb = float(eng_b.train_steps(n))
        done += n
        assert abs(lb - la[done - 1]) < 1e-4, (done, lb, la[done - 1])
    assert done == steps and eng_b.events_done == eng_a.events_done
    torch.testing.assert_close(eng_a.memory, eng_b.memory, rtol=1e-2, atol=1e-3)
    assert torch.equal(eng_a.e_id, eng_b.e_id) and torch.equal(eng_a.last_update, eng_b.last_update)
    assert torch.equal(eng_a.t_ring, eng_b.t_ring)


def test_grouped_host_feeding_equals_resident_path():
    """stage_group + train_group_logged (three batches per pinned H2D copy, per captured graph and per loss
    read-back; six slots in two groups) followed by single-step leftovers == the resident-array path."""
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 26          # 8 groups of 3 + 2 single steps
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 17, True, lr=1e-6)
    _, eng_g, _ = _setup(N, De, D, K, B, B * steps, 17, True, lr=1e-6)
    la = [float(eng_a.train_step(from_device=True)) for _ in range(steps)]
    G = eng_g.group_size
    ngroups = (steps + G - 1) // G
    host = torch.zeros((ngroups + 1, eng_g.group_nbytes()), dtype=torch.uint8).pin_memory()
    for g in range(ngroups):
        batches = []
        for i in range(G):
            s = min(g * G + i, steps - 1)                   # the tail group is padded with a repeated batch
            sl = slice(s * B, (s + 1) * B)
            batches.append((ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl], ev["msg"][sl]))
        eng_g.pack_host_group(host[g], batches)
    got = []
    eng_g.stage_group(host[0], ahead=False)
    full = steps // G
    for g in range(full):
        eng_g.stage_group(host[g + 1])
        prev = eng_g.train_group_logged()
        if prev is not None:
            got += prev
    got += eng_g.flush_group_losses()
    for _ in range(steps - full * G):                        # leftovers: already staged by the last stage_group
        got.append(float(eng_g.train_step(from_device=False, lookahead=True, _capture=False)))
    assert len(got) == steps and eng_g.events_done == eng_a.events_done
    for s, (a, b) in enumerate(zip(la, got)):
        assert abs(a - b) < 1e-4, (s, a, b)
    torch.testing.assert_close(eng_a.memory, eng_g.memory, rtol=1e-2, atol=1e-3)
    assert torch.equal(eng_a.e_id, eng_g.e_id) and torch.equal(eng_a.t_ring, eng_g.t_ring)
    assert torch.equal(eng_a.last_update, eng_g.last_update)


def test_log_growth_drops_graphs_and_keeps_training():
    """re-allocating the message-store log used to leave captured graphs replaying against freed
    buffers.  _grow_log drops every graph; the run continues bit-compatibly with an engine that never grew."""
    N, De, D, K, B, steps = 300, 8, 16, 4, 32, 60
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 23, True, lr=1e-6)
    _, eng_b, _ = _setup(N, De, D, K, B, B * steps, 23, True, lr=1e-6)
    for s in range(steps):
        if s == 44:      # nine slots x (three eager warm-up calls + the capture): every slot has its graph by now
            assert any(k[0] == "train" for k in eng_b._graphs if isinstance(k, tuple))
            old_ptr = eng_b.store.ev_src.data_ptr()
            eng_b._grow_log(4 * eng_b.store.capacity)
            assert eng_b._graphs == {} and eng_b.store.ev_src.data_ptr() != old_ptr
        la, lb = float(eng_a.train_step()), float(eng_b.train_step())
        assert abs(la - lb) < 1e-4, (s, la, lb)
    eng_b.check_device_errors()
    assert torch.equal(eng_a.last_update, eng_b.last_update) and torch.equal(eng_a.e_id, eng_b.e_id)
    torch.testing.assert_close(eng_a.memory, eng_b.memory, rtol=1e-2, atol=1e-3)


def test_stepping_past_the_resident_events_is_refused():
    """edge features are events['msg'][e_id]; an e_id past the resident arrays must not be read."""
    from tgn_b200 import _cabi
    N, De, D, K, B = 300, 8, 16, 4, 32
    _, eng, ev = _setup(N, De, D, K, B, B * 3, 29, False)
    for _ in range(3):
        eng.train_step()
    with pytest.raises(_cabi.TgnError, match="resident events"):
        eng.train_step()
    with pytest.raises(_cabi.TgnError, match="resident events"):
        eng.eval_batch(ev["src"][:B], ev["dst"][:B], ev["neg"][:B].view(B, 1), ev["t"][:B], ev["msg"][:B])


def test_device_error_word_flags_out_of_range_event_rows():
    """tgn_edge_attr_ld with a row id outside [0, num_events): zero row + TGN_DEVERR_EVENT_RANGE, no fault."""
    from tgn_b200 import _cabi, ops
    L = _cabi.lib()
    L.tgn_device_errors(1)
    E, De, Dt, n_ev = 6, 3, 4, 10
    lu = torch.zeros(4, dtype=torch.long, device=DEV)
    nbr = torch.zeros(E, dtype=torch.long, device=DEV)
    t_ev = torch.arange(n_ev, dtype=torch.long, device=DEV)
    msg = torch.ones((n_ev, De), device=DEV)
    rows = torch.tensor([0, 1, 2, 10, 3, -1], dtype=torch.long, device=DEV)
    tw, tb = torch.ones(Dt, device=DEV), torch.zeros(Dt, device=DEV)
    ea = torch.full((E, 8), 7.0, device=DEV)
    p = ops._p
    _cabi.check(L.tgn_edge_attr_ld(p(lu), p(nbr), p(t_ev), p(msg), p(rows), n_ev, E, None, De, Dt, p(tw), p(tb), 8,
                                   p(ea), None, None, ops._stream()))
    torch.cuda.synchronize()
    assert L.tgn_device_errors(0) & 2
    assert L.tgn_device_errors(1) & 2 and L.tgn_device_errors(0) == 0
    assert float(ea[3].abs().sum()) == 0.0 and float(ea[5].abs().sum()) == 0.0
    assert torch.equal(ea[0, Dt:Dt + De], torch.ones(De, device=DEV))


def test_tail_batch_engine_shares_state():
    """A second step geometry on the same training state (TGNEngine(share=...)): the events that do not fill a
    whole batch are trained as one shorter batch, like the reference's last DataLoader batch (ADVICE r1:
    run_tgn.py --engine used to drop them, shifting every later e_id)."""
    from tgn_b200.engine import TGNEngine
    N, De, D, K, B, tail = 400, 12, 32, 5, 50, 17
    steps = 4
    E = B * steps + tail
    ref, eng, ev = _setup(N, De, D, K, B, E, 31, True)
    tail_eng = TGNEngine(N, De, D, K, tail, device=DEV, lr=1e-3, dropout=0.0, use_graph=False, share=eng)
    loader = orc.TorchNeighborLoader(N, K)
    opt = torch.optim.Adam(orc.model_parameters(ref), lr=1e-3)
    strip = lambda sd: {k: v for k, v in sd.items() if k not in ("memory", "last_update", "_assoc")}
    lo = 0
    for s in range(steps + 1):
        n = B if s < steps else tail
        sl = slice(lo, lo + n)
        lo += n
        if s < steps:
            loss = float(eng.train_step(from_device=True))
        else:
            eng.handover()
            loss = float(tail_eng.train_step(from_device=True))
        loss_ref = orc.train_step(ref, loader, opt, ev["src"][sl], ev["dst"][sl], ev["neg"][sl], ev["t"][sl],
                                  ev["msg"][sl], ev["t"], ev["msg"], dropout=False)
        assert abs(loss - loss_ref) < 1e-4 * max(1.0, abs(loss_ref)), (s, loss, loss_ref)
        eng.load_state(strip(ref["memory"].state_dict()), ref["gnn"].state_dict(), ref["link_pred"].state_dict())
    assert eng.events_done == E and tail_eng.ring_pos == E and float(eng.adam_step_dev) == steps + 1
    assert torch.equal(eng.last_update.cpu(), ref["memory"].last_update)
    assert torch.equal(eng.e_id.cpu(), loader.e_id)
    torch.testing.assert_close(eng.memory.cpu(), ref["memory"].memory.detach(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("Q,dense", [(20, False), (37, True)])
def test_sharded_embedding_eval_equals_replicated_eval(Q, dense):
    """eval_batch_dp (roots dealt to ranks, decoder-projected rows gathered, own negative columns scored) with a
    single rank == eval_batch: same scores, same TGB counts, same state afterwards.  `dense` = the candidates
    cover the graph several times, every node is treated as a root (no unique / relabel of the candidate list)."""
    from tgn_b200 import synth
    N, De, D, K, B, steps = (60 if dense else 300), 8, 16, 5, 40, 5
    _, eng_a, ev = _setup(N, De, D, K, B, B * steps, 21, False)
    _, eng_b, _ = _setup(N, De, D, K, B, B * steps, 21, False)
    for e in (eng_a, eng_b):
        e.flush_to_eval()
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        src, dst, t, msg = ev["src"][sl], ev["dst"][sl], ev["t"][sl], ev["msg"][sl]
        neg = torch.from_numpy(synth.eval_negatives(src.numpy(), dst.numpy(), N, Q, seed=s, dst_lo=N // 2))
        pos_a, _, gt_a, ge_a = eng_a.eval_batch(src, dst, neg, t, msg, want_neg_scores=False)
        pos_b, gt_b, ge_b = eng_b.eval_batch_dp(src, dst, neg, t, msg, 0, 1)
        assert eng_b._eval_ctx_dp(B, Q, 0, 1).dense == dense
        torch.testing.assert_close(pos_a, pos_b, rtol=1e-5, atol=1e-6)
        assert torch.equal(gt_a, gt_b) and torch.equal(ge_a, ge_b), s
        torch.cuda.synchronize()
        assert torch.equal(eng_a.last_update, eng_b.last_update) and torch.equal(eng_a.e_id, eng_b.e_id)
        torch.testing.assert_close(eng_a.memory, eng_b.memory, rtol=1e-5, atol=1e-6)
    eng_b.check_device_errors()


@pytest.mark.parametrize("dropout", [0.0, 0.3])
def test_attention_fused_decoder_equals_separate_launches(dropout):
    """tgn_dec_attn_fused (attention forward of every (event, endpoint) occurrence computed inside the decoder
    launch) against tgn_attn_core_fwd + tgn_dec_fused on the same batches, from the same weights: same losses,
    same memory / last_update / ring, same weights after Adam -- with attention dropout as well (the Philox mask is
    keyed by (edge, head, step), so both routes draw the same mask).  Duplicated nodes inside a batch (hub
    sources) exercise the per-occurrence recomputation."""
    N, De, D, K, B, steps = 300, 6, 32, 5, 40, 8
    engs = []
    for fused in (True, False):
        _, eng, _ = _setup(N, De, D, K, B, B * (steps + 2), 33, False)
        eng.dropout = dropout
        assert eng.fused_attn_dec
        eng.fused_attn_dec = fused
        engs.append(eng)
    a, b = engs
    for s in range(steps):
        la = float(a.train_step(from_device=True))
        lb = float(b.train_step(from_device=True))
        assert abs(la - lb) <= 2e-6 * max(1.0, abs(lb)), (s, la, lb)
        # the two routes' gradients agree to rounding (the atomic accumulation order differs from run to run on
        # either route); Adam amplifies that on near-zero gradients, so the weights are re-synced after the check
        torch.testing.assert_close(a.flat_grad, b.flat_grad, rtol=1e-3, atol=2e-5)
        torch.testing.assert_close(a.memory, b.memory, rtol=1e-5, atol=2e-6)
        for x, y in ((a.flat, b.flat), (a.exp_avg, b.exp_avg), (a.exp_avg_sq, b.exp_avg_sq)):
            x.copy_(y)
    assert torch.equal(a.last_update, b.last_update) and torch.equal(a.e_id, b.e_id)
    a.check_device_errors()


def test_peer_bcast_writes_a_block_into_every_table():
    """tgn_peer_bcast with the 'peers' being three tables of this device: the block lands at the same offset of
    each, nothing else is touched (the multi-rank use is covered by tests/dist_eval_check.py on 2 GPUs)."""
    import ctypes
    from tgn_b200 import _cabi
    L = _cabi.lib()
    tables = [torch.full((64,), float(i), device=DEV) for i in range(3)]
    src = torch.arange(16, dtype=torch.float32, device=DEV) + 100
    ptrs = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in tables])
    _cabi.check(L.tgn_peer_bcast(src.data_ptr(), 64, ptrs, 32 * 4, 3, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for i, t in enumerate(tables):
        assert torch.equal(t[32:48], src) and bool((t[:32] == i).all()) and bool((t[48:] == i).all())
    rc = L.tgn_peer_bcast(src.data_ptr(), 60, ptrs, 0, 3, None)          # not a multiple of 16 bytes
    assert rc == _cabi.TGN_EINVAL


def test_group_views_layout_equals_pack_host_group():
    """The numpy views a host loader fills (TGNEngine.group_views) address exactly the layout pack_host_group
    writes -- the layout stage_group copies to the device."""
    _, eng, ev = _setup(100, 3, 16, 5, 8, 8 * 12, 5, False)
    G, B = eng.group_size, 8
    a = torch.zeros(eng.group_nbytes(), dtype=torch.uint8)
    b = torch.zeros(eng.group_nbytes(), dtype=torch.uint8)
    batches = [tuple(ev[k][i * B:(i + 1) * B] for k in ("src", "dst", "neg", "t", "msg")) for i in range(G)]
    eng.pack_host_group(a, batches)
    for (vs, vd, vn, vt, vm), (s_, d_, n_, t_, m_) in zip(eng.group_views(b), batches):
        vs[:], vd[:], vn[:], vt[:], vm[:] = s_.numpy(), d_.numpy(), n_.numpy(), t_.numpy(), m_.numpy()
    assert torch.equal(a, b)
