"""Fine-tune step parity (BASELINE config 4): loss, KL, every parameter
gradient and the post-step parameters of bayeslms_b200.trainer.FineTuner against autograd through the
CPU oracle (oracle.finetune_loss, fp32) on the same batch and the same injected noise, followed by
torch's clip_grad_norm_ + SGD(momentum=0.9).  Dropout is off on both sides (p = 0 step).

Tolerances (precise bf16x3 mode): loss / CE within 2e-4 absolute, KL within 1e-4 relative (north_star);
each gradient tensor within 2e-3 of its largest reference magnitude (+1e-7); parameters after the
step within 2e-3 * lr * max|update|."""
import pytest
import torch

from oracle import bayeslm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
V, D, NHEAD, FF = 500, 128, 2, 256   # head_dim 64 (the training attention kernels' size)


def _build(family, nlayers, dropout=0.0, **flag):
    from bayeslms_b200 import model as M
    torch.manual_seed(7)
    if family == "bayes_tm":
        net = M.BayesTransformerModel(V, D, NHEAD, FF, nlayers, dropout, True, flag["bayes_pos"])
    elif family == "gauss_tm":
        net = M.GaussTransformerModel(V, D, NHEAD, FF, nlayers, dropout, True, flag["gauss_pos"])
    elif family == "std_tm":
        net = M.TransformerModel(V, D, NHEAD, FF, nlayers, dropout, flag.get("activation", "gelu"), True)
    else:
        net = M.VTransformerModel(V, D, NHEAD, FF, nlayers, dropout, True, flag["v_pos"])
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
        for n, p in net.named_parameters():   # non-trivial LayerNorm / bias values so every gradient path is exercised
            if "norm" in n or n.endswith(".bias"):
                p.add_(torch.randn_like(p) * 0.05)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.Config(family=family, ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=nlayers, **flag)
    return net, sd, cfg


def _batch_away_from_the_relu_kink(sd, cfg, x, T, B, g, margin=1e-4):
    """ReLU's derivative is discontinuous: a pre-activation within the arithmetic error of zero (~1e-5 in precise mode)
    takes the other branch on the other side and moves a whole row of a weight gradient (seen: |z| = 2e-6, 15 % of one
    tensor's largest entry).  That is a property of the function, not of either implementation, so the parity batch
    is redrawn until no FFN pre-activation of the oracle's forward pass lies within ``margin`` of the kink."""
    import torch.nn.functional as F
    for _ in range(200):
        zs = []
        real = F.linear

        def spy(inp, w, b=None):
            out = real(inp, w, b)
            if w.shape[0] == cfg.nhid and w.shape[1] == cfg.ninp:
                zs.append(out.detach())
            return out
        F.linear = spy
        try:
            with torch.no_grad():
                O.transformer_forward(sd, x, cfg)
        finally:
            F.linear = real
        assert zs, "no FFN pre-activation seen"
        if min(float(z.abs().min()) for z in zs) > margin:
            return x
        x = torch.randint(0, V, (T, B), generator=g)
    raise AssertionError("no batch found away from the ReLU kink")


def _oracle_step(sd, cfg, x, y, eps, kl_scale, lr, clip, masks=None):
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k != "pos_encoder.pe" else v)
            for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]   # tied
    loss, ce, kl = O.finetune_loss(leaf, x, y, cfg, eps, kl_scale, masks=masks)
    loss.backward()
    params = [(k, v) for k, v in leaf.items() if k != "decoder.weight" and isinstance(v, torch.Tensor) and v.requires_grad]
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in params}
    with_grad = [v for _, v in params if v.grad is not None]
    torch.nn.utils.clip_grad_norm_(with_grad, clip)
    opt = torch.optim.SGD(with_grad, lr=lr, momentum=0.9, weight_decay=0)
    opt.step()
    new = {k: v.detach().clone() for k, v in params}
    return float(loss.detach()), float(ce.detach()), float(kl.detach()) if torch.is_tensor(kl) else float(kl), grads, new


CASES = [
    ("bayes_tm", 2, {"bayes_pos": "FFN"}, 12),
    ("bayes_tm", 2, {"bayes_pos": "MHA"}, 12),
    ("bayes_tm", 2, {"bayes_pos": "EMB"}, 12),
    ("gauss_tm", 2, {"gauss_pos": 3}, 12),
    ("gauss_tm", 2, {"gauss_pos": 1}, 12),
    ("v_tm", 4, {"v_pos": 3}, 100),
    ("v_tm", 3, {"v_pos": 1}, 100),
    ("std_tm", 2, {}, 12),
    ("std_tm", 2, {"activation": "relu"}, 12),       # TransformerModel's constructor default (model.py:124)
]


@pytest.mark.parametrize("family,nlayers,flag,T", CASES)
@pytest.mark.parametrize("sampled", [True, False])
def test_finetune_step_matches_oracle_autograd(family, nlayers, flag, T, sampled):
    from bayeslms_b200.trainer import FineTuner
    if family == "v_tm" and not sampled:
        pytest.skip("the variational layer always adds its noise in training")
    if family == "std_tm" and sampled:
        pytest.skip("the baseline Transformer has nothing to sample")
    net, sd, cfg = _build(family, nlayers, **flag)
    B, kl_scale, lr, clip = 4, 0.37, 0.05, 0.25
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    if flag.get("activation") == "relu":
        x = _batch_away_from_the_relu_kink(sd, cfg, x, T, B, g)
    eps = None
    if sampled:
        if family == "v_tm":
            kinds = O.tm_layer_kinds(cfg)
            eps = {f"layer{i}": torch.randn(T, B, D, generator=g) * 0.1 for i, k in enumerate(kinds) if k == "v"}
        else:
            eps = O.draw_eps(sd, cfg, 99)
    if family == "gauss_tm":
        net.transformerlayers[0].gpnn.sample = bool(sampled)
    want_loss, want_ce, want_kl, want_g, want_p = _oracle_step(sd, cfg, x, y.view(-1), eps, kl_scale, lr, clip)

    net = net.to(DEV).train()
    ft = FineTuner(net, lr, clip=clip, prec="bf16x3")
    loss, ce, kl = ft.forward_backward(x.to(DEV), y.to(DEV), kl_scale, eps=eps)
    assert abs(float(ce) - want_ce) < 2e-4, (float(ce), want_ce)
    assert abs(float(kl) - want_kl) <= 1e-4 * abs(want_kl) + 1e-7, (float(kl), want_kl)
    assert abs(float(loss) - want_loss) < 2e-4 + 1e-4 * abs(want_loss), (float(loss), want_loss)
    bad = []
    for name, ref in want_g.items():
        got = ft.g[name].detach().cpu()
        tol = 2e-3 * ref.abs().max().item() + 1e-7
        err = (got - ref).abs().max().item()
        if not err <= tol:
            bad.append((name, err, ref.abs().max().item()))
    assert not bad, bad
    ft.apply_gradients()
    named = dict(net.named_parameters())
    for name, ref in want_p.items():
        got = named[name].detach().cpu()
        upd = (ref - sd[name]).abs().max().item()
        assert (got - ref).abs().max().item() <= 2e-3 * upd + 1e-7, (name, (got - ref).abs().max().item(), upd)


def _tm_masks(cfg, T, B, p_model, g):
    """The step's nn.Dropout draws as multiplier tensors in the oracle's layout; layer 0 of the Bayesian FFN / MHA
    models drops with the hard-coded 0.2 (model.py:1202,1207)."""
    def mk(shape, p):
        return (torch.rand(shape, generator=g) >= p).float() / (1.0 - p)
    masks = {"pe": mk((T, B, D), p_model)}
    for i, kind in enumerate(O.tm_layer_kinds(O.canonical({}, cfg)[1])):
        p = 0.2 if (i == 0 and kind in ("bayes_ffn", "bayes_mha")) else p_model
        masks[f"layer{i}"] = {"attn": mk((B * NHEAD, T, T), p), "d1": mk((T, B, D), p), "ffn": mk((T, B, FF), p),
                              "d2": mk((T, B, D), p)}
    return masks


DROP_CASES = [("bayes_tm", 2, {"bayes_pos": "FFN"}, 12), ("bayes_tm", 2, {"bayes_pos": "MHA"}, 10),
              ("bayes_tm", 2, {"bayes_pos": "EMB"}, 12), ("gauss_tm", 2, {"gauss_pos": 3}, 12),
              ("v_tm", 4, {"v_pos": 3}, 100), ("std_tm", 2, {}, 12)]


@pytest.mark.parametrize("family,nlayers,flag,T", DROP_CASES)
def test_finetune_step_with_dropout_matches_oracle_autograd(family, nlayers, flag, T):
    """The reference trains with dropout 0.2 (run_nnlm_ami_tm.sh:96).  Every nn.Dropout draw of the training forward
    (embedding + positional output, attention probabilities, both residual branches, FFN activation; the oracle's
    sites are pinned on the reference by tests/golden/dropout_*.pt) is injected as a multiplier tensor on both sides;
    loss, KL and every gradient must match autograd through the oracle at the usual tolerances."""
    from bayeslms_b200.trainer import FineTuner
    p_model = 0.3
    net, sd, cfg = _build(family, nlayers, dropout=p_model, **flag)
    assert net.pos_encoder.p == p_model
    if family == "bayes_tm" and flag["bayes_pos"] in ("FFN", "MHA"):
        assert net.transformerlayers[0].p_drop == 0.2 and net.transformerlayers[1].p_drop == p_model
    B, kl_scale, lr, clip = 4, 0.37, 0.05, 0.25
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    if family == "v_tm":
        eps = {f"layer{i}": torch.randn(T, B, D, generator=g) * 0.1
               for i, k in enumerate(O.tm_layer_kinds(cfg)) if k == "v"}
    elif family == "std_tm":
        eps = None
    else:
        eps = O.draw_eps(sd, cfg, 99)
    if family == "gauss_tm":
        net.transformerlayers[0].gpnn.sample = True
    masks = _tm_masks(cfg, T, B, p_model, g)
    want_loss, want_ce, want_kl, want_g, _ = _oracle_step(sd, cfg, x, y.view(-1), eps, kl_scale, lr, clip, masks)
    nodrop_loss = _oracle_step(sd, cfg, x, y.view(-1), eps, kl_scale, lr, clip)[0]
    assert abs(want_loss - nodrop_loss) > 1e-3                     # the masks matter
    ft = FineTuner(net.to(DEV).train(), lr, clip=clip, prec="bf16x3")
    loss, ce, kl = ft.forward_backward(x.to(DEV), y.to(DEV), kl_scale, eps=eps, masks=masks)
    assert abs(float(ce) - want_ce) < 2e-4, (float(ce), want_ce)
    assert abs(float(kl) - want_kl) <= 1e-4 * abs(want_kl) + 1e-7, (float(kl), want_kl)
    assert abs(float(loss) - want_loss) < 2e-4 + 1e-4 * abs(want_loss), (float(loss), want_loss)
    bad = []
    for name, ref in want_g.items():
        got = ft.g[name].detach().cpu()
        tol = 2e-3 * ref.abs().max().item() + 1e-7
        err = (got - ref).abs().max().item()
        if not err <= tol:
            bad.append((name, err, ref.abs().max().item()))
    assert not bad, bad


def test_philox_dropout_equals_its_exported_masks_and_the_oracle():
    """--dropout 0.2 with device Philox masks (what train.py runs): the step equals, bit for bit, the same step with
    the multipliers exported from the same Philox streams and injected -- and therefore the oracle with those masks.
    A second seed gives another loss; p = 0 modules give the dropout-free step."""
    from bayeslms_b200 import engine, ops, trainer as TR
    from bayeslms_b200.trainer import FineTuner
    p_model, T, B, seed = 0.25, 12, 4, 4242
    net, sd, cfg = _build("bayes_tm", 2, dropout=p_model, bayes_pos="FFN")
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
    l_a = float(ft.forward_backward(x.to(DEV), y.to(DEV), 0.37, seed=seed)[0])
    g_a = {k: v.clone() for k, v in ft.g.items()}
    dev = torch.device(DEV)

    def export(p, tid, shape):           # the kernel's own layout is sequence-major: [B, T, w]
        n = 1
        for s_ in shape:
            n *= s_
        m, _ = ops.dropout(None, ops.Drop(p, seed=seed, stream_id=engine._stream_id(tid, 0)), n=n, device=dev)
        return m.view(shape)

    masks = {"pe": export(p_model, TR._TID_DROP_PE, (B, T, D)).permute(1, 0, 2).contiguous().cpu()}
    for li in range(2):
        p = 0.2 if li == 0 else p_model
        base = TR._TID_DROP_LAYER + 4 * li
        masks[f"layer{li}"] = {
            "attn": export(p, base + 0, (B * NHEAD, T, T)).cpu(),
            "d1": export(p, base + 1, (B, T, D)).permute(1, 0, 2).contiguous().cpu(),
            "ffn": export(p, base + 2, (B, T, FF)).permute(1, 0, 2).contiguous().cpu(),
            "d2": export(p, base + 3, (B, T, D)).permute(1, 0, 2).contiguous().cpu()}
    # weight noise of the same step, exported the same way (Philox stream of the FFN weight, sample 0)
    w = net.transformerlayers[0].linear2.weight_lgstd
    eps = {"layer0": ops.philox_normal(seed, engine._stream_id(engine._TID["ffn_w2"], 0), w.numel(), dev).view(w.shape).cpu()}
    l_b = float(ft.forward_backward(x.to(DEV), y.to(DEV), 0.37, eps=eps, masks=masks)[0])
    assert l_a == l_b
    for k, v in ft.g.items():
        if k in ("encoder.weight", "decoder.weight"):
            assert torch.allclose(g_a[k], v, rtol=0, atol=1e-6 * float(v.abs().max()))
        else:
            assert torch.equal(g_a[k], v), k
    want_loss, _, _, want_g, _ = _oracle_step(sd, cfg, x, y.view(-1), eps, 0.37, 0.05, 0.25, masks)
    assert abs(l_a - want_loss) < 2e-4 + 1e-4 * abs(want_loss)
    for name, ref in want_g.items():
        assert (ft.g[name].detach().cpu() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-7, name
    assert float(ft.forward_backward(x.to(DEV), y.to(DEV), 0.37, seed=seed + 1)[0]) != l_a


def test_finetune_step_fast_mode_gradients_are_aligned():
    """bf16 operands / fp32 accumulate (the bench mode): loss within 2e-2, every sizeable gradient
    tensor within 1 % (relative L2) of the fp32 autograd gradient."""
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build("v_tm", 4, v_pos=3)
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (100, 4), generator=g)
    y = torch.randint(0, V, (100, 4), generator=g)
    eps = {f"layer{i}": torch.randn(100, 4, D, generator=g) * 0.1 for i in (0, 1)}
    want_loss, _, _, want_g, _ = _oracle_step(sd, cfg, x, y.view(-1), eps, 0.37, 0.05, 0.25)
    ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16")
    loss, _, _ = ft.forward_backward(x.to(DEV), y.to(DEV), 0.37, eps=eps)
    assert abs(float(loss) - want_loss) < 2e-2
    for name, ref in want_g.items():
        if ref.norm() < 1e-6:
            continue
        got = ft.g[name].detach().cpu()
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 2e-2, (name, rel)


def test_finetune_reduces_the_loss_and_philox_noise_is_reproducible():
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build("v_tm", 4, v_pos=3)
    net = net.to(DEV).train()
    ft = FineTuner(net, 0.05, clip=0.25, prec="bf16x3")
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (100, 4), generator=g).to(DEV)
    y = torch.randint(0, V, (100, 4), generator=g).to(DEV)
    l0 = float(ft.forward_backward(x, y, 0.01, seed=5)[0])
    g0 = {k: v.clone() for k, v in ft.g.items()}
    l0b = float(ft.forward_backward(x, y, 0.01, seed=5)[0])
    assert l0 == l0b                                          # same seed -> same noise -> same loss
    for k, v in ft.g.items():
        if k in ("encoder.weight", "decoder.weight"):         # embedding scatter: fp32 atomics, order varies
            assert torch.allclose(g0[k], v, rtol=0, atol=1e-6 * float(v.abs().max()))
        else:
            assert torch.equal(g0[k], v), k                   # every other reduction has a fixed order
    losses = [float(ft.step(x, y, 0.01, seed=5 + i)[0]) for i in range(8)]
    assert losses[-1] < l0 - 0.05, (l0, losses)
    # the rescoring path sees the updated weights (cached bf16 copies were invalidated)
    from bayeslms_b200.engine import PackedBatch
    net.eval()
    s = net.score(PackedBatch.from_lists([x[:, 0].tolist()], [y[:, 0].tolist()], DEV), prec="bf16x3")
    assert torch.isfinite(s).all()


@pytest.mark.parametrize("family,nlayers,flag,T,dropout", [
    ("v_tm", 4, {"v_pos": 3}, 100, 0.0), ("v_tm", 4, {"v_pos": 3}, 100, 0.2),
    ("bayes_tm", 2, {"bayes_pos": "FFN"}, 16, 0.2), ("bayes_tm", 2, {"bayes_pos": "MHA"}, 16, 0.0),
    ("bayes_tm", 2, {"bayes_pos": "EMB"}, 16, 0.2), ("gauss_tm", 2, {"gauss_pos": 3}, 16, 0.0)])
def test_captured_step_equals_eager_step(family, nlayers, flag, T, dropout):
    """The CUDA-graph replay (noise refreshed outside the graph with the same Philox streams, dropout keyed by a device
    seed word) takes the same steps as the eager path with seed-driven noise -- for every family that samples, so a
    sampled tensor without a static noise buffer (which would silently train on the posterior mean) cannot recur."""
    from bayeslms_b200.trainer import FineTuner
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (T, 4), generator=g).to(DEV)
    y = torch.randint(0, V, (T, 4), generator=g).to(DEV)
    outs = []
    for captured in (False, True):
        net, _, _ = _build(family, nlayers, dropout=dropout, **flag)
        if family == "gauss_tm":
            net.transformerlayers[0].gpnn.sample = True
        ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
        if captured:
            ft.capture(T, 4, 0.01)
        losses = []
        for i in range(3):
            out = ft.step_captured(x, y, 5 + i) if captured else ft.step(x, y, 0.01, seed=5 + i)
            losses.append(float(out[0]))
        outs.append((losses, ft.flat_p.clone()))
    (l0, p0), (l1, p1) = outs
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (l0, l1)
    assert (p0 - p1).abs().max().item() < 1e-6
    assert len(set(l0)) == 3              # the noise really changes from step to step
    if family != "v_tm":                 # ... and the step is not the posterior-mean step
        net, _, _ = _build(family, nlayers, dropout=dropout, **flag)
        ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
        mean_loss = float(ft.forward_backward(x, y, 0.01)[0])
        assert abs(mean_loss - l1[0]) > 1e-4, (mean_loss, l1)


# ------------------------------------------------------------------------------ LSTM families
def _build_lstm(pos, H=128, dropout=0.0):
    from bayeslms_b200 import model as M
    torch.manual_seed(7)
    if pos == "std":
        net = M.RNNModel("LSTM", V, H, H, 2, dropout, True)
        cfg = O.Config(family="std_lstm", ntoken=V, ninp=H, nhid=H, nlayers=2)
    else:
        net = M.BayesRNNModel("LSTM", V, H, H, 2, dropout, True, pos)
        cfg = O.Config(family="bayes_lstm", ntoken=V, ninp=H, nhid=H, nlayers=2, bayes_pos=pos)
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net, sd, cfg


def _oracle_lstm_step(sd, cfg, x, y, eps, kl_scale, hidden, masks=None):
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v)
            for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]
    loss, ce, kl = O.finetune_loss(leaf, x, y, cfg, eps, kl_scale, hidden=hidden, masks=masks)
    loss.backward()
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in leaf.items() if k != "decoder.weight" and v.requires_grad}
    with torch.no_grad():
        _, new_hidden = O.rnn_forward(sd, x, hidden, cfg, eps, masks=masks)
    return float(loss.detach()), float(ce.detach()), float(kl.detach()) if torch.is_tensor(kl) else float(kl), grads, new_hidden


@pytest.mark.parametrize("pos,sampled,dropout", [(3, True, 0.0), (1, True, 0.0), (4, False, 0.0), (0, False, 0.0),
                                                 ("std", False, 0.0), (3, True, 0.3), ("std", False, 0.3)])
def test_lstm_finetune_step_matches_oracle_autograd(pos, sampled, dropout):
    """Bayes-LSTM (train.py:319-340 with model.rnn.kl_divergence()): loss, KL, every gradient and the
    carried-out hidden state against autograd through the oracle's LSTM, from a non-zero carried-in state.
    Same tolerances as the Transformer cases."""
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build_lstm(pos, dropout=dropout)
    T, B, H, kl_scale = 9, 4, 128, 0.37
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    hidden = (torch.randn(2, B, H, generator=g) * 0.3, torch.randn(2, B, H, generator=g) * 0.3)
    eps = O.draw_eps(sd, cfg, 99) if sampled else None
    masks = None
    if dropout:      # self.drop on the embedding and on the LSTM output (model.py:218,220), injected on both sides
        masks = {k: (torch.rand(T, B, H, generator=g) >= dropout).float() / (1.0 - dropout) for k in ("emb", "out")}
    want_loss, want_ce, want_kl, want_g, want_hidden = _oracle_lstm_step(sd, cfg, x, y.view(-1), eps, kl_scale, hidden, masks)

    ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
    loss, ce, kl = ft.forward_backward(x.to(DEV), y.to(DEV), kl_scale, eps=eps, masks=masks,
                                       hidden=(hidden[0].to(DEV), hidden[1].to(DEV)))
    assert abs(float(ce) - want_ce) < 2e-4, (float(ce), want_ce)
    assert abs(float(kl) - want_kl) <= 1e-4 * abs(want_kl) + 1e-7, (float(kl), want_kl)
    assert abs(float(loss) - want_loss) < 2e-4 + 1e-4 * abs(want_loss), (float(loss), want_loss)
    for got, want in zip(ft.hidden, want_hidden):
        assert (got.cpu() - want).abs().max().item() < 1e-4
    bad = []
    for name, ref in want_g.items():
        got = ft.g[name].detach().cpu()
        tol = 2e-3 * ref.abs().max().item() + 1e-7
        err = (got - ref).abs().max().item()
        if not err <= tol:
            bad.append((name, err, ref.abs().max().item()))
    assert not bad, bad


def test_lstm_finetune_reduces_the_loss_with_carried_state():
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build_lstm(3)
    ft = FineTuner(net.to(DEV).train(), 0.5, clip=0.25, prec="bf16")
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (20, 8), generator=g).to(DEV)
    y = torch.randint(0, V, (20, 8), generator=g).to(DEV)
    l0 = float(ft.step(x, y, 0.01, seed=1)[0])
    hidden = None
    for i in range(12):
        last = float(ft.step(x, y, 0.01, seed=2 + i, hidden=hidden)[0])
        hidden = ft.hidden
    assert last < l0 - 0.05, (l0, last)


# ------------------------------------------------------------------------------ GP-LSTM / Variational-LSTM cells
def _build_cells(family, pos, H=128, dropout=0.0):
    from bayeslms_b200 import model as M
    torch.manual_seed(7)
    if family == "v_lstm":
        net = M.VariationalRNNModel("LSTM", V, H, H, 2, dropout, True, pos)
        cfg = O.Config(family="v_lstm", ntoken=V, ninp=H, nhid=H, nlayers=2, v_pos=pos)
    else:
        net = M.GaussRNNModel("LSTM", V, H, H, 2, dropout, True, pos)
        cfg = O.Config(family="gauss_lstm", ntoken=V, ninp=H, nhid=H, nlayers=2, gauss_pos=pos)
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
        for n, p in net.named_parameters():      # the cells initialise their biases to zero: make every path non-trivial
            if "bias" in n:
                p.add_(torch.randn_like(p) * 0.05)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net, sd, cfg


@pytest.mark.parametrize("family,pos,sampled,dropout", [
    ("v_lstm", "11", True, 0.0), ("v_lstm", "01", True, 0.0), ("v_lstm", "11", False, 0.0), ("v_lstm", "11", True, 0.3),
    ("gauss_lstm", "31", False, 0.0), ("gauss_lstm", "31", True, 0.0), ("gauss_lstm", "333", True, 0.0),
    ("gauss_lstm", "3330", True, 0.0), ("gauss_lstm", "23", True, 0.3), ("gauss_lstm", "12", False, 0.0),
    ("gauss_lstm", "4131", True, 0.0),
    # gate type 7: the GP unit replaces the input product (one mixture GEMM over all steps + blm_gp3_bwd)
    ("gauss_lstm", "73", True, 0.0), ("gauss_lstm", "71", False, 0.0), ("gauss_lstm", "732", True, 0.0),
    ("gauss_lstm", "7333", True, 0.3),
    # gate types 5 / 6: the GP unit on the cell state / in place of the recurrent product, one mixture GEMM per step
    ("gauss_lstm", "53", True, 0.0), ("gauss_lstm", "61", False, 0.0), ("gauss_lstm", "6353", True, 0.0),
    ("gauss_lstm", "532", True, 0.3), ("gauss_lstm", "62", True, 0.0)])
def test_cell_families_finetune_step_matches_oracle_autograd(family, pos, sampled, dropout):
    """GaussRNNModel / VariationalRNNModel fine-tune step (train.py:319-377): loss, the KL of the GP units / of the VNNs,
    every gradient and the carried-out state against autograd through the oracle's training-mode cells (pinned on the
    reference by tests/golden/*_train_*.pt), with injected noise: per-step (T, 1, H) VNN noise, or the GP units'
    (coef, weights, bias) draws with GPNN.sample set; and with injected dropout masks."""
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build_cells(family, pos, dropout=dropout)
    T, B, H, kl_scale = 9, 4, 128, 0.37
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    hidden = (torch.randn(2, B, H, generator=g) * 0.3, torch.randn(2, B, H, generator=g) * 0.3)
    eps = None
    if sampled:
        eps = {}
        if family == "v_lstm":
            for mi in range(2):
                if pos[mi] == "1":
                    eps[f"cell{mi}"] = torch.randn(T, 1, H, generator=g) * 0.1
        else:
            for mi, (kind, _, gt) in enumerate(O.gp_lstm_layout(pos)):
                if kind == "gp":
                    net.rnn.rnn[mi].gpnn.sample = True
                    pre, e = f"rnn.rnn.{mi}.gpnn.", {}
                    if gt in (1, 3):
                        e["coef"] = torch.randn(sd[pre + "coef_mean"].shape, generator=g)
                    if gt in (2, 3):
                        e["weights"] = torch.randn(sd[pre + "weights_mean"].shape, generator=g)
                        e["bias"] = torch.randn(sd[pre + "bias_mean"].shape, generator=g)
                    eps[f"cell{mi}"] = e
    masks = None
    if dropout:
        masks = {k: (torch.rand(T, B, H, generator=g) >= dropout).float() / (1.0 - dropout) for k in ("emb", "out")}
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]
    loss, ce, kl = O.finetune_loss(leaf, x, y.view(-1), cfg, eps, kl_scale, hidden=hidden, masks=masks)
    loss.backward()
    want_g = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
              for k, v in leaf.items() if k != "decoder.weight" and v.requires_grad}
    with torch.no_grad():
        _, want_hidden = O.cell_rnn_forward(sd, x, hidden, cfg, eps, masks)

    ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
    l, c, k = ft.forward_backward(x.to(DEV), y.to(DEV), kl_scale, eps=eps, masks=masks,
                                  hidden=(hidden[0].to(DEV), hidden[1].to(DEV)))
    assert abs(float(c) - float(ce)) < 2e-4, (float(c), float(ce))
    assert abs(float(k) - float(kl)) <= 1e-4 * abs(float(kl)) + 1e-7, (float(k), float(kl))
    assert abs(float(l) - float(loss)) < 2e-4 + 1e-4 * abs(float(loss)), (float(l), float(loss))
    for got, want in zip(ft.hidden, want_hidden):
        assert (got.cpu() - want).abs().max().item() < 1e-4
    bad = []
    for name, ref in want_g.items():
        got = ft.g[name].detach().cpu()
        tol = 2e-3 * ref.abs().max().item() + 1e-7
        err = (got - ref).abs().max().item()
        if not err <= tol:
            bad.append((name, err, ref.abs().max().item()))
    assert not bad, bad
    # the accessors train.py:360-377 calls after the forward
    if family == "v_lstm":
        acc = sum(float(net.rnn.rnn[mi].vnn.kl_divergence()) for mi in range(2) if pos[mi] == "1")
    else:
        acc = sum(float(c.gpnn.kl_divergence()) for c in net.rnn.rnn if hasattr(c, "gpnn"))
    assert abs(acc - float(kl)) <= 1e-4 * abs(float(kl)) + 1e-7


@pytest.mark.parametrize("family,pos", [("v_lstm", "11"), ("gauss_lstm", "31")])
def test_cell_families_training_reduces_the_loss(family, pos):
    from bayeslms_b200.trainer import FineTuner
    net, _, _ = _build_cells(family, pos, dropout=0.1)
    ft = FineTuner(net.to(DEV).train(), 0.5, clip=0.25, prec="bf16")
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (20, 8), generator=g).to(DEV)
    y = torch.randint(0, V, (20, 8), generator=g).to(DEV)
    l0 = float(ft.step(x, y, 0.01, seed=1)[0])
    hidden = None
    for i in range(12):
        last = float(ft.step(x, y, 0.01, seed=2 + i, hidden=hidden)[0])
        hidden = ft.hidden
    assert last < l0 - 0.05, (l0, last)
