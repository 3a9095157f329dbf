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


def _build(family, nlayers, **flag):
    from bayeslms_b200 import model as M
    torch.manual_seed(7)
    if family == "bayes_tm":
        net = M.BayesTransformerModel(V, D, NHEAD, FF, nlayers, 0.0, True, flag["bayes_pos"])
    elif family == "gauss_tm":
        net = M.GaussTransformerModel(V, D, NHEAD, FF, nlayers, 0.0, True, flag["gauss_pos"])
    else:
        net = M.VTransformerModel(V, D, NHEAD, FF, nlayers, 0.0, True, flag["v_pos"])
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
        for n, p in net.named_parameters():   # non-trivial LayerNorm / bias values so every gradient path is exercised
            if "norm" in n or n.endswith(".bias"):
                p.add_(torch.randn_like(p) * 0.05)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.Config(family=family, ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=nlayers, **flag)
    return net, sd, cfg


def _oracle_step(sd, cfg, x, y, eps, kl_scale, lr, clip):
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k != "pos_encoder.pe" else v)
            for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]   # tied
    loss, ce, kl = O.finetune_loss(leaf, x, y, cfg, eps, kl_scale)
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
]


@pytest.mark.parametrize("family,nlayers,flag,T", CASES)
@pytest.mark.parametrize("sampled", [True, False])
def test_finetune_step_matches_oracle_autograd(family, nlayers, flag, T, sampled):
    from bayeslms_b200.trainer import FineTuner
    if family == "v_tm" and not sampled:
        pytest.skip("the variational layer always adds its noise in training")
    net, sd, cfg = _build(family, nlayers, **flag)
    B, kl_scale, lr, clip = 4, 0.37, 0.05, 0.25
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
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


def test_captured_step_equals_eager_step():
    """The CUDA-graph replay (noise refreshed outside the graph with the same Philox streams) takes the
    same steps as the eager path with seed-driven noise."""
    from bayeslms_b200.trainer import FineTuner
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, V, (100, 4), generator=g).to(DEV)
    y = torch.randint(0, V, (100, 4), generator=g).to(DEV)
    outs = []
    for captured in (False, True):
        net, _, _ = _build("v_tm", 4, v_pos=3)
        ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
        if captured:
            ft.capture(100, 4, 0.01)
        losses = []
        for i in range(3):
            out = ft.step_captured(x, y, 5 + i) if captured else ft.step(x, y, 0.01, seed=5 + i)
            losses.append(float(out[0]))
        outs.append((losses, ft.flat_p.clone()))
    (l0, p0), (l1, p1) = outs
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-5, (l0, l1)
    assert (p0 - p1).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------ LSTM families
def _build_lstm(pos, H=128):
    from bayeslms_b200 import model as M
    torch.manual_seed(7)
    net = M.BayesRNNModel("LSTM", V, H, H, 2, 0.0, True, pos)
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.Config(family="bayes_lstm", ntoken=V, ninp=H, nhid=H, nlayers=2, bayes_pos=pos)
    return net, sd, cfg


def _oracle_lstm_step(sd, cfg, x, y, eps, kl_scale, hidden):
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v)
            for k, v in sd.items() if k != "decoder.weight"}
    leaf["decoder.weight"] = leaf["encoder.weight"]
    loss, ce, kl = O.finetune_loss(leaf, x, y, cfg, eps, kl_scale, hidden=hidden)
    loss.backward()
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in leaf.items() if k != "decoder.weight" and v.requires_grad}
    with torch.no_grad():
        _, new_hidden = O.rnn_forward(sd, x, hidden, cfg, eps)
    return float(loss.detach()), float(ce.detach()), float(kl.detach()) if torch.is_tensor(kl) else float(kl), grads, new_hidden


@pytest.mark.parametrize("pos,sampled", [(3, True), (1, True), (4, False), (0, False)])
def test_lstm_finetune_step_matches_oracle_autograd(pos, sampled):
    """Bayes-LSTM (train.py:319-340 with model.rnn.kl_divergence()): loss, KL, every gradient and the
    carried-out hidden state against autograd through the oracle's LSTM, from a non-zero carried-in state.
    Same tolerances as the Transformer cases."""
    from bayeslms_b200.trainer import FineTuner
    net, sd, cfg = _build_lstm(pos)
    T, B, H, kl_scale = 9, 4, 128, 0.37
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, V, (T, B), generator=g)
    y = torch.randint(0, V, (T, B), generator=g)
    hidden = (torch.randn(2, B, H, generator=g) * 0.3, torch.randn(2, B, H, generator=g) * 0.3)
    eps = O.draw_eps(sd, cfg, 99) if sampled else None
    want_loss, want_ce, want_kl, want_g, want_hidden = _oracle_lstm_step(sd, cfg, x, y.view(-1), eps, kl_scale, hidden)

    ft = FineTuner(net.to(DEV).train(), 0.05, clip=0.25, prec="bf16x3")
    loss, ce, kl = ft.forward_backward(x.to(DEV), y.to(DEV), kl_scale, eps=eps,
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
