"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``/root/reference/steps/pytorchnn/model.py`` read-only, builds every model
family of the hot path at a small size with the reference's own initialisers under
``torch.manual_seed(1111)`` (train.py:88), and records weights, inputs and the outputs of
  * eval-mode forward (posterior mean; what the scorer runs, score.py:225),
  * train-mode forward under a fixed generator seed with every Dropout neutralised
    (the parameter noise the reference draws is then reproducible: SURVEY.md 8c),
  * the family's ``kl_divergence``,
  * the restated scorer loop (score.py:206-280 is not runnable on CPU as shipped because of
    its unconditional .cuda() calls) driven through the reference modules.
The fixtures are small (< 3 MB in total) and committed; the GPU box has no /root/reference.
"""
import io
import contextlib
import math
import os
import sys

import torch
import torch.nn as nn

REF = "/root/reference/steps/pytorchnn"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    import model as ref  # noqa: E402

V, NHEAD = 60, 4
D, FF, H = 32, 64, 32
TM = dict(ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=3)
LSTM = dict(ntoken=V, ninp=H, nhid=H, nlayers=2)
NOISE_SEED = 4242


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def neutralise_dropout(m):
    for mod in m.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
    for mod in m.modules():
        if hasattr(mod, "dropout") and isinstance(getattr(mod, "dropout"), float):
            mod.dropout = 0.0


def perturb(m):
    """Make zero-initialised biases / unit LayerNorm weights non-trivial so every term is exercised."""
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("bias") or "bias_mean" in name:
                p.add_(torch.empty_like(p).uniform_(-0.1, 0.1, generator=g))
            if "norm" in name and name.endswith("weight"):
                p.add_(torch.empty_like(p).uniform_(-0.1, 0.1, generator=g))


def tokens(T, B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, V, (T, B), generator=g)


def _strip(obj):
    # pos_encoder.pe is a deterministic 5000-row buffer (model.py:97-103): not stored, the
    # oracle and the product both rebuild it and test_oracle_golden checks the formula.
    if isinstance(obj, dict):
        return {k: _strip(v) for k, v in obj.items() if k != "pos_encoder.pe"}
    return obj


def save(name, obj):
    torch.save(_strip(obj), os.path.join(HERE, name))
    print("wrote", name)


def tm_case(name, ctor, cfg, train_T=7, sample_attr=None):
    torch.manual_seed(1111)
    m = quiet(ctor)
    perturb(m)
    neutralise_dropout(m)
    x = tokens(9, 3, 1)
    m.eval()
    with torch.no_grad():
        out_eval = m(x)
    rec = {"cfg": cfg, "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
           "x": x, "logits_eval": out_eval}
    # train mode with seeded parameter noise
    m.train()
    if sample_attr is not None:
        sample_attr(m)
    xt = tokens(train_T, 2, 2)
    torch.manual_seed(NOISE_SEED)
    with torch.no_grad():
        out_train = m(xt)
    rec.update({"x_train": xt, "noise_seed": NOISE_SEED, "logits_train": out_train})
    return m, rec


def main():
    # ---------------- Bayesian Transformer: FFN / MHA / EMB / none
    for pos in ("FFN", "MHA", "EMB", "none"):
        cfg = dict(family="bayes_tm", bayes_pos=pos, **TM)
        m, rec = tm_case(f"bayes_tm_{pos}", lambda: ref.BayesTransformerModel(V, D, NHEAD, FF, 3, 0.5, True, pos), cfg)
        with torch.no_grad():
            if pos == "FFN":
                rec["kl"] = m.transformerlayers[0].linear2.kl_divergence()
            elif pos == "MHA":
                rec["kl"] = m.transformerlayers[0].self_attn.o_net.kl_divergence()
            elif pos == "EMB":
                rec["kl"] = m.embed_kl_divergence()
        save(f"bayes_tm_{pos}.pt", rec)

    # ---------------- GP Transformer types 0..3 (sample False = as shipped, and True)
    for g in (0, 1, 2, 3):
        cfg = dict(family="gauss_tm", gauss_pos=g, **TM)

        def on(m):
            m.transformerlayers[0].gpnn.sample = True

        m, rec = tm_case(f"gauss_tm_{g}", lambda: ref.GaussTransformerModel(V, D, NHEAD, FF, 3, 0.5, True, g), cfg,
                         sample_attr=on)
        with torch.no_grad():
            kl = m.transformerlayers[0].gpnn.kl_divergence()
            rec["kl"] = kl if torch.is_tensor(kl) else torch.tensor(float(kl))
        # as shipped: sample=False in train mode -> deterministic
        m.transformerlayers[0].gpnn.sample = False
        torch.manual_seed(NOISE_SEED)
        with torch.no_grad():
            rec["logits_train_nosample"] = m(rec["x_train"])
        save(f"gauss_tm_{g}.pt", rec)

    # ---------------- Variational Transformer: eval only is runnable in the reference
    for v in (0, 1, 2, 3):
        cfg = dict(family="v_tm", v_pos=v, **TM)
        torch.manual_seed(1111)
        m = quiet(lambda: ref.VTransformerModel(V, D, NHEAD, FF, 4, 0.5, True, v))
        cfg["nlayers"] = 4
        perturb(m)
        m.eval()
        x = tokens(9, 3, 1)
        with torch.no_grad():
            out = m(x)
        save(f"v_tm_{v}.pt", {"cfg": cfg, "state_dict": {k: t.clone() for k, t in m.state_dict().items()},
                              "x": x, "logits_eval": out, "n_layers_built": len(m.transformerlayers)})

    # ---------------- Bayesian LSTM pos 0..4
    for pos in (0, 1, 3, 4):
        cfg = dict(family="bayes_lstm", bayes_pos=pos, **LSTM)
        torch.manual_seed(1111)
        m = quiet(lambda: ref.BayesRNNModel("LSTM", V, H, H, 2, 0.5, True, pos))
        perturb(m)
        neutralise_dropout(m)
        m.eval()
        x = tokens(8, 3, 3)
        g = torch.Generator().manual_seed(5)
        h0 = (torch.randn(2, 3, H, generator=g) * 0.3, torch.randn(2, 3, H, generator=g) * 0.3)
        with torch.no_grad():
            out_eval, hid_eval = m(x, h0)
        rec = {"cfg": cfg, "state_dict": {k: t.clone() for k, t in m.state_dict().items()}, "x": x,
               "h0": h0, "logits_eval": out_eval, "hidden_eval": hid_eval}
        if pos:
            m.train()
            torch.manual_seed(NOISE_SEED)
            with torch.no_grad():
                out_train, hid_train = m(x, h0)
                rec["kl"] = m.rnn.kl_divergence()
            rec.update({"noise_seed": NOISE_SEED, "logits_train": out_train, "hidden_train": hid_train})
        save(f"bayes_lstm_{pos}.pt", rec)

    # ---------------- scorer loop through the reference modules (Transformer + LSTM)
    vocab_words = ["<s>", "<unk>"] + [f"w{i:03d}" for i in range(V - 2)]
    rng = torch.Generator().manual_seed(11)
    nbest_lines = []
    for u in range(4):
        n_hyp = [3, 1, 4, 2][u]
        for n in range(n_hyp):
            L = int(torch.randint(0 if (u == 1) else 1, 9, (1,), generator=rng))
            words = [vocab_words[int(torch.randint(2, V, (1,), generator=rng))] for _ in range(L)]
            if u == 2 and n == 1:
                words.append("oov-word")
            nbest_lines.append(f"utt-a_{u}-{n + 1} " + " ".join(words) if words else f"utt-a_{u}-{n + 1}")
    crit = nn.CrossEntropyLoss()

    def ref_loop(m, is_rnn):
        # restatement of score.py:206-280 with .cuda() removed; arithmetic is the reference module's
        vocab = {w: i for i, w in enumerate(vocab_words)}
        from collections import OrderedDict
        nbest = OrderedDict()
        for line in nbest_lines:
            line = line.strip()
            try:
                key, hyp = line.split(" ", 1)
            except ValueError:
                key, hyp = line, " "
            nbest.setdefault(key.rsplit("-", 1)[0], []).append(hyp)
        m.eval()
        scores = []
        hidden = m.init_hidden(1) if is_rnn else None
        with torch.no_grad():
            for key, hyps in nbest.items():
                cached = []
                for hyp in hyps:
                    inp = [vocab.get(w, vocab["<unk>"]) for w in ("<s> " + hyp).split()]
                    tgt = [vocab.get(w, vocab["<unk>"]) for w in (hyp + " <s>").split()]
                    data = torch.LongTensor(inp).view(-1, 1)
                    target = torch.LongTensor(tgt).view(-1)
                    if is_rnn:
                        out, nh = m(data, hidden)
                        cached.append(nh)
                    else:
                        out = m(data)
                    loss = crit(out.view(-1, V), target)
                    scores.append(len(inp) * loss.item())
                if is_rnn:
                    hidden = cached[0]
        return scores

    torch.manual_seed(1111)
    m_tm = quiet(lambda: ref.BayesTransformerModel(V, D, NHEAD, FF, 3, 0.5, True, "FFN"))
    perturb(m_tm)
    torch.manual_seed(1111)
    m_rnn = quiet(lambda: ref.BayesRNNModel("LSTM", V, H, H, 2, 0.5, True, 3))
    perturb(m_rnn)
    save("pos_encoding.pt", {"pe_first_40": ref.PositionalEncoding(D, 0.0).pe[:40].clone()})
    save("scorer_loop.pt", {
        "vocab_words": vocab_words, "nbest_lines": nbest_lines,
        "tm_cfg": dict(family="bayes_tm", bayes_pos="FFN", **TM),
        "tm_state_dict": {k: t.clone() for k, t in m_tm.state_dict().items()},
        "tm_scores": ref_loop(m_tm, False),
        "lstm_cfg": dict(family="bayes_lstm", bayes_pos=3, **LSTM),
        "lstm_state_dict": {k: t.clone() for k, t in m_rnn.state_dict().items()},
        "lstm_scores": ref_loop(m_rnn, True),
    })


if __name__ == "__main__":
    main()
