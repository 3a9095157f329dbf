"""Host-side training loop (bayeslms_b200/train.py = the epoch loop of steps/pytorchnn/train.py).
CPU: corpus / batchify / get_batch semantics (data.py:14-54, train.py:164-183, 293-297).
GPU: evaluate() against the oracle's restatement of train.py:440-457, the lr-halving / best-checkpoint
schedule (train.py:470-512), and an end-to-end run that learns a synthetic Markov corpus."""
import math
import os

import pytest
import torch

from oracle import bayeslm_oracle as O

DEV = "cuda:0"


def _write_corpus(d, n_words=60, n_train=6000, seed=0):
    g = torch.Generator().manual_seed(seed)
    words = ["<s>", "<unk>"] + [f"w{i:03d}" for i in range(n_words)]
    with open(os.path.join(d, "words.txt"), "w") as f:
        for i, w in enumerate(words):
            f.write(f"{w} {i}\n")
    # first-order Markov chain with a sharp transition table: learnable well below ln(V)
    nxt = torch.randint(0, n_words, (n_words, 3), generator=g)
    for name, n in (("train", n_train), ("valid", n_train // 8), ("test", n_train // 8)):
        cur, lines, line = 0, [], []
        for _ in range(n):
            cur = int(nxt[cur, int(torch.randint(0, 3, (1,), generator=g))])
            line.append(f"w{cur:03d}")
            if len(line) == 12:
                lines.append(" ".join(line))
                line = []
        lines.append("zzz_oov " + " ".join(line))      # one OOV word -> <unk>
        with open(os.path.join(d, name + ".txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
    return words


def test_corpus_batchify_get_batch(tmp_path):
    from bayeslms_b200 import train as T
    d = str(tmp_path)
    words = _write_corpus(d, n_train=500)
    c = T.Corpus(d)
    assert len(c) == len(words)
    first = open(os.path.join(d, "train.txt")).readline().split()
    want = [c.word2idx[w] for w in first] + [c.word2idx["<s>"]]
    assert c.train[:len(want)].tolist() == want
    assert c.word2idx["<unk>"] in c.train.tolist()                      # the OOV word
    data = torch.arange(23)
    b = T.batchify(data, 4)                                             # 5 rows x 4 columns, 3 elements trimmed
    assert b.shape == (5, 4) and b[:, 0].tolist() == [0, 1, 2, 3, 4] and b[0].tolist() == [0, 5, 10, 15]
    x, y = T.get_batch(b, 0, 3)
    assert x.tolist() == b[0:3].tolist() and y.tolist() == b[1:4].reshape(-1).tolist()
    x, y = T.get_batch(b, 3, 3)                                         # ragged last batch: len(source) - 1 - i rows
    assert x.shape == (1, 4) and y.tolist() == b[4].tolist()
    ref_path = "/root/reference/steps/pytorchnn/data.py"
    if os.path.exists(ref_path):                                        # build container only: the reference itself
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_data", ref_path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        rc = ref.Corpus(d)
        assert torch.equal(rc.train, c.train) and torch.equal(rc.valid, c.valid) and torch.equal(rc.test, c.test)
        assert rc.dictionary.word2idx == c.word2idx


def _oracle_evaluate(sd, cfg, source, seq_len):
    """train.py:440-457 on the oracle."""
    total = 0.0
    hidden = O.init_hidden(cfg, source.size(1)) if cfg.family.endswith("lstm") else None
    with torch.no_grad():
        for i in range(0, source.size(0) - 1, seq_len):
            n = min(seq_len, len(source) - 1 - i)
            data, targets = source[i:i + n], source[i + 1:i + 1 + n].reshape(-1)
            if hidden is None:
                out = O.transformer_forward(sd, data, cfg)
            else:
                out, hidden = O.rnn_forward(sd, data, hidden, cfg)
            total += len(data) * float(torch.nn.functional.cross_entropy(out.view(-1, out.shape[-1]), targets))
    return total / (len(source) - 1)


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["bayes_tm", "bayes_lstm"])
def test_evaluate_matches_oracle(family):
    from bayeslms_b200 import model as M, train as T
    torch.manual_seed(5)
    V = 300
    if family == "bayes_tm":
        net = M.BayesTransformerModel(V, 128, 2, 256, 2, 0.0, True, "FFN")
        cfg = O.Config(family="bayes_tm", bayes_pos="FFN", ntoken=V, ninp=128, nhead=2, nhid=256, nlayers=2)
    else:
        net = M.BayesRNNModel("LSTM", V, 128, 128, 2, 0.0, True, 3)
        cfg = O.Config(family="bayes_lstm", ntoken=V, ninp=128, nhid=128, nlayers=2, bayes_pos=3)
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    source = T.batchify(torch.randint(0, V, (20 * 47 + 3,)), 20)        # 47 rows: 35 + ragged 11
    want = _oracle_evaluate(sd, cfg, source, 35)
    got = T.evaluate(net.to(DEV), source.to(DEV), 35, prec="bf16x3")
    assert abs(got - want) < 1e-4, (got, want)


@pytest.mark.gpu
def test_schedule_halves_lr_and_reloads_best(tmp_path, monkeypatch):
    """train.py:494-512 with scripted validation losses: save on improvement; otherwise lr /= 2, momentum reset,
    best weights reloaded; stop at the `patience`-th reload; best weights loaded at the end."""
    from bayeslms_b200 import model as M, train as T
    torch.manual_seed(1)
    net = M.BayesTransformerModel(100, 128, 2, 256, 1, 0.0, True, "FFN").to(DEV)
    vals = iter([3.0, 3.5, 2.5, 2.6, 2.7, 9.9, 9.9])
    snaps = []

    def fake_epoch(ft, *a, **k):
        with torch.no_grad():
            ft.flat_p.add_(0.01)           # "training" moves every parameter
            ft.flat_v.fill_(1.0)
        snaps.append(ft.flat_p.clone())
        return 0.0

    monkeypatch.setattr(T, "train_epoch", fake_epoch)
    monkeypatch.setattr(T, "evaluate", lambda *a, **k: next(vals))
    save = str(tmp_path / "m.pt")
    dummy = torch.zeros(40, 4, dtype=torch.int64, device=DEV)
    best, hist = T.fit(net, dummy, dummy, lr=0.8, epochs=7, seq_len=10, clip=0.25, save=save, patience=3, log=lambda *a: None)
    assert best == 2.5
    assert [h["reloaded"] for h in hist] == [False, True, False, True, True]      # early stop at the 3rd reload
    assert [h["lr"] for h in hist] == [0.8, 0.4, 0.4, 0.2, 0.1]
    ckpt = torch.load(save, map_location="cpu", weights_only=True)
    for k, v in net.state_dict().items():                                           # best = the epoch-3 weights
        assert torch.equal(v.cpu(), ckpt[k]), k
    named = dict(net.named_parameters())
    assert named["encoder.weight"].data_ptr() >= 0 and torch.equal(named["encoder.weight"].detach().cpu(), ckpt["encoder.weight"])


@pytest.mark.gpu
@pytest.mark.parametrize("model,unc,flags", [("Transformer", "Bayesian", ["--T_bayes_pos", "FFN", "--nhead", "2"]),
                                             ("LSTM", "Bayesian", ["--L_bayes_pos", "3"]),
                                             ("LSTM", "Gaussian", ["--L_gauss_pos", "63"])])
def test_training_run_learns_markov_corpus(tmp_path, model, unc, flags, capsys):
    """End to end through main(): corpus files -> 3 epochs -> checkpoint; validation perplexity far below uniform."""
    from bayeslms_b200 import train as T
    d = str(tmp_path)
    words = _write_corpus(d, n_train=12000)
    save = os.path.join(d, "model.pt")
    rc = T.main(["--data", d, "--model", model, "--uncertainty", unc, *flags, "--emsize", "128", "--nhid",
                 "256" if model == "Transformer" else "128", "--nlayers", "2", "--lr", "1.0" if model == "LSTM" else "0.3",
                 "--batch-size", "16", "--seq_len", "20", "--epochs", "3", "--clip", "0.5", "--save", save, "--tied",
                 "--precision", "bf16", "--log-interval", "0"])
    assert rc == 0 and os.path.exists(save)
    out = capsys.readouterr().out
    vals = [float(l.split("valid loss")[1].split("|")[0]) for l in out.splitlines() if "valid loss" in l]
    assert len(vals) == 3 and vals[-1] < vals[0] and vals[-1] < math.log(len(words)) - 1.0, vals
    ckpt = torch.load(save, map_location="cpu", weights_only=True)
    assert "encoder.weight" in ckpt and ckpt["encoder.weight"].shape[0] == len(words)
