"""GPU parity of the LSTM path (persistent recurrence kernel + session scheduler) against the
reference-generated golden fixtures and the CPU oracle."""
import math
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import bayeslm_oracle as O
from tests.util import load_golden_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("pos", [0, 1, 3, 4, "std"])
def test_forward_matches_reference_golden(golden, pos):
    rec = golden("std_lstm.pt" if pos == "std" else f"bayes_lstm_{pos}.pt")   # std = RNNModel (nn.LSTM keys)
    net = load_golden_model(rec, DEV)
    h0 = tuple(t.to(DEV) for t in rec["h0"])
    out, (h, c) = net(rec["x"].to(DEV), h0)
    assert (out.cpu() - rec["logits_eval"]).abs().max().item() < 1e-3
    assert (h.cpu() - rec["hidden_eval"][0]).abs().max().item() < 1e-4
    assert (c.cpu() - rec["hidden_eval"][1]).abs().max().item() < 1e-4


@pytest.mark.parametrize("pos", [1, 3, 4])
def test_kl_matches_reference_golden(golden, pos):
    rec = golden(f"bayes_lstm_{pos}.pt")
    net = load_golden_model(rec, DEV)
    kl, ref = float(net.rnn.kl_divergence()), float(rec["kl"])
    assert abs(kl - ref) <= 1e-4 * abs(ref)


def _sessions_from(nbest, vocab):
    from bayeslms_b200.scorer import ids_for
    return [[[ids_for(h, vocab) for h in hyps] for hyps in nbest.values()]]


def test_session_scoring_matches_reference_loop(golden):
    """The reference loop (hidden carried from hypothesis #0 of the previous utterance) through the
    reference modules, recorded in scorer_loop.pt, against the two-phase batched scheduler."""
    from bayeslms_b200.scorer import Rescorer
    rec = golden("scorer_loop.pt")
    vocab = {w: i for i, w in enumerate(rec["vocab_words"])}
    nbest = OrderedDict()
    for line in rec["nbest_lines"]:
        key, _, hyp = line.partition(" ")
        nbest.setdefault(key.rsplit("-", 1)[0], []).append(hyp or " ")
    net = load_golden_model({"cfg": rec["lstm_cfg"], "state_dict": rec["lstm_state_dict"]}, DEV)
    for prec, tol in (("bf16x3", 1e-3), ("bf16", 5e-2)):
        got = Rescorer(net, prec=prec).score_sessions(_sessions_from(nbest, vocab))
        assert np.abs(got - np.asarray(rec["lstm_scores"])).max() < tol, prec


def test_logit_interpolation_matches_reference_loop(golden, tmp_path):
    """LSTM logit interpolation (score.py:418-447): the second, untied plain LSTM LM carries its own hidden chain; its
    states join the vocabulary sweep as extra K segments.  Against the reference-generated loop scores, through the
    Rescorer and through the CLI (--interpolation_flag 1 --inter_path ...)."""
    from bayeslms_b200 import scorer as S
    from bayeslms_b200.scorer import Rescorer
    rec, loop = golden("interp_lstm.pt"), golden("scorer_loop.pt")
    vocab = {w: i for i, w in enumerate(loop["vocab_words"])}
    nbest = OrderedDict()
    for line in loop["nbest_lines"]:
        key, _, hyp = line.partition(" ")
        nbest.setdefault(key.rsplit("-", 1)[0], []).append(hyp or " ")
    net1 = load_golden_model({"cfg": rec["cfg1"], "state_dict": rec["sd1"]}, DEV)
    from bayeslms_b200 import model as M
    c2 = rec["cfg2"]
    net2 = M.BayesRNNModel("LSTM", c2["ntoken"], c2["ninp"], c2["nhid"], 2, 0.5, False, 0)     # untied, score.py:422
    net2.load_state_dict(rec["sd2"])
    net2 = net2.to(DEV).eval()
    want = np.asarray(rec["scores"])
    for prec, tol in (("bf16x3", 1e-3), ("bf16", 5e-2)):
        got = Rescorer(net1, prec=prec, inter_model=net2, inter_alpha=rec["alpha"]).score_sessions(_sessions_from(nbest, vocab))
        assert np.abs(got - want).max() < tol, (prec, np.abs(got - want).max())
    # K = 2 posterior samples of model 1 mixed with the (mean) second model, against the oracle
    sd1, cfg1 = rec["sd1"], O.Config(rec["cfg1"])
    eps_list = [O.draw_eps(sd1, cfg1, 40 + k) for k in range(2)]
    ref = O.compute_scores(nbest, vocab, sd1, cfg1, eps_list=eps_list, sd2=rec["sd2"], cfg2=O.Config(c2), alpha=0.6)
    ref = np.asarray([s for items in ref.values() for _, s in items])
    got = Rescorer(net1, prec="bf16x3", eps_list=eps_list, inter_model=net2, inter_alpha=0.6).score_sessions(_sessions_from(nbest, vocab))
    assert np.abs(got - ref).max() < 1e-3
    # the CLI
    vp, npth, ck1, ck2, out = (tmp_path / n for n in ("words.txt", "words_text", "m1.pt", "m2.pt", "lmwt.nn"))
    vp.write_text("".join(f"{w} {i}\n" for i, w in enumerate(loop["vocab_words"])))
    npth.write_text("\n".join(loop["nbest_lines"]) + "\n")
    torch.save(rec["sd1"], ck1), torch.save(rec["sd2"], ck2)
    rc = S.main(["--nbest-list", str(npth), "--outfile", str(out), "--vocabulary", str(vp), "--model-path", str(ck1),
                 "--model", "LSTM", "--emsize", str(c2["ninp"]), "--nhid", str(c2["nhid"]), "--nlayers", "2",
                 "--uncertainty", "Bayesian", "--L_bayes_pos", "3", "--interpolation_flag", "1", "--inter_path", str(ck2),
                 "--inter_alpha", str(rec["alpha"])])
    assert rc == 0
    got = np.asarray([float(l.split()[1]) for l in out.read_text().splitlines()])
    assert np.abs(got - want).max() < 1e-3 + 5e-5


@pytest.mark.parametrize("pos", [1, 3])
def test_injected_eps_and_k_samples_match_oracle(golden, pos):
    from bayeslms_b200.scorer import Rescorer
    rec = golden(f"bayes_lstm_{pos}.pt")
    cfg, sd = O.Config(rec["cfg"]), rec["state_dict"]
    net = load_golden_model(rec, DEV)
    # train-mode golden of the reference: same eps -> same logits / hidden
    eps = O.draw_eps(sd, cfg, rec["noise_seed"])
    rs = np.random.RandomState(5)
    vocab = {"<s>": 0, "<unk>": 1, **{f"w{i}": i for i in range(2, cfg.ntoken)}}
    nbest = OrderedDict((f"u{u}", [" ".join(f"w{rs.randint(2, cfg.ntoken)}" for _ in range(rs.randint(0, 9))) or " "
                                   for _ in range(rs.randint(1, 5))]) for u in range(5))
    for eps_list in ([eps], [O.draw_eps(sd, cfg, 300 + k) for k in range(3)]):
        want = O.compute_scores(nbest, vocab, sd, cfg, eps_list=eps_list)
        flat = np.asarray([s for items in want.values() for _, s in items])
        got = Rescorer(net, prec="bf16x3", eps_list=eps_list).score_sessions(_sessions_from(nbest, vocab))
        assert np.abs(got - flat).max() < 1e-3


def test_many_rows_ragged_lengths_two_sessions():
    """H=256, 2 sessions x 6 utterances x up to 70 hypotheses (> 128 rows per batch, ragged lengths)."""
    from bayeslms_b200 import model as M
    from bayeslms_b200.scorer import Rescorer
    torch.manual_seed(5)
    V, H = 500, 256
    net = M.BayesRNNModel("LSTM", V, H, H, 2, 0.5, True, 3)
    with torch.no_grad():
        net.decoder.bias.uniform_(-0.1, 0.1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.Config(family="bayes_lstm", bayes_pos=3, ntoken=V, ninp=H, nhid=H, nlayers=2)
    net = net.to(DEV).eval()
    rs = np.random.RandomState(1)
    vocab = {"<s>": 0, "<unk>": 1, **{f"w{i}": i for i in range(2, V)}}
    sessions_txt = []
    for s in range(2):
        nb = OrderedDict((f"s{s}u{u}", [" ".join(f"w{rs.randint(2, V)}" for _ in range(rs.randint(0, 20))) or " "
                                        for _ in range(rs.randint(1, 70))]) for u in range(6 - s))
        sessions_txt.append(nb)
    want = np.concatenate([[sc for items in O.compute_scores(nb, vocab, sd, cfg).values() for _, sc in items]
                           for nb in sessions_txt])
    sessions = [_sessions_from(nb, vocab)[0] for nb in sessions_txt]
    got = Rescorer(net, prec="bf16x3", max_tokens=3000).score_sessions(sessions)
    assert np.abs(got - want).max() < 1e-3
    got_fast = Rescorer(net, prec="bf16", max_tokens=3000).score_sessions(sessions)
    assert np.abs(got_fast - want).max() < 5e-2 + 2e-3 * np.abs(want).max()


CELL_FIXTURES = ["gauss_lstm_31", "gauss_lstm_23", "gauss_lstm_13", "gauss_lstm_43", "gauss_lstm_333",
                 "gauss_lstm_3330", "gauss_lstm_51", "gauss_lstm_62", "gauss_lstm_73", "gauss_lstm_6350", "v_lstm_11", "v_lstm_01"]


@pytest.mark.parametrize("name", CELL_FIXTURES)
def test_gp_and_variational_cells_match_reference_golden(golden, name):
    """GP-LSTM (gate replaced by the GP unit, doubled bias_ih) and Variational-LSTM cells,
    eval mode, against the reference's own outputs."""
    rec = golden(name + ".pt")
    net = load_golden_model(rec, DEV)
    h0 = tuple(t.to(DEV) for t in rec["h0"])
    out, (h, c) = net(rec["x"].to(DEV), h0)
    assert (out.cpu() - rec["logits_eval"]).abs().max().item() < 1e-3
    assert (h.cpu() - rec["hidden_eval"][0]).abs().max().item() < 1e-4
    assert (c.cpu() - rec["hidden_eval"][1]).abs().max().item() < 1e-4
    if "kl" in rec and rec["kl"]:
        cells = [m for m in net.rnn.rnn if hasattr(m, "gpnn")]
        for cell, ref in zip(cells, rec["kl"]):
            if ref:
                assert abs(float(cell.gpnn.kl_divergence()) - ref) <= 1e-4 * abs(ref)


@pytest.mark.parametrize("name", ["gauss_lstm_31", "gauss_lstm_3330", "gauss_lstm_73", "gauss_lstm_6350", "v_lstm_11"])
def test_cell_models_session_scoring_matches_oracle(golden, name):
    """The session scheduler (hidden carry through hypothesis #0, ragged lock-step batches) with GP / V cells
    against the oracle's restatement of the reference scoring loop."""
    from bayeslms_b200.scorer import Rescorer, ids_for
    rec = golden(name + ".pt")
    loop = golden("scorer_loop.pt")
    vocab = {w: i for i, w in enumerate(loop["vocab_words"])}
    nbest = OrderedDict()
    for line in loop["nbest_lines"]:
        key, _, hyp = line.partition(" ")
        nbest.setdefault(key.rsplit("-", 1)[0], []).append(hyp or " ")
    cfg = O.Config(rec["cfg"])
    want = O.compute_scores(nbest, vocab, rec["state_dict"], cfg)
    want = np.asarray([s for items in want.values() for _, s in items])
    net = load_golden_model(rec, DEV)
    sessions = [[[ids_for(h, vocab) for h in hyps] for hyps in nbest.values()]]
    got = Rescorer(net, prec="bf16x3").score_sessions(sessions)
    assert np.abs(got - want).max() < 1e-3
