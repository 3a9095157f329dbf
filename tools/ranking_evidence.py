#!/usr/bin/env python
"""Ranking evidence on a PEAKED model: fine-tune a Bayesian Transformer (BASELINE layer sizes)
on a synthetic Markov corpus for a few hundred steps, then score chain-based n-best lists in fast (bf16) and precise
(bf16x3) mode and with the CPU oracle, and report per-utterance full-ranking / 1-best / pair-order agreement, without
and with the stage-7 combination (graph + w nn + (1-w) oldlm, lmrescore_nbest_pytorchnn_cuda.sh:221-229).

    python tools/ranking_evidence.py [--vocab 2000] [--layers 2] [--steps 300] [--utts 60] [--nbest 20] > profiles/...json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(vocab=2000, layers=2, steps=300, utts=60, nbest=20, lr=0.1, clip=1.0, d=512, nhead=8, ff=4096, with_oracle=True,
        branching=3, dev="cuda:0", log=lambda *a: None):
    from bayeslms_b200 import model as M, synth, train as T
    from bayeslms_b200.engine import PackedBatch
    from bayeslms_b200.trainer import FineTuner
    torch.manual_seed(1111)
    mk = synth.make_markov(vocab, branching)
    net = M.BayesTransformerModel(vocab, d, nhead, ff, layers, 0.0, True, "FFN").to(dev)
    ft = FineTuner(net, lr, clip=clip, prec="bf16")
    ids = torch.from_numpy(mk.stream(32 * (100 * min(steps, 400) + 1)))
    t0 = time.time()
    losses = T.train_steps(ft, ids, 32, 100, steps)
    torch.cuda.synchronize()
    log(f"trained {steps} steps in {time.time() - t0:.1f} s: loss {losses[0]:.3f} -> {losses[-1]:.3f} (ln V = {np.log(vocab):.3f})")
    net.eval()
    data = mk.nbest(utts, nbest, seed=5)
    tok, tgt, pos, offs = data.flat_host()
    mk_t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    batch = PackedBatch(mk_t(tok), mk_t(tgt), mk_t(pos), mk_t(offs), int(np.diff(offs).max()), int(offs[-1]), len(offs) - 1)
    per_utt = lambda s: [np.asarray(s[i * nbest:(i + 1) * nbest]) for i in range(utts)]  # noqa: E731
    fast = net.score(batch, prec="bf16").float().cpu().numpy()
    precise = net.score(batch, prec="bf16x3").float().cpu().numpy()
    out = {"model": f"BayesTransformerModel FFN {layers}L d{d} FF{ff} h{nhead} V{vocab}, {steps} fine-tune steps of 32 x 100 "
                    f"tokens (bf16, lr {lr}) on a {branching}-successor Markov corpus",
           "train_loss_first": losses[0], "train_loss_last": losses[-1], "ln_vocab": float(np.log(vocab)),
           "lists": f"{utts} utterances x {nbest}-best, chain sentences + 1-3 random edits",
           "mean_nll_per_token": float(precise.sum() / offs[-1]),
           "fast_vs_precise": synth.ranking_agreement(per_utt(fast), per_utt(precise)),
           "max_abs_diff_fast_vs_precise": float(np.abs(fast - precise).max())}
    stage7 = lambda s: [data.graph[u] + 0.8 * per_utt(s)[u] + 0.2 * data.oldlm[u] for u in range(utts)]  # noqa: E731
    out["fast_vs_precise_stage7"] = synth.ranking_agreement(stage7(fast), stage7(precise))
    if with_oracle:
        from oracle import bayeslm_oracle as O
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        cfg = O.Config(family="bayes_tm", bayes_pos="FFN", ntoken=vocab, ninp=d, nhead=nhead, nhid=ff, nlayers=layers)
        t0 = time.time()
        want = []
        with torch.no_grad():
            for utt in data.tokenised():
                for x, y in utt:
                    want.append(O.sentence_nll(O.transformer_forward(sd, torch.tensor(x).view(-1, 1), cfg), torch.tensor(y)))
        want = np.asarray(want)
        log(f"oracle scored {len(want)} hypotheses in {time.time() - t0:.1f} s")
        out["oracle"] = {"precise_max_abs_err": float(np.abs(precise - want).max()),
                         "fast_max_abs_err": float(np.abs(fast - want).max()),
                         "precise_vs_oracle": synth.ranking_agreement(per_utt(precise), per_utt(want)),
                         "fast_vs_oracle": synth.ranking_agreement(per_utt(fast), per_utt(want)),
                         "precise_vs_oracle_stage7": synth.ranking_agreement(stage7(precise), stage7(want)),
                         "fast_vs_oracle_stage7": synth.ranking_agreement(stage7(fast), stage7(want)),
                         "smallest_gap_between_distinct_scores": float(min(
                             np.diff(np.unique(np.round(w, 4))).min() if len(np.unique(np.round(w, 4))) > 1 else np.inf
                             for w in per_utt(want)))}
        out["_scores"] = {"fast": fast, "precise": precise, "oracle": want}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocab", type=int, default=2000)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--utts", type=int, default=60)
    ap.add_argument("--nbest", type=int, default=20)
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--no-oracle", action="store_true")
    a = ap.parse_args()
    res = run(a.vocab, a.layers, a.steps, a.utts, a.nbest, lr=a.lr, with_oracle=not a.no_oracle,
              log=lambda *x: print(*x, file=sys.stderr))
    res.pop("_scores", None)
    print(json.dumps(res, indent=1))
