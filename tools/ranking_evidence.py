#!/usr/bin/env python
"""Ranking evidence on a PEAKED model: bayeslms_b200.evidence fine-tunes a Bayesian Transformer
(BASELINE layer sizes) on a synthetic Markov corpus, then chain-based n-best lists are scored in fast (bf16) and
precise (bf16x3) mode and by the CPU oracle; reported: per-utterance full-ranking / 1-best / pair-order agreement,
without and with the stage-7 combination (lmrescore_nbest_pytorchnn_cuda.sh:221-229), the largest oracle gap of a pair
the GPU orders differently, and the smallest gap between distinct oracle scores (scores are compared on the %.4f grid
of lmwt.nn, ties broken by hypothesis index).

    python tools/ranking_evidence.py [--vocab 500] [--layers 2] [--steps 1000] [--utts 60] [--nbest 20] > profiles/...json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocab", type=int, default=500)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--utts", type=int, default=60)
    ap.add_argument("--nbest", type=int, default=20)
    ap.add_argument("--lr", type=float, default=0.1)
    a = ap.parse_args()
    from bayeslms_b200 import evidence as E, synth
    from oracle import bayeslm_oracle as O
    t0 = time.time()
    net, mk, losses = E.peaked_model(a.vocab, a.layers, a.steps, a.lr)
    torch.cuda.synchronize()
    print(f"trained {a.steps} steps in {time.time() - t0:.1f} s: loss {losses[0]:.3f} -> {losses[-1]:.3f}", file=sys.stderr)
    data = mk.nbest(a.utts, a.nbest, seed=5)
    rep, fast, precise = E.fast_vs_precise(net, data)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    cfg = O.Config(family="bayes_tm", bayes_pos="FFN", ntoken=a.vocab, ninp=512, nhead=8, nhid=4096, nlayers=a.layers)
    want = []
    with torch.no_grad():
        for utt in data.tokenised():
            for x, y in utt:
                want.append(O.sentence_nll(O.transformer_forward(sd, torch.tensor(x).view(-1, 1), cfg), torch.tensor(y)))
    want = np.asarray(want)
    pu = lambda s: E.per_utterance(s, data)  # noqa: E731
    gaps = [np.diff(np.unique(np.round(w, 4))) for w in pu(want)]
    out = {"model": f"BayesTransformerModel FFN {a.layers}L d512 FF4096 h8 V{a.vocab}, {a.steps} fine-tune steps of 32 x 100 "
                    f"tokens (bf16, lr {a.lr}) on a 3-successor Markov corpus (entropy ln 3 = 1.099 nats)",
           "train_loss_first": losses[0], "train_loss_last": losses[-1], "ln_vocab": float(np.log(a.vocab)),
           "lists": f"{a.utts} utterances x {a.nbest}-best, chain sentences + 1-3 random edits", "fast_vs_precise": rep,
           "oracle": {"precise_max_abs_err": float(np.abs(precise - want).max()),
                      "fast_max_abs_err": float(np.abs(fast - want).max()),
                      "precise_vs_oracle": synth.ranking_agreement(pu(precise), pu(want)),
                      "fast_vs_oracle": synth.ranking_agreement(pu(fast), pu(want)),
                      "precise_vs_oracle_stage7": synth.ranking_agreement(E.stage7(precise, data), E.stage7(want, data)),
                      "fast_vs_oracle_stage7": synth.ranking_agreement(E.stage7(fast, data), E.stage7(want, data)),
                      "smallest_gap_between_distinct_scores": float(min(g.min() for g in gaps if len(g))),
                      "median_gap_between_neighbours": float(np.median(np.concatenate(gaps)))}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
