"""Peaked models for ranking evidence.  A random-init LM has a near-uniform softmax, so its n-best scores are
~ln(V) per token and every ranking is a length ranking; "identical n-best ranking" (BASELINE.json north_star) only
means something on a model whose next-word distributions are sharp.  There are no checkpoints or corpora offline, so
one is made on the spot: a Bayesian Transformer with the BASELINE layer sizes is fine-tuned (this package's own CUDA
fine-tune step) on a synthetic first-order Markov corpus until its loss is close to the chain's entropy, and n-best
lists are drawn from the same chain (reference sentence = a chain path, competitors = random edits of it).
Used by bench.py (fast vs precise mode), tools/ranking_evidence.py and tests/test_gpu_ranking.py (both vs the oracle).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import synth


def peaked_model(vocab: int = 500, layers: int = 2, steps: int = 1000, lr: float = 0.1, clip: float = 1.0, d: int = 512,
                 nhead: int = 8, ff: int = 4096, branching: int = 3, device="cuda:0", seed: int = 1111):
    """(model in eval mode, MarkovCorpus, training losses).  Defaults: 1000 steps of 32 x 100 tokens in bf16 mode take
    ~1.5 s on a B200 and bring the loss from ~19 to ~1.5 nats per token (chain entropy ln 3 = 1.1; measured,
    profiles/r02_markov_sweep.txt -- larger learning rates stall at the unigram plateau)."""
    from . import model as M, train as T
    from .trainer import FineTuner
    torch.manual_seed(seed)
    mk = synth.make_markov(vocab, branching, seed)
    net = M.BayesTransformerModel(vocab, d, nhead, ff, layers, 0.0, True, "FFN").to(device)
    ft = FineTuner(net, lr, clip=clip, prec="bf16", data_parallel=False)
    ids = torch.from_numpy(mk.stream(32 * (100 * min(steps, 400) + 1)))
    losses = T.train_steps(ft, ids, 32, 100, steps, seed=seed)
    net.eval()
    return net, mk, losses


def score_lists(net, data: "synth.SynthNbest", prec: str) -> np.ndarray:
    from .engine import PackedBatch
    dev = next(net.parameters()).device
    tok, tgt, pos, offs = data.flat_host()
    mk = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    batch = PackedBatch(mk(tok), mk(tgt), mk(pos), mk(offs), int(np.diff(offs).max()), int(offs[-1]), len(offs) - 1)
    return net.score(batch, prec=prec).float().cpu().numpy()


def per_utterance(scores: np.ndarray, data: "synth.SynthNbest") -> List[np.ndarray]:
    out, at = [], 0
    for u in data.hyps:
        out.append(np.asarray(scores[at:at + len(u)]))
        at += len(u)
    return out


def stage7(scores: np.ndarray, data: "synth.SynthNbest", w: float = 0.8) -> List[np.ndarray]:
    """graph + w * nn + (1 - w) * oldlm per hypothesis (lmrescore_nbest_pytorchnn_cuda.sh:221-229)."""
    return [data.graph[u].astype(np.float64) + w * s.astype(np.float64) + (1.0 - w) * data.oldlm[u].astype(np.float64)
            for u, s in enumerate(per_utterance(scores, data))]


def fast_vs_precise(net, data: "synth.SynthNbest") -> Tuple[dict, np.ndarray, np.ndarray]:
    fast, precise = score_lists(net, data, "bf16"), score_lists(net, data, "bf16x3")
    rep = {"plain": synth.ranking_agreement(per_utterance(fast, data), per_utterance(precise, data)),
           "stage7": synth.ranking_agreement(stage7(fast, data), stage7(precise, data)),
           "max_abs_score_diff": float(np.abs(fast - precise).max()),
           "mean_nll_per_token": float(precise.sum() / data.n_tokens())}
    return rep, fast, precise
