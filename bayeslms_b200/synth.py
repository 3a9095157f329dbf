"""Deterministic synthetic n-best lists: there is no network for real
lattices, so benchmarks and full-size parity tests use these.

Vocabulary ``<s>``=0, ``<unk>``=1, ``w00002``...; utterance u has a reference sentence of
L_u ~ U{5..25} words drawn Zipf(s=1) over [2, V); its N-best list is the reference plus N-1
perturbations (1-3 random substitutions / deletions / insertions each) in a fixed shuffled
order; every hypothesis carries synthetic ``graph`` / ``oldlm`` scores ~ 2 N(0,1) so the
stage-7 interpolation of ``lmrescore_nbest_pytorchnn_cuda.sh:221-229`` and a WER against the
reference sentence can be computed.  Seed 1111 (the reference default, train.py:88).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass
class SynthNbest:
    vocab_size: int
    refs: List[np.ndarray]              # per utterance: reference word ids
    hyps: List[List[np.ndarray]]        # per utterance: N hypotheses (word ids, no <s>)
    graph: List[np.ndarray]             # per utterance: [N] synthetic graph scores
    oldlm: List[np.ndarray]             # per utterance: [N] synthetic old-LM scores

    @property
    def n_utts(self):
        return len(self.hyps)

    def n_tokens(self, lo=0, hi=None) -> int:
        """scored positions = words + 1 per hypothesis (score.py:146,255)."""
        return sum(len(h) + 1 for u in self.hyps[lo:hi] for h in u)

    def tokenised(self, lo=0, hi=None) -> List[List[Tuple[List[int], List[int]]]]:
        """[(input ids, target ids)] per hypothesis per utterance: <s>+w / w+<s> (score.py:103-104)."""
        return [[([0] + h.tolist(), h.tolist() + [0]) for h in u] for u in self.hyps[lo:hi]]

    def flat_host(self, lo=0, hi=None):
        """Flat int32 host arrays (tokens, targets, pos, offsets) for utterances [lo, hi)."""
        hyps = [h for u in self.hyps[lo:hi] for h in u]
        lens = np.fromiter((len(h) + 1 for h in hyps), dtype=np.int64, count=len(hyps))
        offs = np.zeros(len(hyps) + 1, dtype=np.int32)
        np.cumsum(lens, out=offs[1:])
        M = int(offs[-1])
        tok = np.zeros(M, dtype=np.int32)
        tgt = np.zeros(M, dtype=np.int32)
        for i, h in enumerate(hyps):
            a = offs[i]
            tok[a + 1:a + 1 + len(h)] = h
            tgt[a:a + len(h)] = h
        pos = (np.arange(M, dtype=np.int32) - np.repeat(offs[:-1], lens)).astype(np.int32)
        return tok, tgt, pos, offs

    def words_text(self, lo=0, hi=None) -> List[str]:
        out = []
        for u in range(lo, self.n_utts if hi is None else hi):
            for n, h in enumerate(self.hyps[u], 1):
                out.append(f"utt{u:06d}-{n} " + " ".join(word(i) for i in h) if len(h) else f"utt{u:06d}-{n}")
        return out


def word(i: int) -> str:
    return "<s>" if i == 0 else "<unk>" if i == 1 else f"w{i:05d}"


def vocab_lines(V: int) -> List[str]:
    return [f"{word(i)} {i}" for i in range(V)]


def make_nbest(n_utts: int, n_best: int, vocab_size: int = 30000, seed: int = 1111,
               min_len: int = 5, max_len: int = 25) -> SynthNbest:
    rng = np.random.RandomState(seed)
    ranks = np.arange(2, vocab_size, dtype=np.float64)
    p = 1.0 / (ranks - 1.0)
    cdf = np.cumsum(p / p.sum())

    def zipf(n):
        return (np.searchsorted(cdf, rng.random_sample(n), side="left") + 2).astype(np.int32).clip(2, vocab_size - 1)

    refs, hyps, graph, oldlm = [], [], [], []
    for _ in range(n_utts):
        L = rng.randint(min_len, max_len + 1)
        ref = zipf(L)
        cand = [ref]
        for _ in range(n_best - 1):
            h = ref.tolist()
            for _ in range(rng.randint(1, 4)):
                op = rng.randint(3)
                if op == 0 and h:
                    h[rng.randint(len(h))] = int(zipf(1)[0])
                elif op == 1 and len(h) > 1:
                    del h[rng.randint(len(h))]
                else:
                    h.insert(rng.randint(len(h) + 1), int(zipf(1)[0]))
            cand.append(np.asarray(h, dtype=np.int32))
        order = rng.permutation(n_best)
        refs.append(ref)
        hyps.append([cand[i] for i in order])
        graph.append((rng.standard_normal(n_best) * 2).astype(np.float32))
        oldlm.append((rng.standard_normal(n_best) * 2).astype(np.float32))
    return SynthNbest(vocab_size, refs, hyps, graph, oldlm)


def pick_best(graph: np.ndarray, oldlm: np.ndarray, nn: np.ndarray, w: float = 0.8) -> int:
    """Stage 7 of the pipeline: total = graph + w * nn + (1 - w) * oldlm; lowest cost wins."""
    return int(np.argmin(graph.astype(np.float64) + w * nn.astype(np.float64) + (1.0 - w) * oldlm.astype(np.float64)))


def edit_distance(a, b) -> int:
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def wer(data: SynthNbest, nn_scores: List[np.ndarray], w: float = 0.8, lo: int = 0) -> Tuple[float, List[int]]:
    """Synthetic WER of the 1-best picked after interpolation, and the picks themselves."""
    errs, words, picks = 0, 0, []
    for i, s in enumerate(nn_scores):
        u = lo + i
        k = pick_best(data.graph[u], data.oldlm[u], np.asarray(s), w)
        picks.append(k)
        errs += edit_distance(data.hyps[u][k].tolist(), data.refs[u].tolist())
        words += len(data.refs[u])
    return errs / max(words, 1), picks


# ------------------------------------------------------------------ a learnable corpus (peaked models)
@dataclass
class MarkovCorpus:
    """First-order Markov chain over word ids [2, V) with ``branching`` equally likely successors per word: a model
    fine-tuned on it for a few hundred steps has sharp next-word distributions (entropy ~ ln(branching) nats), which
    random-init weights do not -- rankings of n-best lists then depend on the words, not on the lengths."""
    vocab_size: int
    nxt: np.ndarray      # [V, branching] successor table
    seed: int

    def stream(self, n: int, sent_len: int = 12, seed: int = 0) -> np.ndarray:
        """n ids of chain sentences, each followed by <s> (= 0), like data.py:14-54 lays a corpus out."""
        rng = np.random.RandomState(self.seed * 7919 + seed)
        out = np.empty(n, dtype=np.int64)
        cur, k = int(rng.randint(2, self.vocab_size)), 0
        pick = rng.randint(0, self.nxt.shape[1], size=n)
        for i in range(n):
            if k == sent_len:
                out[i], k = 0, 0
                cur = int(rng.randint(2, self.vocab_size))
            else:
                cur = int(self.nxt[cur, pick[i]])
                out[i] = cur
                k += 1
        return out

    def nbest(self, n_utts: int, n_best: int, seed: int = 0, min_len: int = 5, max_len: int = 25) -> SynthNbest:
        """n-best lists whose reference sentences follow the chain; the competitors are 1-3 random substitutions /
        deletions / insertions of it (random words: off-chain, so a trained model separates them clearly)."""
        rng = np.random.RandomState(self.seed * 104729 + seed)
        V = self.vocab_size
        refs, hyps, graph, oldlm = [], [], [], []
        for _ in range(n_utts):
            L = rng.randint(min_len, max_len + 1)
            cur, ref = int(rng.randint(2, V)), []
            for _ in range(L):
                cur = int(self.nxt[cur, rng.randint(self.nxt.shape[1])])
                ref.append(cur)
            ref = np.asarray(ref, dtype=np.int32)
            cand = [ref]
            for _ in range(n_best - 1):
                h = ref.tolist()
                for _ in range(rng.randint(1, 4)):
                    op = rng.randint(3)
                    if op == 0 and h:
                        h[rng.randint(len(h))] = int(rng.randint(2, V))
                    elif op == 1 and len(h) > 1:
                        del h[rng.randint(len(h))]
                    else:
                        h.insert(rng.randint(len(h) + 1), int(rng.randint(2, V)))
                cand.append(np.asarray(h, dtype=np.int32))
            order = rng.permutation(n_best)
            refs.append(ref)
            hyps.append([cand[i] for i in order])
            graph.append((rng.standard_normal(n_best) * 2).astype(np.float32))
            oldlm.append((rng.standard_normal(n_best) * 2).astype(np.float32))
        return SynthNbest(V, refs, hyps, graph, oldlm)


def make_markov(vocab_size: int, branching: int = 3, seed: int = 1111) -> MarkovCorpus:
    rng = np.random.RandomState(seed)
    return MarkovCorpus(vocab_size, rng.randint(2, vocab_size, size=(vocab_size, branching)), seed)


def ranking_agreement(a: List[np.ndarray], b: List[np.ndarray], min_gap: float = 0.0):
    """Per-utterance comparison of two score lists (lower = better, the pipeline's convention).  Scores are compared
    on the output grid of ``lmwt.nn`` (``%.4f``, score.py:302) with ties broken by hypothesis index -- the tie policy
    of a stable sort over the file order.  Returns a dict: fraction of utterances with the identical FULL ranking, with
    the identical 1-best, the fraction of hypothesis pairs ordered the same way among pairs whose gap in ``b`` exceeds
    ``min_gap``, and the largest gap in ``b`` of a pair that ``a`` orders differently.  ``one_best_within_gap`` also
    counts an utterance whose two 1-best candidates are closer than ``min_gap`` in ``b`` (a tie at the stated
    tolerance: either choice is a correct 1-best)."""
    full = best = best_gap = 0
    pairs = agree = 0
    worst = 0.0
    for x, y in zip(a, b):
        qx, qy = np.round(np.asarray(x, dtype=np.float64), 4), np.round(np.asarray(y, dtype=np.float64), 4)
        rx, ry = np.argsort(qx, kind="stable"), np.argsort(qy, kind="stable")
        full += int(np.array_equal(rx, ry))
        best += int(rx[0] == ry[0])
        best_gap += int(rx[0] == ry[0] or abs(qy[rx[0]] - qy[ry[0]]) <= min_gap)
        dx, dy = qx[:, None] - qx[None, :], qy[:, None] - qy[None, :]
        iu = np.triu_indices(len(qx), 1)
        gx, gy = dx[iu], dy[iu]
        sel = np.abs(gy) > min_gap
        same = np.sign(gx) == np.sign(gy)
        pairs += int(sel.sum())
        agree += int((same & sel).sum())
        flipped = (~same) & (np.sign(gx) * np.sign(gy) < 0)
        if flipped.any():
            worst = max(worst, float(np.abs(gy[flipped]).max()))
    n = max(len(a), 1)
    return {"full_ranking": full / n, "one_best": best / n, "one_best_within_gap": best_gap / n, "pair_order": agree / max(pairs, 1),
            "largest_flipped_gap": worst, "utterances": len(a), "pairs": pairs}
