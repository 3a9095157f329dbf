"""N-best rescoring driver: the B200 counterpart of the reference's
``steps/pytorchnn/compute_sentence_scores_bayes_jianwei.py`` (``score.py`` below).

Same command line (score.py:310-359), same input (``words_text``: ``<utt>-<n> w1 w2 ...``,
vocabulary ``word idx``) and output (``lmwt.nn``: ``<utt>-<n> %.4f``) formats, so it drops into
stage 6 of ``lmrescore_nbest_pytorchnn_cuda.sh``.  What changes is how the work is laid out:

* the reference scores one hypothesis at a time at batch size 1 with two host syncs per
  hypothesis (score.py:148-170); here all hypotheses are packed into token-major batches of
  up to ``max_tokens`` tokens, ids go to the device once per batch and scores come back once;
* Transformer hypotheses are independent -> any packing order is valid;
* LSTM: all hypotheses of an utterance share their initial state and the next utterance
  starts from the state after hypothesis #0 (score.py:271-274), which makes a *session*
  (the reference's per-JOB archive, state reset at score.py:232) the independent unit:
  phase 1 runs the hypothesis-#0 chain of every session in lock step, phase 2 scores all
  hypotheses of all utterances in parallel from the recorded states;
* multi-GPU: utterances (Transformer) or sessions (LSTM) are sharded over ranks with
  replicated weights; the only communication is one final gather of the score vector.
"""
from __future__ import annotations

import argparse
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine
from .engine import PackedBatch


# ----------------------------------------------------------------------- file formats
def read_vocab(path: str) -> Dict[str, int]:
    """``word idx`` per line; the id is the order of first appearance, not the second column
    (score.py:63-84)."""
    vocab: Dict[str, int] = {}
    with open(path, "r", encoding="utf-8") as f:
        for ln, line in enumerate(f, 1):
            fields = line.split()
            if len(fields) != 2:
                raise ValueError(f"{path}:{ln}: expected 'word index', got {line!r}")
            vocab.setdefault(fields[0], len(vocab))
    return vocab


def load_nbest(path: str) -> "OrderedDict[str, List[str]]":
    """``<utt>-<n> words...`` -> {utt: [hyp, ...]} in file order; a line with no words is the
    empty hypothesis ' ' (score.py:20-51)."""
    nbest: "OrderedDict[str, List[str]]" = OrderedDict()
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            key, sep, hyp = line.partition(" ")
            if not sep:
                hyp = " "
            nbest.setdefault(key.rsplit("-", 1)[0], []).append(hyp)
    return nbest


def ids_for(hyp: str, vocab: Dict[str, int]) -> Tuple[List[int], List[int]]:
    """input = <s> + words, target = words + <s>, OOV -> <unk> (score.py:87-120)."""
    bos = vocab["<s>"]
    unk = vocab.get("<unk>")
    words = []
    for w in hyp.split():
        i = vocab.get(w, unk)
        if i is None:
            raise KeyError(f"word {w!r} is out of vocabulary and the vocabulary has no <unk>")
        words.append(i)
    return [bos] + words, words + [bos]


def write_scores(scores: "OrderedDict[str, List[Tuple[str, float]]]", path: str) -> None:
    """``<utt>-<idx from 1> %.4f`` (score.py:283-303)."""
    with open(path, "w", encoding="utf-8") as f:
        for key, items in scores.items():
            for idx, (_, s) in enumerate(items, 1):
                f.write("%s-%d %.4f\n" % (key, idx, s))


# ------------------------------------------------------ file formats, fast path (C helpers of libbayeslm_b200.so)
class Vocab:
    """``words.txt`` held by the library's byte-string hash (blm_vocab_from_text): same ids as :func:`read_vocab`."""

    def __init__(self, path: str):
        import ctypes as C
        with open(path, "rb") as f:
            data = f.read()
        self._lib = _lib.lib()
        self._h = self._lib.blm_vocab_from_text(data, len(data))
        if not self._h:
            raise ValueError(f"{path}: {self._lib.blm_last_error().decode('utf-8', 'replace')}")
        self._C = C

    def __len__(self) -> int:
        return int(self._lib.blm_vocab_size(self._h))

    def id(self, word: str) -> int:
        b = word.encode("utf-8")
        return int(self._lib.blm_vocab_id(self._h, b, len(b)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.blm_vocab_free(h)


class NbestText:
    """A ``words_text`` file scanned by the C helpers: per-line byte offsets, scored positions (words + 1), utterance
    ids in order of first appearance and the 1-based hypothesis index inside the utterance (score.py:20-51, 283-303).
    ``tokenize`` writes the ids of a line range straight into caller-provided (pinned) int32 buffers."""

    def __init__(self, path: str, threads: Optional[int] = None):
        import ctypes as C
        self._C, self._lib = C, _lib.lib()
        with open(path, "rb") as f:
            self.data = f.read()
        self.threads = threads or min(8, os.cpu_count() or 1)
        n_lines, n_tokens, flags = C.c_int64(), C.c_int64(), C.c_int32()
        # room for the worst case (every byte a line); np.empty maps the pages lazily, only the used part is touched
        nl_cap = len(self.data) + 1
        self.line_begin = np.empty(nl_cap + 1, dtype=np.int64)
        self.line_tokens = np.empty(nl_cap, dtype=np.int32)
        _lib.check(self._lib.blm_nbest_scan(self.data, len(self.data), nl_cap, self._p(self.line_begin),
                                            self._p(self.line_tokens), C.byref(n_lines), C.byref(n_tokens),
                                            C.byref(flags), self.threads), "blm_nbest_scan")
        n = self.n_lines = int(n_lines.value)
        self.n_tokens = int(n_tokens.value)
        self.needs_slow_path = bool(flags.value & 1)          # Unicode-only whitespace in the text
        self.line_begin, self.line_tokens = self.line_begin[:n + 1].copy(), self.line_tokens[:n].copy()
        self.offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.line_tokens, out=self.offs[1:])
        self.utt_of_line = np.empty(n, dtype=np.int32)
        self.idx_in_utt = np.empty(n, dtype=np.int32)
        self.key_begin = np.empty(n, dtype=np.int64)
        self.key_len = np.empty(n, dtype=np.int32)
        n_utts = C.c_int64()
        _lib.check(self._lib.blm_nbest_group(self.data, self._p(self.line_begin), n, self._p(self.utt_of_line),
                                             self._p(self.idx_in_utt), self._p(self.key_begin), self._p(self.key_len),
                                             C.byref(n_utts)), "blm_nbest_group")
        self.n_utts = int(n_utts.value)
        # the reference writes utterances in order of first appearance; with every utterance's lines contiguous (what
        # the Kaldi pipeline produces) that is the file order
        self.contiguous = bool(n == 0 or (np.diff(self.utt_of_line) >= 0).all())

    def _p(self, a: np.ndarray):
        return a.ctypes.data_as(self._C.c_void_p)

    def utt_line_ranges(self) -> np.ndarray:
        """[n_utts + 1] first line of every utterance (contiguous files)."""
        return np.concatenate([np.searchsorted(self.utt_of_line, np.arange(self.n_utts), side="left"), [self.n_lines]])

    def tokenize(self, vocab: Vocab, l0: int, l1: int, tok: np.ndarray, tgt: np.ndarray, pos: Optional[np.ndarray]):
        for a in (tok, tgt) + ((pos,) if pos is not None else ()):
            assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"] and a.size >= self.offs[l1] - self.offs[l0]
        _lib.check(self._lib.blm_nbest_tokenize(vocab._h, self.data, self._p(self.line_begin), self._p(self.offs), l0, l1,
                                                self._p(tok), self._p(tgt), None if pos is None else self._p(pos),
                                                self.threads), "blm_nbest_tokenize")

    def format_scores(self, scores: np.ndarray) -> bytes:
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        assert scores.size == self.n_lines
        order = None if self.contiguous else np.argsort(self.utt_of_line, kind="stable").astype(np.int64)
        cap = int(self.key_len.sum()) + 72 * self.n_lines
        buf = self._C.create_string_buffer(cap)
        n = self._lib.blm_scores_format(self.data, self._p(self.key_begin), self._p(self.key_len), self._p(self.idx_in_utt),
                                        None if order is None else self._p(order), self.n_lines, self._p(scores), buf, cap)
        assert 0 <= n <= cap
        return buf.raw[:n]


# ------------------------------------------------------------------------- sharding
def shard_ranges(weights: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split items 0..n into ``world`` contiguous ranges with near-equal total weight."""
    n = len(weights)
    csum = np.concatenate([[0], np.cumsum(np.asarray(weights, dtype=np.int64))])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        bounds.append(int(np.searchsorted(csum, target, side="left")))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def wave_quantum() -> int:
    """Tokens in one full wave of 128-row tiles over the CTA pairs of the LayerNorm-fused GEMM (128 * SMs / 2), or 0
    before the library is initialised."""
    try:
        sms = _lib.lib().blm_num_sms()
    except Exception:
        return 0
    return 128 * (sms // 2) if sms > 1 else 0


def _greedy_fill(csum0: np.ndarray, limit: int) -> List[Tuple[int, int]]:
    """Greedy contiguous fill (every chunk <= limit tokens, at least one item) on the prefix sums csum0 (csum0[0] = 0):
    one searchsorted per chunk instead of a Python loop per hypothesis."""
    out, start, n = [], 0, len(csum0) - 1
    while start < n:
        end = int(np.searchsorted(csum0, csum0[start] + limit, side="right")) - 1
        end = min(max(end, start + 1), n)
        out.append((start, end))
        start = end
    return out


def _chunks_by_tokens(lengths: Sequence[int], max_tokens: int, quantum: int = 0) -> List[Tuple[int, int]]:
    """Balanced cuts (see :func:`_balanced_chunks`); with ``quantum`` > 0 also tries chunks filled to the largest
    multiple of ``quantum`` tokens that fits (whole waves of row tiles, remainder last) and keeps whichever needs
    fewer waves in total."""
    lengths = np.asarray(lengths, dtype=np.int64)
    base = _balanced_chunks(lengths, max_tokens)
    target = (max_tokens // quantum) * quantum if quantum > 0 else 0
    if target <= 0 or len(base) <= 1:
        return base
    csum = np.concatenate([[0], np.cumsum(np.asarray(lengths, dtype=np.int64))])
    alt = _greedy_fill(csum, target)

    def waves(cuts):
        return sum(-(-(-(-int(csum[b] - csum[a]) // 128)) // (quantum // 128)) for a, b in cuts)

    return alt if (len(alt) <= len(base) and waves(alt) < waves(base)) else base


def _balanced_chunks(lengths: Sequence[int], max_tokens: int) -> List[Tuple[int, int]]:
    """Contiguous hypothesis ranges of at most ``max_tokens`` tokens each, BALANCED: the fewest chunks that
    fit, cut at equal shares of the token count (a greedy fill leaves a small tail batch whose GEMMs run on a
    fraction of the SMs)."""
    if not len(lengths):
        return []
    # greedy fill: the fewest-chunk reference (and the answer when single items are close to the limit)
    greedy = _greedy_fill(np.concatenate([[0], np.cumsum(np.asarray(lengths, dtype=np.int64))]), max_tokens)
    csum = np.cumsum(np.asarray(lengths, dtype=np.int64))
    total = int(csum[-1])
    for n in range(max(1, -(-total // max_tokens)), len(greedy) + 1):
        cuts = [0]
        for j in range(1, n):
            i = int(np.searchsorted(csum, j * total / n, side="left")) + 1   # first prefix reaching the share
            cuts.append(min(max(i, cuts[-1] + 1), len(lengths)))
        cuts.append(len(lengths))
        cuts = sorted(set(cuts))
        sizes = [int(csum[b - 1] - (csum[a - 1] if a else 0)) for a, b in zip(cuts[:-1], cuts[1:])]
        if max(sizes) <= max_tokens:
            return list(zip(cuts[:-1], cuts[1:]))
    return greedy


# -------------------------------------------------------------------------- scoring
class Rescorer:
    """Scores tokenised n-best lists with one model replica on one GPU."""

    def __init__(self, model, *, prec: str = "bf16", K: int = 0, seed: Optional[int] = None,
                 max_tokens: int = 65536, eps_list=None, inter_model=None, inter_alpha: float = 0.8):
        self.model, self.prec, self.K, self.seed, self.max_tokens = model, prec, K, seed, max_tokens
        self.eps_list = eps_list
        self.inter_model, self.inter_alpha = inter_model, inter_alpha
        if inter_model is not None and model.family.endswith("lstm") != inter_model.family.endswith("lstm"):
            raise ValueError("logit interpolation needs two models of the same kind (score.py:385-447)")
        self.device = next(model.parameters()).device
        self.is_rnn = model.family.endswith("lstm")
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._stage = None   # pinned int32 staging for ids (grown on demand, reused across calls)
        self._out = None     # pinned fp32 landing buffer for the scores

    def _score(self, batch):
        kw = dict(K=self.K, seed=self.seed, eps_list=self.eps_list, prec=self.prec)
        if self.inter_model is not None:
            kw.update(inter_model=self.inter_model, inter_alpha=self.inter_alpha)
        return self.model.score(batch, **kw)

    # hyps: list of (input_ids, target_ids); returns fp32 numpy [n_hyp]
    def score_transformer(self, hyps: Sequence[Tuple[Sequence[int], Sequence[int]]]) -> np.ndarray:
        lengths = [len(x) for x, _ in hyps]
        outs = []
        for a, b in _chunks_by_tokens(lengths, self.max_tokens, wave_quantum()):
            batch = PackedBatch.from_lists([h[0] for h in hyps[a:b]], [h[1] for h in hyps[a:b]], self.device)
            self.h2d_bytes += batch.h2d_bytes
            outs.append(self._score(batch))
        res = torch.cat(outs) if outs else torch.empty(0, device=self.device)
        host = res.cpu()
        self.d2h_bytes += host.numel() * 4
        return host.numpy()

    def score_packed_host(self, tokens: np.ndarray, targets: np.ndarray, pos: np.ndarray,
                          offsets: np.ndarray) -> np.ndarray:
        """Flat host arrays in, host scores out -- the plain-pointer form of the scoring call
        (int32 ids [M], int32 offsets [n_hyp + 1]).  Hypotheses are cut into batches of at most
        ``max_tokens`` tokens; each batch is one pinned H2D copy, and all scores return in one D2H."""
        n_hyp = len(offsets) - 1
        lengths = np.diff(offsets)
        chunks = _chunks_by_tokens(lengths, self.max_tokens, wave_quantum())
        need = 3 * int(offsets[-1]) + n_hyp + len(chunks)
        if self._stage is None or self._stage.numel() < need:
            self._stage = torch.empty(max(need, 1 << 16), dtype=torch.int32, pin_memory=True)
        stage, at = self._stage.numpy(), 0
        outs = []
        for a, b in chunks:
            t0, t1 = int(offsets[a]), int(offsets[b])
            M, n = t1 - t0, 3 * (t1 - t0) + (b - a + 1)
            buf = stage[at:at + n]
            buf[:M], buf[M:2 * M], buf[2 * M:3 * M] = tokens[t0:t1], targets[t0:t1], pos[t0:t1]
            buf[3 * M:] = offsets[a:b + 1] - offsets[a]
            dev = self._stage[at:at + n].to(self.device, non_blocking=True)
            at += n
            batch = PackedBatch(dev[:M], dev[M:2 * M], dev[2 * M:3 * M], dev[3 * M:], int(lengths[a:b].max()), M, b - a)
            self.h2d_bytes += batch.h2d_bytes
            outs.append(self._score(batch))
        res = torch.cat(outs) if outs else torch.empty(0, device=self.device)
        if self._out is None or self._out.numel() < n_hyp:
            self._out = torch.empty(max(n_hyp, 1 << 12), dtype=torch.float32, pin_memory=True)
        self._out[:n_hyp].copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # scores are on the host; staging buffers reusable
        self.d2h_bytes += n_hyp * 4
        return self._out[:n_hyp].numpy().copy()

    def score_text_lines(self, nb: "NbestText", vocab: "Vocab", l0: int, l1: int) -> np.ndarray:
        """Lines [l0, l1) of a scanned ``words_text`` file -> fp32 scores (host).  Each batch is tokenised by the C
        helper directly into its slice of the pinned staging buffer and sent to the device while the kernels of the
        previous batch run (launches are asynchronous), and all scores return in one D2H copy."""
        n_hyp = l1 - l0
        if n_hyp <= 0:
            return np.zeros(0, dtype=np.float32)
        lengths = nb.line_tokens[l0:l1]
        chunks = _chunks_by_tokens(lengths, self.max_tokens, wave_quantum())
        need = 3 * int(nb.offs[l1] - nb.offs[l0]) + n_hyp + len(chunks)
        if self._stage is None or self._stage.numel() < need:
            self._stage = torch.empty(max(need, 1 << 16), dtype=torch.int32, pin_memory=True)
        stage, at = self._stage.numpy(), 0
        outs = []
        for a, b in chunks:
            la, lb = l0 + a, l0 + b
            M = int(nb.offs[lb] - nb.offs[la])
            n = 3 * M + (b - a + 1)
            buf = stage[at:at + n]
            nb.tokenize(vocab, la, lb, buf[:M], buf[M:2 * M], buf[2 * M:3 * M])
            buf[3 * M:] = nb.offs[la:lb + 1] - nb.offs[la]
            dev = self._stage[at:at + n].to(self.device, non_blocking=True)
            at += n
            batch = PackedBatch(dev[:M], dev[M:2 * M], dev[2 * M:3 * M], dev[3 * M:], int(lengths[a:b].max()), M, b - a)
            self.h2d_bytes += batch.h2d_bytes
            outs.append(self._score(batch))
        res = torch.cat(outs)
        if self._out is None or self._out.numel() < n_hyp:
            self._out = torch.empty(max(n_hyp, 1 << 12), dtype=torch.float32, pin_memory=True)
        self._out[:n_hyp].copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.d2h_bytes += n_hyp * 4
        return self._out[:n_hyp].numpy().copy()

    def score_sessions(self, sessions):
        """sessions: list of sessions, each a list of utterances, each a list of (input, target)."""
        return engine.lstm_score_sessions(self, sessions)

    def score_sessions_flat(self, tok, tgt, offs, sess_of, utt_of):
        """LSTM rescoring from flat host arrays (see :func:`bayeslms_b200.engine.lstm_score_flat`): rows ordered
        (session, utterance, hypothesis); returns fp32 numpy scores in row order."""
        return engine.lstm_score_flat(self, tok, tgt, offs, sess_of, utt_of)


def score_nbest(model, nbest: "OrderedDict[str, List[str]]", vocab: Dict[str, int], *, prec: str = "bf16",
                K: int = 0, seed: Optional[int] = None, max_tokens: int = 65536, eps_list=None,
                session_size: Optional[int] = None, rank: int = 0, world: int = 1, group=None,
                rescorer: Optional[Rescorer] = None, inter_model=None, inter_alpha: float = 0.8):
    """Score every hypothesis; returns {utt: [(hyp, score), ...]} in input order (score.py:206-280).

    With ``world > 1`` each rank scores its shard and the full result is assembled on every rank
    by one all-gather of the fp32 score vector.  ``session_size`` groups consecutive utterances
    into LSTM sessions (None = the whole list is one session, like one reference JOB).
    """
    rs = rescorer or Rescorer(model, prec=prec, K=K, seed=seed, max_tokens=max_tokens, eps_list=eps_list,
                              inter_model=inter_model, inter_alpha=inter_alpha)
    keys = list(nbest.keys())
    tokenised = [[ids_for(h, vocab) for h in nbest[k]] for k in keys]
    counts = [len(u) for u in tokenised]
    n_total = sum(counts)
    if rs.is_rnn:
        size = session_size or len(keys) or 1
        sessions = [list(range(s, min(s + size, len(keys)))) for s in range(0, len(keys), size)]
        weights = [sum(sum(len(x) for x, _ in tokenised[u]) for u in sess) for sess in sessions]
        lo, hi = shard_ranges(weights, world)[rank]
        mine_utts = [u for sess in sessions[lo:hi] for u in sess]
        local = rs.score_sessions([[tokenised[u] for u in sess] for sess in sessions[lo:hi]])
    else:
        weights = [sum(len(x) for x, _ in u) for u in tokenised]
        lo, hi = shard_ranges(weights, world)[rank]
        mine_utts = list(range(lo, hi))
        local = rs.score_transformer([h for u in mine_utts for h in tokenised[u]])
    utt_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if world > 1:
        import torch.distributed as dist
        full = torch.zeros(n_total, dtype=torch.float32, device=rs.device if dist.get_backend(group) == "nccl" else "cpu")
        if len(mine_utts):
            a, b = int(utt_start[mine_utts[0]]), int(utt_start[mine_utts[-1] + 1])
            full[a:b] = torch.from_numpy(np.ascontiguousarray(local)).to(full.device)
        dist.all_reduce(full, group=group)  # disjoint slices: the sum is the gather
        flat = full.cpu().numpy()
    else:
        flat = local
    out: "OrderedDict[str, List[Tuple[str, float]]]" = OrderedDict()
    for u, k in enumerate(keys):
        a = int(utt_start[u])
        out[k] = [(h, float(flat[a + i])) for i, h in enumerate(nbest[k])]
    return out


def score_files(model, nbest_path: str, vocab_path: str, out_path: Optional[str], *, prec: str = "bf16",
                K: int = 0, seed: Optional[int] = None, max_tokens: int = 65536, session_size: Optional[int] = None,
                rank: int = 0, world: int = 1, group=None, rescorer: Optional[Rescorer] = None, inter_model=None,
                inter_alpha: float = 0.8, vocab: Optional[Vocab] = None) -> np.ndarray:
    """``words_text`` + ``words.txt`` on disk -> ``lmwt.nn`` on disk: what stage 6 of the pipeline runs
    (lmrescore_nbest_pytorchnn_cuda.sh:197-219), with the text work done by the library's C helpers -- one pass over
    the file bytes, ids tokenised batch by batch straight into the pinned staging buffer while the GPU scores the
    previous batch, scores formatted in one pass.  Returns the fp32 scores in file order (every rank).

    Files whose utterances are not contiguous, or that contain Unicode-only whitespace, go through the per-hypothesis
    Python path (:func:`score_nbest`), which follows the reference's dict semantics literally."""
    rs = rescorer or Rescorer(model, prec=prec, K=K, seed=seed, max_tokens=max_tokens, inter_model=inter_model,
                              inter_alpha=inter_alpha)
    vocab_job = None
    if vocab is None:      # the vocabulary is parsed on a second host thread while this one scans the n-best text
        import concurrent.futures
        pool = concurrent.futures.ThreadPoolExecutor(1)
        vocab_job = pool.submit(Vocab, vocab_path)
        pool.shutdown(wait=False)
    nb = NbestText(nbest_path)
    if vocab_job is not None:
        vocab = vocab_job.result()
    if nb.needs_slow_path or not nb.contiguous:
        res = score_nbest(model, load_nbest(nbest_path), read_vocab(vocab_path), session_size=session_size, rank=rank,
                          world=world, group=group, rescorer=rs)
        if rank == 0 and out_path:
            write_scores(res, out_path)
        return np.asarray([sc for items in res.values() for _, sc in items], dtype=np.float32)
    n = nb.n_lines
    utt_first = nb.utt_line_ranges()                        # [n_utts + 1]
    utt_tokens = np.diff(nb.offs[utt_first])
    if rs.is_rnn:
        size = session_size or nb.n_utts or 1
        n_sess = -(-nb.n_utts // size)
        sess_first_utt = np.minimum(np.arange(n_sess + 1) * size, nb.n_utts)
        weights = np.diff(nb.offs[utt_first[sess_first_utt]])
        lo, hi = shard_ranges(weights, world)[rank]
        l0, l1 = int(utt_first[sess_first_utt[lo]]), int(utt_first[sess_first_utt[hi]])
        local = np.zeros(0, dtype=np.float32)
        if l1 > l0:
            m = int(nb.offs[l1] - nb.offs[l0])
            tok, tgt = np.empty(m, dtype=np.int32), np.empty(m, dtype=np.int32)
            nb.tokenize(vocab, l0, l1, tok, tgt, None)
            utt = nb.utt_of_line[l0:l1].astype(np.int64)
            local = rs.score_sessions_flat(tok, tgt, nb.offs[l0:l1 + 1] - nb.offs[l0], (utt // size - lo).astype(np.int32),
                                           (utt % size).astype(np.int32))
    else:
        lo, hi = shard_ranges(utt_tokens, world)[rank]
        l0, l1 = int(utt_first[lo]), int(utt_first[hi])
        local = rs.score_text_lines(nb, vocab, l0, l1)
    if world > 1:
        import torch.distributed as dist
        full = torch.zeros(n, dtype=torch.float32, device=rs.device if dist.get_backend(group) == "nccl" else "cpu")
        if l1 > l0:
            full[l0:l1] = torch.from_numpy(np.ascontiguousarray(local)).to(full.device)
        dist.all_reduce(full, group=group)                  # disjoint slices: the sum is the gather
        flat = full.cpu().numpy()
    else:
        flat = local
    if rank == 0 and out_path:
        with open(out_path, "wb") as f:
            f.write(nb.format_scores(flat))
    return flat


# ------------------------------------------------------------------------------ CLI
def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Compute sentence scores of n-best lists with a Bayesian / GP / "
                                            "Variational neural LM on B200.")
    p.add_argument("--nbest-list", type=str, required=True)
    p.add_argument("--outfile", type=str, required=True)
    p.add_argument("--vocabulary", type=str, required=True)
    p.add_argument("--model-path", type=str, required=True)
    p.add_argument("--model", type=str, default="LSTM")
    p.add_argument("--emsize", type=int, default=1024)
    p.add_argument("--nhid", type=int, default=1024)
    p.add_argument("--nlayers", type=int, default=2)
    p.add_argument("--nhead", type=int, default=8)
    p.add_argument("--uncertainty", type=str, default="none")
    p.add_argument("--T_bayes_pos", type=str, default="none")
    p.add_argument("--L_bayes_pos", type=int, default=0)
    p.add_argument("--L_gauss_pos", type=str, default="00")
    p.add_argument("--T_gauss_pos", type=int, default=3)
    p.add_argument("--L_v_pos", type=str, default="11")
    p.add_argument("--T_v_pos", type=str, default="0", help="0..3, or the per-layer bit string 00/01/10/11")
    p.add_argument("--interpolation_flag", type=int, default=0)
    p.add_argument("--inter_path", type=str, default=None)
    p.add_argument("--inter_alpha", type=float, default=0.8)
    # additions (all optional; defaults reproduce the reference's posterior-mean scoring)
    p.add_argument("--precision", type=str, default="bf16x3", choices=["bf16", "bf16x3"])
    p.add_argument("--num-samples", type=int, default=0, help="K posterior samples (0 = posterior mean)")
    p.add_argument("--seed", type=int, default=1111)
    p.add_argument("--max-tokens", type=int, default=65536)
    p.add_argument("--session-size", type=int, default=0, help="LSTM: utterances per session (0 = one session)")
    return p


_DERIVED_KEYS = ("pos_encoder.pe",)     # deterministic buffer (model.py:97-103): rebuilt, may be absent from a checkpoint


def load_checkpoint(model, path: str, *, partial: bool = False) -> dict:
    """Load the checkpoint keys the model knows (score.py:457-462).

    The reference intersects silently, which turns a checkpoint of the wrong architecture (e.g. a ``TransformerModel``
    file given to ``BayesTransformerModel(..., 'none')``) into a randomly initialised model that scores garbage.  Here
    that is an error: every parameter of the model must come from the file and every tensor of the file must have a
    home, unless ``partial=True`` -- the deliberate intersect-and-keep-the-rest semantics of ``--prior True``
    (train.py:239-258), which returns what was left out instead of raising.  Shape mismatches always raise."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(ckpt, dict):
        raise TypeError(f"{path}: expected a state_dict, got {type(ckpt).__name__}")
    own = model.state_dict()
    missing = [k for k in own if k not in ckpt and k not in _DERIVED_KEYS]
    unexpected = [k for k in ckpt if k not in own]
    if (missing or unexpected) and not partial:
        def show(keys):
            return ", ".join(keys[:6]) + (f", ... ({len(keys)} in all)" if len(keys) > 6 else "")
        raise KeyError(f"{path} does not match {type(model).__name__}: "
                       + (f"the model's {show(missing)} are not in the file; " if missing else "")
                       + (f"the file's {show(unexpected)} have no place in the model; " if unexpected else "")
                       + "check --model / --uncertainty / --*_pos against the flags the checkpoint was trained with")
    for k, v in ckpt.items():
        if k in own and tuple(v.shape) != tuple(own[k].shape):
            raise ValueError(f"{path}: {k} has shape {tuple(v.shape)}, the model expects {tuple(own[k].shape)}")
    own.update({k: v for k, v in ckpt.items() if k in own})
    model.load_state_dict(own)
    model.__dict__.pop("_blm_plans", None)
    return {"missing": missing, "unexpected": unexpected}


def main(argv=None) -> int:
    from . import model as models
    args = build_parser().parse_args(argv)
    for pth in (args.nbest_list, args.vocabulary, args.model_path):
        if not os.path.exists(pth):
            raise FileNotFoundError(pth)
    if args.interpolation_flag == 1 and not args.inter_path:
        raise ValueError("--interpolation_flag 1 needs an explicit --inter_path (the reference overwrites the flag with "
                         "a path on its authors' cluster, score.py:451-455)")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    vocab = Vocab(args.vocabulary)
    net = models.build_model(args, len(vocab))
    load_checkpoint(net, args.model_path)
    net = net.cuda().eval()
    net2 = None
    if args.interpolation_flag == 1 and args.uncertainty != "none":   # (--uncertainty none has no model_2, score.py:379,416)
        if args.model == "Transformer":
            # the second model is the standard Transformer, tied, same sizes (score.py:385-410)
            net2 = models.BayesTransformerModel(len(vocab), args.emsize, args.nhead, args.nhid, args.nlayers, 0.5, True,
                                                "none")
        else:
            # ... or the plain two-layer LSTM, built UNTIED (score.py:422-423, 432-433, 443-444)
            net2 = models.BayesRNNModel(args.model, len(vocab), args.emsize, args.nhid, args.nlayers, 0.5, False, 0)
        load_checkpoint(net2, args.inter_path)
        net2 = net2.cuda().eval()
    score_files(net, args.nbest_list, args.vocabulary, args.outfile, prec=args.precision, K=args.num_samples,
                seed=args.seed if args.num_samples else None, max_tokens=args.max_tokens,
                session_size=args.session_size or None, rank=rank, world=world, inter_model=net2,
                inter_alpha=args.inter_alpha, vocab=vocab)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
