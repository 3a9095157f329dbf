"""Tensor-level wrappers over the C ABI: torch owns the memory and the stream, the
kernels in ``libbayeslm_b200.so`` do all the arithmetic.

Precision modes (``prec``):
  ``"bf16"``    one bf16 product per GEMM, fp32 accumulation (fast path);
  ``"bf16x3"``  every fp32 operand is carried as ``hi + lo`` bf16 pairs and each GEMM
                accumulates ``hi*hi + hi*lo + lo*hi`` in fp32 -- about 16 mantissa bits,
                the mode held to the reference's fp32 results within 1e-3.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_FAST, ACT_GPMIX_FAST, ACT_GELU_GRAD, ACT_GPMIX, ACT_GPMIX_GRAD, ACT_NONE, ACT_SOFTMAX_GRAD, EPS_NONE, EPS_PHILOX,
                   EPS_PTR, DropoutDesc, GemmDesc, GemmLnDesc, GemmSampledDesc,
                   VocabNllDesc,
                   check, lib)

PRECISIONS = ("bf16", "bf16x3")
PRECISE_K_CHUNK = 128   # K elements per tensor-core accumulation in bf16x3 mode (blm_gemm_desc.k_chunk)


class _Stats:
    """Launch accounting for bench.py: how many of OUR kernels were launched, and (when
    ``timing`` is a dict) CUDA-event brackets around every launch on the launching stream."""
    launches = 0
    timing = None   # None, or {op name: [(start_event, end_event, work), ...]}
    bytes = {}      # op name -> algorithmic bytes of the bracketed launches (HBM-bound kernels)


STATS = _Stats()


class _op:
    def __init__(self, name: str, kernels: int = 1, work: float = 0.0, bytes_: float = 0.0):
        self.name, self.kernels, self.work, self.bytes = name, kernels, work, bytes_

    def __enter__(self):
        STATS.launches += self.kernels
        if STATS.timing is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STATS.timing is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STATS.timing.setdefault(self.name, []).append((self.e0, e1, self.work))
            if self.bytes:
                STATS.bytes[self.name] = STATS.bytes.get(self.name, 0.0) + self.bytes
        return False


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.BlmError("bayeslms_b200 kernels need CUDA tensors (there is no CPU fallback)")


@dataclass
class Split:
    """An fp32 matrix carried as bf16 ``hi`` (+ optional ``lo`` residual), both [rows, cols]."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.hi.shape

    def float(self) -> torch.Tensor:
        x = self.hi.float()
        return x if self.lo is None else x + self.lo.float()


def empty_split(rows: int, cols: int, prec: str, device) -> Split:
    hi = torch.empty(rows, cols, dtype=torch.bfloat16, device=device)
    lo = torch.empty(rows, cols, dtype=torch.bfloat16, device=device) if prec == "bf16x3" else None
    return Split(hi, lo)


def split(x: torch.Tensor, prec: str = "bf16x3") -> Split:
    """fp32 [rows, cols] -> Split."""
    _require_cuda(x)
    x = x.detach().contiguous().float()
    out = empty_split(x.shape[0] if x.dim() == 2 else 1, x.shape[-1] if x.dim() == 2 else x.numel(), prec, x.device)
    with _op("split_bf16", 1):
        check(lib().blm_split_bf16(_ptr(x), _ptr(out.hi), _ptr(out.lo), x.numel(), _stream()), "blm_split_bf16")
    if x.dim() != 2:
        out = Split(out.hi.view(x.shape), None if out.lo is None else out.lo.view(x.shape))
    return out


def _segments(a: Split, b: Split, prec: str):
    if prec == "bf16":
        return [(a.hi, b.hi)]
    if prec == "bf16x3":
        if a.lo is None or b.lo is None:
            raise _lib.BlmError("bf16x3 needs (hi, lo) operands")
        # small cross terms first, hi*hi last: the tensor core truncates on every accumulate, and the
        # truncation is relative to the running sum
        return [(a.hi, b.lo), (a.lo, b.hi), (a.hi, b.hi)]
    raise _lib.BlmError(f"unknown precision {prec!r}")


def gemm(a: Split, b: Split, *, prec: str = "bf16", bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         coef: Optional[torch.Tensor] = None, col_scale: float = 1.0, col_scale_cols: int = 0,
         resid: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
         out: Optional[Split] = None, extra: Sequence = (), tag: str = "", out_pre: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None,
         targets: Optional[torch.Tensor] = None, grad_scale: float = 1.0, k_chunk: Optional[int] = None,
         a_f16: bool = False, a_mn: bool = False, b_mn: bool = False, fast_act: bool = False,
         f32_rows32: bool = False, drop: Optional["Drop"] = None):
    """``epilogue(a @ b.T)`` with a [M, K], b [N, K].  ``extra`` appends more (A, B) Split pairs
    accumulated into the same output (K-concatenation).  ``out_pre`` receives the fp32 value before
    the activation; ``aux`` is the saved pre-activation of the ``*_GRAD`` epilogues; ``lse`` /
    ``targets`` / ``grad_scale`` feed ``ACT_SOFTMAX_GRAD``.  ``k_chunk`` (default: 128 in bf16x3 mode)
    bounds the length of one tensor-core accumulation (see ``blm_gemm_desc.k_chunk``).
    ``a_mn`` / ``b_mn``: the operand is given MN-major, a as [K, M] / b as [K, N] (e.g. ``dW = gemm(dY, X, a_mn=True,
    b_mn=True)`` with dY [tokens, N], X [tokens, K]; ``dX = gemm(dY, W, b_mn=True)`` with W [N, K]).
    ``f32_rows32``: ``out_f32`` (``rows32_empty(M, N)``) is written in 32-row blocks (``blm_gemm_desc.f32_rows32``).
    ``drop``: dropout fused into the epilogue (``blm_gemm_desc.drop``; N % 32 == 0): forward ``resid + m * act(.)``,
    ``*_GRAD`` epilogues ``(m * acc) * act'(aux)`` with ``out_pre = m * acc``."""
    segs = _segments(a, b, prec)
    for (a2, b2) in extra:
        segs += _segments(a2, b2, prec)
    M = a.hi.shape[1] if a_mn else a.hi.shape[0]
    N = b.hi.shape[1] if b_mn else b.hi.shape[0]
    d = GemmDesc()
    d.M, d.N, d.nseg, d.act = M, N, len(segs), act
    for i, (x, w) in enumerate(segs):
        _require_cuda(x, w)
        want = torch.float16 if a_f16 else torch.bfloat16     # fp16 mode: BOTH operands (mixed types are illegal on sm_100a)
        assert x.dtype == want and w.dtype == want
        kx, kw = x.shape[0 if a_mn else 1], w.shape[0 if b_mn else 1]
        assert kx == kw and x.stride(1) == 1 and w.stride(1) == 1
        d.A[i], d.B[i] = x.data_ptr(), w.data_ptr()
        d.K[i], d.lda[i], d.ldb[i] = kx, x.stride(0), w.stride(0)
    d.bias, d.coef = _ptr(bias), _ptr(coef)
    d.col_scale, d.col_scale_cols = col_scale, col_scale_cols
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.stride(1) == 1
        d.resid, d.ldr = _ptr(resid), resid.stride(0)
    ldc = None
    for t in (out_f32, None if out is None else out.hi, None if out is None else out.lo):
        if t is not None:
            assert t.stride(1) == 1 and (ldc is None or ldc == t.stride(0))
            ldc = t.stride(0)
    d.out_f32 = _ptr(out_f32)
    d.out_hi = _ptr(None if out is None else out.hi)
    d.out_lo = _ptr(None if out is None else out.lo)
    d.ldc = ldc if ldc is not None else N
    if out_pre is not None:
        assert out_pre.dtype == torch.float32 and out_pre.stride(1) == 1 and out_pre.stride(0) == d.ldc
        d.out_pre = _ptr(out_pre)
    if aux is not None:
        assert aux.dtype == torch.float32 and aux.stride(1) == 1
        d.aux, d.ldaux = _ptr(aux), aux.stride(0)
    if lse is not None:
        d.lse, d.targets, d.grad_scale = _ptr(lse), _ptr(targets), grad_scale
    d.k_chunk = (PRECISE_K_CHUNK if prec == "bf16x3" else 0) if k_chunk is None else k_chunk
    d.a_f16, d.a_mn, d.b_mn, d.fast_act = int(a_f16), int(a_mn), int(b_mn), int(fast_act)
    if f32_rows32:
        assert out_f32 is not None and out_f32.is_contiguous() and out_f32.shape == (_pad32(M), N)
        d.f32_rows32 = 1
    dd = None
    if drop is not None:
        assert N % 32 == 0 and (drop.mask is None or drop.mask.numel() == M * N)
        dd = drop.desc()                 # stays alive until the call returns
        d.drop = C.addressof(dd)
    with _op("gemm:" + tag if tag else "gemm", 1, 2.0 * M * N * sum(x.shape[0 if a_mn else 1] for x, _ in segs)):
        check(lib().blm_gemm(C.byref(d), _stream()), "blm_gemm")


def _pad32(n: int) -> int:
    return (n + 31) // 32 * 32


def rows32_empty(M: int, N: int, device) -> torch.Tensor:
    """fp32 buffer of a [M, N] matrix in the 32-row-block layout [ceil(M / 32)][N / 4][32 rows][4 floats]."""
    assert N % 4 == 0
    return torch.empty(_pad32(M), N, dtype=torch.float32, device=device)


def rows32_to_dense(x: torch.Tensor, M: int) -> torch.Tensor:
    """The [M, N] row-major matrix held by a 32-row-block buffer (tests, debugging)."""
    N = x.shape[1]
    return x.view(-1, N // 4, 32, 4).permute(0, 2, 1, 3).reshape(-1, N)[:M]


GEMM_LN_WIDTHS = (128, 256, 384, 512)   # d_model values the LayerNorm-fused GEMM covers (blm_gemm_ln)


def gemm_ln(a: Split, b: Split, *, bias: Optional[torch.Tensor], resid: torch.Tensor, gamma: torch.Tensor,
            beta: torch.Tensor, eps: float, want_f32: bool = True, tag: str = ""):
    """``LayerNorm(resid + a @ b.T + bias) * gamma + beta`` in one kernel (bf16 operands, fp32 statistics):
    returns (fp32 [M, N] or None, Split(hi)).  Fast mode only; N must be one of GEMM_LN_WIDTHS."""
    x, w = a.hi, b.hi
    _require_cuda(x, w, resid, gamma, beta, bias)
    M, N = x.shape[0], w.shape[0]
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.shape[1] == w.shape[1]
    assert x.stride(1) == 1 and w.stride(1) == 1 and resid.dtype == torch.float32 and resid.stride(1) == 1
    y = torch.empty(M, N, dtype=torch.float32, device=x.device) if want_f32 else None
    s = empty_split(M, N, "bf16", x.device)
    d = GemmLnDesc()
    d.M, d.N, d.K = M, N, x.shape[1]
    d.A, d.lda, d.B, d.ldb = x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0)
    d.bias, d.resid, d.ldr = _ptr(bias), _ptr(resid), resid.stride(0)
    d.gamma, d.beta, d.eps = _ptr(gamma), _ptr(beta), eps
    d.out_f32, d.out_hi, d.ldc = _ptr(y), _ptr(s.hi), N
    with _op("gemm_ln:" + tag if tag else "gemm_ln", 1, 2.0 * M * N * x.shape[1]):
        check(lib().blm_gemm_ln(C.byref(d), _stream()), "blm_gemm_ln")
    return y, s


def sigma_bf16(lgstd: torch.Tensor) -> torch.Tensor:
    """bf16(exp(lgstd)), the cached sample-independent scale of the fused sampled GEMM."""
    lgstd = lgstd.detach().contiguous().float()
    out = torch.empty(lgstd.shape, dtype=torch.bfloat16, device=lgstd.device)
    with _op("sigma_bf16", 1):
        check(lib().blm_sigma_bf16(_ptr(lgstd), _ptr(out), lgstd.numel(), _stream()), "blm_sigma_bf16")
    return out


def gemm_sampled(a: Split, mu: Optional[torch.Tensor], sigma: Optional[torch.Tensor], *, eps: Optional[torch.Tensor] = None,
                 mu_f32: Optional[torch.Tensor] = None, lgstd_f32: Optional[torch.Tensor] = None,
                 seed: Optional[int] = None, stream_id: int = 0, bias: Optional[torch.Tensor] = None,
                 act: int = ACT_NONE, coef: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
                 out_f32: Optional[torch.Tensor] = None, out: Optional[Split] = None, tag: str = "sampled",
                 how: str = "once"):
    """``epilogue(a @ bf16(mu + sigma * eps).T)`` with the sampled weight built inside the GEMM launch.
    mu / sigma: bf16 [N, K] (sigma dense).  eps: explicit fp32 tensor, Philox(seed, stream_id), or none (mean).
    ``how="once"``: every CTA draws 1/grid of W~ once into an L2-resident scratch, then the pipelined GEMM
    (default); ``how="tile"``: tile-stationary kernel, W~ regenerated per group of M tiles and never stored."""
    x = a.hi
    M, K = x.shape
    mode = EPS_PTR if eps is not None else (EPS_PHILOX if seed is not None else EPS_NONE)
    d = GemmSampledDesc()
    if mu_f32 is not None:
        # fp32 parameters straight into the generate-once kernel: same rounding as reparam(prec="bf16")
        assert how == "once" and mode != EPS_NONE and lgstd_f32 is not None
        assert mu_f32.dtype == torch.float32 and mu_f32.stride(1) == 1 and mu_f32.shape[1] == K
        assert lgstd_f32.dtype == torch.float32 and lgstd_f32.is_contiguous() and lgstd_f32.shape == mu_f32.shape
        _require_cuda(mu_f32, lgstd_f32)
        N = mu_f32.shape[0]
        d.mu_f32, d.ldmu_f32, d.lgstd_f32 = mu_f32.data_ptr(), mu_f32.stride(0), lgstd_f32.data_ptr()
    else:
        N = mu.shape[0]
    d.M, d.N, d.K = M, N, K
    d.A, d.lda = x.data_ptr(), x.stride(0)
    if mu is not None:
        assert mu.shape[1] == K and mu.stride(1) == 1 and mu.dtype == torch.bfloat16
        d.mu, d.ldmu = mu.data_ptr(), mu.stride(0)
    if sigma is not None:
        assert sigma.dtype == torch.bfloat16 and sigma.is_contiguous() and sigma.shape == mu.shape
        d.sigma = sigma.data_ptr()
    if eps is not None:
        eps = eps.contiguous().float()
        d.eps = eps.data_ptr()
    d.eps_mode, d.act, d.seed, d.stream_id = mode, act, int(seed or 0), int(stream_id)
    d.bias, d.coef = _ptr(bias), _ptr(coef)
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.stride(1) == 1
        d.resid, d.ldr = _ptr(resid), resid.stride(0)
    ldc = None
    for t in (out_f32, None if out is None else out.hi, None if out is None else out.lo):
        if t is not None:
            assert t.stride(1) == 1 and (ldc is None or ldc == t.stride(0))
            ldc = t.stride(0)
    d.out_f32 = _ptr(out_f32)
    d.out_hi = _ptr(None if out is None else out.hi)
    d.out_lo = _ptr(None if out is None else out.lo)
    d.ldc = ldc if ldc is not None else N
    if how == "once" and (eps is not None or seed is not None):
        nbytes = lib().blm_gemm_sampled_workspace_bytes(N, K)
        ws = _workspace("gemm_sampled", nbytes, x.device, zero=True)   # zeroed once, the kernel re-arms its counters
        d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    with _op("gemm_sampled:" + tag, 1, 2.0 * M * N * K):
        check(lib().blm_gemm_sampled(C.byref(d), _stream()), "blm_gemm_sampled")


def vocab_nll(h: Split, e: Split, bias: Optional[torch.Tensor], targets: torch.Tensor, *, prec: str = "bf16",
              extra: Sequence = (), out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-row ``-log softmax(h @ e.T + bias)[target]`` without materialising the logits.
    ``lse`` [M] optionally receives the row log-sum-exp (saved for the backward pass)."""
    segs = _segments(h, e, prec)
    for (h2, e2) in extra:
        segs += _segments(h2, e2, prec)
    M, V = h.hi.shape[0], e.hi.shape[0]
    dev = h.hi.device
    nbytes = lib().blm_vocab_nll_workspace_bytes(M, V)
    ws = _workspace("vocab_nll", max(nbytes, 1 << 20), dev)
    if out is None:
        out = torch.empty(M, dtype=torch.float32, device=dev)
    assert targets.dtype == torch.int32 and targets.numel() == M
    d = VocabNllDesc()
    d.M, d.V, d.nseg = M, V, len(segs)
    for i, (x, w) in enumerate(segs):
        _require_cuda(x, w)
        d.H[i], d.E[i] = x.data_ptr(), w.data_ptr()
        d.K[i], d.ldh[i], d.lde[i] = x.shape[1], x.stride(0), w.stride(0)
    d.bias, d.targets, d.nll = _ptr(bias), _ptr(targets), _ptr(out)
    d.workspace, d.workspace_bytes = _ptr(ws), ws.numel()
    d.lse = _ptr(lse)
    with _op("vocab_nll", 2, 2.0 * M * V * sum(x.shape[1] for x, _ in segs)):
        check(lib().blm_vocab_nll(C.byref(d), _stream()), "blm_vocab_nll")
    return out


def segment_sum(x: torch.Tensor, offsets: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n = offsets.numel() - 1
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=x.device)
    assert offsets.dtype == torch.int32
    with _op("segment_sum", 1):
        check(lib().blm_segment_sum(_ptr(x), _ptr(offsets), n, _ptr(out), _stream()), "blm_segment_sum")
    return out


def mc_combine(nll_km: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[K, M] per-sample token NLLs -> [M] Monte-Carlo predictive NLL."""
    K, M = nll_km.shape
    assert nll_km.is_contiguous() and nll_km.dtype == torch.float32
    if out is None:
        out = torch.empty(M, dtype=torch.float32, device=nll_km.device)
    with _op("mc_combine", 1):
        check(lib().blm_mc_combine(_ptr(nll_km), K, M, _ptr(out), _stream()), "blm_mc_combine")
    return out


def embed(tokens: torch.Tensor, pos: Optional[torch.Tensor], emb: torch.Tensor, pe: Optional[torch.Tensor],
          scale: float, *, prec: str = "bf16", want_f32: bool = True):
    """Returns (x_f32 or None, Split)."""
    M, d = tokens.numel(), emb.shape[1]
    dev = emb.device
    x = torch.empty(M, d, dtype=torch.float32, device=dev) if want_f32 else None
    s = empty_split(M, d, prec, dev)
    with _op("embed", 1):
        check(lib().blm_embed(_ptr(tokens), _ptr(pos), _ptr(emb), _ptr(pe), scale, M, d, _ptr(x), _ptr(s.hi),
                              _ptr(s.lo), _stream()), "blm_embed")
    return x, s


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, prec: str = "bf16",
              want_f32: bool = True):
    M, d = x.shape
    y = torch.empty_like(x) if want_f32 else None
    s = empty_split(M, d, prec, x.device)
    with _op("layernorm", 1):
        check(lib().blm_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), eps, M, d, _ptr(y), _ptr(s.hi), _ptr(s.lo),
                                  _stream()), "blm_layernorm")
    return y, s


def reparam(mu: torch.Tensor, lgstd: Optional[torch.Tensor], *, eps: Optional[torch.Tensor] = None,
            seed: Optional[int] = None, stream_id: int = 0, prec: str = "bf16", want_f32: bool = False,
            out: Optional[Split] = None, out_f32: Optional[torch.Tensor] = None):
    """``mu + exp(lgstd) * eps`` for a 2-D row-slice view ``mu`` (unit column stride).
    eps: explicit tensor, or Philox(seed, stream_id) when ``seed`` is given, or none (mean).
    ``out`` / ``out_f32`` write into caller-provided dense [rows, cols] buffers (e.g. the gate-row
    block of a full LSTM weight copy)."""
    if mu.dim() == 1:
        mu = mu.view(1, -1)
        lgstd = None if lgstd is None else lgstd.view(1, -1)
        eps = None if eps is None else eps.view(1, -1)
    rows, cols = mu.shape
    assert mu.stride(1) == 1
    mode = EPS_PTR if eps is not None else (EPS_PHILOX if seed is not None else EPS_NONE)
    dev = mu.device
    w = out_f32
    if w is None and want_f32:
        w = torch.empty(rows, cols, dtype=torch.float32, device=dev)
    s = out
    if s is None and (out_f32 is None or not want_f32):
        s = empty_split(rows, cols, prec, dev) if out_f32 is None else None
    if w is not None:
        assert w.is_contiguous() and w.numel() == rows * cols
    if s is not None:
        assert s.hi.is_contiguous() and s.hi.numel() == rows * cols and (s.lo is None or s.lo.is_contiguous())
    if lgstd is not None:
        lgstd = lgstd.contiguous()
    if eps is not None:
        eps = eps.contiguous().float()
    with _op("reparam", 1):
        check(lib().blm_reparam(_ptr(mu), mu.stride(0), _ptr(lgstd), _ptr(eps), mode, int(seed or 0), int(stream_id),
                                rows, cols, _ptr(w), _ptr(None if s is None else s.hi),
                                _ptr(None if s is None else s.lo), _stream()), "blm_reparam")
    return w, s


def philox_normal(seed: int, stream_id: int, n: int, device, *, out: Optional[torch.Tensor] = None,
                  scale: float = 1.0) -> torch.Tensor:
    """scale * N(0,1) from the library's Philox stream (seed, stream_id): element i = counter i/4, lane i%4."""
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=device)
    assert out.is_contiguous() and out.numel() == n
    with _op("philox_normal", 1):
        check(lib().blm_philox_normal_scaled(int(seed), int(stream_id), n, scale, _ptr(out), _stream()),
              "blm_philox_normal_scaled")
    return out


def mha_causal(qkv: torch.Tensor, seq_offsets: torch.Tensor, nhead: int, max_len: int, *, prec: str = "bf16",
               want_f32: bool = False):
    M, d3 = qkv.shape
    d = d3 // 3
    nseq = seq_offsets.numel() - 1
    o = torch.empty(M, d, dtype=torch.float32, device=qkv.device) if want_f32 else None
    s = empty_split(M, d, prec, qkv.device)
    with _op("mha_causal", 1):
        check(lib().blm_mha_causal(_ptr(qkv), _ptr(seq_offsets), nseq, nhead, d // nhead, max_len, _ptr(o), _ptr(s.hi),
                                   _ptr(s.lo), _stream()), "blm_mha_causal")
    return o, s


@dataclass
class Drop:
    """One dropout site of the fine-tune step: drop probability ``p`` and where the keep multipliers (0 or 1/(1-p)) come
    from -- an explicit fp32 tensor ``mask`` (parity with masks injected into the oracle) or Philox(seed + *seed_dev,
    stream_id) on the device.  ``seed_dev``: a device int64 [1] added to the key, so captured graphs replay with fresh
    masks."""
    p: float
    mask: Optional[torch.Tensor] = None
    seed: int = 0
    stream_id: int = 0
    seed_dev: Optional[torch.Tensor] = None

    def desc(self) -> DropoutDesc:
        d = DropoutDesc()
        if self.mask is not None:
            assert self.mask.dtype == torch.float32 and self.mask.is_contiguous() and self.mask.is_cuda
            d.mask = self.mask.data_ptr()
        d.p, d.seed, d.stream_id = float(self.p), int(self.seed) & 0xFFFFFFFFFFFFFFFF, int(self.stream_id)
        if self.seed_dev is not None:
            assert self.seed_dev.dtype == torch.int64 and self.seed_dev.is_cuda
            d.seed_dev = self.seed_dev.data_ptr()
        return d


def dropout(x: Optional[torch.Tensor], drop: Drop, *, resid: Optional[torch.Tensor] = None, prec: Optional[str] = None,
            want_f32: bool = True, out_f32: Optional[torch.Tensor] = None, n: Optional[int] = None, device=None):
    """``x * m (+ resid)`` with the multipliers of ``drop``; returns (fp32 or None, Split or None).  ``prec`` asks for
    the bf16 (hi[, lo]) copy as well; ``x=None`` with ``n`` exports the multipliers themselves."""
    if x is not None:
        assert x.dtype == torch.float32 and x.is_contiguous()
        n, device, shape = x.numel(), x.device, x.shape
    else:
        shape = (n,)
    if drop.mask is not None:
        assert drop.mask.numel() == n
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.is_contiguous() and resid.numel() == n
    y = out_f32 if out_f32 is not None else (torch.empty(shape, dtype=torch.float32, device=device) if want_f32 else None)
    sp = None
    if prec is not None:
        sp = Split(torch.empty(shape, dtype=torch.bfloat16, device=device),
                   torch.empty(shape, dtype=torch.bfloat16, device=device) if prec == "bf16x3" else None)
    d = drop.desc()
    with _op("dropout", 1):
        check(lib().blm_dropout(_ptr(x), n, C.byref(d), _ptr(resid), _ptr(y), _ptr(None if sp is None else sp.hi),
                                _ptr(None if sp is None else sp.lo), _stream()), "blm_dropout")
    return y, sp


def mha_causal_bf16(qkv: Split, seq_offsets: torch.Tensor, nhead: int, max_len: int, *, prec: str = "bf16",
                    want_f32: bool = False, drop: Optional[Drop] = None):
    """Causal attention on the bf16 (hi[, lo]) output of the QKV projection (tensor-core kernel).  ``drop``: dropout
    on the attention probabilities (mask layout [n_seq * nhead, L, L], L = max_len rounded up to 4)."""
    M, d3 = qkv.hi.shape
    d = d3 // 3
    nseq = seq_offsets.numel() - 1
    lo = qkv.lo if prec == "bf16x3" else None
    if prec == "bf16x3" and lo is None:
        raise _lib.BlmError("bf16x3 attention needs the lo part of qkv")
    assert qkv.hi.stride(1) == 1 and (lo is None or lo.stride() == qkv.hi.stride())
    o = torch.empty(M, d, dtype=torch.float32, device=qkv.hi.device) if want_f32 else None
    s = empty_split(M, d, prec, qkv.hi.device)
    with _op("mha_causal", 1, 0.0):
        if drop is not None:
            dd = drop.desc()
            check(lib().blm_mha_causal_bf16_dropout(_ptr(qkv.hi), _ptr(lo), qkv.hi.stride(0), _ptr(seq_offsets), nseq, nhead,
                                                    d // nhead, max_len, C.byref(dd), _ptr(o), _ptr(s.hi), _ptr(s.lo), d,
                                                    _stream()), "blm_mha_causal_bf16_dropout")
        else:
            check(lib().blm_mha_causal_bf16(_ptr(qkv.hi), _ptr(lo), qkv.hi.stride(0), _ptr(seq_offsets), nseq, nhead,
                                            d // nhead, max_len, _ptr(o), _ptr(s.hi), _ptr(s.lo), d, _stream()),
                  "blm_mha_causal_bf16")
    return o, s


def kl_gauss(mu: torch.Tensor, lgstd: torch.Tensor, out: torch.Tensor, *, minus_one: bool = False,
             scale: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out[0] (+)= scale * 0.5 * mean(mu^2 - 2 lgstd + exp(2 lgstd) [- 1]) over a 2-D row-slice view."""
    if mu.dim() == 1:
        mu, lgstd = mu.view(1, -1), lgstd.view(1, -1)
    rows, cols = mu.shape
    assert mu.stride(1) == 1 and lgstd.is_contiguous() and lgstd.shape == mu.shape
    ws = _workspace("kl", lib().blm_kl_workspace_bytes(), mu.device, zero=True)
    with _op("kl_gauss", 1, 4.0 * 2 * rows * cols):
        check(lib().blm_kl_gauss(_ptr(mu), mu.stride(0), _ptr(lgstd), rows, cols, int(minus_one), scale,
                                 int(accumulate), _ptr(out), _ptr(ws), _stream()), "blm_kl_gauss")
    return out


def lstm_layer(gates_x: torch.Tensor, w_hh: Split, h0: torch.Tensor, c0: torch.Tensor, lengths: torch.Tensor,
               T: int, B: int, H: int, *, prec: str = "bf16", want_f32: bool = False, want_split: bool = True,
               c_seq: Optional[torch.Tensor] = None, gx_rows32: bool = False, tag: str = ""):
    """One LSTM layer over [T, B] lock-stepped rows.  Returns (out_f32 or None, out Split or None, hT, cT).
    ``c_seq`` [T * B, H] fp32 (optional) receives the cell state after every live step.  ``gx_rows32``: ``gates_x``
    is a ``rows32_empty(T * B, 4 * H)`` buffer filled by ``gemm(..., f32_rows32=True)``."""
    dev = gates_x.device
    assert gates_x.is_contiguous() and lengths.dtype == torch.int32
    assert gates_x.numel() == (_pad32(T * B) if gx_rows32 else T * B) * 4 * H
    h0, c0 = h0.contiguous(), c0.contiguous()
    out32 = torch.empty(T * B, H, dtype=torch.float32, device=dev) if want_f32 else None
    outs = empty_split(T * B, H, prec, dev) if want_split else None
    hT = torch.empty(B, H, dtype=torch.float32, device=dev)
    cT = torch.empty(B, H, dtype=torch.float32, device=dev)
    nbytes = lib().blm_lstm_workspace_bytes(B, H)
    ws = _workspace("lstm", max(nbytes, 1 << 20), dev)
    if c_seq is not None:
        assert c_seq.dtype == torch.float32 and c_seq.is_contiguous() and c_seq.numel() == T * B * H
    # algorithmic bytes per step = W_hh once (4H x H bf16) + gates_x read (B x 4H fp32) + h, c write
    with _op("lstm_layer:" + tag if tag else "lstm_layer", 1, 2.0 * T * B * 4 * H * H, T * (4.0 * H * H * 2 + B * 4.0 * H * 4 + 2.0 * B * H * 4)):
        check(lib().blm_lstm_layer_seq(_ptr(gates_x), int(gx_rows32), _ptr(w_hh.hi), _ptr(w_hh.lo if prec == "bf16x3" else None), _ptr(h0),
                                       _ptr(c0), _ptr(lengths), T, B, H, _ptr(out32),
                                       _ptr(None if outs is None else outs.hi), _ptr(None if outs is None else outs.lo),
                                       _ptr(hT), _ptr(cT), _ptr(c_seq), _ptr(ws), _stream()), "blm_lstm_layer")
    return out32, outs, hT, cT


def gp_lstm_cell(acc5: torch.Tensor, coef: torch.Tensor, gate_type: int, lengths: torch.Tensor, t: int, c: torch.Tensor,
                 h: torch.Tensor, h_op: Split, out_f32: Optional[torch.Tensor], out: Optional[Split]) -> None:
    """One timestep of the GP-LSTM cell (see ``blm_gp_lstm_cell``); c, h updated in place."""
    B, H = h.shape
    assert acc5.stride(1) == 1 and coef.is_contiguous() and c.is_contiguous() and h.is_contiguous()
    with _op("gp_lstm_cell", 1):
        check(lib().blm_gp_lstm_cell(_ptr(acc5), acc5.stride(0), _ptr(coef), coef.shape[0], gate_type, _ptr(lengths), t, B, H,
                                     _ptr(c), _ptr(h), _ptr(h_op.hi), _ptr(h_op.lo), _ptr(out_f32),
                                     _ptr(None if out is None else out.hi), _ptr(None if out is None else out.lo),
                                     _stream()), "blm_gp_lstm_cell")


def lstm_cell_step(acc4: torch.Tensor, lengths: torch.Tensor, t: int, c: torch.Tensor, h: torch.Tensor, h_op: Split,
                   out_f32: Optional[torch.Tensor], out: Optional[Split], c_in: Optional[torch.Tensor] = None) -> None:
    """One timestep of the plain cell from pre-activations acc4 [B, 4H] (see ``blm_lstm_cell_step``); c, h in place."""
    B, H = h.shape
    assert acc4.stride(1) == 1 and acc4.shape[1] >= 4 * H and c.is_contiguous() and h.is_contiguous()
    assert c_in is None or (c_in.is_contiguous() and c_in.shape == c.shape)
    with _op("lstm_cell_step", 1):
        check(lib().blm_lstm_cell_step(_ptr(acc4), acc4.stride(0), _ptr(c_in), _ptr(lengths), t, B, H, _ptr(c), _ptr(h),
                                       _ptr(h_op.hi), _ptr(h_op.lo), _ptr(out_f32),
                                       _ptr(None if out is None else out.hi), _ptr(None if out is None else out.lo),
                                       _stream()), "blm_lstm_cell_step")


# ------------------------------------------------------------------ fine-tune step (backward twins)
def _ld8(n: int) -> int:
    return (n + 7) // 8 * 8


def transpose_split(x: torch.Tensor, prec: str = "bf16x3") -> Split:
    """fp32 [R, C] -> Split [C, R] (row stride rounded up to 8 elements; use ``[:, :R]`` views)."""
    R, C = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    ld = _ld8(R)
    hi = torch.empty(C, ld, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(C, ld, dtype=torch.bfloat16, device=x.device) if prec == "bf16x3" else None
    with _op("transpose", 1):
        check(lib().blm_transpose_split(_ptr(x), x.stride(0), R, C, _ptr(hi), _ptr(lo), ld, _stream()), "blm_transpose_split")
    return Split(hi[:, :R], None if lo is None else lo[:, :R])


def split_transpose(x: torch.Tensor, prec: str = "bf16x3"):
    """fp32 [R, C] -> (Split [R, C], Split [C, R]) in one pass over x."""
    x = x.detach()
    R, C = x.shape
    assert x.dtype == torch.float32 and x.stride(1) == 1
    ld, ldt = _ld8(C), _ld8(R)
    two = prec == "bf16x3"
    mk = lambda r, l: torch.empty(r, l, dtype=torch.bfloat16, device=x.device)  # noqa: E731
    hi, lo = mk(R, ld), (mk(R, ld) if two else None)
    thi, tlo = mk(C, ldt), (mk(C, ldt) if two else None)
    with _op("split_transpose", 1):
        check(lib().blm_split_transpose(_ptr(x), x.stride(0), R, C, _ptr(hi), _ptr(lo), ld, _ptr(thi), _ptr(tlo), ldt,
                                        _stream()), "blm_split_transpose")
    return (Split(hi[:, :C], None if lo is None else lo[:, :C]), Split(thi[:, :R], None if tlo is None else tlo[:, :R]))


def transpose_bf16(x: Split, prec: str = "bf16x3") -> Split:
    """Split [R, C] -> Split [C, R]."""
    R, C = x.hi.shape
    assert x.hi.stride(1) == 1
    ld = _ld8(R)
    hi = torch.empty(C, ld, dtype=torch.bfloat16, device=x.hi.device)
    want_lo = prec == "bf16x3" and x.lo is not None
    lo = torch.empty(C, ld, dtype=torch.bfloat16, device=x.hi.device) if want_lo else None
    with _op("transpose", 1):
        check(lib().blm_transpose_bf16(_ptr(x.hi), _ptr(x.lo if want_lo else None), x.hi.stride(0), R, C, _ptr(hi), _ptr(lo),
                                       ld, _stream()), "blm_transpose_bf16")
    return Split(hi[:, :R], None if lo is None else lo[:, :R])


def colsum(x, out: torch.Tensor, *, scale: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out[n] (+)= scale * sum_m x[m, n]; x fp32 [M, N] or a Split."""
    first = x.hi if isinstance(x, Split) else x
    M, N = first.shape
    ws = _workspace("colsum", lib().blm_colsum_workspace_bytes(M, N), first.device, zero=True)
    with _op("colsum", 1):
        if isinstance(x, Split):
            check(lib().blm_colsum_bf16(_ptr(x.hi), _ptr(x.lo), x.hi.stride(0), M, N, scale, int(accumulate), _ptr(out),
                                        _ptr(ws), _stream()), "blm_colsum_bf16")
        else:
            check(lib().blm_colsum(_ptr(x), x.stride(0), M, N, scale, int(accumulate), _ptr(out), _ptr(ws), _stream()),
                  "blm_colsum")
    return out


_ws_cache = {}
CAPTURE_KEEP = []
_capture_ws = {}


def end_capture() -> None:
    """Forget the scratch blocks handed out during a CUDA-graph capture (call when the capture has ended, or before a new
    one begins): they belong to that graph's private pool."""
    CAPTURE_KEEP.clear()
    _capture_ws.clear()


def _workspace(kind: str, nbytes: int, device, zero: bool = False) -> torch.Tensor:
    """Scratch memory of one kernel family, cached per (device, stream).  While a CUDA graph is being captured the cache is
    a separate one that lives only as long as the capture (``end_capture``): a tensor allocated then belongs to the
    graph's private pool, which is released with the graph, so keeping it for the NEXT capture on the same (re-used)
    capture stream hands out a dangling pointer (found as an illegal address in a later graph's replay).  Inside one
    capture, launches of a family on the same stream are ordered, so they share their block -- a fresh zeroed block per
    launch was 27 fill kernels per fine-tune step."""
    if torch.cuda.is_current_stream_capturing():
        key = (kind, device.index, torch.cuda.current_stream().cuda_stream)
        ws = _capture_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=device)
            # held until the capture ends: a capture may span two streams, and a block returned to the pool here could be
            # re-issued to the other stream while this launch is still unordered against it
            CAPTURE_KEEP.append(ws)
            _capture_ws[key] = ws
        return ws
    key = (kind, device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = (torch.zeros if zero else torch.empty)(max(nbytes, 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float, dgamma: torch.Tensor,
                  dbeta: torch.Tensor, *, accumulate: bool = False, prec: Optional[str] = None,
                  dxsum: Optional[torch.Tensor] = None, dxsum_accumulate: bool = False):
    """Returns dx; x is the LayerNorm input.  dgamma / dbeta are written (or accumulated) in place.  ``prec``: also
    return the bf16 (hi[, lo]) copy of dx -> (dx, Split); ``dxsum`` [d] (+)= column sums of dx (the bias gradient of
    the projection in front of the residual add)."""
    M, d = x.shape
    assert dy.is_contiguous() and x.is_contiguous()
    dx = torch.empty_like(x)
    sp = empty_split(M, d, prec, x.device) if prec is not None else None
    ws = _workspace("ln_bwd", lib().blm_layernorm_bwd_workspace_bytes(M, d), x.device)
    with _op("layernorm_bwd", 2):
        check(lib().blm_layernorm_bwd_ex(_ptr(dy), _ptr(x), _ptr(gamma), eps, M, d, _ptr(dx),
                                         _ptr(None if sp is None else sp.hi), _ptr(None if sp is None else sp.lo),
                                         _ptr(dgamma), _ptr(dbeta), int(accumulate), _ptr(dxsum), int(dxsum_accumulate),
                                         _ptr(ws), _stream()), "blm_layernorm_bwd_ex")
    return dx if prec is None else (dx, sp)


def mha_causal_bwd(qkv: torch.Tensor, dout: torch.Tensor, seq_offsets: torch.Tensor, nhead: int, max_len: int,
                   q_scale: float, prec: Optional[str] = None, drop: Optional[Drop] = None) -> torch.Tensor:
    """Gradient of the causal attention w.r.t. the (q-scaled) qkv projection.  ``prec`` = "bf16" / "bf16x3" selects
    the tensor-core kernel (head_dim 64); None (or BLM_ATTN_BWD_SIMT=1) the fp32 SIMT kernel."""
    M, d3 = qkv.shape
    d = d3 // 3
    dqkv = torch.empty_like(qkv)
    if drop is not None:
        if prec is None or d // nhead != 64 or max_len > 128:
            raise _lib.BlmError("attention dropout runs on the tensor-core kernels (head_dim 64, length <= 128)")
        dd = drop.desc()
        with _op("mha_causal_bwd", 1):
            check(lib().blm_mha_causal_bwd_tc_dropout(_ptr(qkv), qkv.stride(0), _ptr(dout), dout.stride(0), _ptr(seq_offsets),
                                                      seq_offsets.numel() - 1, nhead, 64, max_len, q_scale,
                                                      int(prec == "bf16x3"), C.byref(dd), _ptr(dqkv), dqkv.stride(0),
                                                      _stream()), "blm_mha_causal_bwd_tc_dropout")
        return dqkv
    if prec is not None and d // nhead == 64 and max_len <= 128 and os.environ.get("BLM_ATTN_BWD_SIMT") is None:
        with _op("mha_causal_bwd", 1):
            check(lib().blm_mha_causal_bwd_tc(_ptr(qkv), qkv.stride(0), _ptr(dout), dout.stride(0), _ptr(seq_offsets),
                                              seq_offsets.numel() - 1, nhead, 64, max_len, q_scale, int(prec == "bf16x3"),
                                              _ptr(dqkv), dqkv.stride(0), _stream()), "blm_mha_causal_bwd_tc")
        return dqkv
    with _op("mha_causal_bwd", 1):
        check(lib().blm_mha_causal_bwd(_ptr(qkv), qkv.stride(0), _ptr(dout), dout.stride(0), _ptr(seq_offsets),
                                       seq_offsets.numel() - 1, nhead, d // nhead, max_len, q_scale, _ptr(dqkv),
                                       dqkv.stride(0), _stream()), "blm_mha_causal_bwd")
    return dqkv


def gpmix_dcoef(z: torch.Tensor, dh: torch.Tensor, dcoef: torch.Tensor, *, accumulate: bool = False) -> torch.Tensor:
    M, N = z.shape
    assert z.stride() == dh.stride() and dcoef.is_contiguous()
    with _op("gpmix_dcoef", 1):
        check(lib().blm_gpmix_dcoef(_ptr(z), _ptr(dh), z.stride(0), M, N, int(accumulate), _ptr(dcoef), _stream()),
              "blm_gpmix_dcoef")
    return dcoef


def gp3_bwd(dh: torch.Tensor, z: torch.Tensor, coef: torch.Tensor, dcoef: torch.Tensor, prec: str,
            out: Optional[torch.Tensor] = None, out_split: Optional[Split] = None):
    """Backward of a stand-alone GP unit with acts (sigmoid, tanh, relu) (``blm_gp3_bwd``): returns (dz fp32, Split);
    ``dcoef`` [3, N] is accumulated in place.  ``out`` / ``out_split``: write into these (row slices of larger buffers)."""
    M, N = z.shape
    assert dh.shape == z.shape and dh.stride() == z.stride() and z.stride(1) == 1 and z.stride(0) == N and coef.is_contiguous()
    assert dcoef.is_contiguous() and dcoef.shape == coef.shape == (3, N)
    dz = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=z.device)
    sp = out_split if out_split is not None else empty_split(M, N, prec, z.device)
    assert dz.is_contiguous() and sp.hi.is_contiguous() and dz.shape == (M, N)
    with _op("gp3_bwd", 1):
        check(lib().blm_gp3_bwd(_ptr(dh), _ptr(z), _ptr(coef), N, M, N, _ptr(dz), _ptr(sp.hi), _ptr(sp.lo), _ptr(dcoef),
                                _stream()), "blm_gp3_bwd")
    return dz, sp


def _eps_args(eps, seed):
    if eps is not None:
        return _ptr(eps), EPS_PTR, 0
    if seed is None:
        raise _lib.BlmError("noise needs an explicit tensor or a Philox seed")
    return None, EPS_PHILOX, int(seed)


def vnoise_fwd(f: torch.Tensor, rho: torch.Tensor, B: int, T: int, *, eps: Optional[torch.Tensor] = None,
               seed: Optional[int] = None, stream_id: int = 0, noise_std: float = 0.1,
               resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp = f + e * exp(f * rho[t]) (+ resid) for sequence-major rows (row = b*T + t); rho [T, d]."""
    d = f.shape[1]
    assert f.is_contiguous() and rho.is_contiguous() and f.shape[0] == B * T
    fp = torch.empty_like(f)
    e, mode, sd = _eps_args(eps, seed)
    with _op("vnoise_fwd", 1):
        check(lib().blm_vnoise_fwd(_ptr(f), _ptr(rho), e, mode, sd, int(stream_id), noise_std, _ptr(resid), B, T, d,
                                   _ptr(fp), _stream()), "blm_vnoise_fwd")
    return fp


def vnoise_bwd(dfp: Optional[torch.Tensor], f: torch.Tensor, rho: torch.Tensor, mean_p: torch.Tensor, B: int, T: int,
               kl_scale: float, drho: torch.Tensor, dmean_p: torch.Tensor, *, eps: Optional[torch.Tensor] = None,
               seed: Optional[int] = None, stream_id: int = 0, noise_std: float = 0.1):
    """Returns (df, klpart [T, d]); writes drho / dmean_p."""
    d = f.shape[1]
    df = torch.empty_like(f)
    klpart = torch.empty(T, d, dtype=torch.float32, device=f.device)
    e, mode, sd = _eps_args(eps, seed)
    with _op("vnoise_bwd", 1):
        check(lib().blm_vnoise_bwd(_ptr(dfp), _ptr(f), _ptr(rho), _ptr(mean_p), e, mode, sd, int(stream_id), noise_std, B, T,
                                   d, kl_scale, _ptr(df), _ptr(drho), _ptr(dmean_p), _ptr(klpart), _stream()),
              "blm_vnoise_bwd")
    return df, klpart


def embed_bwd(dx: torch.Tensor, tokens: torch.Tensor, scale: float, dE: torch.Tensor) -> None:
    M, d = dx.shape
    assert dx.is_contiguous() and dE.is_contiguous() and tokens.dtype == torch.int32
    with _op("embed_bwd", 1):
        check(lib().blm_embed_bwd(_ptr(dx), _ptr(tokens), scale, M, d, _ptr(dE), _stream()), "blm_embed_bwd")


def kl_gauss_bwd(mu: torch.Tensor, lgstd: torch.Tensor, scale: float, dmu: torch.Tensor, dlgstd: torch.Tensor) -> None:
    """dmu += c mu, dlgstd += c (exp(2 lgstd) - 1), c = scale / numel; mu / dmu may be row-slice views."""
    if mu.dim() == 1:
        mu, lgstd, dmu, dlgstd = mu.view(1, -1), lgstd.view(1, -1), dmu.view(1, -1), dlgstd.view(1, -1)
    rows, cols = mu.shape
    assert lgstd.is_contiguous() and dlgstd.is_contiguous() and mu.stride(1) == 1 and dmu.stride(1) == 1
    with _op("kl_gauss_bwd", 1):
        check(lib().blm_kl_gauss_bwd(_ptr(mu), mu.stride(0), _ptr(lgstd), rows, cols, scale, _ptr(dmu), dmu.stride(0),
                                     _ptr(dlgstd), _stream()), "blm_kl_gauss_bwd")


def reparam_bwd(G: torch.Tensor, lgstd: torch.Tensor, dmu: torch.Tensor, dlgstd: torch.Tensor, *,
                eps: Optional[torch.Tensor] = None, seed: Optional[int] = None, stream_id: int = 0,
                accumulate: bool = False) -> None:
    """dmu (+)= G, dlgstd (+)= G * eps * exp(lgstd); G may alias dmu."""
    if G.dim() == 1:
        G, lgstd, dmu, dlgstd = G.view(1, -1), lgstd.view(1, -1), dmu.view(1, -1), dlgstd.view(1, -1)
        eps = None if eps is None else eps.view(1, -1)
    rows, cols = G.shape
    assert lgstd.is_contiguous() and dlgstd.is_contiguous() and G.stride(1) == 1 and dmu.stride(1) == 1
    if eps is not None:
        eps = eps.contiguous().float()
    e, mode, sd = _eps_args(eps, seed)
    with _op("reparam_bwd", 1):
        check(lib().blm_reparam_bwd(_ptr(G), G.stride(0), _ptr(lgstd), e, mode, sd, int(stream_id), rows, cols,
                                    int(accumulate), _ptr(dmu), dmu.stride(0), _ptr(dlgstd), _stream()), "blm_reparam_bwd")


def lstm_gates_act(gates: torch.Tensor, c0: torch.Tensor, T: int, B: int, H: int) -> torch.Tensor:
    """Pre-activations [T*B, 4H] -> activated gates (in place); returns the cell states [T*B, H]."""
    assert gates.is_contiguous() and gates.shape == (T * B, 4 * H) and c0.is_contiguous() and c0.shape == (B, H)
    c_all = torch.empty(T * B, H, dtype=torch.float32, device=gates.device)
    with _op("lstm_gates_act", 1):
        check(lib().blm_lstm_gates_act(_ptr(gates), _ptr(c0), T, B, H, _ptr(c_all), _stream()), "blm_lstm_gates_act")
    return c_all


def lstm_bwd_step(gates_t, c_prev, c_t, dout_t, dh_rec, dc, dc_is_zero: bool, dg32_t, dg_t: Split) -> None:
    """One step of the LSTM backward recurrence over rows [B] (see ``blm_lstm_bwd_step``)."""
    B, H = c_t.shape
    for x in (gates_t, c_prev, c_t, dout_t, dc, dg32_t, dg_t.hi):
        assert x.is_contiguous()
    with _op("lstm_bwd_step", 1):
        check(lib().blm_lstm_bwd_step(_ptr(gates_t), _ptr(c_prev), _ptr(c_t), _ptr(dout_t), _ptr(dh_rec), _ptr(dc),
                                      int(dc_is_zero), B, H, _ptr(dg32_t), _ptr(dg_t.hi), _ptr(dg_t.lo), _stream()),
              "blm_lstm_bwd_step")


def reduce_sum(x: torch.Tensor, out: torch.Tensor, *, squares: bool = False, scale: float = 1.0,
               accumulate: bool = False) -> torch.Tensor:
    """out[0] (+)= scale * sum(x) or scale * sum(x^2), deterministic."""
    assert x.is_contiguous() and x.dtype == torch.float32
    ws = _workspace("reduce", lib().blm_reduce_workspace_bytes(), x.device, zero=True)
    with _op("reduce", 1):
        check(lib().blm_reduce(_ptr(x), x.numel(), int(squares), scale, int(accumulate), _ptr(out), _ptr(ws), _stream()),
              "blm_reduce")
    return out


def sgd_momentum(p: torch.Tensor, g: torch.Tensor, v: torch.Tensor, lr: float, momentum: float,
                 norm_sq: Optional[torch.Tensor], max_norm: float, grad_scale: float = 1.0,
                 out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None) -> None:
    """clip + SGD momentum over flat buffers; ``out_hi`` / ``out_lo``: bf16 operand copies of the updated parameters."""
    assert p.is_contiguous() and g.is_contiguous() and v.is_contiguous() and p.numel() == g.numel() == v.numel()
    if out_hi is not None:
        assert out_hi.dtype == torch.bfloat16 and out_hi.numel() == p.numel() and out_hi.is_contiguous()
        with _op("sgd_momentum", 1):
            check(lib().blm_sgd_momentum_split(_ptr(p), _ptr(g), _ptr(v), p.numel(), lr, momentum, _ptr(norm_sq), max_norm,
                                               grad_scale, _ptr(out_hi), _ptr(out_lo), _stream()), "blm_sgd_momentum_split")
        return
    with _op("sgd_momentum", 1):
        check(lib().blm_sgd_momentum(_ptr(p), _ptr(g), _ptr(v), p.numel(), lr, momentum, _ptr(norm_sq), max_norm,
                                     grad_scale, _stream()), "blm_sgd_momentum")


# ------------------------------------------------------------------ training-mode Variational / GP LSTM cells
def vnn_noise(e: torch.Tensor, rho: torch.Tensor) -> torch.Tensor:
    """n[t, j] = e[t, j] * exp(rho[j]); e [T, H], rho [H]."""
    T, H = e.shape
    assert e.is_contiguous() and rho.is_contiguous() and rho.numel() == H
    n = torch.empty_like(e)
    with _op("vnn_noise", 1):
        check(lib().blm_vnn_noise(_ptr(e), _ptr(rho), T, H, _ptr(n), _stream()), "blm_vnn_noise")
    return n


def rowgroup_add(x: torch.Tensor, r: torch.Tensor, G: int, B: int, *, prec: Optional[str] = None, want_f32: bool = True,
                 out_f32: Optional[torch.Tensor] = None):
    """out[g*B + b, :] = x[g*B + b, :] + r[g, :]; returns (fp32 or None, Split or None)."""
    W = x.shape[-1]
    assert x.is_contiguous() and r.is_contiguous() and x.numel() == G * B * W and r.numel() == G * W
    y = out_f32 if out_f32 is not None else (torch.empty_like(x) if want_f32 else None)
    sp = None
    if prec is not None:
        sp = Split(torch.empty(x.shape, dtype=torch.bfloat16, device=x.device),
                   torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if prec == "bf16x3" else None)
    with _op("rowgroup_add", 1):
        check(lib().blm_rowgroup_add(_ptr(x), _ptr(r), G, B, W, _ptr(y), _ptr(None if sp is None else sp.hi),
                                     _ptr(None if sp is None else sp.lo), _stream()), "blm_rowgroup_add")
    return y, sp


def rowgroup_sum(x: torch.Tensor, G: int, B: int) -> torch.Tensor:
    """out[g, :] = sum_b x[g*B + b, :]."""
    W = x.shape[-1]
    assert x.is_contiguous() and x.numel() == G * B * W
    out = torch.empty(G, W, dtype=torch.float32, device=x.device)
    with _op("rowgroup_sum", 1):
        check(lib().blm_rowgroup_sum(_ptr(x), G, B, W, _ptr(out), _stream()), "blm_rowgroup_sum")
    return out


def vnn_kl(h: torch.Tensor, rho: torch.Tensor, kl_scale: float, kl_out: Optional[torch.Tensor],
           dh: Optional[torch.Tensor], drho: Optional[torch.Tensor]) -> None:
    """VNN.kl_divergence on the pure last-step hidden h [B, H] (+= into kl_out[0]) and its scaled gradients (+=)."""
    B, H = h.shape
    assert h.is_contiguous() and rho.is_contiguous() and (dh is None or dh.is_contiguous())
    with _op("vnn_kl", 1):
        check(lib().blm_vnn_kl(_ptr(h), _ptr(rho), B, H, float(kl_scale), _ptr(kl_out), _ptr(dh), _ptr(drho), _stream()),
              "blm_vnn_kl")


def vnn_drho(dn: torch.Tensor, e: torch.Tensor, rho: torch.Tensor, drho: torch.Tensor) -> None:
    """drho[j] += exp(rho[j]) sum_t dn[t, j] e[t, j]."""
    T, H = e.shape
    assert dn.is_contiguous() and e.is_contiguous() and dn.shape == e.shape
    with _op("vnn_drho", 1):
        check(lib().blm_vnn_drho(_ptr(dn), _ptr(e), _ptr(rho), T, H, _ptr(drho), _stream()), "blm_vnn_drho")


def gp_lstm_bwd_step(acc5, coef, gate_type: int, c_prev, c_t, dout_t, dh_rec, dc, dc_is_zero: bool, dacc, dacc_s: Split,
                     dcoef) -> None:
    """Backward twin of :func:`gp_lstm_cell` for one timestep (see ``blm_gp_lstm_bwd_step``)."""
    B, H = c_t.shape
    for x in (c_prev, c_t, dout_t, dc, coef, dcoef):
        assert x.is_contiguous()
    assert acc5.stride(1) == 1 and dacc.stride(1) == 1 and dacc_s.hi.stride() == dacc.stride()
    with _op("gp_lstm_bwd_step", 1):
        check(lib().blm_gp_lstm_bwd_step(_ptr(acc5), acc5.stride(0), _ptr(coef), coef.shape[0], gate_type, _ptr(c_prev),
                                         _ptr(c_t), _ptr(dout_t), _ptr(dh_rec), _ptr(dc), int(dc_is_zero), B, H, _ptr(dacc),
                                         _ptr(dacc_s.hi), _ptr(dacc_s.lo), dacc.stride(0), _ptr(dcoef), _stream()),
              "blm_gp_lstm_bwd_step")
