"""Fine-tune step: the B200 counterpart of one iteration of the reference's ``train()`` loop
(``steps/pytorchnn/train.py:306-438``; ``train.py`` below) for the Transformer families.

    loss = CE(mean over the batch's tokens) + KL * kl_scale,  kl_scale = seq_len / len(train_data)
    clip_grad_norm_(parameters, clip);  SGD(lr, momentum=0.9, weight_decay=0)       (train.py:412-420, 466)

Which KL is added follows train.py:335-399: Bayes-FFN -> layer 0 ``linear2``; Bayes-MHA -> layer 0
``self_attn.o_net``; GP -> layer 0 ``gpnn`` (gauss_pos 1..3); Variational -> the V layers of ``T_v_pos``.

How the work is laid out on the GPU
  * the batch (T, B) is packed sequence-major ([M = B*T, width] activations, row = b*T + t), so the
    forward pass is the rescoring pass of :mod:`bayeslms_b200.engine` with the activations kept;
  * every contraction of the backward pass (dgrad dX = dY W, wgrad dW = dY^T X) is a ``blm_gemm`` call on
    bf16 (hi[, lo]) operands -- transposed copies come from ``blm_transpose_*`` -- with the activation
    derivative (GELU / GP mixture) fused into the dgrad epilogue and the softmax gradient fused into the
    decoder product (the [M, V] logits never exist; dZ = softmax - onehot is produced directly as bf16);
  * parameters, gradients and momentum live in three flat fp32 buffers: one reduction gives the global
    gradient norm, one kernel applies clip + momentum + update, and data-parallel training needs a single
    NCCL all-reduce of the gradient buffer (replicated weights, batch sharded over ranks).
Dropout (train.py:75, model.py:116,913,1039-1045,218-221): every nn.Dropout of the reference's training forward is
applied -- embedding + positional output, attention probabilities (inside the attention kernels, forward and
backward), the two residual branches and the FFN activation of every layer (layer 0 of the Bayesian FFN / MHA models
with the hard-coded 0.2 of model.py:1202,1207), and the LSTM's input / output.  Keep masks are Philox(seed, site,
element) on the device, re-derived in the backward pass, or injected multiplier tensors (``masks``, the oracle's
layout) for parity.  Noise is injected (``eps``) or Philox (``seed``).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Optional

import torch

from . import _lib, engine, ops
from .ops import (ACT_GELU, ACT_GELU_GRAD, ACT_GPMIX, ACT_GPMIX_GRAD, ACT_NONE, ACT_SOFTMAX_GRAD, Split)

# Weight gradients dW = dY^T X and input gradients dX = dY W consume their "transposed" operands MN-major straight
# from the [tokens, features] activations / [N, K] weights (blm_gemm_desc.a_mn / b_mn, tcgen05 MN-major shared
# memory descriptors): the 42 transpose + 21 split-transpose launches of a step (a sixth of its time) are gone.
# BLM_TRAIN_TRANSPOSE=1 restores the materialised transposes (A/B switch).
_MN = os.environ.get("BLM_TRAIN_TRANSPOSE") is None
# the optimiser kernel writes the bf16 operand copies of the updated weights (BLM_TRAIN_NO_MIRROR=1: split per step)
_MIRROR = os.environ.get("BLM_TRAIN_NO_MIRROR") is None
# Data parallel, opt-in (BLM_TRAIN_OVERLAP=1): all-reduce of the upper layers' gradients overlapped with the rest of
# the backward pass (two CUDA graphs).  Correct (tools/ddp_train_check.py) but MEASURED SLOWER on 8 GPUs: 4.11 vs
# 4.01 ms per step -- only 38 % of the bytes can start early (the tied embedding gradient completes last), the NCCL
# kernel takes SMs from a backward pass made of 20-70 us kernels, and three small all-reduces pay three latencies.
_OVERLAP = os.environ.get("BLM_TRAIN_OVERLAP") is not None
# Captured step: the weight / bias gradients of the plain projections (dW = dY^T X, db = colsum dY) are off the critical
# path of the backward pass (the dX chain), and at 3200 tokens every GEMM of the step leaves SMs idle (16-128 CTAs on
# 148 SMs) -- so they are captured on a second stream and fill those SMs while the dX chain proceeds
# (BLM_TRAIN_WGRAD_STREAM=0 is the A/B switch).
_WGRAD_STREAM = os.environ.get("BLM_TRAIN_WGRAD_STREAM", "1") != "0"
# Data parallel, captured step: the NCCL all-reduce of each layer's gradient range is captured INSIDE the backward graph,
# on the second stream, as soon as that layer's gradients are final -- it runs while the dX chain works through the lower
# layers; only the embedding / decoder range (final at the very end) is reduced on the critical path
# (BLM_TRAIN_NCCL_GRAPH=0: one eager all-reduce of the whole buffer between the two graphs, the r01 scheme).
_NCCL_IN_GRAPH = os.environ.get("BLM_TRAIN_NCCL_GRAPH", "0") != "0"


class _T:
    """An operand that stands for the TRANSPOSE of the Split it holds."""
    __slots__ = ("s",)

    def __init__(self, s: Split):
        self.s = s

    def rows(self, sl: slice) -> "_T":
        """Rows ``sl`` of the transpose = columns of the held tensor."""
        return _T(Split(self.s.hi[:, sl], None if self.s.lo is None else self.s.lo[:, sl]))


def _gemm(a, b, **kw):
    a_mn, b_mn = isinstance(a, _T), isinstance(b, _T)
    return ops.gemm(a.s if a_mn else a, b.s if b_mn else b, a_mn=a_mn, b_mn=b_mn, **kw)


def _tsplit(x: torch.Tensor, prec: str, have: Optional[Split] = None):
    """Transpose of an fp32 tensor as a GEMM operand (``have``: an existing Split of x to reuse)."""
    if _MN:
        return _T(have if have is not None else ops.split(x, prec))
    return ops.transpose_split(x, prec)


def _tbf16(s: Split, prec: str):
    return _T(s) if _MN else ops.transpose_bf16(s, prec)

_TID = engine._TID
# Philox streams of the dropout sites: embedding / positional output, LSTM input and output, then four per layer
_TID_DROP_PE, _TID_DROP_EMB, _TID_DROP_OUT, _TID_DROP_LAYER = 48, 49, 50, 64
_DROP_SITES = {"attn": 0, "d1": 1, "ffn": 2, "d2": 3}
_TID_VNN = 40              # Philox streams of the Variational-LSTM per-step noise: 40 + cell index
_TID_GPCELL = 24           # ... and of the GP-LSTM cells' unit samples: 24 + 3 * cell index + {coef, weights, bias}
V_NOISE_STD = engine.V_NOISE_STD
_TID_VNOISE = engine.V_NOISE_TID   # Philox stream ids of the V-layer noise: 32 + layer index


def _fusable(drop, n_cols: int):
    """The dropout site as a GEMM-epilogue argument (``blm_gemm_desc.drop`` needs N % 32 == 0)."""
    if drop is not None and n_cols % 32 != 0:
        raise _lib.BlmError(f"dropout fused into the projection needs a width that is a multiple of 32, got {n_cols}")
    return drop


class FineTuner:
    """Owns the flat parameter / gradient / momentum buffers of ``model`` and runs training steps."""

    def __init__(self, model, lr: float, *, momentum: float = 0.9, clip: float = 0.25, prec: str = "bf16x3",
                 group=None, data_parallel: Optional[bool] = None):
        if model.family not in ("bayes_tm", "gauss_tm", "v_tm", "std_tm", "bayes_lstm", "std_lstm", "gauss_lstm", "v_lstm"):
            raise NotImplementedError(f"no fine-tune step for the {model.family} family")
        self.hidden = None
        self.model, self.lr, self.momentum, self.clip, self.prec = model, float(lr), float(momentum), float(clip), prec
        self.group = group
        self.world, self.rank = 1, 0
        dist_up = torch.distributed.is_available() and torch.distributed.is_initialized()
        if data_parallel is None:
            data_parallel = group is not None or dist_up     # default: data parallel whenever a process group exists
        if data_parallel:
            if not dist_up:
                raise _lib.BlmError("data_parallel=True needs an initialised torch.distributed process group")
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        self._keep, self._forked, self._side = [], False, None     # side-stream launches of the captured step (_aside)
        self._ar_in_graph, self._ar_works = False, []
        self._drop = None          # per-step dropout state, see _begin_dropout
        self._seed_dev = None      # device int64 [1] added to the dropout key inside captured graphs
        named = list(model.named_parameters())          # tied encoder / decoder weight appears once
        dev = named[0][1].device
        if dev.type != "cuda":
            raise _lib.BlmError("the model must live on a CUDA device: bayeslms_b200 has no CPU path")
        _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
        self.device = dev
        offs, total = {}, 0
        for name, p in named:
            offs[name] = total
            total += (p.numel() + 7) // 8 * 8          # 16-byte aligned views of the fp32 AND the bf16 mirror buffers
        self._offs = {name: (offs[name], (p.numel() + 7) // 8 * 8) for name, p in named}
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for name, p in named:
                o, n = offs[name], p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
                self.g[name] = self.flat_g[o:o + n].view(p.shape)
        if model.decoder.weight is model.encoder.weight:
            self.g["decoder.weight"] = self.g["encoder.weight"]
        # the baseline models keep torch's key names (model.py:23-171); the step addresses gradients by the names of
        # the Bayesian containers, so alias them
        for name, t in list(self.g.items()):
            alias = _engine_name(name)
            if alias != name:
                self.g[alias] = t
        # bf16 (hi[, lo]) mirrors of the parameters: written by the optimiser kernel together with the update, so a
        # step splits no weight (21 launches); re-made from scratch whenever a parameter was written from outside
        self.flat_hi = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.flat_lo = torch.zeros(total, dtype=torch.bfloat16, device=dev) if prec == "bf16x3" else None
        self._params = [p for _, p in named]
        self._seen_version = None
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_buf = torch.zeros(3, dtype=torch.float32, device=dev)   # ce, kl, loss

    # ------------------------------------------------------------------ helpers
    def refresh_weights(self) -> None:
        """Re-make the bf16 mirrors from the fp32 parameters (after load_state_dict or any external write)."""
        with torch.no_grad():
            _lib.check(_lib.lib().blm_split_bf16(ops._ptr(self.flat_p), ops._ptr(self.flat_hi), ops._ptr(self.flat_lo),
                                                 self.flat_p.numel(), ops._stream()), "blm_split_bf16")
        self._seen_version = sum(p._version for p in self._params)

    def _check_mirrors(self) -> None:
        if self._seen_version != sum(p._version for p in self._params):
            self.refresh_weights()

    def _mirror(self, w: torch.Tensor) -> Optional[Split]:
        """The (hi[, lo]) mirror views of a parameter tensor that lives in the flat buffer, else None."""
        if w.dtype != torch.float32 or not w.is_contiguous() or not _MIRROR:
            return None
        off = w.data_ptr() - self.flat_p.data_ptr()
        if off < 0 or off + 4 * w.numel() > 4 * self.flat_p.numel() or off % 32:
            return None
        o, n = off // 4, w.numel()
        return Split(self.flat_hi[o:o + n].view(w.shape), None if self.flat_lo is None else self.flat_lo[o:o + n].view(w.shape))

    def _w(self, w: torch.Tensor) -> Split:
        return self._mirror(w.detach()) or ops.split(w.detach(), self.prec)

    def _w2(self, w: torch.Tensor):
        """(Split [N, K], Split [K, N]) of an fp32 weight: forward B operand and dgrad B operand, one pass."""
        if _MN:
            sp = self._w(w.detach().float().contiguous())
            return sp, _T(sp)
        return ops.split_transpose(w.detach().float().contiguous(), self.prec)

    def _wt(self, w: torch.Tensor) -> Split:
        """[N, K] fp32 weight -> Split of its transpose [K, N] (B operand of the dgrad product)."""
        return _tsplit(w.detach().float().contiguous(), self.prec)

    def _reparam32(self, mu, lgstd, tid, eps_t, seed):
        """fp32 sample mu + exp(lgstd) * eps of one tensor (injected eps or Philox stream (tid, 0))."""
        if eps_t is not None:
            w, _ = ops.reparam(mu, lgstd, eps=eps_t.to(self.device).float(), prec="bf16", want_f32=True)
        else:
            w, _ = ops.reparam(mu, lgstd, seed=seed, stream_id=engine._stream_id(tid, 0), prec="bf16", want_f32=True)
        return w.view(mu.shape)

    def _f32(self, rows, cols):
        return torch.empty(rows, cols, dtype=torch.float32, device=self.device)

    def _wgrad(self, dy_t: Split, x_t: Split, out: torch.Tensor, tag: str):
        """out[N, K] = dY^T X, both operands given transposed ([N, M] and [K, M])."""
        _gemm(dy_t, x_t, prec=self.prec, out_f32=out, tag="wgrad:" + tag)

    def _aside(self, *keep):
        """Context for launches that are off the critical path: inside a graph capture they go to a second stream,
        ordered after everything queued so far on the main stream, and are joined by :meth:`_join_aside`.  ``keep``:
        every tensor those launches touch -- held until the join so that the allocator cannot hand their memory to a
        main-stream tensor while the side stream still uses it."""
        import contextlib
        if not (_WGRAD_STREAM and getattr(self, "_capturing", False) and torch.cuda.is_current_stream_capturing()):
            return contextlib.nullcontext()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(torch.cuda.current_stream())
        self._keep.extend(keep)
        self._forked = True
        return torch.cuda.stream(self._side)

    def _layer_range(self, li: int):
        """[lo, hi) of layer ``li``'s parameters in the flat buffers (named_parameters keeps a layer contiguous)."""
        pre = ("transformerlayers.layers." if self.model.family == "std_tm" else "transformerlayers.") + f"{li}."
        spans = [(o, o + n) for name, (o, n) in self._offs.items() if name.startswith(pre)]
        return min(a for a, _ in spans), max(b for _, b in spans)

    def _allreduce_in_graph(self, lo: int, hi: int, final: bool = False):
        """Captured all-reduce of flat_g[lo:hi].  Non-final ranges go to the second stream (asynchronous: the compute
        of the lower layers overlaps the transfer); everything is waited for when the final range has been issued."""
        dist = torch.distributed
        if not final:
            with self._aside():
                self._ar_works.append(dist.all_reduce(self.flat_g[lo:hi], group=self._background_group(), async_op=True))
            return
        if hi > lo:
            self._ar_works.append(dist.all_reduce(self.flat_g[lo:hi], group=self.group, async_op=True))
        for w in self._ar_works:
            w.wait()
        self._ar_works = []

    def _background_group(self):
        """Communicator of the overlapped all-reduces: few NCCL CTAs (BLM_TRAIN_NCCL_CTAS, default 4), so that the
        transfer trickles along behind the backward kernels instead of taking their SMs."""
        if getattr(self, "_bg_group", None) is None:
            dist = torch.distributed
            n = int(os.environ.get("BLM_TRAIN_NCCL_CTAS", "4"))
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = n
                opts.config.min_ctas = 1
                self._bg_group = dist.new_group(pg_options=opts)
                dist.all_reduce(torch.zeros(1, device=self.device), group=self._bg_group)   # communicator built now
                torch.cuda.synchronize()
            except Exception:
                self._bg_group = self.group
        return self._bg_group

    def _join_aside(self):
        if getattr(self, "_forked", False):
            torch.cuda.current_stream().wait_stream(self._side)
            self._forked = False
        self._keep.clear()

    def _static_noise(self, e, what: str):
        """While a CUDA graph is being captured every sampled tensor must read its noise from a static buffer that
        step_captured() refills: a missing one would silently train on the posterior mean."""
        if e is None and getattr(self, "_capturing", False):
            raise _lib.BlmError(f"capture(): no static noise buffer for the sampled tensor {what}")
        return e

    # ------------------------------------------------------------------ dropout sites
    def _begin_dropout(self, masks, seed, layout: str = "tbd"):
        """Dropout of this step: injected multiplier tensors (``masks``, oracle layout), else Philox keyed by ``seed``
        (or by the device seed word while a graph is being captured); off when neither is given."""
        if masks is not None:
            self._drop = {"masks": _to_device(masks, self.device), "seed": None, "layout": layout}
        elif getattr(self, "_capturing", False):
            self._drop = {"masks": None, "seed": 0, "seed_dev": self._seed_dev}
        elif seed is not None:
            self._drop = {"masks": None, "seed": int(seed), "seed_dev": None}
        else:
            self._drop = None

    def _site(self, p: float, tid: int, key=None, rows=None) -> Optional[ops.Drop]:
        """The dropout of one site, or None when it is the identity.  ``key``: path into the injected masks;
        ``rows`` = (T, B) re-orders an injected (T, B, w) mask to this step's sequence-major rows."""
        st = self._drop
        if st is None:
            return None
        if st["masks"] is not None:
            m = st["masks"]
            for k in key:
                m = m.get(k) if isinstance(m, dict) else None
                if m is None:
                    return None
            if rows is not None and st["layout"] == "tbd":
                T, B = rows
                m = m.view(T, B, -1).permute(1, 0, 2)
            return ops.Drop(p, mask=m.contiguous())
        if p <= 0.0:
            return None
        # per-activation noise: the rank is part of the stream, so data-parallel shards draw different masks
        return ops.Drop(p, seed=st["seed"], stream_id=engine._stream_id(tid, self.rank), seed_dev=st.get("seed_dev"))

    def _layer_site(self, layer, li: int, name: str, rows=None) -> Optional[ops.Drop]:
        dr = self._site(layer.p_drop, _TID_DROP_LAYER + 4 * li + _DROP_SITES[name], (f"layer{li}", name), rows)
        if name == "attn" and dr is not None and dr.mask is not None and dr.mask.shape[-1] % 4:
            # the attention kernels index their [pairs, L, L] multipliers with L = T rounded up to a multiple of 4
            pad = -dr.mask.shape[-1] % 4
            dr.mask = torch.nn.functional.pad(dr.mask, (0, pad, 0, pad)).contiguous()
        return dr

    # ------------------------------------------------------------------ one step
    def forward_backward(self, tokens_tb: torch.Tensor, targets_tb: torch.Tensor, kl_scale: float, *,
                         eps: Optional[dict] = None, seed: Optional[int] = None, v_eps_layout: str = "tbd",
                         hidden=None, on_layer_done=None, masks: Optional[dict] = None):
        """Fills the gradient buffer for the batch (T, B); returns (loss, ce, kl) as 0-dim device tensors.
        ``eps``: injected noise in the oracle's layout ({'layer<i>': ...}; V layers: (T, B, d) tensors
        already scaled by 0.1, or (B, T, d) with ``v_eps_layout="btd"``); ``seed``: Philox noise instead.
        LSTM families: ``hidden`` = (h, c) carried in from the previous batch (zeros if None); the state
        after the batch is left in ``self.hidden``.
        Dropout follows the modules' probabilities: Philox masks keyed by ``seed``, or ``masks`` = injected multiplier
        tensors in the oracle's layout ({'pe': (T,B,d), 'layer<i>': {'attn': (B*nhead,T,T), 'd1','ffn','d2': (T,B,.)}};
        LSTM: {'emb','out'}); with neither (``eps``-only parity calls) the step runs without dropout."""
        self._check_mirrors()
        self._begin_dropout(masks, seed)
        if self.model.family in ("bayes_lstm", "std_lstm"):
            return self._lstm_forward_backward(tokens_tb, targets_tb, kl_scale, hidden, eps, seed)
        if self.model.family in ("gauss_lstm", "v_lstm"):
            return self._cell_forward_backward(tokens_tb, targets_tb, kl_scale, hidden, eps, seed)
        m, prec, dev = self.model, self.prec, self.device
        T, B = tokens_tb.shape
        M, d, nhead = T * B, m.ninp, m.nhead
        V = m.decoder.weight.shape[0]
        if d // nhead != 64:
            raise NotImplementedError("the training attention kernels need head_dim 64")
        eps = _to_device(eps or {}, dev)
        tok = tokens_tb.t().contiguous().view(-1).to(torch.int32)
        tgt = targets_tb.view(T, B).t().contiguous().view(-1).to(torch.int32)
        pos = torch.arange(T, dtype=torch.int32, device=dev).repeat(B)
        offs = torch.arange(0, (B + 1) * T, T, dtype=torch.int32, device=dev)
        scale_q = float(d // nhead) ** -0.5
        self.flat_g.zero_()
        g = self.g

        # ---------------------------------------------------------------- forward, activations kept
        pe = m.pos_encoder.pe.detach()[:, 0, :].float().contiguous()
        emb_variant = getattr(m, "bayes_embed", False)
        E_in = {}
        if emb_variant:
            # x = (E[tok] sqrt(d)) W~^T + pe with W~ = embed_mean + exp(embed_lgstd) eps (model.py:1284-1293)
            _, x0s = ops.embed(tok, None, m.encoder.weight.detach().float(), None, math.sqrt(d), prec=prec, want_f32=False)
            pe_rows, _ = ops.embed(pos, None, pe, None, 1.0, prec="bf16", want_f32=True)
            ee = self._static_noise(eps.get("embed"), "embed_mean")
            sampled = ee is not None or seed is not None
            w_in32 = (self._reparam32(m.embed_mean.detach(), m.embed_lgstd.detach(), _TID["embed"], ee, seed)
                      if sampled else m.embed_mean.detach())
            w_in, w_in_t = self._w2(w_in32)
            x32 = self._f32(M, d)
            xs = ops.empty_split(M, d, prec, dev)
            _gemm(x0s, w_in, prec=prec, resid=pe_rows, out_f32=x32, out=xs, tag="embed_in")
            E_in = {"x0s": x0s, "w_in_t": w_in_t, "eps": ee, "sampled": sampled}
        else:
            x32, xs = ops.embed(tok, pos, m.encoder.weight.detach().float(), pe, math.sqrt(d), prec=prec)
        drop_pe = self._site(m.pos_encoder.p, _TID_DROP_PE, ("pe",), (T, B))      # PositionalEncoding.dropout, model.py:116
        if drop_pe is not None:
            x32, xs = ops.dropout(x32, drop_pe, prec=prec, out_f32=x32)
        saved = []
        for li, layer in enumerate(m.transformerlayers):
            kind, pre = layer.kind, f"transformerlayers.{li}."
            a = layer.self_attn
            S = {"kind": kind, "x32": x32, "xs": xs}
            qkv32 = self._f32(M, 3 * d)
            qkvs = ops.empty_split(M, 3 * d, prec, dev)
            if kind == "bayes_mha":   # separate q / k / v projections, Bayesian bias-free o_net (model.py:931-1019)
                wqkv32 = torch.cat([a.q_net.weight.detach(), a.k_net.weight.detach(), a.v_net.weight.detach()], 0)
                bqkv = torch.cat([a.q_net.bias.detach(), a.k_net.bias.detach(), a.v_net.bias.detach()], 0)
                le = self._static_noise(eps.get(f"layer{li}"), pre + "self_attn.o_net.weight")
                S["wo_eps"], S["wo_sampled"] = le, (le is not None or seed is not None)
                wo32 = (self._reparam32(a.o_net.weight_mean.detach(), a.o_net.weight_lgstd.detach(), _TID["mha_o"], le, seed)
                        if S["wo_sampled"] else a.o_net.weight_mean.detach())
                bo = None
            else:
                wqkv32, bqkv, wo32, bo = a.qkv_net.weight, a.qkv_net.bias.detach(), a.o_net.weight, a.o_net.bias.detach()
            wqkv, S["wqkv_t"] = self._w2(wqkv32)
            _gemm(xs, wqkv, prec=prec, bias=bqkv, col_scale=scale_q,
                     col_scale_cols=d, out_f32=qkv32, out=qkvs, tag="qkv")
            S["drop_attn"] = self._layer_site(layer, li, "attn")                 # on the probabilities, model.py:913
            _, atts = ops.mha_causal_bf16(qkvs, offs, nhead, T, prec=prec, drop=S["drop_attn"])
            y1 = self._f32(M, d)
            wo, S["wo_t"] = self._w2(wo32)
            S["drop_d1"] = self._layer_site(layer, li, "d1", (T, B))              # dropout1, model.py:1039
            # x + dropout1(attention output) in the projection's epilogue
            _gemm(atts, wo, prec=prec, bias=bo, resid=x32, out_f32=y1, tag="o_net", drop=_fusable(S["drop_d1"], d))
            x1_32, x1s = ops.layernorm(y1, layer.norm1.weight.detach(), layer.norm1.bias.detach(), layer.norm1.eps, prec=prec)
            S.update(qkv32=qkv32, atts=atts, y1=y1, x1_32=x1_32, x1s=x1s)
            # first FFN projection (+ GELU or the GP mixture), pre-activation kept
            F = layer.linear2.in_features
            z1 = self._f32(M, F)
            hs = ops.empty_split(M, F, prec, dev)
            if kind == "gauss":
                gp = layer.gpnn
                le = self._static_noise(eps.get(f"layer{li}"), pre + "gpnn") if gp.sample else None
                use_noise = gp.sample and (le is not None or seed is not None)
                w1_32, b1, coef = gp.weights_mean.detach(), gp.bias_mean.detach(), gp.coef_mean.detach()
                S["gp_noise"] = use_noise
                if use_noise:
                    ge = (lambda k: None) if le is None else le.get
                    if gp.gpnn_type in (1, 3):
                        coef = self._reparam32(gp.coef_mean.detach(), gp.coef_lgstd.detach(), _TID["gp_coef"], ge("coef"), seed)
                    if gp.gpnn_type in (2, 3):
                        w1_32 = self._reparam32(gp.weights_mean.detach(), gp.weights_lgstd.detach(), _TID["gp_w"],
                                                ge("weights"), seed)
                        b1 = self._reparam32(gp.bias_mean.detach(), gp.bias_lgstd.detach(), _TID["gp_b"], ge("bias"), seed)
                coef = coef.contiguous()
                w1, S["w1_t"] = self._w2(w1_32)
                act1, bias1, S["coef"] = ACT_GPMIX, b1, coef
            else:
                w1_32 = layer.linear1.weight.detach()
                w1, S["w1_t"] = self._w2(w1_32)
                act1, bias1, coef = ACT_GELU, layer.linear1.bias.detach(), None
                if getattr(layer, "activation", "gelu") == "relu":     # ReLU = the mixture epilogue with the relu row only
                    coef = torch.zeros(4, w1_32.shape[0], dtype=torch.float32, device=dev)
                    coef[2] = 1.0
                    act1, S["coef"] = ACT_GPMIX, coef
            S["drop_ffn"] = self._layer_site(layer, li, "ffn", (T, B))            # dropout on the activation, model.py:1043
            # h = dropout(act(z1)) in the epilogue; z1 (saved for the backward pass) is the unmasked pre-activation
            _gemm(x1s, w1, prec=prec, bias=bias1, act=act1, coef=coef, out=hs, out_pre=z1, tag="ffn1",
                  drop=_fusable(S["drop_ffn"], F))
            S.update(z1=z1, hs=hs)
            # second FFN projection
            if kind == "bayes_ffn":
                le = self._static_noise(eps.get(f"layer{li}"), pre + "linear2.weight")
                S["w2_eps"] = le
                if le is not None or seed is not None:
                    w2_32 = self._reparam32(layer.linear2.weight_mean.detach(), layer.linear2.weight_lgstd.detach(),
                                            _TID["ffn_w2"], le, seed)
                    S["w2_sampled"] = True
                else:
                    w2_32, S["w2_sampled"] = layer.linear2.weight_mean.detach(), False
                b2 = None
            else:
                w2_32, b2 = layer.linear2.weight.detach(), layer.linear2.bias.detach()
            w2, S["w2_t"] = self._w2(w2_32)
            y2 = self._f32(M, d)
            v_active = kind == "v"
            if v_active:
                if T != 100:
                    raise _lib.BlmError("the variational layer is defined for sequence length 100 only "
                                        "(its parameters are (100, 1, d), model.py:2754-2761)")
                f = self._f32(M, d)
                _gemm(hs, w2, prec=prec, bias=b2, out_f32=f, tag="ffn2")
                le = self._static_noise(eps.get(f"layer{li}"), pre + "hiddens")
                if le is None:
                    e_bt = None
                elif v_eps_layout == "btd":
                    e_bt = le.view(M, d)
                else:
                    e_bt = le.permute(1, 0, 2).contiguous().view(M, d)
                rho = layer.hiddens_lgstd.detach().view(T, d)
                S["drop_d2"] = self._layer_site(layer, li, "d2", (T, B))         # dropout2 AFTER the noise, model.py:2801-2803
                S["v_stream"] = engine._stream_id(_TID_VNOISE + li, self.rank)   # per-token noise: differs per DP rank
                if S["drop_d2"] is not None:
                    fp = ops.vnoise_fwd(f, rho, B, T, eps=e_bt, seed=seed, stream_id=S["v_stream"], noise_std=V_NOISE_STD)
                    ops.dropout(fp, S["drop_d2"], resid=x1_32, out_f32=y2)
                else:
                    y2 = ops.vnoise_fwd(f, rho, B, T, eps=e_bt, seed=seed, stream_id=S["v_stream"],
                                        noise_std=V_NOISE_STD, resid=x1_32)   # y2 = x1 + fp
                S.update(f=f, v_eps=e_bt)
                layer._v_state = {"f": f, "B": B, "T": T, "eps": e_bt, "seed": seed, "stream_id": S["v_stream"]}
            else:
                S["drop_d2"] = self._layer_site(layer, li, "d2", (T, B))         # dropout2, model.py:1045
                _gemm(hs, w2, prec=prec, bias=b2, resid=x1_32, out_f32=y2, tag="ffn2", drop=_fusable(S["drop_d2"], d))
            x32, xs = ops.layernorm(y2, layer.norm2.weight.detach(), layer.norm2.bias.detach(), layer.norm2.eps, prec=prec)
            S["y2"] = y2
            saved.append(S)

        if emb_variant:   # F.linear(x, embed_mean.t()): mean only (model.py:1303)
            xs_pre = xs
            em, em_t = self._w2(m.embed_mean.detach())            # em [k, n] (dgrad operand), em_t = embed_mean^T (forward)
            xs = ops.empty_split(M, d, prec, dev)
            _gemm(xs_pre, em_t, prec=prec, out=xs, tag="embed_out")

        dx, ce, kl, loss = self._loss_and_decoder_grads(xs, tgt)
        if emb_variant:   # back through x @ embed_mean: d embed_mean += x^T dout, dx = dout @ embed_mean^T
            dout = dx
            self._wgrad(_tbf16(xs_pre, prec), _tsplit(dout, prec), g["embed_mean"], "embed_out")
            dx = self._f32(M, d)
            _gemm(ops.split(dout, prec), em, prec=prec, out_f32=dx, tag="dgrad:embed_out")

        # ---------------------------------------------------------------- backward: layers
        for li in range(len(saved) - 1, -1, -1):
            S, layer, pre = saved[li], m.transformerlayers[li], f"transformerlayers.{li}."
            kind, a = S["kind"], layer.self_attn
            # without dropout2 / the variational noise dy2 IS the gradient of linear2's output: the LayerNorm backward then
            # also writes its bf16 operand copy and its column sums (= d linear2.bias)
            ln2_fused = S["drop_d2"] is None and kind != "v"
            bias2_done = ln2_fused and kind != "bayes_ffn"
            dy2 = ops.layernorm_bwd(dx, S["y2"], layer.norm2.weight.detach(), layer.norm2.eps, g[pre + "norm2.weight"],
                                    g[pre + "norm2.bias"], prec=prec if ln2_fused else None,
                                    dxsum=g[pre + "linear2.bias"] if bias2_done else None)
            # gradient of the FFN branch = dropout2's mask on dy2 (the residual branch keeps dy2 itself)
            dfs = None
            if ln2_fused:
                dy2, dfs = dy2
                dbr = dy2
            elif S["drop_d2"] is None:
                dbr = dy2
            elif kind == "v":
                dbr = ops.dropout(dy2, S["drop_d2"])[0]
            else:
                dbr, dfs = ops.dropout(dy2, S["drop_d2"], prec=prec)      # masked gradient and its operand copy in one pass
            if kind == "v":
                df, klpart = ops.vnoise_bwd(dbr, S["f"], layer.hiddens_lgstd.detach().view(T, d),
                                            layer.hiddens_mean_p.detach().view(T, d), B, T, kl_scale,
                                            g[pre + "hiddens_lgstd"].view(T, d), g[pre + "hiddens_mean_p"].view(T, d),
                                            eps=S["v_eps"], seed=seed, stream_id=S["v_stream"], noise_std=V_NOISE_STD)
                ops.reduce_sum(klpart.view(-1), kl, scale=0.5 / (M * d), accumulate=True)
            else:
                df = dbr
            if dfs is None:
                dfs = ops.split(df, prec)
            # FFN2: dgrad fused with the activation derivative (and the mask of the FFN dropout), wgrad, bias
            dz1 = self._f32(M, S["z1"].shape[1])
            dz1s = ops.empty_split(M, S["z1"].shape[1], prec, dev)
            relu = kind != "gauss" and getattr(layer, "activation", "gelu") == "relu"
            if relu:
                _gemm(dfs, S["w2_t"], prec=prec, act=ACT_GPMIX_GRAD, aux=S["z1"], coef=S["coef"], out_f32=dz1, out=dz1s,
                      tag="dgrad:ffn2", drop=_fusable(S["drop_ffn"], dz1.shape[1]))
            elif kind == "gauss":
                dh = torch.empty_like(dz1)
                # h = m . act(z1): the mask multiplies dL/dh in the epilogue, before act'(z1) and before dh is stored
                _gemm(dfs, S["w2_t"], prec=prec, act=ACT_GPMIX_GRAD, aux=S["z1"], coef=S["coef"], out_f32=dz1,
                         out=dz1s, out_pre=dh, tag="dgrad:ffn2", drop=_fusable(S["drop_ffn"], dz1.shape[1]))
            else:
                _gemm(dfs, S["w2_t"], prec=prec, act=ACT_GELU_GRAD, aux=S["z1"], out_f32=dz1, out=dz1s,
                      fast_act=(prec == "bf16"), tag="dgrad:ffn2", drop=_fusable(S["drop_ffn"], dz1.shape[1]))
            dft, ht = _tsplit(df, prec, dfs), _tbf16(S["hs"], prec)
            if kind == "bayes_ffn":
                G = g[pre + "linear2.weight_mean"]
                self._wgrad(dft, ht, G, "ffn2")
                lin = layer.linear2
                if S["w2_sampled"]:
                    ops.reparam_bwd(G, lin.weight_lgstd.detach(), G, g[pre + "linear2.weight_lgstd"], eps=S["w2_eps"],
                                    seed=seed, stream_id=engine._stream_id(_TID["ffn_w2"], 0))
                ops.kl_gauss(lin.weight_mean.detach(), lin.weight_lgstd.detach(), kl, accumulate=True)
                ops.kl_gauss_bwd(lin.weight_mean.detach(), lin.weight_lgstd.detach(), kl_scale, G,
                                 g[pre + "linear2.weight_lgstd"])
            else:
                with self._aside(dfs, df, S["hs"]):
                    self._wgrad(dft, ht, g[pre + "linear2.weight"], "ffn2")
                    if not bias2_done:
                        ops.colsum(df, g[pre + "linear2.bias"])
            # FFN1
            dx1 = self._f32(M, d)
            _gemm(dz1s, S["w1_t"], prec=prec, resid=dy2, out_f32=dx1, tag="dgrad:ffn1")
            dz1t, x1t = _tsplit(dz1, prec, dz1s), _tbf16(S["x1s"], prec)
            if kind == "gauss":
                self._gp_backward(layer, pre, S, dz1, dz1t, x1t, dh, kl, kl_scale, eps.get(f"layer{li}"), seed)
            else:
                with self._aside(dz1, dz1s, S["x1s"]):
                    self._wgrad(dz1t, x1t, g[pre + "linear1.weight"], "ffn1")
                    ops.colsum(dz1, g[pre + "linear1.bias"])
            # LayerNorm 1, output projection, attention, QKV projection
            ln1_fused = S["drop_d1"] is None
            bias_o_done = ln1_fused and kind != "bayes_mha"
            dy1 = ops.layernorm_bwd(dx1, S["y1"], layer.norm1.weight.detach(), layer.norm1.eps, g[pre + "norm1.weight"],
                                    g[pre + "norm1.bias"], prec=prec if ln1_fused else None,
                                    dxsum=g[pre + "self_attn.o_net.bias"] if bias_o_done else None)
            if S["drop_d1"] is not None:     # gradient of the attention branch = dropout1's mask on dy1
                do1, dy1s = ops.dropout(dy1, S["drop_d1"], prec=prec)
            else:
                dy1, dy1s = dy1
                do1 = dy1
            datt = self._f32(M, d)
            _gemm(dy1s, S["wo_t"], prec=prec, out_f32=datt, tag="dgrad:o_net")
            dy1t, attt = _tsplit(do1, prec, dy1s), _tbf16(S["atts"], prec)
            if kind == "bayes_mha":
                G, lin = g[pre + "self_attn.o_net.weight_mean"], a.o_net
                self._wgrad(dy1t, attt, G, "o_net")
                if S["wo_sampled"]:
                    ops.reparam_bwd(G, lin.weight_lgstd.detach(), G, g[pre + "self_attn.o_net.weight_lgstd"], eps=S["wo_eps"],
                                    seed=seed, stream_id=engine._stream_id(_TID["mha_o"], 0))
                ops.kl_gauss(lin.weight_mean.detach(), lin.weight_lgstd.detach(), kl, accumulate=True)
                ops.kl_gauss_bwd(lin.weight_mean.detach(), lin.weight_lgstd.detach(), kl_scale, G,
                                 g[pre + "self_attn.o_net.weight_lgstd"])
            else:
                with self._aside(dy1s, do1, S["atts"]):
                    self._wgrad(dy1t, attt, g[pre + "self_attn.o_net.weight"], "o_net")
                    if not bias_o_done:
                        ops.colsum(do1, g[pre + "self_attn.o_net.bias"])
            dqkv = ops.mha_causal_bwd(S["qkv32"], datt, offs, nhead, T, scale_q, prec=prec, drop=S["drop_attn"])
            dx = self._f32(M, d)
            dqkvs = ops.split(dqkv, prec)
            _gemm(dqkvs, S["wqkv_t"], prec=prec, resid=dy1, out_f32=dx, tag="dgrad:qkv")
            dqkvt, xst = _tsplit(dqkv, prec, dqkvs), _tbf16(S["xs"], prec)
            if kind == "bayes_mha":
                for k, nm in enumerate(("q_net", "k_net", "v_net")):
                    rows = slice(k * d, (k + 1) * d)
                    part = dqkvt.rows(rows) if _MN else Split(dqkvt.hi[rows], None if dqkvt.lo is None else dqkvt.lo[rows])
                    self._wgrad(part, xst, g[pre + f"self_attn.{nm}.weight"], nm)
                    ops.colsum(dqkv[:, rows], g[pre + f"self_attn.{nm}.bias"])
            else:
                with self._aside(dqkv, dqkvs, S["xs"]):
                    self._wgrad(dqkvt, xst, g[pre + "self_attn.qkv_net.weight"], "qkv")
                    ops.colsum(dqkv, g[pre + "self_attn.qkv_net.bias"])
            if on_layer_done is not None:
                on_layer_done(li)      # every gradient of layers >= li is final (capture() splits the graph here)
            if self._ar_in_graph:
                self._allreduce_in_graph(*self._layer_range(li))
        if drop_pe is not None:
            dx, _ = ops.dropout(dx, drop_pe, out_f32=dx)
        if emb_variant:   # back through x0 W~^T: G = dx^T x0 (-> embed_mean, embed_lgstd), dx0 = dx W~
            dxt, x0t = _tsplit(dx, prec), _tbf16(E_in["x0s"], prec)
            if E_in["sampled"]:
                G = self._f32(d, d)
                self._wgrad(dxt, x0t, G, "embed_in")
                ops.reparam_bwd(G, m.embed_lgstd.detach(), g["embed_mean"], g["embed_lgstd"], eps=E_in["eps"], seed=seed,
                                stream_id=engine._stream_id(_TID["embed"], 0), accumulate=True)
            else:   # added onto the output-side gradient already in place (residual operand = output)
                _gemm(dxt, x0t, prec=prec, resid=g["embed_mean"], out_f32=g["embed_mean"], tag="wgrad:embed_in")
            ops.kl_gauss(m.embed_mean.detach(), m.embed_lgstd.detach(), kl, accumulate=True)
            ops.kl_gauss_bwd(m.embed_mean.detach(), m.embed_lgstd.detach(), kl_scale, g["embed_mean"], g["embed_lgstd"])
            dx0 = self._f32(M, d)
            _gemm(ops.split(dx, prec), E_in["w_in_t"], prec=prec, out_f32=dx0, tag="dgrad:embed_in")
            dx = dx0
        # embedding: scatter-add on top of the decoder's weight gradient when the weights are tied (so the side stream's
        # weight gradients are joined first)
        self._join_aside()
        ops.embed_bwd(dx, tok, math.sqrt(d), g["encoder.weight"])
        # loss = ce + kl * kl_scale
        ops.reduce_sum(ce, loss)
        ops.reduce_sum(kl, loss, scale=float(kl_scale), accumulate=True)
        if self._ar_in_graph:
            # what is not a layer (embedding / decoder, positional-free extras) sits in [0, first layer) and
            # [last layer end, total): final only now.  Two ranges -> the second one closes the wait list.
            nl = len(m.transformerlayers)
            first_lo = self._layer_range(0)[0] if nl else self.flat_g.numel()
            last_hi = self._layer_range(nl - 1)[1] if nl else self.flat_g.numel()
            if first_lo > 0:
                self._ar_works.append(torch.distributed.all_reduce(self.flat_g[:first_lo], group=self.group, async_op=True))
            self._allreduce_in_graph(last_hi, self.flat_g.numel(), final=True)
        return self.loss_buf[2], self.loss_buf[0], self.loss_buf[1]

    def _loss_and_decoder_grads(self, xs: Split, tgt: torch.Tensor):
        """Cross entropy (mean over the M rows) through the vocabulary-streaming kernel, then the decoder's
        backward: dZ = (softmax - onehot) / M produced directly as bf16 by the fused product (the [M, V]
        logits never exist), dx = dZ E, dE += dZ^T x, db = colsum(dZ).  Returns (dx fp32 [M, d], ce, kl, loss)
        with kl zeroed."""
        m, prec, dev, g = self.model, self.prec, self.device, self.g
        M, d = xs.hi.shape
        V = m.decoder.weight.shape[0]
        E32 = m.decoder.weight.detach().float()
        Es, Et = self._w2(E32)
        dec_b = m.decoder.bias.detach()
        lse = torch.empty(M, dtype=torch.float32, device=dev)
        nll = ops.vocab_nll(xs, Es, dec_b, tgt, prec=prec, lse=lse)
        ce, kl, loss = self.loss_buf[0:1], self.loss_buf[1:2], self.loss_buf[2:3]
        ops.reduce_sum(nll, ce, scale=1.0 / M)
        kl.zero_()
        ldv = ops._ld8(V)
        dZ = Split(torch.empty(M, ldv, dtype=torch.bfloat16, device=dev)[:, :V],
                   torch.empty(M, ldv, dtype=torch.bfloat16, device=dev)[:, :V] if prec == "bf16x3" else None)
        _gemm(xs, Es, prec=prec, bias=dec_b, act=ACT_SOFTMAX_GRAD, lse=lse, targets=tgt, grad_scale=1.0 / M, out=dZ,
                 tag="dlogits")
        dx = self._f32(M, d)
        _gemm(dZ, Et, prec=prec, out_f32=dx, tag="dgrad:decoder")
        with self._aside(dZ, xs):
            self._wgrad(_tbf16(dZ, prec), _tbf16(xs, prec), g["decoder.weight"], "decoder")
            ops.colsum(dZ, g["decoder.bias"])
        return dx, ce, kl, loss

    # ------------------------------------------------------------------ LSTM families
    def _lstm_layer_params(self, layer: int, eps: dict, seed, sampled: bool):
        """fp32 (W_ih, W_hh, b_ih + b_hh) of one layer for this step: the means, with the Bayesian gate's
        rows replaced by mu + exp(lgstd) eps when sampling (model.py:668-725)."""
        r = self.model.rnn
        rows = r.gate_rows() if sampled else None
        out = {}
        for name in ("ih", "hh"):
            mu = getattr(r, f"weight_{name}_mean_{layer}").detach()
            if sampled:
                key = f"weight_{name}_{layer}"
                w = mu.clone()
                sub = self._reparam32(mu[rows], getattr(r, f"weight_{name}_lgstd_{layer}").detach(),
                                      _TID["lstm"] + engine.LSTM_EPS_ORDER.index(key), eps.get(key), seed)
                w[rows] = sub
                mu = w
            out["w_" + name] = mu
        bias = getattr(r, f"bias_ih_mean_{layer}").detach() + getattr(r, f"bias_hh_mean_{layer}").detach()
        if sampled:
            cur = bias[rows].contiguous()
            for name in ("ih", "hh"):
                key = f"bias_{name}_{layer}"
                cur = self._reparam32(cur, getattr(r, f"bias_{name}_lgstd_{layer}").detach(),
                                      _TID["lstm"] + engine.LSTM_EPS_ORDER.index(key), eps.get(key), seed)
            bias[rows] = cur
        out["bias"] = bias
        return out

    def _lstm_forward_backward(self, tokens_tb, targets_tb, kl_scale, hidden, eps, seed):
        """BayesRNNModel step (model.py:217-222, train.py:319-340): embedding -> two LSTM layers -> decoder.
        Rows are time-major (row = t*B + b).  Forward keeps each layer's input and output; the backward pass
        rebuilds the gate pre-activations with one GEMM per layer, walks the recurrence backwards (one small
        kernel + one [B, 4H] x [4H, H] product per step) and finishes with three GEMMs per layer (input
        gradient and the two weight gradients over all T*B rows at once).  The incoming hidden state is a
        constant (train.py repackages it); the outgoing one is kept in ``self.hidden``."""
        m, prec, dev, g = self.model, self.prec, self.device, self.g
        r = m.rnn
        T, B = tokens_tb.shape
        M, H = T * B, m.nhid
        eps = _to_device(eps or {}, dev)
        if hidden is None:
            hidden = m.init_hidden(B)
        h0, c0 = hidden[0].detach().float().contiguous(), hidden[1].detach().float().contiguous()
        bayes = 1 <= r.position <= 4
        sampled = bayes and (len(eps) > 0 or seed is not None)
        tok = tokens_tb.reshape(-1).to(torch.int32)
        tgt = targets_tb.reshape(-1).to(torch.int32)
        lengths = torch.full((B,), T, dtype=torch.int32, device=dev)
        self.flat_g.zero_()

        # self.drop on the embedding and on the LSTM output (model.py:218,220); rows are time-major like the oracle's
        drop_emb = self._site(m.p_drop, _TID_DROP_EMB, ("emb",))
        drop_out = self._site(m.p_drop, _TID_DROP_OUT, ("out",))
        x32, x = ops.embed(tok, None, m.encoder.weight.detach().float(), None, 1.0, prec=prec, want_f32=drop_emb is not None)
        if drop_emb is not None:
            _, x = ops.dropout(x32, drop_emb, prec=prec, want_f32=False)
        saved, hT, cT = [], [], []
        for li in range(2):
            P = self._lstm_layer_params(li + 1, eps, seed, sampled)
            w_ih, w_ih_t = self._w2(P["w_ih"])
            w_hh, w_hh_t = self._w2(P["w_hh"])
            gates = self._f32(M, 4 * H)
            _gemm(x, w_ih, prec=prec, bias=P["bias"], out_f32=gates, tag=f"lstm_in{li + 1}")
            out32, out, h_last, c_last = ops.lstm_layer(gates, w_hh, h0[li], c0[li], lengths, T, B, H, prec=prec,
                                                        want_f32=(li == 1 and drop_out is not None), want_split=True)
            saved.append({"x": x, "out": out, "gates": gates, "w_ih_t": w_ih_t, "w_hh": w_hh, "w_hh_t": w_hh_t})
            hT.append(h_last)
            cT.append(c_last)
            x = out
        self.hidden = (torch.stack(hT), torch.stack(cT))
        if drop_out is not None:
            _, x = ops.dropout(out32, drop_out, prec=prec, want_f32=False)

        dout, ce, kl, loss = self._loss_and_decoder_grads(x, tgt)
        if drop_out is not None:
            dout, _ = ops.dropout(dout, drop_out, out_f32=dout)

        for li in (1, 0):
            S, layer = saved[li], li + 1
            h0s = ops.split(h0[li], prec)
            hprev = Split(torch.cat([h0s.hi, S["out"].hi[:M - B]], 0),
                          None if h0s.lo is None else torch.cat([h0s.lo, S["out"].lo[:M - B]], 0))
            gates = S["gates"]
            _gemm(hprev, S["w_hh"], prec=prec, resid=gates, out_f32=gates, tag="lstm_rebuild")
            c_all = ops.lstm_gates_act(gates, c0[li], T, B, H)
            dG = self._f32(M, 4 * H)
            dGs = ops.empty_split(M, 4 * H, prec, dev)
            dc = self._f32(B, H)
            dh = [self._f32(B, H), self._f32(B, H)]
            rec = None
            for t in range(T - 1, -1, -1):
                sl = slice(t * B, (t + 1) * B)
                c_prev = c0[li] if t == 0 else c_all[(t - 1) * B:t * B]
                dGt = Split(dGs.hi[sl], None if dGs.lo is None else dGs.lo[sl])
                ops.lstm_bwd_step(gates[sl], c_prev, c_all[sl], dout[sl], rec, dc, t == T - 1, dG[sl], dGt)
                if t > 0:
                    rec = dh[t & 1]
                    _gemm(dGt, S["w_hh_t"], prec=prec, out_f32=rec, tag="lstm_dh")
            # input gradient, weight gradients, biases
            dGT = _tsplit(dG, prec)
            pre = "rnn."
            G_ih, G_hh = g[f"{pre}weight_ih_mean_{layer}"], g[f"{pre}weight_hh_mean_{layer}"]
            self._wgrad(dGT, _tbf16(S["x"], prec), G_ih, f"lstm_ih{layer}")
            self._wgrad(dGT, _tbf16(hprev, prec), G_hh, f"lstm_hh{layer}")
            gb_ih, gb_hh = g[f"{pre}bias_ih_mean_{layer}"], g[f"{pre}bias_hh_mean_{layer}"]
            ops.colsum(dG, gb_ih)
            ops.colsum(dG, gb_hh)
            dx = self._f32(M, S["x"].hi.shape[1])
            _gemm(dGs, S["w_ih_t"], prec=prec, out_f32=dx, tag="dgrad:lstm_in")
            dout = dx
            if bayes:
                rows = r.gate_rows()
                terms = [(f"weight_hh_{layer}", G_hh, H / float(H + r.input_size)),
                         (f"weight_ih_{layer}", G_ih, r.input_size / float(H + r.input_size)),
                         (f"bias_hh_{layer}", gb_hh, 0.5), (f"bias_ih_{layer}", gb_ih, 0.5)]
                for key, G, frac in terms:
                    name = key.rsplit("_", 1)[0]
                    lg = getattr(r, f"{name}_lgstd_{layer}").detach()
                    mu = getattr(r, f"{name}_mean_{layer}").detach()
                    g_lg = g[f"{pre}{name}_lgstd_{layer}"]
                    if sampled:
                        ops.reparam_bwd(G[rows], lg, G[rows], g_lg, eps=eps.get(key), seed=seed,
                                        stream_id=engine._stream_id(_TID["lstm"] + engine.LSTM_EPS_ORDER.index(key), 0))
                    if layer == 1:   # the reference's KL only sees the layer-1 tensors (model.py:736-765)
                        ops.kl_gauss(mu[rows], lg, kl, scale=frac, accumulate=True)
                        ops.kl_gauss_bwd(mu[rows], lg, kl_scale * frac, G[rows], g_lg)
        if drop_emb is not None:
            dout, _ = ops.dropout(dout, drop_emb, out_f32=dout)
        self._join_aside()
        ops.embed_bwd(dout, tok, 1.0, g["encoder.weight"])
        ops.reduce_sum(ce, loss)
        ops.reduce_sum(kl, loss, scale=float(kl_scale), accumulate=True)
        return self.loss_buf[2], self.loss_buf[0], self.loss_buf[1]

    # ------------------------------------------------------------------ GP-LSTM / Variational-LSTM cell families
    def _cell_layers(self):
        """The recurrent stack of a GaussRNNModel / VariationalRNNModel as a flat list of layers (model.py:1619-1636,
        2436-2437): nn.LSTM members contribute plain layers, GPLSTMCell / VLSTMCell one layer each."""
        layers = []
        for mi, member in enumerate(self.model.rnn.rnn):
            pre = f"rnn.rnn.{mi}."
            if isinstance(member, torch.nn.LSTM):
                for l in range(member.num_layers):
                    layers.append({"kind": "plain", "pre": pre, "mi": mi, "mod": member,
                                   "names": {k: f"{pre}{k}_l{l}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")}})
            elif getattr(member, "kind", "") == "gp":
                if member.gate_type in (5, 6):   # GP unit on the cell state / in place of the recurrent product: per-step GEMMs
                    layers.append({"kind": "gp56", "pre": pre, "mi": mi, "mod": member})
                    continue
                if member.gate_type == 7:      # gates = GPNN(x) + W_hh h + b_ih: a plain layer behind a GP input transform
                    layers.append({"kind": "gp7", "pre": pre, "mi": mi, "mod": member,
                                   "names": {"weight_hh": pre + "weights_hh", "bias_ih": pre + "bias_ih"}})
                    continue
                layers.append({"kind": "gp", "pre": pre, "mi": mi, "mod": member})
            else:
                layers.append({"kind": "v", "pre": pre, "mi": mi, "mod": member,
                               "names": {"weight_ih": pre + "weights_ih", "weight_hh": pre + "weights_hh",
                                         "bias_ih": pre + "bias_ih", "bias_hh": pre + "bias_hh"}})
        return layers

    def _cell_forward_backward(self, tokens_tb, targets_tb, kl_scale, hidden, eps, seed):
        """One step of GaussRNNModel / VariationalRNNModel (model.py:1354-1359, 2411-2416; KL terms of train.py:360-377).
        Rows are time-major (row = t*B + b).
          * plain layers and variational cells run on the persistent recurrence kernel.  A variational cell built with
            vnn_type 1 adds n_t = e_t exp(hidden_lgstd), e_t ~ N(0, 0.1^2) of shape (1, H), to h after every step
            (model.py:2506-2507): additive noise commutes with the recurrent product, so W_hh n_{t-1} enters the hoisted
            input gates as a per-step bias row and the kernel runs on the pure hidden state; its KL
            (VNN.kl_divergence, model.py:2545-2551) is taken on the pure hidden of the last step;
          * a GP cell steps through one [B, 5H] product per timestep on [W_hh; W_g(h part)] (gate replaced by the GP
            mixture, bias_ih doubled, model.py:1743-1777), keeps every pre-activation, and is walked backwards by
            blm_gp_lstm_bwd_step; its unit samples (coef, weights, bias) once per step only when GPNN.sample is set.
        ``eps``: {'cell<mi>': (T, 1, H) noise (variational) | {'coef','weights','bias'} (GP)}; else Philox(``seed``)."""
        m, prec, dev, g = self.model, self.prec, self.device, self.g
        T, B = tokens_tb.shape
        M, H = T * B, m.nhid
        eps = _to_device(eps or {}, dev)
        if hidden is None:
            hidden = m.init_hidden(B)
        h0, c0 = hidden[0].detach().float().contiguous(), hidden[1].detach().float().contiguous()
        tok = tokens_tb.reshape(-1).to(torch.int32)
        tgt = targets_tb.reshape(-1).to(torch.int32)
        lengths = torch.full((B,), T, dtype=torch.int32, device=dev)
        self.flat_g.zero_()
        layers = self._cell_layers()
        drop_emb = self._site(m.p_drop, _TID_DROP_EMB, ("emb",))
        drop_out = self._site(m.p_drop, _TID_DROP_OUT, ("out",))
        x32, x = ops.embed(tok, None, m.encoder.weight.detach().float(), None, 1.0, prec=prec, want_f32=drop_emb is not None)
        if drop_emb is not None:
            _, x = ops.dropout(x32, drop_emb, prec=prec, want_f32=False)
        hT, cT, out32 = [], [], None
        for li, L in enumerate(layers):
            last = li == len(layers) - 1
            mod, mi = L["mod"], L["mi"]
            L["x"] = x
            if L["kind"] == "gp":
                out32, x, h_l, c_l = self._gp_cell_forward(L, x, h0[li], c0[li], lengths, T, B, H, eps.get(f"cell{mi}"), seed)
            elif L["kind"] == "gp56":
                out32, x, h_l, c_l = self._gp56_forward(L, x, h0[li], c0[li], lengths, T, B, H, eps.get(f"cell{mi}"), seed)
            elif L["kind"] == "gp7":
                # gate type 7 (model.py:1749-1750): the GP unit over x is ONE GEMM with the mixture epilogue for all steps
                # (pre-activation z saved), + b_ih, then the plain recurrence on W_hh
                coef, coef4, wg, bg = self._gp_unit_sample(L, eps.get(f"cell{mi}"), seed)
                wgs, L["wg_t"] = self._w2(wg)
                L["w_hh"], L["w_hh_t"] = self._w2(_param(m, L["names"]["weight_hh"]))
                gates, z = self._f32(M, 4 * H), self._f32(M, 4 * H)
                _gemm(x, wgs, prec=prec, bias=bg, act=ACT_GPMIX, coef=coef4, out_f32=gates, out_pre=z, tag="gplstm_in")
                ops.rowgroup_add(gates, _param(m, L["names"]["bias_ih"]).detach().float().view(1, -1).contiguous(), 1, M,
                                 out_f32=gates)
                need32 = last and drop_out is not None
                out32, out, h_l, c_l = ops.lstm_layer(gates, L["w_hh"], h0[li], c0[li], lengths, T, B, H, prec=prec,
                                                      want_f32=need32, want_split=True)
                L.update(gates=gates, out=out, z=z, coef3=coef, out_fed=out)
                x = out
            else:
                named = L.setdefault("params", {k: _param(m, n) for k, n in L["names"].items()})
                w_ih, L["w_ih_t"] = self._w2(named["weight_ih"])
                L["w_hh"], L["w_hh_t"] = self._w2(named["weight_hh"])
                bias = (2.0 * named["bias_ih"]) if L["kind"] == "v" else (named["bias_ih"] + named["bias_hh"])
                gates = self._f32(M, 4 * H)
                _gemm(x, w_ih, prec=prec, bias=bias.contiguous(), out_f32=gates, tag=f"lstm_in{li + 1}")
                noisy = L["kind"] == "v" and mod.vnn_type == 1 and (f"cell{mi}" in eps or seed is not None)
                L["noisy"] = noisy
                if noisy:
                    rho = mod.vnn.hidden_lgstd.detach().view(-1)
                    e = eps.get(f"cell{mi}")
                    E = (e.reshape(T, H).contiguous() if e is not None else
                         ops.philox_normal(seed, engine._stream_id(_TID_VNN + mi, 0), T * H, dev, scale=V_NOISE_STD).view(T, H))
                    N = ops.vnn_noise(E, rho)
                    n_prev = torch.cat([torch.zeros(1, H, dtype=torch.float32, device=dev), N[:-1]], 0)
                    g_n = self._f32(T, 4 * H)
                    _gemm(ops.split(n_prev, prec), L["w_hh"], prec=prec, out_f32=g_n, tag="vnn_bias")   # W_hh n_{t-1}
                    ops.rowgroup_add(gates, g_n, T, B, out_f32=gates)
                    L.update(E=E, N=N)
                need32 = noisy or (last and drop_out is not None)
                out32, out, h_l, c_l = ops.lstm_layer(gates, L["w_hh"], h0[li], c0[li], lengths, T, B, H, prec=prec,
                                                      want_f32=need32, want_split=True)
                L.update(gates=gates, out=out, h_pure=h_l)
                if L["kind"] == "v":
                    mod.vnn.hidden_mean = h_l
                if noisy:
                    out32, x = ops.rowgroup_add(out32, L["N"], T, B, prec=prec)
                    h_l, _ = ops.rowgroup_add(h_l, L["N"][T - 1:T].contiguous(), 1, B)
                else:
                    x = out
                L["out_fed"] = x              # what the next layer / step consumed (noised for a variational cell)
            hT.append(h_l)
            cT.append(c_l)
        self.hidden = (torch.stack(hT), torch.stack(cT))
        if drop_out is not None:
            _, x = ops.dropout(out32, drop_out, prec=prec, want_f32=False)

        dout, ce, kl, loss = self._loss_and_decoder_grads(x, tgt)
        if drop_out is not None:
            dout, _ = ops.dropout(dout, drop_out, out_f32=dout)

        for li in range(len(layers) - 1, -1, -1):
            L = layers[li]
            if L["kind"] == "gp":
                dout = self._gp_cell_backward(L, dout, h0[li], c0[li], T, B, H, kl, kl_scale, eps.get(f"cell{L['mi']}"), seed)
                continue
            if L["kind"] == "gp56":
                dout = self._gp56_backward(L, dout, h0[li], c0[li], T, B, H, kl, kl_scale, eps.get(f"cell{L['mi']}"), seed)
                continue
            mod, named, nm = L["mod"], L.get("params"), L["names"]
            h0s = ops.split(h0[li], prec)
            cat = lambda s: Split(torch.cat([h0s.hi, s.hi[:M - B]], 0),  # noqa: E731
                                  None if h0s.lo is None else torch.cat([h0s.lo, s.lo[:M - B]], 0))
            gates = L["gates"]
            _gemm(cat(L["out"]), L["w_hh"], prec=prec, resid=gates, out_f32=gates, tag="lstm_rebuild")
            c_all = ops.lstm_gates_act(gates, c0[li], T, B, H)
            if L["kind"] == "v" and mod.vnn_type == 1:
                rho = mod.vnn.hidden_lgstd.detach().view(-1)
                g_rho = g[L["pre"] + "vnn.hidden_lgstd"].view(-1)
                s1 = ops.rowgroup_sum(dout, T, B) if L["noisy"] else None       # d n_t from the consumers of h_t + n_t
                # KL on the pure last hidden (train.py:372-377): value, and gradient into the last step's dh
                ops.vnn_kl(L["h_pure"], rho, kl_scale, kl, dout[M - B:], g_rho)
            dG = self._f32(M, 4 * H)
            dGs = ops.empty_split(M, 4 * H, prec, dev)
            dc = self._f32(B, H)
            dh = [self._f32(B, H), self._f32(B, H)]
            rec = None
            for t in range(T - 1, -1, -1):
                sl = slice(t * B, (t + 1) * B)
                c_prev = c0[li] if t == 0 else c_all[(t - 1) * B:t * B]
                dGt = Split(dGs.hi[sl], None if dGs.lo is None else dGs.lo[sl])
                ops.lstm_bwd_step(gates[sl], c_prev, c_all[sl], dout[sl], rec, dc, t == T - 1, dG[sl], dGt)
                if t > 0:
                    rec = dh[t & 1]
                    _gemm(dGt, L["w_hh_t"], prec=prec, out_f32=rec, tag="lstm_dh")
            dGT = _tsplit(dG, prec)
            if L["kind"] == "gp7":
                dout = self._gp7_backward(L, dG, dGT, cat(L["out_fed"]), kl, kl_scale, eps.get(f"cell{L['mi']}"), seed)
                continue
            self._wgrad(dGT, _tbf16(L["x"], prec), g[nm["weight_ih"]], f"lstm_ih{li + 1}")
            self._wgrad(dGT, _tbf16(cat(L["out_fed"]), prec), g[nm["weight_hh"]], f"lstm_hh{li + 1}")
            if L["kind"] == "v":        # bias_ih enters both products, bias_hh never (model.py:2519)
                ops.colsum(dG, g[nm["bias_ih"]], scale=2.0)
            else:
                ops.colsum(dG, g[nm["bias_ih"]])
                ops.colsum(dG, g[nm["bias_hh"]])
            if L["kind"] == "v" and L["noisy"]:
                # d n_t = sum_b (dout_t + dh_rec_t) = s1[t] + (sum_b dG_{t+1}) W_hh ; d rho += exp(rho) sum_t d n_t e_t
                s2 = ops.rowgroup_sum(dG, T, B)
                s2 = torch.cat([s2[1:], torch.zeros(1, 4 * H, dtype=torch.float32, device=dev)], 0)
                dn = self._f32(T, H)
                _gemm(ops.split(s2, prec), L["w_hh_t"], prec=prec, resid=s1, out_f32=dn, tag="vnn_dn")
                ops.vnn_drho(dn, L["E"], rho, g_rho)
            dx = self._f32(M, L["x"].hi.shape[1])
            _gemm(dGs, L["w_ih_t"], prec=prec, out_f32=dx, tag="dgrad:lstm_in")
            dout = dx
        if drop_emb is not None:
            dout, _ = ops.dropout(dout, drop_emb, out_f32=dout)
        self._join_aside()
        ops.embed_bwd(dout, tok, 1.0, g["encoder.weight"])
        ops.reduce_sum(ce, loss)
        ops.reduce_sum(kl, loss, scale=float(kl_scale), accumulate=True)
        return self.loss_buf[2], self.loss_buf[0], self.loss_buf[1]

    def _gp7_backward(self, L, dG, dGT, hprev: Split, kl, kl_scale, le, seed):
        """Tail of the backward pass of a gate-type-7 GP cell: dG [M, 4H] is the gradient of the gate pre-activations
        gates = GPNN(x) + W_hh h + b_ih.  Returns dx."""
        prec, g, mod, pre, mi = self.prec, self.g, L["mod"], L["pre"], L["mi"]
        gp = mod.gpnn
        self._wgrad(dGT, _tbf16(hprev, prec), g[pre + "weights_hh"], "gp7_hh")
        ops.colsum(dG, g[pre + "bias_ih"])                       # bias_ih enters once here, bias_hh and weights_ih never
        gc, gw, gb = g[pre + "gpnn.coef_mean"], g[pre + "gpnn.weights_mean"], g[pre + "gpnn.bias_mean"]
        dz, dzs = ops.gp3_bwd(dG, L["z"], L["coef3"], gc, prec)
        self._wgrad(_tsplit(dz, prec, dzs), _tbf16(L["x"], prec), gw, "gp7_w")
        ops.colsum(dz, gb)
        self._gp_unit_param_grads(L, kl, kl_scale, le, seed)
        dx = self._f32(dz.shape[0], L["x"].hi.shape[1])
        _gemm(dzs, L["wg_t"], prec=prec, out_f32=dx, tag="dgrad:gplstm_in")
        return dx

    def _gp_unit_sample(self, L, le, seed):
        """(coef [3, N], W_g, b_g) of a cell's GP unit for this step: posterior means, or one draw when GPNN.sample is set
        (sample_parameters at model.py:1721-1723), and the coefficient table in the GEMM epilogue's order."""
        gp, mi = L["mod"].gpnn, L["mi"]
        coef, wg, bg = gp.coef_mean.detach(), gp.weights_mean.detach(), gp.bias_mean.detach()
        L["noise"] = bool(gp.sample) and (le is not None or seed is not None)
        if L["noise"]:
            ge = (lambda k: None) if le is None else le.get
            tid = _TID_GPCELL + 3 * mi
            if gp.gpnn_type in (1, 3):
                coef = self._reparam32(coef, gp.coef_lgstd.detach(), tid, ge("coef"), seed)
            if gp.gpnn_type in (2, 3):
                wg = self._reparam32(wg, gp.weights_lgstd.detach(), tid + 1, ge("weights"), seed)
                bg = self._reparam32(bg, gp.bias_lgstd.detach(), tid + 2, ge("bias"), seed)
        coef = coef.float().contiguous()                       # rows: sigmoid, tanh, relu
        coef4 = torch.stack([coef[1], coef[0], coef[2], torch.zeros_like(coef[0])]).contiguous()   # tanh, sigmoid, relu, gelu
        return coef, coef4, wg, bg.reshape(-1).float().contiguous()

    def _gp_unit_param_grads(self, L, kl, kl_scale, le, seed):
        """The GP unit's share of the step once d coef / d W_g / d b_g (of the sampled values) sit in the gradient buffer:
        chain rule into the log-sigmas for a sampled unit, KL value and KL gradients (train.py:360-369)."""
        g, pre, mi, gp = self.g, L["pre"], L["mi"], L["mod"].gpnn
        gc, gw, gb = g[pre + "gpnn.coef_mean"], g[pre + "gpnn.weights_mean"], g[pre + "gpnn.bias_mean"]
        sid = lambda k: engine._stream_id(_TID_GPCELL + 3 * mi + k, 0)  # noqa: E731
        t_ = gp.gpnn_type
        if t_ in (1, 3):
            if L["noise"]:
                ops.reparam_bwd(gc, gp.coef_lgstd.detach(), gc, g[pre + "gpnn.coef_lgstd"],
                                eps=None if le is None else le.get("coef"), seed=seed, stream_id=sid(0))
            ops.kl_gauss(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl_scale, gc, g[pre + "gpnn.coef_lgstd"])
        if t_ in (2, 3):
            if L["noise"]:
                ops.reparam_bwd(gw, gp.weights_lgstd.detach(), gw, g[pre + "gpnn.weights_lgstd"],
                                eps=None if le is None else le.get("weights"), seed=seed, stream_id=sid(1))
                ops.reparam_bwd(gb, gp.bias_lgstd.detach(), gb, g[pre + "gpnn.bias_lgstd"],
                                eps=None if le is None else le.get("bias"), seed=seed, stream_id=sid(2))
            ops.kl_gauss(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl_scale, gw, g[pre + "gpnn.weights_lgstd"])
            ops.kl_gauss(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl_scale, gb, g[pre + "gpnn.bias_lgstd"])

    def _gp56_forward(self, L, x: Split, h0, c0, lengths, T, B, H, le, seed):
        """Gate types 5 ("cell": c <- GPNN(c) before the update, model.py:1763-1764) and 6 ("hidden": gates = W_ih x + b_ih +
        GPNN(h), model.py:1747-1748): the input product is hoisted, the GP unit is one GEMM with the mixture epilogue
        per step (pre-activation saved), then the plain cell update."""
        prec, dev, mod = self.prec, self.device, L["mod"]
        gate, M = mod.gate_type, T * B
        coef, coef4, wg, bg = self._gp_unit_sample(L, le, seed)
        wgs, L["wg_t"] = self._w2(wg)
        w_ih, L["w_ih_t"] = self._w2(mod.weights_ih)
        b_ih = mod.bias_ih.detach().float()
        pre = self._f32(M, 4 * H)
        _gemm(x, w_ih, prec=prec, bias=((2.0 if gate == 5 else 1.0) * b_ih).contiguous(), out_f32=pre, tag="gplstm_in")
        NZ = 4 * H if gate == 6 else H
        acc, zs, c_all = self._f32(M, 4 * H), self._f32(M, NZ), self._f32(M, H)
        out32 = self._f32(M, H)
        outs = ops.empty_split(M, H, prec, dev)
        cgp = self._f32(M, H) if gate == 5 else None
        if gate == 5:
            w_hh, L["w_hh_t"] = self._w2(mod.weights_hh)
        h = h0.clone()
        h_op = ops.split(h, prec)
        c_prev = c0.contiguous()
        for t in range(T):
            rows = slice(t * B, (t + 1) * B)
            if gate == 6:
                _gemm(h_op, wgs, prec=prec, bias=bg, act=ACT_GPMIX, coef=coef4, resid=pre[rows], out_f32=acc[rows],
                      out_pre=zs[rows], tag="gplstm_rec")
                c_in = c_prev
            else:
                _gemm(ops.split(c_prev, prec), wgs, prec=prec, bias=bg, act=ACT_GPMIX, coef=coef4, out_f32=cgp[rows],
                      out_pre=zs[rows], tag="gplstm_cell")
                _gemm(h_op, w_hh, prec=prec, resid=pre[rows], out_f32=acc[rows], tag="gplstm_rec")
                c_in = cgp[rows]
            ops.lstm_cell_step(acc[rows], lengths, t, c_all[rows], h, h_op, out32[rows],
                               Split(outs.hi[rows], None if outs.lo is None else outs.lo[rows]), c_in=c_in)
            c_prev = c_all[rows]
        L.update(acc=acc, zs=zs, c_all=c_all, cgp=cgp, out=outs, coef3=coef)
        return out32, outs, h, c_all[M - B:].clone()

    def _gp56_backward(self, L, dout, h0, c0, T, B, H, kl, kl_scale, le, seed):
        prec, dev, g, mod, pre = self.prec, self.device, self.g, L["mod"], L["pre"]
        gate, M = mod.gate_type, T * B
        acc, zs, c_all, cgp = L["acc"], L["zs"], L["c_all"], L["cgp"]
        ops.lstm_gates_act(acc, c0, T, B, H)          # pre-activations -> gate values in place (its cell states are not used)
        gc = g[pre + "gpnn.coef_mean"]
        NZ = zs.shape[1]
        dG, dGs = self._f32(M, 4 * H), ops.empty_split(M, 4 * H, prec, dev)
        dZ, dZs = self._f32(M, NZ), ops.empty_split(M, NZ, prec, dev)
        dc = self._f32(B, H)
        dh = [self._f32(B, H), self._f32(B, H)]
        rec = None
        for t in range(T - 1, -1, -1):
            sl = slice(t * B, (t + 1) * B)
            c_before = c0 if t == 0 else c_all[(t - 1) * B:t * B]
            dGt = Split(dGs.hi[sl], None if dGs.lo is None else dGs.lo[sl])
            dZt = Split(dZs.hi[sl], None if dZs.lo is None else dZs.lo[sl])
            ops.lstm_bwd_step(acc[sl], cgp[sl] if gate == 5 else c_before, c_all[sl], dout[sl], rec, dc, t == T - 1, dG[sl], dGt)
            if gate == 5:
                # dc holds dL/dc' of this step's transformed cell state: back through the GP unit to dL/dc_{t-1}
                ops.gp3_bwd(dc, zs[sl], L["coef3"], gc, prec, out=dZ[sl], out_split=dZt)
                _gemm(dZt, L["wg_t"], prec=prec, out_f32=dc, tag="gplstm_dc")
                if t > 0:
                    rec = dh[t & 1]
                    _gemm(dGt, L["w_hh_t"], prec=prec, out_f32=rec, tag="gplstm_dh")
            else:
                ops.gp3_bwd(dG[sl], zs[sl], L["coef3"], gc, prec, out=dZ[sl], out_split=dZt)
                if t > 0:
                    rec = dh[t & 1]
                    _gemm(dZt, L["wg_t"], prec=prec, out_f32=rec, tag="gplstm_dh")     # dh_{t-1} = dz_t W_g
        h0s = ops.split(h0, prec)
        hprev = Split(torch.cat([h0s.hi, L["out"].hi[:M - B]], 0),
                      None if h0s.lo is None else torch.cat([h0s.lo, L["out"].lo[:M - B]], 0))
        dGT, dZT = _tsplit(dG, prec, dGs), _tsplit(dZ, prec, dZs)
        self._wgrad(dGT, _tbf16(L["x"], prec), g[pre + "weights_ih"], "gp56_ih")
        ops.colsum(dG, g[pre + "bias_ih"], scale=2.0 if gate == 5 else 1.0)
        if gate == 5:
            self._wgrad(dGT, _tbf16(hprev, prec), g[pre + "weights_hh"], "gp5_hh")
            cprev = ops.split(torch.cat([c0, c_all[:M - B]], 0), prec)
            self._wgrad(dZT, _tbf16(cprev, prec), g[pre + "gpnn.weights_mean"], "gp5_w")
        else:
            self._wgrad(dZT, _tbf16(hprev, prec), g[pre + "gpnn.weights_mean"], "gp6_w")
        ops.colsum(dZ, g[pre + "gpnn.bias_mean"])
        self._gp_unit_param_grads(L, kl, kl_scale, le, seed)
        dx = self._f32(M, L["x"].hi.shape[1])
        _gemm(dGs, L["w_ih_t"], prec=prec, out_f32=dx, tag="dgrad:gplstm_in")
        return dx

    def _gp_cell_forward(self, L, x: Split, h0, c0, lengths, T, B, H, le, seed):
        """GPLSTMCell.forward (model.py:1720-1777) keeping what the backward pass needs."""
        prec, dev, mod = self.prec, self.device, L["mod"]
        gp, nin, mi = mod.gpnn, mod.input_size, L["mi"]
        M = T * B
        coef, wg, bg = gp.coef_mean.detach(), gp.weights_mean.detach(), gp.bias_mean.detach()
        noise = bool(gp.sample) and (le is not None or seed is not None)
        L["noise"] = noise
        if noise:
            ge = (lambda k: None) if le is None else le.get
            tid = _TID_GPCELL + 3 * mi
            if gp.gpnn_type in (1, 3):
                coef = self._reparam32(coef, gp.coef_lgstd.detach(), tid, ge("coef"), seed)
            if gp.gpnn_type in (2, 3):
                wg = self._reparam32(wg, gp.weights_lgstd.detach(), tid + 1, ge("weights"), seed)
                bg = self._reparam32(bg, gp.bias_lgstd.detach(), tid + 2, ge("bias"), seed)
        coef = coef.contiguous()
        w_in32 = torch.cat([mod.weights_ih.detach(), wg[:, :nin]], 0).contiguous()      # [5H, nin]
        w_rec32 = torch.cat([mod.weights_hh.detach(), wg[:, nin:]], 0).contiguous()     # [5H, H]
        bias5 = torch.cat([2.0 * mod.bias_ih.detach(), bg.reshape(-1)], 0).contiguous()
        w_in, L["w_in_t"] = self._w2(w_in32)
        w_rec, L["w_rec_t"] = self._w2(w_rec32)
        pre5 = self._f32(M, 5 * H)
        _gemm(x, w_in, prec=prec, bias=bias5, out_f32=pre5, tag="gplstm_in")
        acc_all, c_all = self._f32(M, 5 * H), self._f32(M, H)
        out32 = self._f32(M, H)
        outs = ops.empty_split(M, H, prec, dev)
        h, c = h0.clone(), c0.clone()
        h_op = ops.split(h, prec)
        for t in range(T):
            rows = slice(t * B, (t + 1) * B)
            _gemm(h_op, w_rec, prec=prec, resid=pre5[rows], out_f32=acc_all[rows], tag="gplstm_rec")
            ops.gp_lstm_cell(acc_all[rows], coef, mod.gate_type, lengths, t, c, h, h_op, out32[rows],
                             Split(outs.hi[rows], None if outs.lo is None else outs.lo[rows]))
            c_all[rows].copy_(c)
        L.update(acc=acc_all, c_all=c_all, out=outs, coef=coef)
        return out32, outs, h, c

    def _gp_cell_backward(self, L, dout, h0, c0, T, B, H, kl, kl_scale, le, seed):
        prec, dev, g, mod = self.prec, self.device, self.g, L["mod"]
        gp, nin, pre, mi = mod.gpnn, mod.input_size, L["pre"], L["mi"]
        M = T * B
        dacc = self._f32(M, 5 * H)
        daccs = ops.empty_split(M, 5 * H, prec, dev)
        dc = self._f32(B, H)
        dh = [self._f32(B, H), self._f32(B, H)]
        gc = g[pre + "gpnn.coef_mean"]
        rec = None
        for t in range(T - 1, -1, -1):
            sl = slice(t * B, (t + 1) * B)
            c_prev = c0 if t == 0 else L["c_all"][(t - 1) * B:t * B]
            dst = Split(daccs.hi[sl], None if daccs.lo is None else daccs.lo[sl])
            ops.gp_lstm_bwd_step(L["acc"][sl], L["coef"], mod.gate_type, c_prev, L["c_all"][sl], dout[sl], rec, dc, t == T - 1,
                                 dacc[sl], dst, gc)
            if t > 0:
                rec = dh[t & 1]
                _gemm(dst, L["w_rec_t"], prec=prec, out_f32=rec, tag="gplstm_dh")
        h0s = ops.split(h0, prec)
        hprev = Split(torch.cat([h0s.hi, L["out"].hi[:M - B]], 0),
                      None if h0s.lo is None else torch.cat([h0s.lo, L["out"].lo[:M - B]], 0))
        # weight gradients: rows [0, 4H) of the stacked products belong to the cell's own matrices, rows [4H, 5H) to the
        # GP unit, whose weight is laid out [H, nin + H] = [x part | h part] (torch.cat([inp, hx]), model.py:1866-1868)
        d4, dz = dacc[:, :4 * H], dacc[:, 4 * H:]
        d4t, dzt = _tsplit(d4, prec), _tsplit(dz, prec)
        xt, ht = _tbf16(L["x"], prec), _tbf16(hprev, prec)
        gw = g[pre + "gpnn.weights_mean"]
        self._wgrad(d4t, xt, g[pre + "weights_ih"], "gp_ih")
        self._wgrad(d4t, ht, g[pre + "weights_hh"], "gp_hh")
        self._wgrad(dzt, xt, gw[:, :nin], "gp_wx")
        self._wgrad(dzt, ht, gw[:, nin:], "gp_wh")
        ops.colsum(d4, g[pre + "bias_ih"], scale=2.0)          # bias_ih is added twice, bias_hh never (model.py:1748-1752)
        gb = g[pre + "gpnn.bias_mean"]
        ops.colsum(dz, gb)
        sid = lambda k: engine._stream_id(_TID_GPCELL + 3 * mi + k, 0)  # noqa: E731
        t_ = gp.gpnn_type
        if t_ in (1, 3):
            if L["noise"]:
                ops.reparam_bwd(gc, gp.coef_lgstd.detach(), gc, g[pre + "gpnn.coef_lgstd"],
                                eps=None if le is None else le.get("coef"), seed=seed, stream_id=sid(0))
            ops.kl_gauss(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl_scale, gc, g[pre + "gpnn.coef_lgstd"])
        if t_ in (2, 3):
            if L["noise"]:
                ops.reparam_bwd(gw, gp.weights_lgstd.detach(), gw, g[pre + "gpnn.weights_lgstd"],
                                eps=None if le is None else le.get("weights"), seed=seed, stream_id=sid(1))
                ops.reparam_bwd(gb, gp.bias_lgstd.detach(), gb, g[pre + "gpnn.bias_lgstd"],
                                eps=None if le is None else le.get("bias"), seed=seed, stream_id=sid(2))
            ops.kl_gauss(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl_scale, gw, g[pre + "gpnn.weights_lgstd"])
            ops.kl_gauss(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl_scale, gb, g[pre + "gpnn.bias_lgstd"])
        dx = self._f32(M, nin)
        _gemm(daccs, L["w_in_t"], prec=prec, out_f32=dx, tag="dgrad:gplstm_in")
        return dx

    def _gp_backward(self, layer, pre, S, dz1, dz1t, x1t, dh, kl, kl_scale, le, seed):
        """Gradients of the GP unit (model.py:1780-1902): weights / bias / coef means, their log-sigmas
        through the reparameterisation when the unit samples, and the unit's KL (the '-1' variant)."""
        g, gp = self.g, layer.gpnn
        t = gp.gpnn_type
        gw, gb, gc = g[pre + "gpnn.weights_mean"], g[pre + "gpnn.bias_mean"], g[pre + "gpnn.coef_mean"]
        self._wgrad(dz1t, x1t, gw, "gpnn")
        ops.colsum(dz1, gb)
        ops.gpmix_dcoef(S["z1"], dh, gc)
        noise = S["gp_noise"]
        sid = lambda name: engine._stream_id(_TID[name], 0)  # noqa: E731
        if t in (1, 3):
            if noise:
                ops.reparam_bwd(gc, gp.coef_lgstd.detach(), gc, g[pre + "gpnn.coef_lgstd"],
                                eps=None if le is None else le.get("coef"), seed=seed, stream_id=sid("gp_coef"))
            ops.kl_gauss(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.coef_mean.detach(), gp.coef_lgstd.detach(), kl_scale, gc, g[pre + "gpnn.coef_lgstd"])
        if t in (2, 3):
            if noise:
                ops.reparam_bwd(gw, gp.weights_lgstd.detach(), gw, g[pre + "gpnn.weights_lgstd"],
                                eps=None if le is None else le.get("weights"), seed=seed, stream_id=sid("gp_w"))
                ops.reparam_bwd(gb, gp.bias_lgstd.detach(), gb, g[pre + "gpnn.bias_lgstd"],
                                eps=None if le is None else le.get("bias"), seed=seed, stream_id=sid("gp_b"))
            ops.kl_gauss(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.weights_mean.detach(), gp.weights_lgstd.detach(), kl_scale, gw, g[pre + "gpnn.weights_lgstd"])
            ops.kl_gauss(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl, minus_one=True, accumulate=True)
            ops.kl_gauss_bwd(gp.bias_mean.detach(), gp.bias_lgstd.detach(), kl_scale, gb, g[pre + "gpnn.bias_lgstd"])

    def apply_gradients(self):
        """All-reduce (data parallel), global-norm clip, SGD momentum; invalidates the cached weight copies."""
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_g, group=self.group)
        ops.reduce_sum(self.flat_g, self.norm_sq, squares=True)
        ops.sgd_momentum(self.flat_p, self.flat_g, self.flat_v, self.lr, self.momentum, self.norm_sq, self.clip,
                         1.0 / self.world, out_hi=self.flat_hi if _MIRROR else None, out_lo=self.flat_lo if _MIRROR else None)
        self.model.__dict__.pop("_blm_plans", None)

    def step(self, tokens_tb, targets_tb, kl_scale, *, eps=None, seed=None, hidden=None):
        out = self.forward_backward(tokens_tb, targets_tb, kl_scale, eps=eps, seed=seed, hidden=hidden)
        self.apply_gradients()
        return out

    # ------------------------------------------------------------------ CUDA-graph replay
    def capture(self, T: int, B: int, kl_scale: float):
        """Capture forward + backward (graph 1) and norm + clip + SGD (graph 2) for batches of shape
        (T, B).  A step is ~250 small launches; replaying them as two graphs removes the host-side launch
        cost that otherwise dominates a 3200-token step.  The noise of the step lives in static device
        buffers that :meth:`step_captured` refills (Philox, outside the graph) before every replay, so the
        graph itself only ever sees injected noise.  The all-reduce stays outside, between the graphs."""
        m, dev = self.model, self.device
        d = m.ninp
        self._check_mirrors()
        self._cap = {"T": T, "B": B, "kl_scale": float(kl_scale),
                     "x": torch.zeros(T, B, dtype=torch.int64, device=dev),
                     "y": torch.zeros(T, B, dtype=torch.int64, device=dev), "eps": {}, "fill": []}
        cap = self._cap
        for li, layer in enumerate(m.transformerlayers):
            if layer.kind == "v":
                buf = torch.zeros(B, T, d, dtype=torch.float32, device=dev)   # the kernel's own (sequence-major) layout
                cap["eps"][f"layer{li}"] = buf
                cap["fill"].append((buf, _TID_VNOISE + li, V_NOISE_STD))
            elif layer.kind == "bayes_ffn":
                buf = torch.zeros_like(layer.linear2.weight_lgstd)
                cap["eps"][f"layer{li}"] = buf
                cap["fill"].append((buf, _TID["ffn_w2"], 1.0))
            elif layer.kind == "bayes_mha":
                buf = torch.zeros_like(layer.self_attn.o_net.weight_lgstd)
                cap["eps"][f"layer{li}"] = buf
                cap["fill"].append((buf, _TID["mha_o"], 1.0))
            elif layer.kind == "gauss" and layer.gpnn.sample:
                e, gp = {}, layer.gpnn
                if gp.gpnn_type in (1, 3):
                    e["coef"] = torch.zeros_like(gp.coef_lgstd)
                    cap["fill"].append((e["coef"], _TID["gp_coef"], 1.0))
                if gp.gpnn_type in (2, 3):
                    e["weights"], e["bias"] = torch.zeros_like(gp.weights_lgstd), torch.zeros_like(gp.bias_lgstd)
                    cap["fill"] += [(e["weights"], _TID["gp_w"], 1.0), (e["bias"], _TID["gp_b"], 1.0)]
                cap["eps"][f"layer{li}"] = e
        if getattr(m, "bayes_embed", False):
            buf = torch.zeros_like(m.embed_lgstd)
            cap["eps"]["embed"] = buf
            cap["fill"].append((buf, _TID["embed"], 1.0))
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)   # dropout key of the replay, written per step
        ops.end_capture()
        self._capturing = True
        try:
            self._capture_graphs(cap, T, B)
        finally:
            self._capturing = False
            ops.end_capture()
        return self

    def _capture_graphs(self, cap, T: int, B: int):
        m = self.model
        self._refill_noise(0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, lazy attributes)
            for _ in range(2):
                self.forward_backward(cap["x"], cap["y"], cap["kl_scale"], eps=cap["eps"], v_eps_layout="btd")
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        nl = len(m.transformerlayers)
        cap["split"] = None
        if self.world > 1 and nl >= 2 and _OVERLAP:
            # Data parallel: the backward pass is captured as TWO graphs, cut after layer `sl`.  The gradients of the
            # layers >= sl are final there, so their all-reduce (one contiguous range of the flat buffer) runs on
            # NCCL's stream while the second graph finishes the backward pass of the lower layers and the embedding.
            sl = nl // 2
            rng = [self._offs[n] for n in self._offs if n.startswith("transformerlayers.")
                   and int(n.split(".")[1]) >= sl]
            lo, hi = min(o for o, _ in rng), max(o + n for o, n in rng)
            inside = [n for n, (o, _) in self._offs.items() if lo <= o < hi]
            if all(n.startswith("transformerlayers.") and int(n.split(".")[1]) >= sl for n in inside):
                cap["split"] = (sl, lo, hi)
        cap["ar_in_graph"] = False
        if cap["split"] is None:
            cap["g1"] = torch.cuda.CUDAGraph()
            in_graph = self.world > 1 and _NCCL_IN_GRAPH and self.model.family not in ("bayes_lstm", "std_lstm", "gauss_lstm", "v_lstm")
            if in_graph:
                # NCCL's watchdog thread polls its events while this thread captures: thread-local capture mode keeps
                # that legal; one eager all-reduce first so that the communicator exists before the capture starts
                torch.distributed.all_reduce(torch.zeros(1, device=self.device), group=self.group)
                torch.cuda.synchronize()
                self._background_group()
                self._ar_in_graph = True
                try:
                    with torch.cuda.graph(cap["g1"], capture_error_mode="thread_local"):
                        cap["out"] = self.forward_backward(cap["x"], cap["y"], cap["kl_scale"], eps=cap["eps"],
                                                           v_eps_layout="btd")
                finally:
                    self._ar_in_graph = False
                cap["ar_in_graph"] = True
            else:
                with torch.cuda.graph(cap["g1"]):
                    cap["out"] = self.forward_backward(cap["x"], cap["y"], cap["kl_scale"], eps=cap["eps"], v_eps_layout="btd")
        else:
            sl = cap["split"][0]
            cap["g1"], cap["g1b"] = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            state = {"cut": False}

            def cut(li):
                if li == sl and not state["cut"]:
                    self._join_aside()
                    cap["g1"].capture_end()
                    cap["g1b"].capture_begin(pool=cap["g1"].pool())
                    state["cut"] = True

            cs = torch.cuda.Stream()
            cs.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cs):
                self.forward_backward(cap["x"], cap["y"], cap["kl_scale"], eps=cap["eps"], v_eps_layout="btd")  # workspaces of this stream
                torch.cuda.synchronize()
                cap["g1"].capture_begin()
                cap["out"] = self.forward_backward(cap["x"], cap["y"], cap["kl_scale"], eps=cap["eps"], v_eps_layout="btd",
                                                   on_layer_done=cut)
                (cap["g1b"] if state["cut"] else cap["g1"]).capture_end()
            torch.cuda.current_stream().wait_stream(cs)
            torch.cuda.synchronize()
            if not state["cut"]:
                cap["split"] = None
        ops.end_capture()      # the optimiser graph has its own pool: no scratch block of the first graph is reused in it
        cap["g2"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cap["g2"]):
            ops.reduce_sum(self.flat_g, self.norm_sq, squares=True)
            ops.sgd_momentum(self.flat_p, self.flat_g, self.flat_v, self.lr, self.momentum, self.norm_sq, self.clip,
                             1.0 / self.world, out_hi=self.flat_hi if _MIRROR else None,
                             out_lo=self.flat_lo if _MIRROR else None)
        return self

    def _refill_noise(self, seed: int):
        for buf, tid, std in self._cap["fill"]:
            # weight noise is shared by the data-parallel ranks (replicated weights); the per-token noise of the
            # variational layers belongs to the rank's own rows
            k = self.rank if tid >= _TID_VNOISE else 0
            ops.philox_normal(seed, engine._stream_id(tid, k), buf.numel(), self.device, out=buf.view(-1), scale=std)
        self._seed_dev.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)

    def step_captured(self, tokens_tb, targets_tb, seed: int):
        """One step through the captured graphs; (loss, ce, kl) are 0-dim device tensors valid until the
        next replay."""
        cap = self._cap
        self._check_mirrors()          # outside the graphs: a parameter written from outside re-makes the bf16 mirrors
        cap["x"].copy_(tokens_tb, non_blocking=True)
        cap["y"].copy_(targets_tb.view(cap["T"], cap["B"]), non_blocking=True)
        self._refill_noise(seed)
        cap["g1"].replay()
        if self.world > 1 and cap.get("split"):
            _, lo, hi = cap["split"]
            dist = torch.distributed
            w1 = dist.all_reduce(self.flat_g[lo:hi], group=self.group, async_op=True)   # overlaps the second graph
            cap["g1b"].replay()
            works = [w1, dist.all_reduce(self.flat_g[:lo], group=self.group, async_op=True)]
            if hi < self.flat_g.numel():
                works.append(dist.all_reduce(self.flat_g[hi:], group=self.group, async_op=True))
            for w in works:
                w.wait()
        elif self.world > 1 and not cap.get("ar_in_graph"):
            torch.distributed.all_reduce(self.flat_g, group=self.group)
        cap["g2"].replay()
        self.model.__dict__.pop("_blm_plans", None)
        return cap["out"]


_TORCH_TM_KEYS = {"self_attn.in_proj_weight": "self_attn.qkv_net.weight", "self_attn.in_proj_bias": "self_attn.qkv_net.bias",
                  "self_attn.out_proj.weight": "self_attn.o_net.weight", "self_attn.out_proj.bias": "self_attn.o_net.bias"}


def _engine_name(name: str) -> str:
    """Parameter name of a TransformerModel / RNNModel (torch's nn.TransformerEncoder / nn.LSTM keys) -> the name the
    same tensor has in the Bayesian containers; other names pass through."""
    m = re.fullmatch(r"transformerlayers\.layers\.(\d+)\.(.*)", name)
    if m:
        return f"transformerlayers.{m[1]}.{_TORCH_TM_KEYS.get(m[2], m[2])}"
    m = re.fullmatch(r"rnn\.(weight|bias)_(ih|hh)_l(\d+)", name)
    if m:
        return f"rnn.{m[1]}_{m[2]}_mean_{int(m[3]) + 1}"
    return name


def _param(model, name: str) -> torch.Tensor:
    obj = model
    for part in name.split("."):
        obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
    return obj.detach()


def _to_device(obj, dev):
    """Injected noise arrives as (nested dicts of) CPU tensors in the oracle's layout."""
    if isinstance(obj, dict):
        return {k: _to_device(v, dev) for k, v in obj.items()}
    if torch.is_tensor(obj):
        return obj.to(dev, non_blocking=True).float().contiguous()
    return obj
