"""Drop-in replacement for the reference's ``steps/pytorchnn/model.py`` on the rescoring
hot path: same class names, same positional constructor arguments, same ``state_dict``
keys and shapes, same ``forward`` signatures -- but the modules are parameter containers
only.  All arithmetic is done by :mod:`bayeslms_b200.engine` through the sm_100a kernels of
``libbayeslm_b200.so``; there is no PyTorch/CPU execution path.

Reference call sites this mirrors (``model.py`` = reference steps/pytorchnn/model.py):
  RNNModel 23-72; TransformerModel 120-171 (their nn.LSTM / nn.TransformerEncoder state_dict keys);
  BayesRNNModel 179-229 / Bayes2LSTM 585-828; MultiheadAttention 836-928;
  BayesMultiheadAttention 931-1019; StandardTransformerEncoderLayer 1022-1046;
  BayesLinear 1049-1134; BayesTransformerEncoderLayer 1137-1176; BayesTransformerModel
  1179-1309; GPNN 1780-1906; GaussTransformerEncoderLayer 2250-2287; GaussTransformerModel
  2290-2364; VTransformerEncoderLayer 2741-2805; VTransformerModel 2808-2897.

Beyond the reference API every model has ``score(batch, ...)`` -- per-hypothesis NLL without
materialising logits -- which is what :mod:`bayeslms_b200.scorer` uses.
"""
from __future__ import annotations

import math
import re
from typing import Optional

import torch
import torch.nn as nn

from . import engine as _engine

_LN_EPS = 1e-5


def _uniform(shape, lo, hi):
    return nn.Parameter(torch.empty(*shape).uniform_(lo, hi))


def _lgstd_like(shape, stdv):
    """log-sigma initialiser shared by every Bayesian tensor: U(2 ln s, ln s) (model.py:657,1073,1845)."""
    return _uniform(shape, 2.0 * math.log(stdv), math.log(stdv))


class PositionalEncoding(nn.Module):
    """Sinusoidal table registered as buffer ``pe`` [max_len, 1, d] (model.py:93-103)."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.p = float(dropout)
        pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(1))


class BayesLinear(nn.Module):
    """Bias-free Gaussian linear layer: weight_mean / weight_lgstd [out, in] (model.py:1049-1134).
    ``bias=True`` is unusable in the reference (its forward raises) and rejected here."""
    kind = "bayes_linear"

    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        if bias:
            raise NotImplementedError("BayesLinear(bias=True) does not run in the reference either")
        self.in_features, self.out_features = in_features, out_features
        self.sample = True
        s = 1.0 / math.sqrt(out_features + 1)
        self.weight_mean = _uniform((out_features, in_features), -s, s)
        self.weight_lgstd = _lgstd_like((out_features, in_features), s)

    def kl_divergence(self, prior=None):
        if prior is not None:
            raise NotImplementedError("prior-centred KL is unused by train.py (train.py:342)")
        return _engine.kl_sum([(self.weight_mean, self.weight_lgstd, 1.0)], minus_one=False)


class MultiheadAttention(nn.Module):
    """qkv_net Linear(d, 3d) + o_net Linear(d, d) (model.py:836-869)."""
    kind = "mha"

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, float(dropout)
        self.qkv_net = nn.Linear(embed_dim, 3 * embed_dim)
        self.o_net = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.qkv_net.weight)
        nn.init.zeros_(self.qkv_net.bias)
        nn.init.zeros_(self.o_net.bias)


class BayesMultiheadAttention(nn.Module):
    """Separate q/k/v projections and a Bayesian, bias-free output projection (model.py:931-961)."""
    kind = "bayes_mha"

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, float(dropout)
        self.q_net = nn.Linear(embed_dim, embed_dim)
        self.k_net = nn.Linear(embed_dim, embed_dim)
        self.v_net = nn.Linear(embed_dim, embed_dim)
        self.o_net = BayesLinear(embed_dim, embed_dim)


class _EncoderLayer(nn.Module):
    kind = "std"

    def __init__(self, d_model, nhead, dim_feedforward, dropout, attn_cls=MultiheadAttention):
        super().__init__()
        self.self_attn = attn_cls(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, eps=_LN_EPS)
        self.norm2 = nn.LayerNorm(d_model, eps=_LN_EPS)
        self.p_drop = float(dropout)


class StandardTransformerEncoderLayer(_EncoderLayer):
    """Post-LN block with exact-erf GELU (model.py:1022-1046)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__(d_model, nhead, dim_feedforward, dropout)


class BayesTransformerEncoderLayer(_EncoderLayer):
    """Layer 0 of the Bayesian Transformer: Bayesian linear2 ('FFN') or Bayesian o_net ('MHA')
    (model.py:1137-1176)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, bayes_pos=None):
        super().__init__(d_model, nhead, dim_feedforward, dropout,
                         attn_cls=BayesMultiheadAttention if bayes_pos == "MHA" else MultiheadAttention)
        self.bayes_pos = bayes_pos
        if bayes_pos == "FFN":
            self.linear2 = BayesLinear(dim_feedforward, d_model)
        self.kind = {"FFN": "bayes_ffn", "MHA": "bayes_mha"}.get(bayes_pos, "std")


class GPNN(nn.Module):
    """Gaussian-process activation unit: z = x W^T + b, out = sum_i coef[i] * act_i(z) with
    acts (tanh, sigmoid, relu, gelu); type 1/3 Bayesian coef, 2/3 Bayesian W,b (model.py:1780-1902)."""

    ACT_SETS = (("tanh", "sigmoid", "relu", "gelu"),      # Transformer GP layer, model.py:2263
                ("sigmoid", "tanh", "relu"),              # GP-LSTM gates i, g, o (GPNN default, model.py:1787,1693-1695)
                ("sigmoid",))                             # GP-LSTM forget gate, model.py:1697

    def __init__(self, input_size, output_size, act_set=("sigmoid", "tanh", "relu"), gpnn_type=0):
        super().__init__()
        if tuple(act_set) not in self.ACT_SETS:
            raise NotImplementedError(f"activation set {tuple(act_set)} is not one the reference builds")
        self.act_set = tuple(act_set)
        self.input_size, self.output_size, self.gpnn_type = input_size, output_size, gpnn_type
        self.sample = False  # as shipped (model.py:1799); train.py never flips it
        s = 1.0 / math.sqrt(output_size)
        n_act = len(self.act_set)
        self.weights_mean = _uniform((output_size, input_size), -s, s)
        self.bias_mean = nn.Parameter(torch.zeros(output_size))
        self.coef_mean = _uniform((n_act, output_size), 0.0, 1.0)
        if gpnn_type in (1, 3):
            self.coef_lgstd = _lgstd_like((n_act, output_size), s)
        if gpnn_type in (2, 3):
            self.weights_lgstd = _lgstd_like((output_size, input_size), s)
            self.bias_lgstd = _lgstd_like((output_size,), s)

    def kl_divergence(self, prior=None):
        if prior is not None:
            raise NotImplementedError("prior-centred KL is not on the train.py path")
        terms = []
        if self.gpnn_type in (1, 3):
            terms.append((self.coef_mean, self.coef_lgstd, 1.0))
        if self.gpnn_type in (2, 3):
            terms.append((self.weights_mean, self.weights_lgstd, 1.0))
            terms.append((self.bias_mean, self.bias_lgstd, 1.0))
        return _engine.kl_sum(terms, minus_one=True) if terms else 0


class GaussTransformerEncoderLayer(_EncoderLayer):
    """Layer 0 of the GP Transformer: act(linear1(x)) is replaced by GPNN(x); ``linear1`` stays in the
    state_dict but is dead in forward (model.py:2250-2287)."""
    kind = "gauss"

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, gauss_pos=None):
        super().__init__(d_model, nhead, dim_feedforward, dropout)
        self.gauss_pos = self.gpnn_type = gauss_pos
        if not 0 <= gauss_pos <= 3:
            raise NotImplementedError("GPNN2 (gauss_pos=4) is broken in the reference (SURVEY.md a16)")
        self.gpnn = GPNN(d_model, dim_feedforward, act_set=("tanh", "sigmoid", "relu", "gelu"), gpnn_type=gauss_pos)


class VTransformerEncoderLayer(_EncoderLayer):
    """Variational layer: standard block + four (100, 1, d) tensors initialised U(0,1)
    (model.py:2741-2761); noise only in training at sequence length 100."""
    kind = "v"

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1):
        super().__init__(d_model, nhead, dim_feedforward, dropout)
        for name in ("hiddens_mean_p", "hiddens_lgstd_p", "hiddens_mean", "hiddens_lgstd"):
            setattr(self, name, nn.Parameter(torch.rand(100, 1, d_model)))
        self._v_state = None     # FFN output + noise source of the last training-mode forward (the reference's self.hidden)

    def kl_divergence(self):
        """mean((h - h * mean_p)^2 - 2 lgstd + exp(2 lgstd)) / 2 on the noised FFN output h of the last training-mode
        forward at sequence length 100 (model.py:2770-2781; called by train.py:380-395)."""
        if self._v_state is None:
            raise RuntimeError("VTransformerEncoderLayer.kl_divergence() needs a training-mode forward at sequence length "
                               "100 first (the reference reads the hidden it stored there, model.py:2772)")
        return _engine.v_layer_kl(self)


class _TransformerLM(nn.Module):
    """Shared container: encoder / pos_encoder / transformerlayers / decoder (tied)."""
    family = "tm"

    def _finish(self, ntoken, ninp, dropout, tie_weights):
        self.model_type = "Transformer"
        self.ninp = ninp
        self.encoder = nn.Embedding(ntoken, ninp)
        self.decoder = nn.Linear(ninp, ntoken)
        if tie_weights:
            self.decoder.weight = self.encoder.weight
        nn.init.uniform_(self.encoder.weight, -0.1, 0.1)
        nn.init.zeros_(self.decoder.bias)
        nn.init.uniform_(self.decoder.weight, -0.1, 0.1)
        self.p_drop = float(dropout)

    @property
    def nhead(self):
        return self.transformerlayers[0].self_attn.num_heads if len(self.transformerlayers) else 1

    def forward(self, src, has_mask=True):
        """(T, B) int64 -> logits (T, B, V); posterior mean in eval mode (reference semantics)."""
        if not has_mask:
            raise NotImplementedError("the rescoring / training callers always use the causal mask")
        return _engine.transformer_logits(self, src)

    def score(self, batch, **kw):
        """Per-hypothesis NLL of a :class:`bayeslms_b200.engine.PackedBatch` (no logits in HBM)."""
        return _engine.transformer_score(self, batch, **kw)


class BayesTransformerModel(_TransformerLM):
    """BayesTransformerModel(ntoken, ninp, nhead, nhid, nlayers, dropout, tie_weights, bayes_pos)
    with bayes_pos in {'none','FFN','MHA','EMB'} (model.py:1182-1241).  Layer 0's dropout is the
    hard-coded 0.2 of model.py:1202,1207."""
    family = "bayes_tm"

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.5, tie_weights=False, bayes_pos=None):
        super().__init__()
        self.bayes_pos = bayes_pos
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        layers = []
        if bayes_pos in ("FFN", "MHA"):
            layers.append(BayesTransformerEncoderLayer(ninp, nhead, nhid, dropout=0.2, bayes_pos=bayes_pos))
            layers += [StandardTransformerEncoderLayer(ninp, nhead, nhid, dropout) for _ in range(nlayers - 1)]
        elif bayes_pos in ("none", "EMB"):
            layers += [StandardTransformerEncoderLayer(ninp, nhead, nhid, dropout) for _ in range(nlayers)]
        self.transformerlayers = nn.ModuleList(layers)
        self.bayes_embed = bayes_pos == "EMB"
        self._finish(ntoken, ninp, dropout, tie_weights)
        if self.bayes_embed:
            s = 1.0 / math.sqrt(ninp + 1)
            self.embed_mean = _uniform((ninp, ninp), -s, s)
            self.embed_lgstd = _lgstd_like((ninp, ninp), s)

    def embed_kl_divergence(self):
        return _engine.kl_sum([(self.embed_mean, self.embed_lgstd, 1.0)], minus_one=False)


class GaussTransformerModel(_TransformerLM):
    """GaussTransformerModel(..., tie_weights, gauss_pos:int) (model.py:2293-2327)."""
    family = "gauss_tm"

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.5, tie_weights=False, gauss_pos=4):
        super().__init__()
        self.gauss_pos = gauss_pos
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        if gauss_pos > 4:
            layers = [StandardTransformerEncoderLayer(ninp, nhead, nhid, dropout) for _ in range(nlayers)]
        else:
            layers = [GaussTransformerEncoderLayer(ninp, nhead, nhid, dropout, gauss_pos)]
            layers += [StandardTransformerEncoderLayer(ninp, nhead, nhid, dropout) for _ in range(nlayers - 1)]
        self.transformerlayers = nn.ModuleList(layers)
        self._finish(ntoken, ninp, dropout, tie_weights)


def normalise_v_pos(v_pos):
    """``--T_v_pos`` is an int in the scripts but the README passes the per-layer bit string '11'
    (SURVEY.md 8c-2): accept 0..3 and '00','01','10','11' -> 0,1,2,3."""
    table = {"00": 0, "01": 1, "10": 2, "11": 3, 0: 0, 1: 1, 2: 2, 3: 3, 10: 2, 11: 3}
    key = v_pos if not isinstance(v_pos, str) else (v_pos if v_pos in table else int(v_pos))
    if key not in table:
        raise ValueError(f"T_v_pos must be one of 0,1,2,3 or a two-bit string, got {v_pos!r}")
    return table[key]


class VTransformerModel(_TransformerLM):
    """VTransformerModel(..., tie_weights, v_pos) (model.py:2811-2857).  v_pos 2 and 3 build
    ``nlayers - 1`` layers in total, like the reference (model.py:2834,2840), so checkpoints load."""
    family = "v_tm"

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.5, tie_weights=False, v_pos=0):
        super().__init__()
        self.v_pos = v_pos = normalise_v_pos(v_pos)
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        std = lambda: StandardTransformerEncoderLayer(ninp, nhead, nhid, dropout)  # noqa: E731
        var = lambda: VTransformerEncoderLayer(ninp, nhead, nhid, dropout)  # noqa: E731
        head = {0: [], 1: [var()], 2: [std(), var()], 3: [var(), var()]}[v_pos]
        n_tail = {0: nlayers, 1: nlayers - 1, 2: nlayers - 3, 3: nlayers - 3}[v_pos]
        self.transformerlayers = nn.ModuleList(head + [std() for _ in range(n_tail)])
        self._finish(ntoken, ninp, dropout, tie_weights)


class _ParamView:
    """(weight, bias) pair presented under the attribute names the engine reads."""
    __slots__ = ("weight", "bias")

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class _TorchMultiheadAttention(nn.Module):
    """Parameter container with ``nn.MultiheadAttention``'s keys (in_proj_weight, in_proj_bias, out_proj.*):
    what ``nn.TransformerEncoderLayer`` puts in a reference ``TransformerModel`` checkpoint (model.py:131-133)."""
    kind = "mha"

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, float(dropout)
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    @property
    def qkv_net(self):
        return _ParamView(self.in_proj_weight, self.in_proj_bias)

    @property
    def o_net(self):
        return self.out_proj


class _TorchEncoderLayer(_EncoderLayer):
    """Post-LN layer with ``nn.TransformerEncoderLayer``'s keys (self_attn.in_proj_*, self_attn.out_proj.*,
    linear1, linear2, norm1, norm2); activation = exact-erf GELU, the one the callers pass (score.py:377,
    train.py:198-201), or ReLU, the constructor's default (model.py:124)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="gelu"):
        super().__init__(d_model, nhead, dim_feedforward, dropout, attn_cls=_TorchMultiheadAttention)
        self.activation = activation


class _TorchEncoder(nn.Module):
    """``nn.TransformerEncoder``'s layout: the layers live under ``.layers`` (keys transformerlayers.layers.<i>.*)."""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def __iter__(self):
        return iter(self.layers)

    def __len__(self):
        return len(self.layers)

    def __getitem__(self, i):
        return self.layers[i]


class TransformerModel(_TransformerLM):
    """TransformerModel(ntoken, ninp, nhead, nhid, nlayers, dropout, activation, tie_weights) (model.py:120-171):
    the baseline model of ``--uncertainty none`` (score.py:377, train.py:198-201), a stack of
    ``nn.TransformerEncoderLayer``s in the reference -- so its checkpoints carry torch's key names, which this
    container reproduces.  Runs on the same kernels as the standard layers of the Bayesian models."""
    family = "std_tm"

    def __init__(self, ntoken, ninp, nhead, nhid, nlayers, dropout=0.5, activation="relu", tie_weights=False):
        super().__init__()
        if activation not in ("gelu", "relu"):
            raise ValueError(f"activation should be relu/gelu, not {activation}")     # torch's own message (model.py:131)
        self.activation = activation
        self.src_mask = None
        self.pos_encoder = PositionalEncoding(ninp, dropout)
        self.transformerlayers = _TorchEncoder([_TorchEncoderLayer(ninp, nhead, nhid, dropout, activation)
                                                for _ in range(nlayers)])
        self._finish(ntoken, ninp, dropout, tie_weights)


class Bayes2LSTM(nn.Module):
    """Two-layer LSTM whose gate ``position`` (1=i, 2=f, 3=g, 4=o) has Gaussian weights in both layers
    (model.py:585-666).  ``weight_ih_mean_2`` is (4H, input_size) like the reference, i.e. only
    meaningful when input_size == hidden_size."""

    def __init__(self, input_size, hidden_size, num_layers=1, position=0, bias=True, dropout=0.0, bayes_pos=0):
        super().__init__()
        if num_layers != 2:
            raise NotImplementedError("Bayes2LSTM is hard-wired to two layers in the reference")
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        self.position, self.bias, self.dropout = position, bias, float(dropout)
        H, s = hidden_size, 1.0 / math.sqrt(hidden_size)
        for layer in (1, 2):
            setattr(self, f"weight_ih_mean_{layer}", _uniform((4 * H, input_size), -s, s))
            setattr(self, f"weight_hh_mean_{layer}", _uniform((4 * H, H), -s, s))
            setattr(self, f"bias_ih_mean_{layer}", _uniform((4 * H,), -s, s))
            setattr(self, f"bias_hh_mean_{layer}", _uniform((4 * H,), -s, s))
        if 1 <= position <= 4:
            for layer in (1, 2):
                setattr(self, f"weight_hh_lgstd_{layer}", _lgstd_like((H, H), s))
                setattr(self, f"weight_ih_lgstd_{layer}", _lgstd_like((H, input_size), s))
                setattr(self, f"bias_hh_lgstd_{layer}", _lgstd_like((H,), s))
                setattr(self, f"bias_ih_lgstd_{layer}", _lgstd_like((H,), s))
        elif position != 0:
            raise NotImplementedError("position 5 is not a supported flag value (score.py:335-336)")

    def gate_rows(self):
        H = self.hidden_size
        return slice((self.position - 1) * H, self.position * H)

    def kl_divergence(self, prior=None):
        """Layer-1 tensors only, mean over cat(hh, ih) -- the reference's exact (if surprising) KL
        (model.py:736-765)."""
        if prior is not None:
            raise NotImplementedError("prior-centred KL is unused by train.py (train.py:338)")
        if not 1 <= self.position <= 4:
            return 0
        rows = self.gate_rows()
        n_hh, n_ih = self.hidden_size, self.input_size
        tot = float(n_hh + n_ih)
        w = _engine.kl_sum([(self.weight_hh_mean_1[rows], self.weight_hh_lgstd_1, n_hh / tot),
                            (self.weight_ih_mean_1[rows], self.weight_ih_lgstd_1, n_ih / tot),
                            (self.bias_hh_mean_1[rows], self.bias_hh_lgstd_1, 0.5),
                            (self.bias_ih_mean_1[rows], self.bias_ih_lgstd_1, 0.5)], minus_one=False)
        return w


class BayesRNNModel(nn.Module):
    """BayesRNNModel(rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights, bayes_pos:int)
    (model.py:181-229)."""
    family = "bayes_lstm"

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False, bayes_pos=0):
        super().__init__()
        if rnn_type != "LSTM":
            raise NotImplementedError("only --model LSTM is on the Bayesian rescoring path")
        self.rnn_type, self.nhid, self.nlayers, self.bayes_pos = rnn_type, nhid, nlayers, bayes_pos
        self.p_drop = float(dropout)
        self.encoder = nn.Embedding(ntoken, ninp)
        self.rnn = Bayes2LSTM(ninp, nhid, nlayers, position=bayes_pos, dropout=dropout)
        self.decoder = nn.Linear(nhid, ntoken)
        if tie_weights:
            if nhid != ninp:
                raise ValueError("When using the tied flag, nhid must be equal to emsize.")
            self.decoder.weight = self.encoder.weight
        nn.init.uniform_(self.encoder.weight, -0.1, 0.1)
        nn.init.zeros_(self.decoder.bias)
        nn.init.uniform_(self.decoder.weight, -0.1, 0.1)

    def init_hidden(self, bsz):
        w = self.encoder.weight
        return (w.new_zeros(self.nlayers, bsz, self.nhid), w.new_zeros(self.nlayers, bsz, self.nhid))

    def forward(self, x, hidden):
        """(T, B) int64, (h, c) each (2, B, H) -> logits (T, B, V), (h, c)."""
        return _engine.lstm_logits(self, x, hidden)

    def score(self, batch, hidden, **kw):
        return _engine.lstm_score(self, batch, hidden, **kw)


class _TorchLSTMParams(nn.Module):
    """Parameter container with ``nn.LSTM``'s keys (weight_ih_l<k>, weight_hh_l<k>, bias_ih_l<k>, bias_hh_l<k>).
    The Bayes2LSTM-style names ``<w>_mean_<layer>`` (layer from 1) resolve to the same tensors, so the engine and
    the trainer treat it as a Bayes2LSTM with no Bayesian gate."""
    position = 0
    _ALIAS = re.compile(r"(weight|bias)_(ih|hh)_mean_(\d+)")

    def __init__(self, input_size, hidden_size, num_layers=1, dropout=0.0):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.dropout = input_size, hidden_size, num_layers, float(dropout)
        H, s = hidden_size, 1.0 / math.sqrt(hidden_size)
        for l in range(num_layers):
            setattr(self, f"weight_ih_l{l}", _uniform((4 * H, input_size if l == 0 else H), -s, s))
            setattr(self, f"weight_hh_l{l}", _uniform((4 * H, H), -s, s))
            setattr(self, f"bias_ih_l{l}", _uniform((4 * H,), -s, s))
            setattr(self, f"bias_hh_l{l}", _uniform((4 * H,), -s, s))

    def __getattr__(self, name):
        m = self._ALIAS.fullmatch(name)
        if m:
            return super().__getattr__(f"{m[1]}_{m[2]}_l{int(m[3]) - 1}")
        return super().__getattr__(name)

    def gate_rows(self):
        return slice(0, 0)

    def kl_divergence(self, prior=None):
        return 0


class RNNModel(nn.Module):
    """RNNModel(rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights) (model.py:23-72): the baseline LSTM LM of
    ``--uncertainty none`` (score.py:413, train.py:204-207) with ``nn.LSTM``'s state_dict keys."""
    family = "std_lstm"

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False):
        super().__init__()
        if rnn_type != "LSTM":
            raise NotImplementedError("only --model LSTM is on the rescoring path (GRU / RNN_TANH / RNN_RELU are not)")
        if nlayers != 2:
            raise NotImplementedError("the recurrence path is laid out for the recipes' two-layer LSTM")
        self.rnn_type, self.nhid, self.nlayers = rnn_type, nhid, nlayers
        self.p_drop = float(dropout)
        self.encoder = nn.Embedding(ntoken, ninp)
        self.rnn = _TorchLSTMParams(ninp, nhid, nlayers, dropout=dropout)
        self.decoder = nn.Linear(nhid, ntoken)
        if tie_weights:
            if nhid != ninp:
                raise ValueError("When using the tied flag, nhid must be equal to emsize.")
            self.decoder.weight = self.encoder.weight
        nn.init.uniform_(self.encoder.weight, -0.1, 0.1)
        nn.init.zeros_(self.decoder.bias)
        nn.init.uniform_(self.decoder.weight, -0.1, 0.1)

    def init_hidden(self, bsz):
        w = self.encoder.weight
        return (w.new_zeros(self.nlayers, bsz, self.nhid), w.new_zeros(self.nlayers, bsz, self.nhid))

    def forward(self, x, hidden):
        """(T, B) int64, (h, c) each (nlayers, B, H) -> logits (T, B, V), (h, c)."""
        return _engine.lstm_logits(self, x, hidden)


class GPLSTMCell(nn.Module):
    """LSTM cell with a GP unit (model.py:1674-1777).  gate_type 1..4 (i, f, g, o): that gate is a GP unit over
    cat[x, h]; 5 ("cell"): the cell state passes through GPNN(H -> H) before the update; 6 ("hidden"): GPNN(h)
    (H -> 4H) replaces the recurrent product; 7 ("inputs"): GPNN(x) (in -> 4H) replaces the input product.  Keeps the
    reference's bias quirk: ``bias_ih`` is used wherever the reference writes it (twice in the plain product form) and
    ``bias_hh`` never (model.py:1745-1752).  GPNN2 (gpnn_type 4) is rejected: its own KL crashes in the reference."""
    kind = "gp"

    def __init__(self, input_size, hidden_size, gate_type=0, gpnn_type=0):
        super().__init__()
        if not (1 <= gate_type <= 7 and 0 <= gpnn_type <= 3):
            raise NotImplementedError(f"GP-LSTM gate_type {gate_type} / gpnn_type {gpnn_type} is outside the B200 hot path "
                                      "(gate types 1-7 with GPNN types 0-3 are)")
        if gate_type in (5, 6) and input_size != hidden_size:
            raise ValueError(f"GP-LSTM gate_type {gate_type} feeds a [*, hidden] tensor to GPNN(input_size, ...): the "
                             f"reference needs input_size == hidden_size, got {input_size} and {hidden_size}")
        self.input_size, self.hidden_size, self.gate_type, self.gpnn_type = input_size, hidden_size, gate_type, gpnn_type
        if gate_type <= 4:
            act_set = ("sigmoid",) if gate_type == 2 else ("sigmoid", "tanh", "relu")
            self.gpnn = GPNN(hidden_size + input_size, hidden_size, act_set=act_set, gpnn_type=gpnn_type)
        elif gate_type == 5:
            self.gpnn = GPNN(input_size, hidden_size, act_set=("sigmoid", "tanh", "relu"), gpnn_type=gpnn_type)
        else:
            self.gpnn = GPNN(input_size, 4 * hidden_size, act_set=("sigmoid", "tanh", "relu"), gpnn_type=gpnn_type)
        s = 1.0 / math.sqrt(hidden_size)
        self.weights_ih = _uniform((4 * hidden_size, input_size), -s, s)
        self.bias_ih = nn.Parameter(torch.zeros(4 * hidden_size))
        self.weights_hh = _uniform((4 * hidden_size, hidden_size), -s, s)
        self.bias_hh = nn.Parameter(torch.zeros(4 * hidden_size))


class GPLSTM(nn.Module):
    """``gpnn_type`` string (the ``--L_gauss_pos`` flag): '<gate><type>' -> GP cell + nn.LSTM; three characters ->
    nn.LSTM + GP cell; four -> two GP cells (gates [0] and [2], GPNN type [1]); '0x' -> plain nn.LSTM
    (model.py:1609-1636).  The nn.LSTM members only hold parameters (keys weight_ih_l0, ...)."""

    def __init__(self, input_size, hidden_size, num_layers=1, bias=True, dropout=0.0, gpnn_type="00"):
        super().__init__()
        if num_layers != 2:
            raise NotImplementedError("the GP-LSTM of the reference recipes has two layers")
        self.input_size, self.hidden_size, self.num_layers, self.gpnn_type = input_size, hidden_size, num_layers, gpnn_type
        t = gpnn_type
        cell = lambda gate: GPLSTMCell(input_size, hidden_size, gate_type=int(gate), gpnn_type=int(t[1]))  # noqa: E731
        plain = lambda n: nn.LSTM(input_size=hidden_size, hidden_size=hidden_size, num_layers=n)  # noqa: E731
        if int(t[0]) == 0:
            members = [plain(num_layers)]
        elif len(t) == 2:
            members = [cell(t[0]), plain(num_layers - 1)]
        elif len(t) == 3:
            members = [plain(num_layers - 1), cell(t[0])]
        else:
            members = [cell(t[0]), cell(t[2])]
        self.rnn = nn.ModuleList(members)


class VNN(nn.Module):
    """hidden_lgstd (1, input_size); noise only in training (model.py:2534-2579)."""

    def __init__(self, input_size):
        super().__init__()
        s = 1.0 / math.sqrt(input_size)
        self.sample = True
        self.hidden_lgstd = _lgstd_like((1, input_size), s)
        self.hidden_mean = None     # pure hidden state of the last step of the last training-mode forward (model.py:2570)

    def kl_divergence(self, prior=None):
        """mean(hidden_mean^2 - 2 hidden_lgstd + exp(2 hidden_mean) - 1) / 2 as the reference writes it
        (model.py:2545-2551; called by train.py:372-377 after a training-mode forward)."""
        if prior is not None:
            raise NotImplementedError("prior-centred KL is not on the train.py path")
        if self.hidden_mean is None:
            raise RuntimeError("VNN.kl_divergence() needs a training-mode forward first (it reads the hidden it stored)")
        return _engine.vnn_kl(self)


class VLSTMCell(nn.Module):
    """LSTM cell with the doubled ``bias_ih`` (model.py:2519) and a VNN that perturbs h in training."""
    kind = "plain"

    def __init__(self, input_size, hidden_size, vnn_type=0):
        super().__init__()
        self.input_size, self.hidden_size, self.vnn_type = input_size, hidden_size, vnn_type
        self.vnn = VNN(input_size)
        s = 1.0 / math.sqrt(hidden_size)
        self.weights_ih = _uniform((4 * hidden_size, input_size), -s, s)
        self.bias_ih = nn.Parameter(torch.zeros(4 * hidden_size))
        self.weights_hh = _uniform((4 * hidden_size, hidden_size), -s, s)
        self.bias_hh = nn.Parameter(torch.zeros(4 * hidden_size))


class VariationalLSTM(nn.Module):
    """Two VLSTMCells; ``vlstm_type`` = the ``--L_v_pos`` bit string (model.py:2426-2468)."""

    def __init__(self, input_size, hidden_size, num_layers=1, bias=True, dropout=0.0, vlstm_type="00"):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers, self.vlstm_type = input_size, hidden_size, num_layers, vlstm_type
        self.rnn = nn.ModuleList([VLSTMCell(input_size, hidden_size, vnn_type=int(vlstm_type[0])),
                                  VLSTMCell(input_size, hidden_size, vnn_type=int(vlstm_type[1]))])


class _CellRNNModel(nn.Module):
    """Shared container of GaussRNNModel / VariationalRNNModel: encoder, rnn, decoder (tied)."""

    def _finish(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights):
        if rnn_type != "LSTM":
            raise NotImplementedError("only --model LSTM is on the rescoring path")
        self.rnn_type, self.nhid, self.nlayers, self.p_drop = rnn_type, nhid, nlayers, float(dropout)
        self.encoder = nn.Embedding(ntoken, ninp)
        self.decoder = nn.Linear(nhid, ntoken)
        if tie_weights:
            if nhid != ninp:
                raise ValueError("When using the tied flag, nhid must be equal to emsize.")
            self.decoder.weight = self.encoder.weight
        nn.init.uniform_(self.encoder.weight, -0.1, 0.1)
        nn.init.zeros_(self.decoder.bias)
        nn.init.uniform_(self.decoder.weight, -0.1, 0.1)

    def init_hidden(self, bsz):
        w = self.encoder.weight
        return (w.new_zeros(self.nlayers, bsz, self.nhid), w.new_zeros(self.nlayers, bsz, self.nhid))

    def forward(self, x, hidden):
        """(T, B) int64, (h, c) each (2, B, H) -> logits (T, B, V), (h, c); eval semantics (posterior means)."""
        return _engine.lstm_logits(self, x, hidden)


class GaussRNNModel(_CellRNNModel):
    """GaussRNNModel(rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights, gauss_pos:str) (model.py:1317-1366)."""
    family = "gauss_lstm"

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False, gauss_pos="00"):
        super().__init__()
        self.gauss_pos = gauss_pos
        self.rnn = GPLSTM(ninp, nhid, nlayers, dropout=dropout, gpnn_type=gauss_pos)
        self._finish(rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights)


class VariationalRNNModel(_CellRNNModel):
    """VariationalRNNModel(..., tie_weights, v_pos:str) (model.py:2373-2423)."""
    family = "v_lstm"

    def __init__(self, rnn_type, ntoken, ninp, nhid, nlayers, dropout=0.5, tie_weights=False, v_pos="00"):
        super().__init__()
        self.v_pos = v_pos
        if nlayers != 2:
            raise NotImplementedError("VariationalLSTM is hard-wired to two cells in the reference")
        self.rnn = VariationalLSTM(ninp, nhid, nlayers, dropout=dropout, vlstm_type=v_pos)
        self._finish(rnn_type, ntoken, ninp, nhid, nlayers, dropout, tie_weights)


def build_model(args, ntokens):
    """The model-selection switch of the scorer / trainer (score.py:374-448, train.py:193-224),
    for the families on the hot path."""
    unc = args.uncertainty
    tied = getattr(args, "tied", True)     # the scorer ties every model_1 (score.py:377-440); the trainer passes --tied
    if args.model == "Transformer":
        common = (ntokens, args.emsize, args.nhead, args.nhid, args.nlayers, getattr(args, "dropout", 0.5), tied)
        if unc == "Bayesian":
            return BayesTransformerModel(*common, args.T_bayes_pos)
        if unc == "none":    # score.py:377, train.py:198: nn.TransformerEncoder keys
            return TransformerModel(*common[:-1], "gelu", common[-1])
        if unc == "Gaussian":
            return GaussTransformerModel(*common, args.T_gauss_pos)
        if unc == "Variational":
            return VTransformerModel(*common, args.T_v_pos)
    elif args.model == "LSTM":
        common = ("LSTM", ntokens, args.emsize, args.nhid, args.nlayers, getattr(args, "dropout", 0.5), tied)
        if unc == "Bayesian":
            return BayesRNNModel(*common, args.L_bayes_pos)
        if unc == "none":    # score.py:413, train.py:204: nn.LSTM keys
            return RNNModel(*common)
        if unc == "Gaussian":   # the scorer builds this family untied (score.py:428), the trainer with --tied (train.py:219)
            return GaussRNNModel(*common[:-1], getattr(args, "tied", False), args.L_gauss_pos)
        if unc == "Variational":
            return VariationalRNNModel(*common, args.L_v_pos)
    raise NotImplementedError(f"--model {args.model} --uncertainty {unc} is outside the B200 hot path")
