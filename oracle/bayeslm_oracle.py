"""CPU oracle for the BayesLMs n-best rescoring / fine-tune hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bayeslms_b200/`` may import this file; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline /
``--impl reference`` legs as the checker and the timed CPU arm.

What it is: a functional restatement, in plain fp32 PyTorch CPU ops over a ``state_dict``
(the reference's own key names), of the algorithm in the reference's
``steps/pytorchnn/model.py`` and ``steps/pytorchnn/compute_sentence_scores_bayes_jianwei.py``
(abbreviated ``model.py`` / ``score.py`` below; line numbers refer to the reference
checkout).  The arithmetic itself lives in a third-party dependency of the reference --
PyTorch, un-pinned (the ``requirements.txt`` its README mentions is absent) -- so the
LSTM cell, LayerNorm, GELU, softmax and cross-entropy follow PyTorch's published
definitions and are anchored on the reference's call sites.

Parity pinning: the reference ships no tests or golden vectors.  This oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the
unmodified reference ``model.py`` (possible only in the build container), runs each
model family in eval mode and in train mode with seeded noise, and commits inputs,
weights and outputs as ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks
every function here against them.

Positions taken on reference defects:
  * sigma = exp(lgstd) everywhere (model.py:670,1086,1876), not softplus;
  * ``GPNN.sample`` gates sampling (model.py:1799,1876); callers pass eps only when
    they mean "sample is True";
  * Bayes2LSTM KL uses layer-1 tensors only and has no "-1" (model.py:737-765);
  * GPNN KL has the "-1" (model.py:1821-1825);
  * V-Transformer training forward (crashes as shipped, model.py:2785 reads an undefined
    attribute): noise is added out of place, ``hidden`` = the noised FFN output, KL is
    taken on it -- forward-value-identical to the aliased in-place code (model.py:2799-2801).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- config
class Config(dict):
    """family: 'bayes_tm' | 'gauss_tm' | 'v_tm' | 'bayes_lstm' | 'gauss_lstm' | 'v_lstm' | 'std_tm' | 'std_lstm'; plus
    ntoken, ninp, nhead, nhid, nlayers and the family's position flag (bayes_pos / gauss_pos / v_pos)."""
    __getattr__ = dict.__getitem__


_TORCH_TM_KEYS = {"self_attn.in_proj_weight": "self_attn.qkv_net.weight", "self_attn.in_proj_bias": "self_attn.qkv_net.bias",
                  "self_attn.out_proj.weight": "self_attn.o_net.weight", "self_attn.out_proj.bias": "self_attn.o_net.bias"}


def canonical(sd: SD, cfg: Config) -> Tuple[SD, Config]:
    """The baseline models of ``--uncertainty none`` are torch library modules in the reference: ``TransformerModel``
    = nn.TransformerEncoder of post-LN nn.TransformerEncoderLayer(activation='gelu') (model.py:131-133),
    ``RNNModel`` = nn.LSTM (model.py:35).  Their published arithmetic is the one restated below for the reference's
    own StandardTransformerEncoderLayer (fused in_proj = qkv_net, q scaled by head_dim^-1/2, additive causal mask,
    softmax, out_proj = o_net, erf GELU, LayerNorm eps 1e-5) and for the Bayesian LSTM at position 0 (gate order
    i, f, g, o, two biases), so they are evaluated by renaming their keys; pinned by tests/golden/std_{tm,lstm}.pt."""
    fam = cfg.family
    if fam == "std_tm":
        out = {}
        for k, v in sd.items():
            if k.startswith("transformerlayers.layers."):
                i, rest = k[len("transformerlayers.layers."):].split(".", 1)
                k = f"transformerlayers.{i}.{_TORCH_TM_KEYS.get(rest, rest)}"
            out[k] = v
        return out, Config(cfg, family="bayes_tm", bayes_pos="none")
    if fam == "std_lstm":
        out = {}
        for k, v in sd.items():
            if k.startswith("rnn.") and "_l" in k:
                name, l = k[4:].rsplit("_l", 1)
                k = f"rnn.{name}_mean_{int(l) + 1}"
            out[k] = v
        return out, Config(cfg, family="bayes_lstm", bayes_pos=0)
    return sd, cfg


def tm_layer_kinds(cfg: Config) -> List[str]:
    """Layer layout of the three Transformer families (model.py:1193-1214, 2304-2312, 2822-2842)."""
    n = cfg.nlayers
    fam = cfg.family
    if fam == "bayes_tm":
        pos = cfg.bayes_pos
        if pos in ("none", "EMB"):
            return ["std"] * n
        if pos == "FFN":
            return ["bayes_ffn"] + ["std"] * (n - 1)
        if pos == "MHA":
            return ["bayes_mha"] + ["std"] * (n - 1)
        return []  # any other string builds no layers at all
    if fam == "gauss_tm":
        g = cfg.gauss_pos
        if g > 4:
            return ["std"] * n
        return ["gauss"] + ["std"] * (n - 1)
    if fam == "v_tm":
        v = cfg.v_pos
        if v == 0:
            return ["std"] * n
        if v == 1:
            return ["v"] + ["std"] * (n - 1)
        if v == 2:
            return ["std", "v"] + ["std"] * (n - 3)   # nlayers-3: one layer short, kept (model.py:2834)
        if v == 3:
            return ["v", "v"] + ["std"] * (n - 3)
        return []  # e.g. T_v_pos=11 as typed in the README builds zero layers
    raise ValueError(fam)


# ----------------------------------------------------------------- transformer pieces
def positional_encoding(max_len: int, d: int) -> Tensor:
    """model.py:97-102: pe[pos, 2i] = sin(pos * w_i), pe[pos, 2i+1] = cos(pos * w_i)."""
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def causal_mask(T: int) -> Tensor:
    """model.py:1258-1262: 0 on and below the diagonal, -inf above."""
    m = torch.full((T, T), float("-inf"))
    return torch.triu(m, diagonal=1)


def _attention_core(q: Tensor, k: Tensor, v: Tensor, nhead: int, mask: Optional[Tensor],
                    drop: Optional[Tensor] = None) -> Tensor:
    """model.py:889-920: heads are contiguous column groups; bmm, additive mask, softmax, [dropout], bmm.
    ``drop``: the nn.Dropout of model.py:913 as an injected multiplier tensor (B * nhead, T, T), values 0 or 1/(1-p)."""
    T, B, d = q.shape
    hd = d // nhead
    q = q.contiguous().view(T, B * nhead, hd).transpose(0, 1)
    k = k.contiguous().view(-1, B * nhead, hd).transpose(0, 1)
    v = v.contiguous().view(-1, B * nhead, hd).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))
    if mask is not None:
        w = w + mask.unsqueeze(0)
    w = F.softmax(w, dim=-1)
    if drop is not None:
        w = w * drop
    o = torch.bmm(w, v)
    return o.transpose(0, 1).contiguous().view(T, B, d)


def mha(x: Tensor, sd: SD, pre: str, nhead: int, mask: Optional[Tensor], drop: Optional[Tensor] = None) -> Tensor:
    """MultiheadAttention.forward, model.py:871-928 (fused qkv_net, q scaled after the bias)."""
    d = x.shape[-1]
    scaling = float(d // nhead) ** -0.5
    q, k, v = F.linear(x, sd[pre + "qkv_net.weight"], sd[pre + "qkv_net.bias"]).chunk(3, dim=-1)
    o = _attention_core(q * scaling, k, v, nhead, mask, drop)
    return F.linear(o, sd[pre + "o_net.weight"], sd[pre + "o_net.bias"])


def bayes_linear_weight(mu: Tensor, lgstd: Tensor, eps: Optional[Tensor]) -> Tensor:
    """BayesLinear._flat_weights, model.py:1098-1102 with sample_weight_diff 1086-1088."""
    return mu if eps is None else mu + eps * torch.exp(lgstd)


def bayes_mha(x: Tensor, sd: SD, pre: str, nhead: int, mask: Optional[Tensor], eps_o: Optional[Tensor],
              drop: Optional[Tensor] = None) -> Tensor:
    """BayesMultiheadAttention.forward, model.py:971-1019: separate q/k/v nets, bias-free Bayesian o_net."""
    d = x.shape[-1]
    scaling = float(d // nhead) ** -0.5
    q = F.linear(x, sd[pre + "q_net.weight"], sd[pre + "q_net.bias"]) * scaling
    k = F.linear(x, sd[pre + "k_net.weight"], sd[pre + "k_net.bias"])
    v = F.linear(x, sd[pre + "v_net.weight"], sd[pre + "v_net.bias"])
    o = _attention_core(q, k, v, nhead, mask, drop)
    w = bayes_linear_weight(sd[pre + "o_net.weight_mean"], sd[pre + "o_net.weight_lgstd"], eps_o)
    return F.linear(o, w)


def _ln(x: Tensor, sd: SD, pre: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)


GP_ACTS = ("tanh", "sigmoid", "relu", "gelu")  # act_set order of model.py:2263


def gpnn(x: Tensor, sd: SD, pre: str, gpnn_type: int, eps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """GPNN.forward, model.py:1863-1902.  eps = {'coef','weights','bias'} when sampling
    (model.py:1876-1883: types 1,3 sample coef; types 2,3 sample weights and bias)."""
    coef, w, b = sd[pre + "coef_mean"], sd[pre + "weights_mean"], sd[pre + "bias_mean"]
    if eps is not None:
        if gpnn_type in (1, 3):
            coef = coef + torch.exp(sd[pre + "coef_lgstd"]) * eps["coef"]
        if gpnn_type in (2, 3):
            w = w + torch.exp(sd[pre + "weights_lgstd"]) * eps["weights"]
            b = b + torch.exp(sd[pre + "bias_lgstd"]) * eps["bias"]
    z = F.linear(x, w, b)
    acts = (torch.tanh(z), torch.sigmoid(z), F.relu(z), F.gelu(z))
    return sum(a * coef[i] for i, a in enumerate(acts))


def tm_layer(x: Tensor, sd: SD, pre: str, kind: str, nhead: int, mask: Optional[Tensor], cfg: Config,
             eps=None, aux: Optional[dict] = None, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """One post-LN encoder layer (model.py:1037-1046, 1162-1176, 2274-2287, 2792-2805).
    ``eps`` is the layer's injected noise (None = posterior mean / eval).  ``masks`` injects the layer's four
    training-mode nn.Dropout draws as multiplier tensors (0 or 1/(1-p)): 'attn' (B*nhead, T, T) on the attention
    probabilities (model.py:913), 'd1' (T, B, d) = dropout1 on the attention branch, 'ffn' (T, B, F) = dropout on the
    activation, 'd2' (T, B, d) = dropout2 on the FFN branch -- after the variational noise, which is added in place
    first (model.py:2799-2803)."""
    masks = masks or {}
    if kind == "bayes_mha":
        a = bayes_mha(x, sd, pre + "self_attn.", nhead, mask, eps, masks.get("attn"))
    else:
        a = mha(x, sd, pre + "self_attn.", nhead, mask, masks.get("attn"))
    if "d1" in masks:
        a = a * masks["d1"]
    x = _ln(x + a, sd, pre + "norm1.")
    if kind == "gauss":
        h = gpnn(x, sd, pre + "gpnn.", cfg.gauss_pos, eps)
    else:
        # nn.TransformerEncoderLayer(activation=...) of TransformerModel (model.py:124, 131-135): 'relu' is the constructor's
        # default, 'gelu' what every caller passes; the reference's own layers are GELU only (model.py:1035)
        act = F.relu if cfg.get("activation", "gelu") == "relu" else F.gelu
        h = act(F.linear(x, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"]))
    if "ffn" in masks:
        h = h * masks["ffn"]
    if kind == "bayes_ffn":
        w2 = bayes_linear_weight(sd[pre + "linear2.weight_mean"], sd[pre + "linear2.weight_lgstd"], eps)
        f = F.linear(h, w2)
    else:
        f = F.linear(h, sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
    if kind == "v" and eps is not None:
        # model.py:2785-2788, 2799-2801: eps' ~ N(0, 0.1^2) scaled by exp(f * hiddens_lgstd); T must be 100
        assert f.shape[0] == 100, "the reference only adds the variational noise at sequence length 100"
        f = f + eps * torch.exp(f * sd[pre + "hiddens_lgstd"])
    if aux is not None and kind == "v":
        aux[pre + "hidden"] = f
    if "d2" in masks:
        f = f * masks["d2"]
    return _ln(x + f, sd, pre + "norm2.")


def transformer_hidden(sd: SD, tokens: Tensor, cfg: Config, eps: Optional[dict] = None,
                       aux: Optional[dict] = None, masks: Optional[dict] = None) -> Tensor:
    """Everything of {Bayes,Gauss,V}TransformerModel.forward before the decoder
    (model.py:1274-1304, 2341-2360, 2871-2891).  tokens: (T, B) int64.
    eps: {'layer<i>': noise for layer i, 'embed': noise for the EMB variant}.
    masks (training-mode dropout as injected multipliers): {'pe': (T, B, d) for PositionalEncoding's dropout
    (model.py:116), 'layer<i>': the layer's masks, see tm_layer}."""
    masks = masks or {}
    sd, cfg = canonical(sd, cfg)
    eps = eps or {}
    T = tokens.shape[0]
    d = cfg.ninp
    mask = causal_mask(T)
    x = F.embedding(tokens, sd["encoder.weight"]) * math.sqrt(d)
    emb_variant = cfg.family == "bayes_tm" and cfg.bayes_pos == "EMB"
    if emb_variant:
        w = bayes_linear_weight(sd["embed_mean"], sd["embed_lgstd"], eps.get("embed"))
        x = F.linear(x, w)
    pe = sd["pos_encoder.pe"] if "pos_encoder.pe" in sd else positional_encoding(5000, d).unsqueeze(1)
    x = x + pe[:T]
    if "pe" in masks:
        x = x * masks["pe"]
    for i, kind in enumerate(tm_layer_kinds(cfg)):
        x = tm_layer(x, sd, f"transformerlayers.{i}.", kind, cfg.nhead, mask, cfg, eps.get(f"layer{i}"), aux,
                     masks.get(f"layer{i}"))
    if emb_variant:
        x = F.linear(x, sd["embed_mean"].t())  # model.py:1303: mean only, transposed
    return x


def transformer_forward(sd: SD, tokens: Tensor, cfg: Config, eps: Optional[dict] = None) -> Tensor:
    """logits (T, B, V) = decoder(hidden), model.py:1304-1306."""
    h = transformer_hidden(sd, tokens, cfg, eps)
    return F.linear(h, sd["decoder.weight"], sd["decoder.bias"])


def transformer_forward_masks(sd: SD, tokens: Tensor, cfg: Config, eps: Optional[dict], masks: Optional[dict]) -> Tensor:
    """Training-mode logits with injected parameter noise and dropout masks."""
    h = transformer_hidden(sd, tokens, cfg, eps, None, masks)
    return F.linear(h, sd["decoder.weight"], sd["decoder.bias"])


# ------------------------------------------------------------------------------ LSTM
LSTM_EPS_ORDER = ("weight_hh_1", "weight_ih_1", "bias_hh_1", "bias_ih_1",
                  "weight_hh_2", "weight_ih_2", "bias_hh_2", "bias_ih_2")  # draw order, model.py:670-700


def lstm_flat_parameters(sd: SD, pos: int, eps: Optional[Dict[str, Tensor]] = None, pre: str = "rnn.") -> Dict[str, Tensor]:
    """Bayes2LSTM.flat_parameters, model.py:705-732: copy the means; in sampling mode add
    eps*exp(lgstd) to rows [(pos-1)H, pos*H) of W_hh, W_ih, b_hh, b_ih of both layers."""
    out = {}
    for layer in (1, 2):
        for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            p = sd[f"{pre}{name}_mean_{layer}"].clone()
            if eps is not None and 1 <= pos <= 4:
                H = sd[f"{pre}weight_hh_mean_1"].shape[1]
                lg = sd[f"{pre}{name}_lgstd_{layer}"]
                p[(pos - 1) * H: pos * H] += eps[f"{name}_{layer}"] * torch.exp(lg)
            out[f"{name}_{layer}"] = p
    return out


def lstm_layer(x: Tensor, h: Tensor, c: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor):
    """One LSTM layer as PyTorch defines it (the ``_VF.lstm`` call of model.py:812): gate order i,f,g,o."""
    outs = []
    for t in range(x.shape[0]):
        gates = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs), h, c


def rnn_forward(sd: SD, tokens: Tensor, hidden: Tuple[Tensor, Tensor], cfg: Config,
                eps: Optional[Dict[str, Tensor]] = None, return_hidden_states: bool = False,
                masks: Optional[Dict[str, Tensor]] = None):
    """BayesRNNModel.forward in eval / injected-noise mode, model.py:217-222 + 783-828.
    tokens (T, B); hidden = (h, c) each (2, B, H).  Returns logits (T, B, V), (h, c).
    The GP / Variational cell families (eval mode) are routed to ``cell_rnn_forward``.
    masks: the two training-mode ``self.drop`` draws of model.py:218,220 as multipliers: 'emb' (T, B, ninp) on the
    embedding, 'out' (T, B, H) on the LSTM output (the inter-layer dropout is the literal 0., model.py:813)."""
    masks = masks or {}
    if cfg.family in ("gauss_lstm", "v_lstm"):
        assert not return_hidden_states
        return cell_rnn_forward(sd, tokens, hidden, cfg, eps, masks)
    sd, cfg = canonical(sd, cfg)
    p = lstm_flat_parameters(sd, cfg.bayes_pos, eps)
    x = F.embedding(tokens, sd["encoder.weight"])
    if "emb" in masks:
        x = x * masks["emb"]
    h0, c0 = hidden
    hs, cs = [], []
    for layer in (1, 2):
        x, h, c = lstm_layer(x, h0[layer - 1], c0[layer - 1], p[f"weight_ih_{layer}"], p[f"weight_hh_{layer}"],
                             p[f"bias_ih_{layer}"], p[f"bias_hh_{layer}"])
        hs.append(h)
        cs.append(c)
    if "out" in masks:
        x = x * masks["out"]
    new_hidden = (torch.stack(hs), torch.stack(cs))
    if return_hidden_states:
        return x, new_hidden
    return F.linear(x, sd["decoder.weight"], sd["decoder.bias"]), new_hidden


# ---------------------------------------------------------------- GP-LSTM / Variational-LSTM cells
LSTM_GP_ACTS = {1: ("sigmoid", "tanh", "relu"), 2: ("sigmoid",), 3: ("sigmoid", "tanh", "relu"),
                4: ("sigmoid", "tanh", "relu"), 5: ("sigmoid", "tanh", "relu"), 6: ("sigmoid", "tanh", "relu"),
                7: ("sigmoid", "tanh", "relu")}   # act_set per gate_type, model.py:1690-1697 (+ GPNN default 1787)


def gp_lstm_layout(gauss_pos: str) -> List[Tuple[str, int, int]]:
    """Members of GPLSTM.rnn for the ``--L_gauss_pos`` string (model.py:1619-1636):
    ('gp', gate_type, gpnn_type) or ('lstm', n_layers, 0)."""
    t = gauss_pos
    if int(t[0]) == 0:
        return [("lstm", 2, 0)]
    if len(t) == 2:
        return [("gp", int(t[0]), int(t[1])), ("lstm", 1, 0)]
    if len(t) == 3:
        return [("lstm", 1, 0), ("gp", int(t[0]), int(t[1]))]
    return [("gp", int(t[0]), int(t[1])), ("gp", int(t[2]), int(t[1]))]


def gp_lstm_cell_layer(x: Tensor, h: Tensor, c: Tensor, sd: SD, pre: str, gate_type: int, gpnn_type: int = 0,
                       eps: Optional[Dict[str, Tensor]] = None):
    """GPLSTMCell.forward / Gplstm, model.py:1720-1777.
    gate_type 1..4: gates = W_ih x + b_ih + W_hh h + b_ih (bias_ih twice, bias_hh never); the chosen gate is replaced by
    GPNN(cat[x, h]) = sum_i coef[i] * act_i(W_g cat[x, h] + b_g).
    gate_type 5 ("cell"): the same gates, and the cell state passes through the GP unit first, c <- GPNN(c)
    (model.py:1763-1764; GPNN(input_size -> H), so input_size must equal H).
    gate_type 6 ("hidden"): gates = W_ih x + b_ih + GPNN(h), GPNN(input_size -> 4H) in place of the recurrent product
    (model.py:1747-1748).  gate_type 7 ("inputs"): gates = GPNN(x) + W_hh h + b_ih (model.py:1749-1750).
    ``eps`` = {'coef','weights','bias'}: the GP unit's parameters sampled ONCE per forward call (sample_parameters at
    model.py:1721-1723; used only when the unit is in training mode with .sample set, model.py:1876-1883); None =
    posterior means."""
    acts = LSTM_GP_ACTS[gate_type]
    coef, wg, bg = sd[pre + "gpnn.coef_mean"], sd[pre + "gpnn.weights_mean"], sd[pre + "gpnn.bias_mean"]
    if eps is not None:
        if gpnn_type in (1, 3):
            coef = coef + torch.exp(sd[pre + "gpnn.coef_lgstd"]) * eps["coef"]
        if gpnn_type in (2, 3):
            wg = wg + torch.exp(sd[pre + "gpnn.weights_lgstd"]) * eps["weights"]
            bg = bg + torch.exp(sd[pre + "gpnn.bias_lgstd"]) * eps["bias"]

    def unit(v):
        z = F.linear(v, wg, bg)
        return sum(getattr(torch, a)(z) * coef[k] for k, a in enumerate(acts))

    w_ih, w_hh, b_ih = sd[pre + "weights_ih"], sd[pre + "weights_hh"], sd[pre + "bias_ih"]
    outs = []
    for t in range(x.shape[0]):
        if gate_type == 6:
            gates = F.linear(x[t], w_ih, b_ih) + unit(h)
        elif gate_type == 7:
            gates = unit(x[t]) + F.linear(h, w_hh, b_ih)
        else:
            gates = F.linear(x[t], w_ih, b_ih) + F.linear(h, w_hh, b_ih)
        i, f, g, o = gates.chunk(4, 1)
        gp = unit(torch.cat([x[t], h], -1)) if gate_type <= 4 else None
        i = gp if gate_type == 1 else torch.sigmoid(i)
        f = gp if gate_type == 2 else torch.sigmoid(f)
        g = gp if gate_type == 3 else torch.tanh(g)
        o = gp if gate_type == 4 else torch.sigmoid(o)
        if gate_type == 5:
            c = unit(c)
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs), h, c


def v_lstm_cell_layer(x: Tensor, h: Tensor, c: Tensor, sd: SD, pre: str, noise: Optional[Tensor] = None):
    """VLSTMCell.forward (model.py:2493-2531): the plain cell with bias_ih on both products; with ``noise`` (T, 1, H)
    -- training mode of a cell built with vnn_type 1 -- every step's h gets noise[t] * exp(vnn.hidden_lgstd) added
    before it is stored and fed back (VNN.forward, model.py:2565-2579; the draw is normal_(0, 0.1), model.py:2561).
    Returns (outputs, h, c, hidden_mean) where hidden_mean is the PURE h of the last step, what VNN.kl_divergence
    reads (model.py:2549, 2570)."""
    outs, h_mean = [], h
    for t in range(x.shape[0]):
        gates = F.linear(x[t], sd[pre + "weights_ih"], sd[pre + "bias_ih"]) + F.linear(h, sd[pre + "weights_hh"], sd[pre + "bias_ih"])
        i, f, g, o = gates.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        h_mean = h
        if noise is not None:
            h = h + noise[t] * torch.exp(sd[pre + "vnn.hidden_lgstd"])
        outs.append(h)
    return torch.stack(outs), h, c, h_mean


def kl_vnn(sd: SD, pre: str, hidden_mean: Tensor) -> Tensor:
    """VNN.kl_divergence as written (model.py:2545-2551): mean(h^2 - 2 lgstd + exp(2 h) - 1) / 2 over (B, H), with h the
    pure hidden state of the last step ("hidden_mean") -- the exponential is of the hidden, not of the log-sigma."""
    ls = sd[pre + "vnn.hidden_lgstd"]
    return torch.mean(hidden_mean ** 2 - ls * 2. + torch.exp(hidden_mean * 2) - 1) / 2.


def cell_rnn_forward(sd: SD, tokens: Tensor, hidden: Tuple[Tensor, Tensor], cfg: Config, eps: Optional[dict] = None,
                     masks: Optional[Dict[str, Tensor]] = None, aux: Optional[dict] = None):
    """GaussRNNModel.forward (model.py:1354-1359, GPLSTM.forward 1638-1671) and VariationalRNNModel.forward
    (model.py:2411-2416, VariationalLSTM.forward 2447-2468).  Returns logits (T, B, V), (h, c).
    eps = None: eval mode.  Training mode: eps = {'cell<mi>': ...} per member of rnn.rnn -- for a GP cell the
    {'coef','weights','bias'} draws (only meaningful when GPNN.sample is set), for a variational cell with vnn_type 1
    the (T, 1, H) per-step noise (already N(0, 0.1^2)); masks = {'emb','out'} dropout multipliers (model.py:1355,1357);
    aux receives 'cell<mi>.hidden_mean' for the VNN KL."""
    eps, masks = eps or {}, masks or {}
    x = F.embedding(tokens, sd["encoder.weight"])
    if "emb" in masks:
        x = x * masks["emb"]
    h0, c0 = hidden
    hs, cs, li = [], [], 0
    if cfg.family == "gauss_lstm":
        for mi, (kind, a, gt) in enumerate(gp_lstm_layout(cfg.gauss_pos)):
            pre = f"rnn.rnn.{mi}."
            if kind == "gp":
                x, h, c = gp_lstm_cell_layer(x, h0[li], c0[li], sd, pre, a, gt, eps.get(f"cell{mi}"))
                hs.append(h), cs.append(c)
                li += 1
            else:
                for l in range(a):
                    x, h, c = lstm_layer(x, h0[li], c0[li], sd[f"{pre}weight_ih_l{l}"], sd[f"{pre}weight_hh_l{l}"],
                                         sd[f"{pre}bias_ih_l{l}"], sd[f"{pre}bias_hh_l{l}"])
                    hs.append(h), cs.append(c)
                    li += 1
    elif cfg.family == "v_lstm":
        for mi in range(2):   # VLSTMCell.lstmcell, model.py:2515-2531: bias_ih on both products
            pre = f"rnn.rnn.{mi}."
            noise = eps.get(f"cell{mi}") if cfg.v_pos[mi] == "1" else None
            x, h, c, h_mean = v_lstm_cell_layer(x, h0[mi], c0[mi], sd, pre, noise)
            if aux is not None:
                aux[f"cell{mi}.hidden_mean"] = h_mean
            hs.append(h), cs.append(c)
    else:
        raise ValueError(cfg.family)
    if "out" in masks:
        x = x * masks["out"]
    return F.linear(x, sd["decoder.weight"], sd["decoder.bias"]), (torch.stack(hs), torch.stack(cs))


def init_hidden(cfg: Config, bsz: int) -> Tuple[Tensor, Tensor]:
    """model.py:224-229."""
    return torch.zeros(cfg.nlayers, bsz, cfg.nhid), torch.zeros(cfg.nlayers, bsz, cfg.nhid)


# -------------------------------------------------------------------------------- KL
def _kl_term(mu: Tensor, lgstd: Tensor, minus_one: bool = False) -> Tensor:
    t = mu ** 2. - lgstd * 2. + torch.exp(lgstd * 2)
    if minus_one:
        t = t - 1
    return torch.mean(t) / 2.


def kl_bayes_lstm(sd: SD, pos: int, pre: str = "rnn.") -> Tensor:
    """Bayes2LSTM.kl_divergence, prior=None, pos 1..4 (model.py:736-765): layer 1 only, no -1."""
    H = sd[pre + "weight_hh_mean_1"].shape[1]
    sl = slice((pos - 1) * H, pos * H)
    w_mu = torch.cat([sd[pre + "weight_hh_mean_1"][sl], sd[pre + "weight_ih_mean_1"][sl]], -1)
    w_ls = torch.cat([sd[pre + "weight_hh_lgstd_1"], sd[pre + "weight_ih_lgstd_1"]], -1)
    b_mu = torch.cat([sd[pre + "bias_hh_mean_1"][sl], sd[pre + "bias_ih_mean_1"][sl]], -1)
    b_ls = torch.cat([sd[pre + "bias_hh_lgstd_1"], sd[pre + "bias_ih_lgstd_1"]], -1)
    return _kl_term(w_mu, w_ls) + _kl_term(b_mu, b_ls)


def kl_bayes_linear(sd: SD, pre: str) -> Tensor:
    """BayesLinear.kl_divergence, model.py:1115 (bias-free in practice)."""
    return _kl_term(sd[pre + "weight_mean"], sd[pre + "weight_lgstd"])


def kl_embed(sd: SD) -> Tensor:
    """BayesTransformerModel.embed_kl_divergence, model.py:1251-1256."""
    return _kl_term(sd["embed_mean"], sd["embed_lgstd"])


def kl_gpnn(sd: SD, pre: str, gpnn_type: int) -> Tensor:
    """GPNN.kl_divergence, model.py:1816-1826 (with the -1)."""
    kl = torch.zeros(())
    if gpnn_type in (1, 3):
        kl = kl + _kl_term(sd[pre + "coef_mean"], sd[pre + "coef_lgstd"], True)
    if gpnn_type in (2, 3):
        kl = kl + _kl_term(sd[pre + "weights_mean"], sd[pre + "weights_lgstd"], True)
        kl = kl + _kl_term(sd[pre + "bias_mean"], sd[pre + "bias_lgstd"], True)
    return kl


def kl_v_layer(sd: SD, pre: str, hidden: Tensor) -> Tensor:
    """VTransformerEncoderLayer.kl_divergence, model.py:2770-2781, on the noised FFN output."""
    prior_mean = hidden * sd[pre + "hiddens_mean_p"]
    ls = sd[pre + "hiddens_lgstd"]
    return torch.mean((hidden - prior_mean) ** 2. - ls * 2. + torch.exp(ls * 2)) / 2.


def model_kl(sd: SD, cfg: Config, aux: Optional[dict] = None) -> Tensor:
    """The KL term train.py adds for each family (train.py:335-399)."""
    fam = cfg.family
    if fam in ("std_tm", "std_lstm"):
        return torch.zeros(())
    if fam == "bayes_lstm":
        # pos 0 = plain LSTM: train.py only adds the KL under --uncertainty Bayesian, and the reference's
        # kl_divergence raises for position 0 (model.py:757-775 falls into the prior branch with prior=None)
        return kl_bayes_lstm(sd, cfg.bayes_pos) if 1 <= cfg.bayes_pos <= 4 else torch.zeros(())
    if fam == "bayes_tm":
        if cfg.bayes_pos == "FFN":
            return kl_bayes_linear(sd, "transformerlayers.0.linear2.")
        if cfg.bayes_pos == "MHA":
            return kl_bayes_linear(sd, "transformerlayers.0.self_attn.o_net.")
        if cfg.bayes_pos == "EMB":
            return kl_embed(sd)
        return torch.zeros(())
    if fam == "gauss_tm":
        return kl_gpnn(sd, "transformerlayers.0.gpnn.", cfg.gauss_pos) if cfg.gauss_pos <= 3 else torch.zeros(())
    if fam == "gauss_lstm":      # train.py:360-371: the GP cells' units, when the position string selects a GP type 1..3
        kl = torch.zeros(())
        for mi, (kind, _, gt) in enumerate(gp_lstm_layout(cfg.gauss_pos)):
            if kind == "gp" and 0 < gt <= 3:
                kl = kl + kl_gpnn(sd, f"rnn.rnn.{mi}.gpnn.", gt)
        return kl
    if fam == "v_lstm":          # train.py:372-377
        kl = torch.zeros(())
        for mi in range(2):
            if cfg.v_pos[mi] == "1":
                kl = kl + kl_vnn(sd, f"rnn.rnn.{mi}.", aux[f"cell{mi}.hidden_mean"])
        return kl
    if fam == "v_tm":
        kl = torch.zeros(())
        for i, kind in enumerate(tm_layer_kinds(cfg)):
            if kind == "v":
                pre = f"transformerlayers.{i}."
                kl = kl + kl_v_layer(sd, pre, aux[pre + "hidden"])
        return kl
    raise ValueError(fam)


# ---------------------------------------------------------------------- noise drawing
def draw_eps(sd: SD, cfg: Config, gen_seed: int) -> dict:
    """Draw every noise tensor of one posterior sample in the reference's order under
    ``torch.manual_seed(gen_seed)`` on the CPU generator -- exactly what the reference's
    train-mode forward consumes (``new_zeros(...).normal_()``, model.py:671-699, 1087, 1246,
    1857-1861), so ``model.train(); torch.manual_seed(s); model(x)`` is reproduced."""
    torch.manual_seed(gen_seed)
    fam = cfg.family
    if fam == "bayes_lstm":
        out = {}
        for key in LSTM_EPS_ORDER:
            name, layer = key.rsplit("_", 1)
            out[key] = torch.zeros_like(sd[f"rnn.{name}_lgstd_{layer}"]).normal_()
        return out
    if fam == "bayes_tm":
        if cfg.bayes_pos == "EMB":
            return {"embed": torch.zeros_like(sd["embed_lgstd"]).normal_()}
        if cfg.bayes_pos == "FFN":
            return {"layer0": torch.zeros_like(sd["transformerlayers.0.linear2.weight_lgstd"]).normal_()}
        if cfg.bayes_pos == "MHA":
            return {"layer0": torch.zeros_like(sd["transformerlayers.0.self_attn.o_net.weight_lgstd"]).normal_()}
        return {}
    if fam == "gauss_tm":
        g = cfg.gauss_pos
        pre = "transformerlayers.0.gpnn."
        e = {}
        if g in (1, 3):
            e["coef"] = torch.zeros_like(sd[pre + "coef_mean"]).normal_()
        if g in (2, 3):
            e["weights"] = torch.zeros_like(sd[pre + "weights_mean"]).normal_()
            e["bias"] = torch.zeros_like(sd[pre + "bias_mean"]).normal_()
        return {"layer0": e} if e else {}
    raise ValueError(f"no parameter noise for family {fam}")


# ------------------------------------------------------------------- scoring contract
def read_vocab(path: str) -> Dict[str, int]:
    """score.py:63-84: index = order of first occurrence; two columns per line."""
    word2idx: Dict[str, int] = {}
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            word = line.split()
            assert len(word) == 2
            if word[0] not in word2idx:
                word2idx[word[0]] = len(word2idx)
    return word2idx


def load_nbest(path: str) -> "OrderedDict[str, List[str]]":
    """score.py:20-51: key = text before the last '-', empty hypothesis -> ' '."""
    nbest: "OrderedDict[str, List[str]]" = OrderedDict()
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            try:
                key, hyp = line.split(" ", 1)
            except ValueError:
                key, hyp = line, " "
            key = key.rsplit("-", 1)[0]
            nbest.setdefault(key, []).append(hyp)
    return nbest


def get_input_and_target(hyp: str, vocab: Dict[str, int]) -> Tuple[List[int], List[int]]:
    """score.py:87-120: input '<s> '+hyp, target hyp+' <s>', OOV -> '<unk>'."""
    unk = vocab.get("<unk>")
    inp = [vocab.get(w, unk) for w in ("<s> " + hyp).split()]
    tgt = [vocab.get(w, unk) for w in (hyp + " <s>").split()]
    if None in inp or None in tgt:
        raise KeyError("<unk>")
    return inp, tgt


def sentence_nll(logits: Tensor, target: Tensor) -> float:
    """score.py:168-170: length * mean cross-entropy."""
    V = logits.shape[-1]
    loss = F.cross_entropy(logits.view(-1, V), target)
    return target.numel() * loss.item()


def token_logprobs(logits: Tensor, target: Tensor) -> Tensor:
    V = logits.shape[-1]
    return -F.cross_entropy(logits.view(-1, V), target, reduction="none")


def compute_scores(nbest, vocab: Dict[str, int], sd: SD, cfg: Config, eps_list: Optional[Sequence[dict]] = None,
                   sd2: Optional[SD] = None, cfg2: Optional[Config] = None, alpha: float = 0.8):
    """score.py:206-280 restated without .cuda()/prints.  One hypothesis at a time, batch 1.
    LSTM: every hypothesis of an utterance starts from the same state; the next utterance
    starts from the state after hypothesis #0 (score.py:271-274); zeros at the start (232).

    eps_list = K injected posterior samples: the harness-defined K-sample score is the
    Monte-Carlo predictive  sum_t -log( 1/K sum_k p_k(w_t) )  (SURVEY.md 8c); for the LSTM
    each sample k carries its own hidden-state chain.  None = posterior mean (the
    reference's eval-mode behaviour).
    sd2/cfg2: second model for logit interpolation alpha*o1 + (1-alpha)*o2 (score.py:157-163).
    """
    is_rnn = cfg.family.endswith("lstm")
    samples = list(eps_list) if eps_list else [None]
    K = len(samples)
    with torch.no_grad():
        if is_rnn:
            hidden = [init_hidden(cfg, 1) for _ in range(K)]
            hidden2 = init_hidden(cfg2, 1) if sd2 is not None else None
        out: "OrderedDict[str, List[Tuple[str, float]]]" = OrderedDict()
        for key, hyps in nbest.items():
            cached, cached2 = [], []
            for hyp in hyps:
                x, y = get_input_and_target(hyp, vocab)
                data = torch.tensor(x, dtype=torch.long).view(-1, 1)
                target = torch.tensor(y, dtype=torch.long).view(-1)
                if sd2 is not None:
                    if is_rnn:
                        logits2, new_h2 = rnn_forward(sd2, data, hidden2, cfg2)
                        cached2.append(new_h2)
                    else:
                        logits2 = transformer_forward(sd2, data, cfg2)
                lps, new_hs = [], []
                for k, eps in enumerate(samples):
                    if is_rnn:
                        logits, nh = rnn_forward(sd, data, hidden[k], cfg, eps)
                        new_hs.append(nh)
                    else:
                        logits = transformer_forward(sd, data, cfg, eps)
                    if sd2 is not None:
                        logits = alpha * logits + (1. - alpha) * logits2
                    if K == 1:
                        score = sentence_nll(logits, target)
                    else:
                        lps.append(token_logprobs(logits, target))
                if K > 1:
                    lp = torch.logsumexp(torch.stack(lps), 0) - math.log(K)
                    score = float(-lp.sum().item())
                if is_rnn:
                    cached.append(new_hs)
                out.setdefault(key, []).append((hyp, score))
            if is_rnn:
                hidden = cached[0]
                if sd2 is not None:
                    hidden2 = cached2[0]
    return out


def write_scores(nbest_and_scores, path: str) -> None:
    """score.py:283-303: '<utt>-<idx from 1> %.4f'."""
    with open(path, "w", encoding="utf-8") as f:
        for key, items in nbest_and_scores.items():
            for idx, (_, score) in enumerate(items, 1):
                f.write("%s %.4f\n" % ("-".join([key, str(idx)]), score))


# --------------------------------------------------- stage-7 combination + synthetic WER
def combine_and_pick(graph: Sequence[float], oldlm: Sequence[float], nn: Sequence[float], w: float) -> int:
    """lmrescore_nbest_pytorchnn_cuda.sh:221-229: total = graph + w*nn + (1-w)*oldlm; best = argmin."""
    tot = [g + w * n + (1. - w) * o for g, o, n in zip(graph, oldlm, nn)]
    return min(range(len(tot)), key=lambda i: (tot[i], i))


def edit_distance(a: Sequence, b: Sequence) -> int:
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


# ------------------------------------------------------------------- fine-tune step
def finetune_loss(sd: SD, tokens: Tensor, targets: Tensor, cfg: Config, eps: Optional[dict], kl_scale: float,
                  hidden=None, masks: Optional[dict] = None):
    """train.py:319-404: loss = CE(mean over tokens) + KL * kl_scale, kl_scale = seq_len / len(train_data).
    ``sd`` may hold tensors requiring grad; autograd through this function is the gradient oracle.
    ``masks``: the step's dropout draws as injected multiplier tensors (see transformer_hidden / rnn_forward)."""
    aux: dict = {}
    if cfg.family in ("gauss_lstm", "v_lstm"):
        logits, _ = cell_rnn_forward(sd, tokens, hidden, cfg, eps, masks, aux)
    elif cfg.family.endswith("lstm"):
        logits, _ = rnn_forward(sd, tokens, hidden, cfg, eps, masks=masks)
    else:
        h = transformer_hidden(sd, tokens, cfg, eps, aux, masks)
        logits = F.linear(h, sd["decoder.weight"], sd["decoder.bias"])
    ce = F.cross_entropy(logits.view(-1, logits.shape[-1]), targets.view(-1))
    kl = model_kl(sd, cfg, aux)
    return ce + kl * kl_scale, ce, kl
