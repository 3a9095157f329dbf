"""Layer-by-layer error of the precise (bf16x3) path against the CPU oracle at BASELINE layer sizes."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from bayeslms_b200 import _lib, ops, engine, model as M
from oracle import bayeslm_oracle as O

_lib.init(0)
dev = torch.device("cuda:0")
V, D, NHEAD, FF, NL = 30000, 512, 8, 4096, 2
torch.manual_seed(1111)
net = M.BayesTransformerModel(V, D, NHEAD, FF, NL, 0.5, True, "FFN")
with torch.no_grad():
    net.decoder.bias.uniform_(-0.1, 0.1)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
cfg = O.Config(family="bayes_tm", bayes_pos="FFN", ntoken=V, ninp=D, nhead=NHEAD, nhid=FF, nlayers=NL)
net = net.to(dev).eval()
g = torch.Generator().manual_seed(5)
toks = [0] + torch.randint(2, V, (26,), generator=g).tolist()
tg = toks[1:] + [0]
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"

def rep(name, got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = (got - ref)
    align = (err * ref).sum() / (ref * ref).sum()
    print(f"{name:28s} max|err| {err.abs().max():.3e}  mean err {err.mean():+.3e}  rel-to-max {err.abs().max()/ref.abs().max():.3e}  align {align:+.3e}  ref max {ref.abs().max():.3e}", flush=True)

with torch.no_grad():
    # oracle intermediates (fp32 CPU) and the same in float64 as ground truth
    def oracle_stages(sd_, dt):
        sdx = {k: v.to(dt) for k, v in sd_.items()}
        x = F.embedding(torch.tensor(toks).view(-1, 1), sdx["encoder.weight"]) * math.sqrt(D)
        x = x + sdx["pos_encoder.pe"][: len(toks)].to(dt) if "pos_encoder.pe" in sdx else x + O.positional_encoding(5000, D).unsqueeze(1)[: len(toks)].to(dt)
        out = [x]
        mask = O.causal_mask(len(toks)).to(dt)
        for i, kind in enumerate(O.tm_layer_kinds(cfg)):
            pre = f"transformerlayers.{i}."
            a = O.mha(x, sdx, pre + "self_attn.", NHEAD, mask)
            x1 = F.layer_norm(x + a, (D,), sdx[pre + "norm1.weight"], sdx[pre + "norm1.bias"], 1e-5)
            h = F.gelu(F.linear(x1, sdx[pre + "linear1.weight"], sdx[pre + "linear1.bias"]))
            if kind == "bayes_ffn":
                f = F.linear(h, sdx[pre + "linear2.weight_mean"])
            else:
                f = F.linear(h, sdx[pre + "linear2.weight"], sdx[pre + "linear2.bias"])
            x = F.layer_norm(x1 + f, (D,), sdx[pre + "norm2.weight"], sdx[pre + "norm2.bias"], 1e-5)
            out += [a, x1, h, f, x]
        lg = F.linear(x, sdx["decoder.weight"], sdx["decoder.bias"])
        out.append(lg)
        nll = F.cross_entropy(lg.view(-1, V), torch.tensor(tg), reduction="none")
        out.append(nll)
        return out
    sd_pe = dict(sd); sd_pe["pos_encoder.pe"] = O.positional_encoding(5000, D).unsqueeze(1)
    o32 = oracle_stages(sd_pe, torch.float32)
    o64 = oracle_stages(sd_pe, torch.float64)
    names = ["x0"] + [f"L{i}.{n}" for i in range(NL) for n in ("attn_out", "x1", "h", "ffn", "x2")] + ["logits", "token_nll"]
    print("---- oracle fp32 vs fp64")
    for n, a, b in zip(names, o32, o64):
        rep(n, a.squeeze(1) if a.dim() == 3 else a, b.squeeze(1) if b.dim() == 3 else b)

    plan = engine.plan_for(net, prec)
    batch = engine.PackedBatch.from_lists([toks], [tg], dev)
    run = engine._TmRun(net, plan, batch)
    carry = run.prefix(None)
    x32, xs = carry["x"]
    print(f"---- GPU {prec} vs oracle fp64")
    rep("x0", x32, o64[0].squeeze(1))
    k = 1
    for i, L in enumerate(plan.layers):
        att = run.part_a(L, xs)
        y = run.f32(D)
        ops.gemm(att, L["o"], prec=prec, bias=L["o_b"], out_f32=y)
        rep(f"L{i}.attn_out", y, o64[k].squeeze(1))
        x1_32, x1s = run.part_b(L, x32, att, L["o"])
        rep(f"L{i}.x1", x1_32, o64[k + 1].squeeze(1))
        h = run.part_c(L, x1s, L["w1"], L["b1"], None)
        rep(f"L{i}.h", h.float(), o64[k + 2].squeeze(1))
        f = run.f32(D)
        ops.gemm(h, L["w2"], prec=prec, bias=L["b2"], out_f32=f)
        rep(f"L{i}.ffn", f, o64[k + 3].squeeze(1))
        x32, xs = run.part_d(L, x1_32, h, L["w2"])
        rep(f"L{i}.x2", x32, o64[k + 4].squeeze(1))
        k += 5
    ld = (V + 7) // 8 * 8
    logits = torch.empty(len(toks), ld, device=dev)
    ops.gemm(xs, plan.E, prec=prec, bias=plan.dec_b, out_f32=logits)
    rep("logits", logits[:, :V], o64[k].squeeze(1))
    nll = ops.vocab_nll(xs, plan.E, plan.dec_b, batch.targets, prec=prec)
    rep("token_nll (kernel)", nll, o64[k + 1])
    ref_from_gpu_logits = F.cross_entropy(logits[:, :V].double(), torch.tensor(tg, device=dev), reduction="none")
    rep("token_nll kernel vs own lg", nll, ref_from_gpu_logits)
    print("sum nll gpu", nll.sum().item(), "oracle32", o32[-1].sum().item(), "oracle64", o64[-1].sum().item())

    # ---- isolated stages: each GPU stage fed with the ORACLE's (fp64 -> fp32) input
    print("---- isolated stages (oracle input -> GPU stage) vs oracle fp64")
    def sp(t):
        return ops.split(t.squeeze(1).float().to(dev).contiguous(), prec)
    k = 1
    x_in = o64[0]
    for i, L in enumerate(plan.layers):
        xs_o = sp(x_in)
        att = run.part_a(L, xs_o)
        y = run.f32(D)
        ops.gemm(att, L["o"], prec=prec, bias=L["o_b"], out_f32=y)
        rep(f"L{i}.attn_out", y, o64[k].squeeze(1))
        pre1 = (x_in + o64[k]).squeeze(1).float().to(dev).contiguous()
        g_, b_, e_ = L["norm1"]
        x1g, _ = ops.layernorm(pre1, g_, b_, e_, prec=prec)
        rep(f"L{i}.x1 (LN only)", x1g, o64[k + 1].squeeze(1))
        h = run.part_c(L, sp(o64[k + 1]), L["w1"], L["b1"], None)
        rep(f"L{i}.h", h.float(), o64[k + 2].squeeze(1))
        f = run.f32(D)
        ops.gemm(sp(o64[k + 2]), L["w2"], prec=prec, bias=L["b2"], out_f32=f)
        rep(f"L{i}.ffn", f, o64[k + 3].squeeze(1))
        x_in = o64[k + 4]
        k += 5
    ops.gemm(sp(x_in), plan.E, prec=prec, bias=plan.dec_b, out_f32=logits)
    rep("logits", logits[:, :V], o64[k].squeeze(1))
    top = o64[k].squeeze(1).argmax(-1)
    zt = logits[:, :V].double().cpu().gather(1, top.view(-1, 1)).squeeze(1) - o64[k].squeeze(1).gather(1, top.view(-1, 1)).squeeze(1)
    print("signed error of the top logit per token (isolated decoder):", [f"{v:+.1e}" for v in zt.tolist()][:12])
