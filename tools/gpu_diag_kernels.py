"""First-contact GPU diagnostics: every kernel against a torch fp32/fp64 reference.
Prints one line per case; exits non-zero on the first gross failure."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops

torch.manual_seed(0)
dev = torch.device("cuda:0")
_lib.init(0)
print("sms", _lib.lib().blm_num_sms(), torch.cuda.get_device_name(0), flush=True)
fails = 0

def report(name, got, ref, tol):
    global fails
    err = (got.double() - ref.double()).abs().max().item()
    scale = ref.double().abs().max().item()
    ok = err <= tol * max(scale, 1e-6) and torch.isfinite(got).all().item()
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.3e} ref_max={scale:.3e} tol_rel={tol:.1e}", flush=True)
    if not ok:
        fails += 1

# split
x = torch.randn(1000, 520, device=dev)
s = ops.split(x, "bf16x3")
report("split hi+lo", s.float(), x, 2e-5)

def gemm_case(M, N, K, prec, act=ops.ACT_NONE, bias=False, resid=False, colscale=False):
    a = torch.randn(M, K, device=dev) * 0.5
    b = torch.randn(N, K, device=dev) * 0.1
    A, B = ops.split(a, prec), ops.split(b, prec)
    bi = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    coef = torch.rand(4, N, device=dev) if act == ops.ACT_GPMIX else None
    out32 = torch.empty(M, N, device=dev)
    out = ops.empty_split(M, N, "bf16x3", dev)
    ops.gemm(A, B, prec=prec, bias=bi, act=act, coef=coef, col_scale=0.125 if colscale else 1.0,
             col_scale_cols=(N // 2) if colscale else 0, resid=r, out_f32=out32, out=out)
    torch.cuda.synchronize()
    if prec == "bf16":
        z = A.hi.double() @ B.hi.double().T
    else:
        z = a.double() @ b.double().T
    if bias: z = z + bi.double()
    if colscale: z[:, : N // 2] *= 0.125
    if act == ops.ACT_GELU: z = torch.nn.functional.gelu(z)
    if act == ops.ACT_GPMIX:
        c = coef.double()
        z = c[0] * torch.tanh(z) + c[1] * torch.sigmoid(z) + c[2] * torch.relu(z) + c[3] * torch.nn.functional.gelu(z)
    if resid: z = z + r.double()
    tol = 4e-5 if prec == "bf16x3" else 2e-5   # bf16 case compares against the bf16-rounded operands
    report(f"gemm M={M} N={N} K={K} {prec} act={act} bias={bias} resid={resid} cs={colscale}", out32, z, tol)
    report(f"   (hi+lo output)", out.float(), z, 6e-5)

t0 = time.time()
gemm_case(128, 128, 64, "bf16")
gemm_case(128, 256, 128, "bf16")
gemm_case(300, 520, 200, "bf16", bias=True)
gemm_case(1000, 1536, 512, "bf16", bias=True, colscale=True)
gemm_case(1000, 1536, 512, "bf16x3", bias=True, colscale=True)
gemm_case(20000, 4096, 512, "bf16", bias=True, act=ops.ACT_GELU)
gemm_case(20000, 512, 4096, "bf16x3", bias=True, resid=True)
gemm_case(5000, 4096, 512, "bf16x3", bias=True, act=ops.ACT_GPMIX)
gemm_case(77, 64, 64, "bf16x3", bias=True, act=ops.ACT_GELU)
print("gemm cases took", time.time() - t0, flush=True)

def nll_case(M, V, K, prec):
    h = torch.randn(M, K, device=dev)
    e = (torch.rand(V, K, device=dev) - 0.5) * 0.2
    b = (torch.rand(V, device=dev) - 0.5) * 0.2
    t = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    H, E = ops.split(h, prec), ops.split(e, prec)
    nll = ops.vocab_nll(H, E, b, t, prec=prec)
    torch.cuda.synchronize()
    if prec == "bf16":
        logits = H.hi.double() @ E.hi.double().T + b.double()
    else:
        logits = h.double() @ e.double().T + b.double()
    ref = torch.logsumexp(logits, -1) - logits.gather(1, t.long().view(-1, 1)).squeeze(1)
    report(f"vocab_nll M={M} V={V} K={K} {prec}", nll, ref, 6e-6)

nll_case(100, 1000, 64, "bf16")
nll_case(100, 1000, 64, "bf16x3")
nll_case(3000, 30000, 512, "bf16")
nll_case(3000, 30000, 512, "bf16x3")
nll_case(40000, 30000, 512, "bf16")
nll_case(257, 30000, 1024, "bf16x3")

# layernorm
x = torch.randn(3000, 512, device=dev) * 3 + 1
g, bb = torch.randn(512, device=dev), torch.randn(512, device=dev)
y, ys = ops.layernorm(x, g, bb, 1e-5, prec="bf16x3")
report("layernorm 512", y, torch.nn.functional.layer_norm(x.double(), (512,), g.double(), bb.double(), 1e-5), 1e-5)
report("layernorm split", ys.float(), y, 2e-5)
x = torch.randn(100, 64, device=dev)
g, bb = torch.randn(64, device=dev), torch.randn(64, device=dev)
y, _ = ops.layernorm(x, g, bb, 1e-5, prec="bf16")
report("layernorm 64", y, torch.nn.functional.layer_norm(x.double(), (64,), g.double(), bb.double(), 1e-5), 1e-5)

# embed
V, d = 1000, 512
emb = torch.randn(V, d, device=dev); pe = torch.randn(200, d, device=dev)
tok = torch.randint(0, V, (777,), device=dev, dtype=torch.int32); pos = torch.randint(0, 200, (777,), device=dev, dtype=torch.int32)
xf, xs = ops.embed(tok, pos, emb, pe, 22.627, prec="bf16x3")
report("embed", xf, emb[tok.long()] * 22.627 + pe[pos.long()], 1e-6)

# attention
def attn_case(lens, nhead, hd):
    d = nhead * hd
    offs = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    M = int(offs[-1])
    qkv = torch.randn(M, 3 * d, device=dev)
    o, os_ = ops.mha_causal(qkv, offs, nhead, max(lens), prec="bf16x3", want_f32=True)
    torch.cuda.synchronize()
    ref = torch.empty(M, d, dtype=torch.float64, device=dev)
    for i, T in enumerate(lens):
        r0 = int(offs[i])
        blk = qkv[r0:r0 + T].double()
        q, k, v = blk[:, :d].view(T, nhead, hd), blk[:, d:2 * d].view(T, nhead, hd), blk[:, 2 * d:].view(T, nhead, hd)
        sc = torch.einsum("ihc,jhc->hij", q, k)
        mask = torch.triu(torch.ones(T, T, device=dev, dtype=torch.bool), 1)
        sc = sc.masked_fill(mask, float("-inf"))
        p = torch.softmax(sc, -1)
        ref[r0:r0 + T] = torch.einsum("hij,jhc->ihc", p, v).reshape(T, d)
    report(f"mha lens={lens[:4]}.. nhead={nhead} hd={hd}", o, ref, 1e-5)

attn_case([1, 5, 17, 26, 33, 64], 8, 64)
attn_case([100, 100, 7], 8, 64)
attn_case([1, 2, 7, 8, 9, 15, 16, 17, 26, 31, 32, 5, 5, 5, 11, 13, 3], 8, 64)
attn_case(list(range(1, 27)) * 3, 8, 64)
attn_case([3, 9, 128], 4, 16)

# KL
mu = torch.randn(4096, 1024, device=dev) * 0.03; ls = torch.rand(1024, 1024, device=dev) * -3.4 - 3.4
out = torch.zeros(1, device=dev)
ops.kl_gauss(mu[2048:3072], ls, out)
ref = ((mu[2048:3072].double() ** 2 - 2 * ls.double() + torch.exp(2 * ls.double())).mean() / 2)
report("kl slice", out, ref.view(1), 1e-6)
ops.kl_gauss(mu[2048:3072], ls, out, minus_one=True, scale=0.5, accumulate=True)
ref2 = ref + 0.5 * ((mu[2048:3072].double() ** 2 - 2 * ls.double() + torch.exp(2 * ls.double()) - 1).mean() / 2)
report("kl accumulate minus_one", out, ref2.view(1), 1e-6)

# reparam
eps = torch.randn(1024, 1024, device=dev)
w, ws = ops.reparam(mu[2048:3072], ls, eps=eps, prec="bf16x3", want_f32=True)
report("reparam eps", w, mu[2048:3072] + torch.exp(ls) * eps, 1e-6)
z = ops.philox_normal(1234, 7, 4_000_000, dev)
print("philox mean/std/kurt", z.mean().item(), z.std().item(), ((z - z.mean()) ** 4).mean().item() / z.var().item() ** 2, flush=True)
w2, _ = ops.reparam(mu[2048:3072], ls, seed=1234, stream_id=7, prec="bf16", want_f32=True)
report("reparam philox == injected philox", w2, mu[2048:3072] + torch.exp(ls) * z[: 1024 * 1024].view(1024, 1024), 1e-6)

# quick GEMM timing
def bench(M, N, K, prec, iters=20):
    A, B = ops.split(torch.randn(M, K, device=dev), prec), ops.split(torch.randn(N, K, device=dev), prec)
    out = ops.empty_split(M, N, "bf16", dev)
    for _ in range(3): ops.gemm(A, B, prec=prec, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(A, B, prec=prec, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"gemm {M}x{N}x{K} {prec}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)

bench(65536, 4096, 512, "bf16"); bench(65536, 512, 4096, "bf16"); bench(65536, 1536, 512, "bf16")
bench(65536, 4096, 512, "bf16x3"); bench(8192, 8192, 8192, "bf16")

def bench_nll(M, V, K, prec, iters=10):
    H, E = ops.split(torch.randn(M, K, device=dev), prec), ops.split(torch.randn(V, K, device=dev) * 0.05, prec)
    t = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    for _ in range(3): ops.vocab_nll(H, E, None, t, prec=prec)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.vocab_nll(H, E, None, t, prec=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"vocab_nll {M}x{V}x{K} {prec}: {ms*1e3:.1f} us  {2*M*V*K/ms/1e9:.1f} TFLOP/s", flush=True)

bench_nll(65536, 30000, 512, "bf16"); bench_nll(65536, 30000, 1024, "bf16")
print("FAILS", fails, flush=True)
sys.exit(1 if fails else 0)
