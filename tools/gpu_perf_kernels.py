"""Times the kernels of the rescoring step in their real configurations (CUDA events, 20 iterations
after 3 warm-ups, inputs >> L2).  usage: python tools/gpu_perf_kernels.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
from bayeslms_b200.ops import ACT_GELU, ACT_NONE, ACT_GPMIX

_lib.init(0)
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, F, V = 512, 4096, 30000
torch.manual_seed(0)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def line(name, ms, flops=None, bytes_=None):
    s = f"{name:44s} {ms*1e3:9.1f} us"
    if flops:
        s += f"  {flops/ms/1e9:8.1f} TFLOP/s"
    if bytes_:
        s += f"  {bytes_/ms/1e6:8.1f} GB/s"
    print(s, flush=True)


print("BLM_EPI_WARPS =", os.environ.get("BLM_EPI_WARPS", "8 (default)"), " M =", M, flush=True)
x = ops.split(torch.randn(M, d, device=dev), "bf16")
x32 = torch.randn(M, d, device=dev)
h = ops.split(torch.randn(M, F, device=dev), "bf16")
wqkv = ops.split(torch.randn(3 * d, d, device=dev) * 0.05, "bf16")
wo = ops.split(torch.randn(d, d, device=dev) * 0.05, "bf16")
w1 = ops.split(torch.randn(F, d, device=dev) * 0.05, "bf16")
w2 = ops.split(torch.randn(d, F, device=dev) * 0.02, "bf16")
b3, b1, bd, bF = (torch.randn(n, device=dev) for n in (3 * d, d, d, F))
coef = torch.rand(4, F, device=dev)
qkv = torch.empty(M, 3 * d, device=dev)
y = torch.empty(M, d, device=dev)
hout = ops.empty_split(M, F, "bf16", dev)

line("gemm qkv   [M,1536,512] bias+qscale f32", timeit(lambda: ops.gemm(x, wqkv, bias=b3, col_scale=0.125, col_scale_cols=d, out_f32=qkv)), 2.0 * M * 3 * d * d)
qkvs = ops.empty_split(M, 3 * d, "bf16", dev)
line("gemm qkv   [M,1536,512] bias+qscale bf16", timeit(lambda: ops.gemm(x, wqkv, bias=b3, col_scale=0.125, col_scale_cols=d, out=qkvs)), 2.0 * M * 3 * d * d)
line("gemm o_net [M,512,512] bias+resid f32", timeit(lambda: ops.gemm(x, wo, bias=bd, resid=x32, out_f32=y)), 2.0 * M * d * d)
line("gemm ffn1  [M,4096,512] bias+GELU bf16", timeit(lambda: ops.gemm(x, w1, bias=bF, act=ACT_GELU, out=hout)), 2.0 * M * F * d)
from bayeslms_b200.ops import ACT_GELU_FAST
line("gemm ffn1  [M,4096,512] bias+GELU(f16x2) bf16", timeit(lambda: ops.gemm(x, w1, bias=bF, act=ACT_GELU_FAST, out=hout)), 2.0 * M * F * d)
line("gemm ffn1  [M,4096,512] bias+GPMIX bf16", timeit(lambda: ops.gemm(x, w1, bias=bF, act=ACT_GPMIX, coef=coef, out=hout)), 2.0 * M * F * d)
from bayeslms_b200.ops import ACT_GPMIX_FAST
line("gemm ffn1  [M,4096,512] bias+GPMIX(f16x2) bf16", timeit(lambda: ops.gemm(x, w1, bias=bF, act=ACT_GPMIX_FAST, coef=coef, out=hout)), 2.0 * M * F * d)
line("gemm ffn1  [M,4096,512] no epilogue bf16", timeit(lambda: ops.gemm(x, w1, out=hout)), 2.0 * M * F * d)
line("gemm ffn2  [M,512,4096] resid f32", timeit(lambda: ops.gemm(h, w2, resid=x32, out_f32=y)), 2.0 * M * d * F)
A8, B8 = ops.split(torch.randn(8192, 8192, device=dev), "bf16"), ops.split(torch.randn(8192, 8192, device=dev), "bf16")
o8 = ops.empty_split(8192, 8192, "bf16", dev)
line("gemm 8192^3 bf16", timeit(lambda: ops.gemm(A8, B8, out=o8)), 2.0 * 8192 ** 3)

E = ops.split(torch.randn(V, d, device=dev) * 0.05, "bf16")
bV = torch.randn(V, device=dev) * 0.1
t = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
line("vocab_nll [M,30000,512] bias", timeit(lambda: ops.vocab_nll(x, E, bV, t), 10), 2.0 * M * V * d)
line("vocab_nll [M,30000,512] no bias", timeit(lambda: ops.vocab_nll(x, E, None, t), 10), 2.0 * M * V * d)
x3 = ops.split(torch.randn(M, d, device=dev), "bf16x3")
E3 = ops.split(torch.randn(V, d, device=dev) * 0.05, "bf16x3")
line("vocab_nll [M,30000,512] bias bf16x3", timeit(lambda: ops.vocab_nll(x3, E3, bV, t, prec="bf16x3"), 5), 2.0 * M * V * d)

# attention over synthetic hypothesis lengths (U{6..26}) and fine-tune length 100
import numpy as np
rng = np.random.default_rng(0)
lens = []
while sum(lens) < M - 26:
    lens.append(int(rng.integers(6, 27)))
offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
Ma = int(offs[-1])
qa = torch.randn(Ma, 3 * d, device=dev)
line(f"mha_causal rescoring lens 6..26 ({Ma} tok)", timeit(lambda: ops.mha_causal(qa, offs, 8, 26)), bytes_=Ma * (3 * d * 4 + d * 2))
qs = ops.split(qa, "bf16"); qs3 = ops.split(qa, "bf16x3")
line(f"mha_causal_bf16 (mma) bf16   ({Ma} tok)", timeit(lambda: ops.mha_causal_bf16(qs, offs, 8, 26)), bytes_=Ma * (3 * d * 2 + d * 2))
line(f"mha_causal_bf16 (mma) bf16x3 ({Ma} tok)", timeit(lambda: ops.mha_causal_bf16(qs3, offs, 8, 26, prec="bf16x3")), bytes_=Ma * (3 * d * 4 + d * 4))
offs100 = torch.arange(0, 32 * 100 + 1, 100, dtype=torch.int32, device=dev)
q100 = ops.split(torch.randn(3200, 3 * d, device=dev), "bf16x3")
line("mha_causal_bf16 (mma) bf16x3 32 x T=100", timeit(lambda: ops.mha_causal_bf16(q100, offs100, 8, 100, prec="bf16x3")), bytes_=3200 * (3 * d * 4 + d * 4))
g = torch.ones(d, device=dev); bt = torch.zeros(d, device=dev)
line("layernorm [M,512] f32 -> f32 + bf16", timeit(lambda: ops.layernorm(x32, g, bt, 1e-5)), bytes_=M * d * (4 + 4 + 2))
tok = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
pos = torch.randint(0, 26, (M,), device=dev, dtype=torch.int32)
emb = torch.randn(V, d, device=dev); pe = torch.randn(5000, d, device=dev)
line("embed [M,512]", timeit(lambda: ops.embed(tok, pos, emb, pe, 22.6)), bytes_=M * d * (4 + 4 + 2))
mu = torch.randn(d, F, device=dev) * 0.03; ls = torch.rand(d, F, device=dev) * -3 - 3
out = torch.zeros(1, device=dev)
line("kl_gauss [512,4096]", timeit(lambda: ops.kl_gauss(mu, ls, out)), bytes_=2 * d * F * 4)
mu2 = torch.randn(4096, 8192, device=dev) * 0.03; ls2 = torch.rand(4096, 8192, device=dev) * -3 - 3
line("kl_gauss [4096,8192] (268 MB)", timeit(lambda: ops.kl_gauss(mu2, ls2, out)), bytes_=2 * 4096 * 8192 * 4)
# sampled FFN2: fused vs materialise + plain GEMM
sig = ops.sigma_bf16(ls)
mub = ops.split(mu, "bf16")
line("gemm_sampled ffn2 [M,512,4096] philox generate-once", timeit(lambda: ops.gemm_sampled(h, mub.hi, sig, seed=1, stream_id=5, resid=x32, out_f32=y, how="once")), 2.0 * M * d * F)
line("gemm_sampled ffn2 [M,512,4096] philox tile-stationary", timeit(lambda: ops.gemm_sampled(h, mub.hi, sig, seed=1, stream_id=5, resid=x32, out_f32=y, how="tile")), 2.0 * M * d * F)
line("gemm_sampled ffn2 [M,512,4096] mean (tile kernel)", timeit(lambda: ops.gemm_sampled(h, mub.hi, None, resid=x32, out_f32=y)), 2.0 * M * d * F)
sigT = ops.sigma_bf16(ls.t().contiguous()); muT = ops.split(mu.t().contiguous(), "bf16")
line("gemm_sampled gpnn [M,4096,512] philox generate-once +GELU bf16", timeit(lambda: ops.gemm_sampled(x, muT.hi, sigT, seed=1, stream_id=6, bias=bF, act=ACT_GELU, out=hout, how="once")), 2.0 * M * d * F)
def mat():
    _, w = ops.reparam(mu, ls, seed=1, stream_id=5, prec="bf16")
    ops.gemm(h, w, resid=x32, out_f32=y)
line("reparam + gemm ffn2 [M,512,4096] philox", timeit(mat), 2.0 * M * d * F)

# ---- LSTM recurrence (one layer, all T steps in one cooperative launch)
H = 1024
for (T_, B_) in ((20, 2048), (26, 1024), (20, 256)):
    gx = torch.randn(T_ * B_, 4 * H, device=dev) * 0.1
    whh = ops.split(torch.randn(4 * H, H, device=dev) * 0.03, "bf16")
    h0 = torch.zeros(B_, H, device=dev); c0 = torch.zeros(B_, H, device=dev)
    lens = torch.full((B_,), T_, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.lstm_layer(gx, whh, h0, c0, lens, T_, B_, H, prec="bf16", want_f32=False, want_split=True), 10)
    # algorithmic bytes per step: W_hh re-read 4H*H*2 + gates_x B*4H*4 + h,c write 2*B*H*4
    byt = T_ * (4 * H * H * 2 + B_ * 4 * H * 4 + 2 * B_ * H * 4)
    line(f"lstm_layer T={T_} B={B_} H=1024 bf16", ms, 2.0 * T_ * B_ * 4 * H * H, byt)
    print(f"    per step {ms*1e3/T_:.1f} us", flush=True)
whh3 = ops.split(torch.randn(4 * H, H, device=dev) * 0.03, "bf16x3")
ms = timeit(lambda: ops.lstm_layer(gx, whh3, h0, c0, lens, T_, B_, H, prec="bf16x3", want_f32=False, want_split=True), 10)
line(f"lstm_layer T={T_} B={B_} H=1024 bf16x3", ms, 2.0 * T_ * B_ * 4 * H * H)
