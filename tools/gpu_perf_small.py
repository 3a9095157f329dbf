"""Warm per-launch time of the small kernels of the fine-tune step (M = 3200 tokens): 50 launches back to back, eager
and replayed from a CUDA graph (what the step does).  ncu's per-launch numbers are cold-cache and serialised."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
from bayeslms_b200.ops import Split

_lib.init(0)
dev = torch.device("cuda:0")
M, d, F, V = 3200, 512, 4096, 30000
N_REP = 50


def timeit(fn, name, flops=0.0, bytes_=0.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N_REP):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / N_REP * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(N_REP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / (5 * N_REP) * 1e3
    extra = ""
    if flops:
        extra += f"  {flops / graph / 1e6:7.1f} TF/s"
    if bytes_:
        extra += f"  {bytes_ / graph / 1e3:7.1f} GB/s"
    print(f"{name:44s} eager {eager:7.1f} us   graph {graph:7.1f} us{extra}", flush=True)


def sp(r, c):
    return ops.split(torch.randn(r, c, device=dev), "bf16")


x512, x4096, x1536 = sp(M, d), sp(M, F), sp(M, 3 * d)
w = {(n, k): sp(n, k) for n, k in [(512, 512), (1536, 512), (4096, 512), (512, 4096), (512, 1536)]}
res = torch.randn(M, d, device=dev)
for (n, k), wt in w.items():
    a = {512: x512, 4096: x4096, 1536: x1536}[k]
    out = torch.empty(M, n, device=dev)
    timeit(lambda: ops.gemm(a, wt, prec="bf16", out_f32=out), f"gemm f32out M{M} N{n} K{k}", 2.0 * M * n * k)
    if n == 512:
        timeit(lambda: ops.gemm(a, wt, prec="bf16", resid=res, out_f32=out), f"gemm f32out+resid M{M} N{n} K{k}", 2.0 * M * n * k)
    outs = ops.empty_split(M, n, "bf16", dev)
    timeit(lambda: ops.gemm(a, wt, prec="bf16", out=outs), f"gemm bf16out M{M} N{n} K{k}", 2.0 * M * n * k)
# wgrad shapes: dW[N, K] = dY^T X, reduction over M tokens, MN-major operands
for n, k in [(512, 512), (1536, 512), (4096, 512), (512, 4096)]:
    dy, xx = sp(M, n), sp(M, k)
    out = torch.empty(n, k, device=dev)
    timeit(lambda: ops.gemm(dy, xx, prec="bf16", out_f32=out, a_mn=True, b_mn=True), f"wgrad dW[{n},{k}] over M{M}", 2.0 * M * n * k)
# dgrad with b_mn
for n_in, n_out in [(512, 4096), (4096, 512), (1536, 512)]:
    dy, wt = sp(M, n_in), sp(n_in, n_out)
    out = torch.empty(M, n_out, device=dev)
    timeit(lambda: ops.gemm(dy, wt, prec="bf16", out_f32=out, b_mn=True), f"dgrad dX[{M},{n_out}] K{n_in}", 2.0 * M * n_in * n_out)
y = torch.randn(M, d, device=dev)
g_, b_ = torch.randn(d, device=dev), torch.randn(d, device=dev)
timeit(lambda: ops.layernorm(y, g_, b_, 1e-5, prec="bf16"), "layernorm [3200,512]", bytes_=M * d * 10.0)
dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
timeit(lambda: ops.layernorm_bwd(y, y, g_, 1e-5, dg, db), "layernorm_bwd [3200,512]", bytes_=M * d * 12.0)
o512, o4096, o1536 = torch.zeros(d, device=dev), torch.zeros(F, device=dev), torch.zeros(3 * d, device=dev)
z4096, z1536 = torch.randn(M, F, device=dev), torch.randn(M, 3 * d, device=dev)
timeit(lambda: ops.colsum(y, o512), "colsum [3200,512]", bytes_=M * d * 4.0)
timeit(lambda: ops.colsum(z4096, o4096), "colsum [3200,4096]", bytes_=M * F * 4.0)
timeit(lambda: ops.colsum(z1536, o1536), "colsum [3200,1536]", bytes_=M * 3 * d * 4.0)
timeit(lambda: ops.split(y, "bf16"), "split [3200,512]", bytes_=M * d * 6.0)
timeit(lambda: ops.split(z4096, "bf16"), "split [3200,4096]", bytes_=M * F * 6.0)
n_par = 42_000_000
p_, gg, vv = (torch.randn(n_par, device=dev) for _ in range(3))
hi = torch.empty(n_par, dtype=torch.bfloat16, device=dev)
nsq = torch.ones(1, device=dev)
timeit(lambda: ops.sgd_momentum(p_, gg, vv, 0.01, 0.9, nsq, 0.25, 1.0, out_hi=hi), "sgd_momentum_split 42 M params", bytes_=n_par * 22.0)
timeit(lambda: ops.reduce_sum(gg, nsq, squares=True), "reduce (grad norm) 42 M", bytes_=n_par * 4.0)
timeit(lambda: gg.zero_(), "flat_g.zero_ 42 M", bytes_=n_par * 4.0)
qkv32 = torch.randn(M, 3 * d, device=dev)
offs = torch.arange(0, M + 1, 100, dtype=torch.int32, device=dev)
dat = torch.randn(M, d, device=dev)
timeit(lambda: ops.mha_causal_bwd(qkv32, dat, offs, 8, 100, 0.125, prec="bf16"), "mha_causal_bwd 32 x 100")
timeit(lambda: ops.mha_causal_bf16(x1536, offs, 8, 100, prec="bf16"), "mha_causal fwd 32 x 100")
