"""A/B: projection + residual + LayerNorm as one kernel (blm_gemm_ln) vs blm_gemm + blm_layernorm.
CUDA events, 20 iterations after 3 warm-ups.  usage: python tools/gpu_perf_gemm_ln.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops

_lib.init(0)
dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, F = 512, 4096
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


x = ops.split(torch.randn(M, d, device=dev), "bf16")
h = ops.split(torch.randn(M, F, device=dev), "bf16")
x32 = torch.randn(M, d, device=dev)
wo = ops.split(torch.randn(d, d, device=dev) * 0.05, "bf16")
w2 = ops.split(torch.randn(d, F, device=dev) * 0.02, "bf16")
bd = torch.randn(d, device=dev)
g = torch.rand(d, device=dev) + 0.5
bt = torch.randn(d, device=dev)
y = torch.empty(M, d, device=dev)


def unfused(a, w):
    ops.gemm(a, w, bias=bd, resid=x32, out_f32=y)
    ops.layernorm(y, g, bt, 1e-5)


print(f"M = {M}")
for name, a, w, K in (("o_net [M,512,512]", x, wo, d), ("ffn2 [M,512,4096]", h, w2, F)):
    t0 = timeit(lambda: unfused(a, w))
    t1 = timeit(lambda: ops.gemm_ln(a, w, bias=bd, resid=x32, gamma=g, beta=bt, eps=1e-5))
    fl = 2.0 * M * d * K
    print(f"{name:22s} gemm + layernorm {t0*1e3:8.1f} us   gemm_ln {t1*1e3:8.1f} us  ({fl/t1/1e9:7.1f} TFLOP/s)", flush=True)
