"""Where does the o_net + norm1 kernel spend its time?  Drop one output pass at a time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops as O
_lib.init(0)
dev = torch.device("cuda:0")
M, d = 52833, 512
torch.manual_seed(0)
x = O.split(torch.randn(M, d, device=dev), "bf16")
h = O.split(torch.randn(M, 4096, device=dev), "bf16")
x32 = torch.randn(M, d, device=dev)
wo = O.split(torch.randn(d, d, device=dev) * 0.05, "bf16")
w2 = O.split(torch.randn(d, 4096, device=dev) * 0.02, "bf16")
bd, g, bt = torch.randn(d, device=dev), torch.rand(d, device=dev) + 0.5, torch.randn(d, device=dev)
y = torch.empty(M, d, device=dev); yh = torch.empty(M, d, device=dev, dtype=torch.bfloat16)

def run(a, w, f32, hi, bias=True):
    dd = O.GemmLnDesc()
    dd.M, dd.N, dd.K = M, d, a.hi.shape[1]
    dd.A, dd.lda, dd.B, dd.ldb = a.hi.data_ptr(), a.hi.stride(0), w.hi.data_ptr(), w.hi.stride(0)
    dd.bias, dd.resid, dd.ldr = O._ptr(bd if bias else None), O._ptr(x32), d
    dd.gamma, dd.beta, dd.eps = O._ptr(g), O._ptr(bt), 1e-5
    dd.out_f32, dd.out_hi, dd.ldc = O._ptr(y if f32 else None), O._ptr(yh if hi else None), d
    O.check(O.lib().blm_gemm_ln(O.C.byref(dd), O._stream()), "blm_gemm_ln")

def t(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for name, a, w in (("o_net K=512", x, wo), ("ffn2 K=4096", h, w2)):
    print(name, "both %.1f  f32 only %.1f  bf16 only %.1f  both/no bias %.1f us" % (
        t(lambda: run(a, w, True, True)), t(lambda: run(a, w, True, False)), t(lambda: run(a, w, False, True)),
        t(lambda: run(a, w, True, True, bias=False))))
