import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
_lib.init(0)
dev = torch.device("cuda:0")
M, d, F = 65536, 512, 4096
torch.manual_seed(0)
h = ops.split(torch.randn(M, F, device=dev), "bf16")
x32 = torch.randn(M, d, device=dev)
mu = torch.randn(d, F, device=dev) * 0.03; ls = torch.rand(d, F, device=dev) * -3 - 3
sig = ops.sigma_bf16(ls); mub = ops.split(mu, "bf16")
y = torch.empty(M, d, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("once", t(lambda: ops.gemm_sampled(h, mub.hi, sig, seed=1, stream_id=5, resid=x32, out_f32=y, how="once")))
ws = ops._ws_cache
for k, v in ws.items():
    if k[0] == "gemm_sampled":
        print("counters", v[:8].view(torch.int32).tolist(), "bytes", v.numel())
        wt = v[256:256 + d * F * 2].view(torch.bfloat16).view(d, F)
        eps = ops.philox_normal(1, 5, d * F, dev).view(d, F)
        ref = torch.addcmul(mub.hi.float(), sig.float(), eps).to(torch.bfloat16)
        print("scratch == expected W~:", torch.equal(wt, ref))
print("tile", t(lambda: ops.gemm_sampled(h, mub.hi, sig, seed=1, stream_id=5, resid=x32, out_f32=y, how="tile")))
_, w = ops.reparam(mu, ls, seed=1, stream_id=5, prec="bf16")
print("plain gemm", t(lambda: ops.gemm(h, w, resid=x32, out_f32=y)))
