"""Generate-once sampled GEMM (FFN shape), tile-stationary sampled GEMM and the KL reduction, for ncu."""
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
mu = torch.randn(d, F, device=dev) * 0.03
ls = torch.rand(d, F, device=dev) * -3 - 3
sig, mub = ops.sigma_bf16(ls), ops.split(mu, "bf16")
y = torch.empty(M, d, device=dev)
out = torch.zeros(1, device=dev)
mu2 = torch.randn(4096, 8192, device=dev) * 0.03
ls2 = torch.rand(4096, 8192, device=dev) * -3 - 3
for _ in range(3):
    ops.gemm_sampled(h, None, None, mu_f32=mu, lgstd_f32=ls, seed=1, stream_id=5, resid=x32, out_f32=y, how="once")
    ops.gemm_sampled(h, mub.hi, sig, seed=1, stream_id=5, resid=x32, out_f32=y, how="tile")
    ops.kl_gauss(mu, ls, out)
    ops.kl_gauss(mu2, ls2, out)
torch.cuda.synchronize()
print("ok", float(y[0, 0]), float(out))
