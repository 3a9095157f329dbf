import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
_lib.init(0)
DEV = "cuda:0"
torch.manual_seed(0)
M, N, K = 1000, 512, 4096
a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.float16)
b = (torch.randn(N, K, device=DEV) * 0.1).to(torch.float16)
out = torch.empty(M, N, device=DEV)
ops.gemm(ops.Split(a), ops.Split(b), prec="bf16", out_f32=out, a_f16=True)
torch.cuda.synchronize()
ref = a.double() @ b.double().T
print("f16 x f16 MMA: max err", (out.double() - ref).abs().max().item(), "ref max", ref.abs().max().item())
