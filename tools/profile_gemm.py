"""One K=512 bf16-output GEMM (FFN1 shape) for ncu.  usage: profile_gemm.py [act]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
_lib.init(0)
dev = torch.device("cuda:0")
act = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M, d, F = 65536, 512, 4096
torch.manual_seed(0)
x = ops.split(torch.randn(M, d, device=dev), "bf16")
w1 = ops.split(torch.randn(F, d, device=dev) * 0.05, "bf16")
bF = torch.randn(F, device=dev)
hout = ops.empty_split(M, F, "bf16", dev)
for _ in range(3):
    ops.gemm(x, w1, bias=bF, act=act, out=hout)
torch.cuda.synchronize()
print("ok", float(hout.hi[0, 0]))
