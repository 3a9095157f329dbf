import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
_lib.init(0)
DEV = "cuda:0"
torch.manual_seed(0)
for (M, N, K) in ((65613, 512, 512), (20000, 512, 4096), (65613, 512, 4096), (9472, 512, 4096), (9472 * 2, 512, 4096), (9472 * 3, 512, 4096)):
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    bi = torch.randn(N, device=DEV)
    r = torch.randn(M, N, device=DEV)
    g = torch.rand(N, device=DEV) + 0.5
    be = torch.randn(N, device=DEV)
    for rep in range(2):
        y32, ys = ops.gemm_ln(A, B, bias=bi, resid=r, gamma=g, beta=be, eps=1e-5)
        z = A.hi.double() @ B.hi.double().T + r.double() + bi.double()
        ref = torch.nn.functional.layer_norm(z, (N,), g.double(), be.double(), 1e-5)
        err = (y32.double() - ref).abs()
        bad = (err > 1e-3)
        rows = bad.any(1).nonzero().flatten()
        print(M, N, K, "rep", rep, "max err", err.max().item(), "bad rows", rows.numel(), "bad elems", int(bad.sum()))
        if rows.numel():
            rr = rows.tolist()
            tiles = sorted(set(x // 128 for x in rr))
            print("   tiles", tiles[:20], "n tiles", len(tiles), " rows in tile", sorted(set(x % 128 for x in rr))[:40])
            cols = bad.any(0).nonzero().flatten().tolist()
            print("   cols", cols[:8], "...", cols[-4:], len(cols))
            r0 = rr[0]
            print("   row", r0, "got", y32[r0, :4].tolist(), "ref", ref[r0, :4].tolist(), "bf16 ok:", bool(torch.equal(ys.hi[r0], y32[r0].to(torch.bfloat16))))
