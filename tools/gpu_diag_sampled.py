"""Tile-fused sampled GEMM vs the materialised path (blm_reparam + blm_gemm) and a float64 reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
torch.manual_seed(0); dev = torch.device("cuda:0"); _lib.init(0)
fails = 0
def case(M, N, K, mode, act=ops.ACT_NONE, bias=False, resid=False):
    global fails
    a = torch.randn(M, K, device=dev) * 0.5
    mu = torch.randn(N, K, device=dev) * 0.05
    ls = torch.rand(N, K, device=dev) * -3.0 - 2.0
    A = ops.split(a, "bf16")
    eps = torch.randn(N, K, device=dev) if mode == "ptr" else None
    seed = 1234 if mode == "philox" else None
    bi = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    coef = torch.rand(4, N, device=dev) if act == ops.ACT_GPMIX else None
    ldc = (N + 7) // 8 * 8
    out = torch.empty(M, ldc, device=dev)
    mu_b = ops.split(mu, "bf16").hi
    sg_b = ops.sigma_bf16(ls)
    ops.gemm_sampled(A, mu_b, sg_b if mode != "mean" else None, eps=eps, seed=seed, stream_id=5, bias=bi, act=act,
                     coef=coef, resid=r, out_f32=out)
    # reference W~ from the same bf16 (mu, sigma) and the same noise
    if mode == "mean":
        wt = mu_b
    else:
        e = eps if mode == "ptr" else ops.philox_normal(seed, 5, N * K, dev).view(N, K)
        wt = torch.addcmul(mu_b.float(), sg_b.float(), e).to(torch.bfloat16)
    out2 = torch.empty(M, ldc, device=dev)
    ops.gemm(A, ops.Split(wt.contiguous()), prec="bf16", bias=bi, act=act, coef=coef, resid=r, out_f32=out2)
    torch.cuda.synchronize()
    z = A.hi.double() @ wt.double().T
    if bias: z = z + bi.double()
    if act == ops.ACT_GELU: z = torch.nn.functional.gelu(z)
    if act == ops.ACT_GPMIX:
        c = coef.double(); z = c[0]*torch.tanh(z) + c[1]*torch.sigmoid(z) + c[2]*torch.relu(z) + c[3]*torch.nn.functional.gelu(z)
    if resid: z = z + r.double()
    e_ref = (out[:, :N].double() - z).abs().max().item(); e_mat = (out[:, :N] - out2[:, :N]).abs().max().item()
    ok = e_ref < 1e-3 * max(1.0, z.abs().max().item()) and e_mat < 1e-3 * max(1.0, z.abs().max().item())
    fails += not ok
    print(("OK " if ok else "BAD"), f"sampled M={M} N={N} K={K} {mode} act={act} bias={bias} resid={resid}: vs f64 {e_ref:.2e}, vs materialised {e_mat:.2e}", flush=True)

case(512, 128, 64, "mean"); case(512, 128, 64, "ptr"); case(512, 128, 64, "philox")
case(77, 60, 72, "ptr", bias=True); case(1000, 520, 200, "philox", bias=True, resid=True)
case(20000, 512, 4096, "ptr", resid=True); case(20000, 512, 4096, "philox", resid=True)
case(5000, 4096, 512, "philox", act=ops.ACT_GPMIX, bias=True); case(3000, 512, 512, "philox")
def bench(M, N, K, mode, iters=10):
    A = ops.split(torch.randn(M, K, device=dev), "bf16"); mu = torch.randn(N, K, device=dev) * 0.05
    ls = torch.rand(N, K, device=dev) * -3.0 - 2.0; out = torch.empty(M, N, device=dev)
    mu_b = ops.split(mu, "bf16").hi; sg_b = ops.sigma_bf16(ls)
    f = lambda: ops.gemm_sampled(A, mu_b, sg_b, seed=(7 if mode == "philox" else None), stream_id=1, out_f32=out)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / iters
    print(f"gemm_sampled {M}x{N}x{K} {mode}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    W = ops.reparam(mu, ls, seed=7, stream_id=1, prec="bf16")[1]
    g = lambda: (ops.reparam(mu, ls, seed=7, stream_id=1, prec="bf16"), ops.gemm(A, W, prec="bf16", out_f32=out))
    for _ in range(3): g()
    e0.record()
    for _ in range(iters): g()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / iters
    print(f"   materialise + gemm        : {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
for M in (65536, 8192, 512, 128):
    bench(M, 512, 4096, "philox"); bench(M, 512, 4096, "mean")
bench(65536, 4096, 512, "philox")
print("FAILS", fails); sys.exit(1 if fails else 0)
