"""Kernel-level GPU parity: every C-ABI kernel against a float64 restatement of the same
operation (torch ops on the device, test-only), at the shapes of the BASELINE configs and at
ragged / edge shapes.  Tolerances: relative to the largest reference magnitude."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from bayeslms_b200 import _lib, ops as _ops
    _lib.init(0)
    torch.manual_seed(0)
    return _ops


def _close(got, ref, tol):
    assert torch.isfinite(got).all()
    err = (got.double() - ref.double()).abs().max().item()
    scale = max(ref.double().abs().max().item(), 1e-6)
    assert err <= tol * scale, (err, scale, tol)


def _gp_mix(z, c):
    return c[0] * torch.tanh(z) + c[1] * torch.sigmoid(z) + c[2] * torch.relu(z) + c[3] * torch.nn.functional.gelu(z)


@pytest.mark.parametrize("M,N,K,prec,act,bias,resid,cs", [
    (128, 128, 64, "bf16", 0, False, False, False),
    (300, 520, 200, "bf16", 0, True, False, False),         # ragged M, N, K
    (1000, 1536, 512, "bf16x3", 0, True, False, True),      # QKV projection, q scaled
    (20000, 4096, 512, "bf16", 1, True, False, False),      # FFN1 + GELU
    (20000, 512, 4096, "bf16x3", 0, True, True, False),     # FFN2 + residual
    (5000, 4096, 512, "bf16x3", 2, True, False, False),     # GP mixture epilogue
    (77, 64, 64, "bf16x3", 1, True, False, False),
    (1, 8, 8, "bf16x3", 0, True, True, False),              # a single row
])
def test_gemm_epilogues(ops, M, N, K, prec, act, bias, resid, cs):
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    A, B = ops.split(a, prec), ops.split(b, prec)
    bi = torch.randn(N, device=DEV) if bias else None
    r = torch.randn(M, N, device=DEV) if resid else None
    coef = torch.rand(4, N, device=DEV) if act == ops.ACT_GPMIX else None
    out32 = torch.empty(M, N, device=DEV)
    out = ops.empty_split(M, N, "bf16x3", DEV)
    ops.gemm(A, B, prec=prec, bias=bi, act=act, coef=coef, col_scale=0.125 if cs else 1.0,
             col_scale_cols=(N // 2) if cs else 0, resid=r, out_f32=out32, out=out)
    z = (A.hi.double() @ B.hi.double().T) if prec == "bf16" else (a.double() @ b.double().T)
    if bias:
        z = z + bi.double()
    if cs:
        z[:, : N // 2] *= 0.125
    if act == ops.ACT_GELU:
        z = torch.nn.functional.gelu(z)
    if act == ops.ACT_GPMIX:
        z = _gp_mix(z, coef.double())
    if resid:
        z = z + r.double()
    _close(out32, z, 4e-5 if prec == "bf16x3" else 2e-5)
    _close(out.float(), z, 6e-5)


@pytest.mark.parametrize("M,V,K,prec", [(100, 1000, 64, "bf16"), (100, 1000, 64, "bf16x3"), (3000, 30000, 512, "bf16"),
                                        (3000, 30000, 512, "bf16x3"), (40000, 30000, 512, "bf16"),
                                        (257, 30000, 1024, "bf16x3"), (5, 33, 8, "bf16x3")])
def test_vocab_nll(ops, M, V, K, prec):
    h = torch.randn(M, K, device=DEV)
    e = (torch.rand(V, K, device=DEV) - 0.5) * 0.2
    b = (torch.rand(V, device=DEV) - 0.5) * 0.2
    t = torch.randint(0, V, (M,), device=DEV, dtype=torch.int32)
    H, E = ops.split(h, prec), ops.split(e, prec)
    nll = ops.vocab_nll(H, E, b, t, prec=prec)
    nll_nb = ops.vocab_nll(H, E, None, t, prec=prec)
    logits = (H.hi.double() @ E.hi.double().T) if prec == "bf16" else (h.double() @ e.double().T)
    for got, lg in ((nll, logits + b.double()), (nll_nb, logits)):
        ref = torch.logsumexp(lg, -1) - lg.gather(1, t.long().view(-1, 1)).squeeze(1)
        _close(got, ref, 6e-6)


def _attention_ref(qkv, offs, nhead):
    M, d3 = qkv.shape
    d = d3 // 3
    hd = d // nhead
    ref = torch.zeros(M, d, dtype=torch.float64, device=qkv.device)
    offs = offs.tolist()
    for i in range(len(offs) - 1):
        r0, T = offs[i], offs[i + 1] - offs[i]
        if T == 0:
            continue
        blk = qkv[r0:r0 + T].double()
        q, k, v = (blk[:, j * d:(j + 1) * d].view(T, nhead, hd) for j in range(3))
        sc = torch.einsum("ihc,jhc->hij", q, k)
        sc = sc.masked_fill(torch.triu(torch.ones(T, T, device=qkv.device, dtype=torch.bool), 1), float("-inf"))
        ref[r0:r0 + T] = torch.einsum("hij,jhc->ihc", torch.softmax(sc, -1), v).reshape(T, d)
    return ref


ATTN_LENS = [[1, 5, 17, 26, 32, 2], [100, 100, 7], [1, 2, 7, 8, 9, 15, 16, 17, 26, 31, 32, 5, 5, 5, 11, 13, 3],
             list(range(1, 27)) * 3, [33, 64, 65, 96, 97, 128, 1],
             [129, 5, 200, 128, 257, 300, 1, 640]]      # > 128 tokens: key / value blocks walked flash-attention style


@pytest.mark.parametrize("lens", ATTN_LENS)
@pytest.mark.parametrize("prec", ["bf16x3", "bf16"])
def test_attention_tensor_core(ops, lens, prec):
    """blm_mha_causal_bf16 (mma.sync on bf16 hi[, lo]) at head_dim 64: short and long variants."""
    nhead, hd = 8, 64
    d = nhead * hd
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=DEV)
    M = int(offs[-1])
    qkv = torch.randn(M, 3 * d, device=DEV)
    qkv[:, :d] *= hd ** -0.5
    qs = ops.split(qkv, prec)
    o32, osplit = ops.mha_causal_bf16(qs, offs, nhead, max(lens), prec=prec, want_f32=True)
    src = qkv if prec == "bf16x3" else qs.hi.float()
    ref = _attention_ref(src, offs, nhead)
    # bf16 mode additionally rounds the softmax weights to bf16 (2^-9 relative)
    _close(o32, ref, 3e-5 if prec == "bf16x3" else 6e-3)
    _close(osplit.float(), ref, 5e-5 if prec == "bf16x3" else 1e-2)


@pytest.mark.parametrize("lens,nhead,hd", [([1, 5, 17, 26, 33, 64], 8, 64), ([100, 100, 7], 8, 64), ([3, 9, 128], 4, 16),
                                            ([1, 2, 3, 4, 26, 32], 4, 8)])
def test_attention_fp32(ops, lens, nhead, hd):
    d = nhead * hd
    offs = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=DEV)
    qkv = torch.randn(int(offs[-1]), 3 * d, device=DEV)
    o, _ = ops.mha_causal(qkv, offs, nhead, max(lens), prec="bf16x3", want_f32=True)
    _close(o, _attention_ref(qkv, offs, nhead), 1e-5)


def test_layernorm_embed_split(ops):
    x = torch.randn(3000, 512, device=DEV) * 3 + 1
    g, bb = torch.randn(512, device=DEV), torch.randn(512, device=DEV)
    y, ys = ops.layernorm(x, g, bb, 1e-5, prec="bf16x3")
    _close(y, torch.nn.functional.layer_norm(x.double(), (512,), g.double(), bb.double(), 1e-5), 1e-5)
    _close(ys.float(), y, 2e-5)
    V, d = 1000, 512
    emb, pe = torch.randn(V, d, device=DEV), torch.randn(200, d, device=DEV)
    tok = torch.randint(0, V, (777,), device=DEV, dtype=torch.int32)
    pos = torch.randint(0, 200, (777,), device=DEV, dtype=torch.int32)
    xf, _ = ops.embed(tok, pos, emb, pe, 22.627, prec="bf16x3")
    _close(xf, emb[tok.long()] * 22.627 + pe[pos.long()], 1e-6)
    s = ops.split(x, "bf16x3")
    _close(s.float(), x, 2e-5)


def test_kl_and_reparam(ops):
    mu = torch.randn(4096, 1024, device=DEV) * 0.03
    ls = torch.rand(1024, 1024, device=DEV) * -3.4 - 3.4
    out = torch.zeros(1, device=DEV)
    ops.kl_gauss(mu[2048:3072], ls, out)
    ref = (mu[2048:3072].double() ** 2 - 2 * ls.double() + torch.exp(2 * ls.double())).mean() / 2
    assert abs(out.item() - ref.item()) <= 1e-6 * abs(ref.item())      # north_star: KL within 1e-4 relative
    ops.kl_gauss(mu[2048:3072], ls, out, minus_one=True, scale=0.5, accumulate=True)
    ref2 = ref + 0.5 * ((mu[2048:3072].double() ** 2 - 2 * ls.double() + torch.exp(2 * ls.double()) - 1).mean() / 2)
    assert abs(out.item() - ref2.item()) <= 1e-6 * abs(ref2.item())
    eps = torch.randn(1024, 1024, device=DEV)
    w, _ = ops.reparam(mu[2048:3072], ls, eps=eps, prec="bf16x3", want_f32=True)
    _close(w, mu[2048:3072] + torch.exp(ls) * eps, 1e-6)
    z = ops.philox_normal(1234, 7, 4_000_000, DEV)
    assert abs(z.mean().item()) < 3e-3 and abs(z.std().item() - 1) < 3e-3
    kurt = ((z - z.mean()) ** 4).mean().item() / z.var().item() ** 2
    assert abs(kurt - 3) < 0.05
    w2, _ = ops.reparam(mu[2048:3072], ls, seed=1234, stream_id=7, prec="bf16", want_f32=True)
    _close(w2, mu[2048:3072] + torch.exp(ls) * z[: 1024 * 1024].view(1024, 1024), 1e-6)


def test_precise_gemm_truncation_bias_is_bounded(ops):
    """The tensor core truncates on every fp32 accumulate (a shrink of ~2e-8 per 16-wide K step).
    Chunked accumulation (k_chunk, the bf16x3 default) keeps the systematic error of a K = 4096
    product of same-sign operands below 5e-7 relative; one long accumulation does not."""
    M, N, K = 512, 512, 4096
    a = torch.rand(M, K, device=DEV) + 0.5
    b = torch.rand(N, K, device=DEV) + 0.5
    A, B = ops.split(a, "bf16x3"), ops.split(b, "bf16x3")
    ref = a.double() @ b.double().T
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A, B, prec="bf16x3", out_f32=out)
    bias_chunked = ((out.double() - ref) / ref).mean().item()
    ops.gemm(A, B, prec="bf16x3", out_f32=out, k_chunk=0)
    bias_single = ((out.double() - ref) / ref).mean().item()
    assert abs(bias_chunked) < 5e-7, (bias_chunked, bias_single)
    assert abs(bias_single) > 4 * abs(bias_chunked), (bias_chunked, bias_single)


@pytest.mark.parametrize("R,C", [(3200, 512), (100, 30000), (77, 130), (1, 8), (64, 64), (65, 63)])
def test_transposes_and_colsum(ops, R, C):
    x = torch.randn(R, C, device=DEV)
    t = ops.transpose_split(x, "bf16x3")
    _close(t.float(), x.t(), 2e-5)
    s, st = ops.split_transpose(x, "bf16x3")
    _close(s.float(), x, 2e-5)
    _close(st.float(), x.t(), 2e-5)
    t2 = ops.transpose_bf16(s, "bf16x3")
    _close(t2.float(), x.t(), 2e-5)
    out = torch.zeros(C, device=DEV)
    ops.colsum(x, out, scale=0.5)
    _close(out, 0.5 * x.double().sum(0), 1e-5)
    ops.colsum(s, out, accumulate=True)
    _close(out, 1.5 * x.double().sum(0), 3e-5)


@pytest.mark.parametrize("M,N,K", [(4500, 520, 328), (2048, 128, 64), (8192, 512, 4096)])
def test_cluster_sampled_gemm_bit_exact(ops, M, N, K):
    """The 4-CTA-cluster variant of blm_gemm_sampled (generated W~ quarters pushed through DSMEM):
    bit-identical to bf16(mu + sigma * eps) fed to the plain GEMM, for injected and Philox noise,
    with ragged M groups / N tiles / K blocks."""
    a = torch.randn(M, K, device=DEV) * 0.5
    mu = torch.randn(N, K, device=DEV) * 0.05
    ls = torch.rand(N, K, device=DEV) * -3.0 - 2.0
    A, mu_b, sg_b = ops.split(a, "bf16"), ops.split(mu, "bf16").hi, ops.sigma_bf16(ls)
    resid = torch.randn(M, N, device=DEV)
    for mode in ("ptr", "philox"):
        eps = torch.randn(N, K, device=DEV) if mode == "ptr" else ops.philox_normal(11, 4, N * K, DEV).view(N, K)
        out = torch.empty(M, N, device=DEV)
        ops.gemm_sampled(A, mu_b, sg_b, eps=eps if mode == "ptr" else None, seed=None if mode == "ptr" else 11,
                         stream_id=4, resid=resid, out_f32=out, how="tile")
        wt = torch.addcmul(mu_b.float(), sg_b.float(), eps).to(torch.bfloat16)
        ref = torch.empty(M, N, device=DEV)
        ops.gemm(A, ops.Split(wt), prec="bf16", resid=resid, out_f32=ref, k_chunk=0)
        assert torch.equal(out, ref), mode


@pytest.mark.parametrize("M,N,K", [(4500, 520, 328), (2048, 128, 64), (65536, 512, 4096), (30000, 4096, 512)])
def test_generate_once_sampled_gemm_bit_exact(ops, M, N, K):
    """blm_gemm_sampled in generate-once mode (W~ drawn once per launch into the L2-resident scratch, grid-wide
    arrival counter, then the pipelined GEMM): bit-identical to bf16(mu + sigma * eps) fed to the plain GEMM for
    injected and Philox noise; repeated launches reuse the self-re-arming workspace; bf16 outputs + GELU too."""
    a = torch.randn(M, K, device=DEV) * 0.5
    mu = torch.randn(N, K, device=DEV) * 0.05
    ls = torch.rand(N, K, device=DEV) * -3.0 - 2.0
    A, mu_b, sg_b = ops.split(a, "bf16"), ops.split(mu, "bf16").hi, ops.sigma_bf16(ls)
    resid = torch.randn(M, N, device=DEV)
    for rep in range(2):
        for mode in ("ptr", "philox"):
            sid = 4 + rep
            eps = torch.randn(N, K, device=DEV) if mode == "ptr" else ops.philox_normal(11, sid, N * K, DEV).view(N, K)
            out = torch.empty(M, N, device=DEV)
            ops.gemm_sampled(A, mu_b, sg_b, eps=eps if mode == "ptr" else None, seed=None if mode == "ptr" else 11,
                             stream_id=sid, resid=resid, out_f32=out, how="once")
            wt = torch.addcmul(mu_b.float(), sg_b.float(), eps).to(torch.bfloat16)
            ref = torch.empty(M, N, device=DEV)
            ops.gemm(A, ops.Split(wt), prec="bf16", resid=resid, out_f32=ref, k_chunk=0)
            assert torch.equal(out, ref), (mode, rep)
    bias = torch.randn(N, device=DEV)
    o1, o2 = ops.empty_split(M, N, "bf16", DEV), ops.empty_split(M, N, "bf16", DEV)
    ops.gemm_sampled(A, mu_b, sg_b, seed=11, stream_id=5, bias=bias, act=ops.ACT_GELU, out=o1, how="once")
    ops.gemm(A, ops.Split(wt), prec="bf16", bias=bias, act=ops.ACT_GELU, out=o2)
    assert torch.equal(o1.hi, o2.hi)


def test_fast_gelu_epilogue(ops):
    """BLM_ACT_GELU_FAST (packed fp16 evaluation, bf16-hi output): within bf16 rounding of the exact GELU."""
    M, N, K = 3000, 4096, 512
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    bias = torch.randn(N, device=DEV)
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    out = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", bias=bias, act=ops.ACT_GELU_FAST, out=out)
    z = A.hi.double() @ B.hi.double().T + bias.double()
    ref = torch.nn.functional.gelu(z)
    err = (out.hi.double() - ref).abs()
    assert (err <= 4.5e-3 * ref.abs() + 1.5e-3).all(), (err.max().item(),)        # bf16 ulp/2 + fp16 evaluation
    exact = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", bias=bias, act=ops.ACT_GELU, out=exact)
    # the two variants agree to within one bf16 ulp almost everywhere
    diff = (out.hi.float() - exact.hi.float()).abs()
    assert (diff <= 8e-3 * exact.hi.float().abs() + 2e-3).all()
    assert (diff == 0).float().mean().item() > 0.8


@pytest.mark.parametrize("M,N,K,act", [(4096, 1536, 512, 0), (3000, 4096, 512, 1), (2500, 696, 200, 0), (20000, 4096, 512, 6)])
def test_bf16_output_gemm_shapes_of_the_pair_kernel(ops, M, N, K, act):
    """bf16-only-output GEMMs at the shapes the CTA-pair (cta_group::2) kernel takes when BLM_GEMM2=1
    (and the 1-CTA kernel otherwise): QKV, FFN1 + GELU, ragged edges."""
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    bias = torch.randn(N, device=DEV)
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    out = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", bias=bias, act=act, col_scale=0.125, col_scale_cols=N // 3 // 8 * 8, out=out)
    z = A.hi.double() @ B.hi.double().T + bias.double()
    z[:, : N // 3 // 8 * 8] *= 0.125
    if act:
        z = torch.nn.functional.gelu(z)
    err = (out.hi.double() - z).abs()
    assert (err <= 4.5e-3 * z.abs() + (1.5e-3 if act == 6 else 1e-5)).all(), err.max().item()


@pytest.mark.parametrize("M,N,K,bias,offset", [
    (128, 128, 64, False, 0.0),
    (300, 256, 200, True, 0.0),          # ragged M and K
    (1000, 384, 512, True, 0.0),
    (20000, 512, 512, True, 0.0),        # o_net + residual + norm1 at the BASELINE width
    (65536 + 77, 512, 4096, True, 0.0),  # linear2 + residual + norm2, more tiles than SMs, ragged last tile
    (4000, 512, 512, True, 300.0),       # |mean| >> std: the shifted statistics must not cancel
])
def test_gemm_ln_matches_float64(ops, M, N, K, bias, offset):
    """blm_gemm_ln (projection + residual + LayerNorm in one kernel) against a float64 restatement on the
    same bf16 operands: fp32 output within 2e-5 of the largest magnitude, bf16 copy = rounded fp32."""
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    bi = torch.randn(N, device=DEV) if bias else None
    r = torch.randn(M, N, device=DEV) + offset
    g = torch.rand(N, device=DEV) + 0.5
    be = torch.randn(N, device=DEV)
    y32, ys = ops.gemm_ln(A, B, bias=bi, resid=r, gamma=g, beta=be, eps=1e-5)
    z = A.hi.double() @ B.hi.double().T + r.double()
    if bias:
        z = z + bi.double()
    ref = torch.nn.functional.layer_norm(z, (N,), g.double(), be.double(), 1e-5)
    # offset case: fp32 input rounding of a 300-sized value (3e-5 abs) is amplified by 1/std ~ 1
    _close(y32, ref, 2e-5 if offset == 0.0 else 2e-4)
    assert torch.equal(ys.hi, y32.to(torch.bfloat16))
    # in place on the residual stream
    r2 = r.clone()
    from bayeslms_b200 import ops as O
    d = O.GemmLnDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.B, d.ldb = A.hi.data_ptr(), A.hi.stride(0), B.hi.data_ptr(), B.hi.stride(0)
    d.bias, d.resid, d.ldr = O._ptr(bi), O._ptr(r2), N
    d.gamma, d.beta, d.eps = O._ptr(g), O._ptr(be), 1e-5
    d.out_f32, d.out_hi, d.ldc = O._ptr(r2), None, N
    O.check(O.lib().blm_gemm_ln(O.C.byref(d), O._stream()), "blm_gemm_ln")
    assert torch.equal(r2, y32)


def test_gemm_ln_rejects_other_widths(ops):
    from bayeslms_b200 import _lib
    a, b = ops.split(torch.randn(64, 64, device=DEV), "bf16"), ops.split(torch.randn(192, 64, device=DEV), "bf16")
    with pytest.raises(_lib.BlmError):
        ops.gemm_ln(a, b, bias=None, resid=torch.zeros(64, 192, device=DEV), gamma=torch.ones(192, device=DEV),
                    beta=torch.zeros(192, device=DEV), eps=1e-5)


def test_fast_gpmix_epilogue(ops):
    """BLM_ACT_GPMIX_FAST (GP activation mixture in packed fp16, tanh / sigmoid through MUFU.TANH.F16, bf16-hi
    output): within bf16 rounding + the fp16 evaluation of the exact mixture (model.py:1893-1899)."""
    M, N, K = 3000, 4096, 512
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    bias = torch.randn(N, device=DEV)
    coef = torch.rand(4, N, device=DEV) - 0.3
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    out = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", bias=bias, act=ops.ACT_GPMIX_FAST, coef=coef, out=out)
    z = A.hi.double() @ B.hi.double().T + bias.double()
    ref = _gp_mix(z, coef.double())
    scale = (coef.double().abs().unsqueeze(1) * torch.stack([torch.tanh(z).abs(), torch.sigmoid(z), torch.relu(z),
                                                               torch.nn.functional.gelu(z).abs()])).sum(0)
    err = (out.hi.double() - ref).abs()
    # bf16 ulp/2 of the result + fp16 evaluation of each term (relative to the terms' magnitudes: they can cancel)
    assert (err <= 4.5e-3 * ref.abs() + 2.5e-3 * scale + 1e-3).all(), (err.max().item(),)
    ref32 = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", bias=bias, act=ops.ACT_GPMIX, coef=coef, out=ref32)
    d = (out.hi.float() - ref32.hi.float()).abs()          # vs the fp32 evaluation: one bf16 ulp + the fp16 terms
    assert (d <= 8e-3 * ref32.hi.float().abs() + 2.5e-3 * scale.float() + 1e-3).all(), d.max().item()


def test_fp16_operands(ops):
    """a_f16: both operands fp16 (tcgen05 kind::f16 with A = B = F16; mixed f16 x bf16 is an illegal instruction)."""
    M, N, K = 1000, 512, 1024
    a = (torch.randn(M, K, device=DEV) * 0.5).to(torch.float16)
    b = (torch.randn(N, K, device=DEV) * 0.1).to(torch.float16)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(ops.Split(a), ops.Split(b), prec="bf16", out_f32=out, a_f16=True)
    _close(out, a.double() @ b.double().T, 1e-5)


@pytest.mark.parametrize("M,N,K,prec,a_mn,b_mn", [
    (512, 256, 128, "bf16", True, True),
    (4096, 512, 3200, "bf16", True, True),        # wgrad FFN1: dW[4096, 512] = dZ^T X over 3200 tokens
    (512, 4096, 3200, "bf16x3", True, True),      # wgrad FFN2, precise (three segments)
    (3200, 512, 4096, "bf16", False, True),       # dgrad FFN1: dX[3200, 512] = dZ[3200, 4096] W1[4096, 512]
    (3200, 4096, 512, "bf16x3", False, True),     # dgrad FFN2
    (1536, 520, 3203, "bf16x3", True, True),      # ragged reduction length and N
    (300, 200, 80, "bf16", True, False),
    (30000, 512, 3200, "bf16", True, True),       # decoder wgrad
])
def test_gemm_mn_major_operands(ops, M, N, K, prec, a_mn, b_mn):
    """MN-major operands (no transposed copies for wgrad / dgrad): a given as [K, M] and / or b as [K, N]."""
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    pad8 = lambda n: (n + 7) // 8 * 8   # noqa: E731
    if a_mn:
        buf = torch.zeros(K, pad8(M), device=DEV)
        buf[:, :M] = a.t()
        A = ops.split(buf, prec)
        A = ops.Split(A.hi[:, :M], None if A.lo is None else A.lo[:, :M])
    else:
        A = ops.split(a, prec)
    if b_mn:
        buf = torch.zeros(K, pad8(N), device=DEV)
        buf[:, :N] = b.t()
        B = ops.split(buf, prec)
        B = ops.Split(B.hi[:, :N], None if B.lo is None else B.lo[:, :N])
    else:
        B = ops.split(b, prec)
    out = torch.empty(M, pad8(N), device=DEV)[:, :N]
    bias = torch.randn(N, device=DEV)
    ops.gemm(A, B, prec=prec, bias=bias, out_f32=out, a_mn=a_mn, b_mn=b_mn)
    if prec == "bf16":
        af = A.hi.double().t() if a_mn else A.hi.double()
        bf = B.hi.double().t() if b_mn else B.hi.double()
        ref = af @ bf.T + bias.double()
        _close(out, ref, 2e-5)
    else:
        _close(out, a.double() @ b.double().T + bias.double(), 2e-5)


@pytest.mark.parametrize("lens,prec,tol", [
    ([100] * 4, "bf16x3", 2e-4), ([100] * 4, "bf16", 2e-2),
    ([1, 7, 16, 17, 32, 33, 64, 65, 100, 128], "bf16x3", 2e-4),
    ([5, 128, 31], "bf16", 2e-2),
])
def test_attention_backward_tensor_core(ops, lens, prec, tol):
    """blm_mha_causal_bwd_tc against float64 autograd of softmax(q k^T + causal mask) v per (sequence, head), and
    against the fp32 SIMT kernel; error relative to the largest gradient magnitude."""
    nhead, hd = 4, 64
    d = nhead * hd
    M = sum(lens)
    offs = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    qkv = torch.randn(M, 3 * d, device=DEV) * 0.7
    dout = torch.randn(M, d, device=DEV)
    got = ops.mha_causal_bwd(qkv, dout, offs, nhead, max(lens), 0.125, prec=prec)
    simt = ops.mha_causal_bwd(qkv, dout, offs, nhead, max(lens), 0.125)
    ref = torch.zeros(M, 3 * d, dtype=torch.float64, device=DEV)
    for s in range(len(lens)):
        a, b = int(offs[s]), int(offs[s + 1])
        x = qkv[a:b].double().clone().requires_grad_(True)
        q, k, v = (x[:, i * d:(i + 1) * d].view(b - a, nhead, hd).transpose(0, 1) for i in range(3))
        sc = q @ k.transpose(1, 2)
        mask = torch.ones(b - a, b - a, dtype=torch.bool, device=DEV).tril()
        p = torch.softmax(sc.masked_fill(~mask, float("-inf")), dim=-1)
        o = (p @ v).transpose(0, 1).reshape(b - a, d)
        (o * dout[a:b].double()).sum().backward()
        ref[a:b] = x.grad
    ref[:, :d] *= 0.125                      # dq is returned w.r.t. the unscaled projection (model.py:877)
    _close(simt, ref, 1e-5)
    _close(got, ref, tol)


@pytest.mark.parametrize("fast", [False, True])
def test_gelu_grad_epilogue_coalesced(ops, fast):
    """BLM_ACT_GELU_GRAD through the store-transpose path (fp32 + bf16 outputs, saved pre-activation read row
    coalesced); ``fast_act``: gelu' in packed fp16 (fast mode)."""
    M, N, K = 3200, 4096, 512
    a = torch.randn(M, K, device=DEV) * 0.5
    b = torch.randn(N, K, device=DEV) * 0.1
    z = torch.randn(M, N, device=DEV) * 1.5
    A, B = ops.split(a, "bf16"), ops.split(b, "bf16")
    out32 = torch.empty(M, N, device=DEV)
    out = ops.empty_split(M, N, "bf16", DEV)
    ops.gemm(A, B, prec="bf16", act=ops.ACT_GELU_GRAD, aux=z, out_f32=out32, out=out, fast_act=fast)
    zz = z.double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    ref = (A.hi.double() @ B.hi.double().T) * zz.grad
    err = (out32.double() - ref).abs()
    acc = (A.hi.double() @ B.hi.double().T).abs()
    assert (err <= (2e-3 if fast else 2e-6) * acc + 1e-6 * acc.max()).all(), err.max().item()
    assert torch.equal(out.hi, out32.to(torch.bfloat16))


# ------------------------------------------------------------------------------ dropout (fine-tune step)
def test_dropout_kernel_philox_and_injected(ops):
    """blm_dropout: out = x * m (+ resid).  Philox multipliers are 0 or 1/(1-p) with keep rate 1-p, a pure function of
    (seed + *seed_dev, stream, element); the exported multipliers re-injected as an explicit mask give the same bits."""
    n, p = 1 << 20, 0.2
    x = torch.randn(n, device=DEV)
    r = torch.randn(n, device=DEV)
    dr = ops.Drop(p, seed=123, stream_id=77)
    m, _ = ops.dropout(None, dr, n=n, device=x.device)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1.25) < 1e-6
    keep = float((m > 0).float().mean())
    assert abs(keep - 0.8) < 2e-3, keep
    y, ys = ops.dropout(x, dr, resid=r, prec="bf16x3")
    assert torch.equal(y, x * m + r)
    _close(ys.float(), y, 2e-5)
    y2, _ = ops.dropout(x, ops.Drop(p, mask=m), resid=r)
    assert torch.equal(y, y2)
    # in place, other stream -> other mask, device seed word adds to the key
    z = x.clone()
    ops.dropout(z, dr, out_f32=z)
    assert torch.equal(z, x * m)
    m2, _ = ops.dropout(None, ops.Drop(p, seed=123, stream_id=78), n=n, device=x.device)
    assert not torch.equal(m, m2)
    word = torch.tensor([23], dtype=torch.int64, device=DEV)
    m3, _ = ops.dropout(None, ops.Drop(p, seed=100, stream_id=77, seed_dev=word), n=n, device=x.device)
    assert torch.equal(m, m3)


def _attention_dropout_ref(qkv, dout, offs, nhead, mask_l, L):
    """float64 forward + autograd of softmax(q k^T + causal) -> * mask -> @ v per (sequence, head)."""
    M, d3 = qkv.shape
    d, hd = d3 // 3, d3 // 3 // nhead
    out = torch.zeros(M, d, dtype=torch.float64, device=qkv.device)
    grad = torch.zeros(M, d3, dtype=torch.float64, device=qkv.device)
    for s in range(len(offs) - 1):
        a, b = int(offs[s]), int(offs[s + 1])
        T = b - a
        x = qkv[a:b].double().clone().requires_grad_(True)
        q, k, v = (x[:, i * d:(i + 1) * d].view(T, nhead, hd).transpose(0, 1) for i in range(3))
        sc = q @ k.transpose(1, 2)
        tril = torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril()
        pr = torch.softmax(sc.masked_fill(~tril, float("-inf")), dim=-1)
        pr = pr * mask_l[s * nhead:(s + 1) * nhead, :T, :T].double()
        o = (pr @ v).transpose(0, 1).reshape(T, d)
        (o * dout[a:b].double()).sum().backward()
        out[a:b], grad[a:b] = o.detach(), x.grad
    return out, grad


@pytest.mark.parametrize("lens,prec,tol_f,tol_b", [([100] * 3, "bf16x3", 5e-5, 2e-4), ([100] * 3, "bf16", 1e-2, 2e-2),
                                                   ([7, 26, 32, 1], "bf16x3", 5e-5, 2e-4), ([5, 128, 31, 66], "bf16x3", 5e-5, 2e-4)])
def test_attention_dropout_forward_and_backward(ops, lens, prec, tol_f, tol_b):
    """Dropout on the attention probabilities (model.py:912-913) inside the tensor-core kernels: injected multipliers
    against float64 autograd, and Philox mode == the same step with its own exported multipliers injected."""
    nhead, hd, p = 4, 64, 0.3
    d = nhead * hd
    M, n_seq = sum(lens), len(lens)
    L = (max(lens) + 3) // 4 * 4
    offs = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    qkv = torch.randn(M, 3 * d, device=DEV) * 0.7
    dout = torch.randn(M, d, device=DEV)
    g = torch.Generator(device="cpu").manual_seed(5)
    mask = ((torch.rand(n_seq * nhead, L, L, generator=g) >= p).float() / (1 - p)).to(DEV)
    qs = ops.split(qkv, prec)
    src = qkv if prec == "bf16x3" else qs.hi.float()
    ref_o, ref_g = _attention_dropout_ref(src, dout, offs.tolist(), nhead, mask, L)
    o32, _ = ops.mha_causal_bf16(qs, offs, nhead, max(lens), prec=prec, want_f32=True, drop=ops.Drop(p, mask=mask))
    _close(o32, ref_o, tol_f)
    # the backward kernel takes the projection BEFORE the q scaling was folded in: feed q as is with scale 1
    got = ops.mha_causal_bwd(qkv, dout, offs, nhead, max(lens), 1.0, prec=prec, drop=ops.Drop(p, mask=mask))
    _, ref_g32 = _attention_dropout_ref(qkv, dout, offs.tolist(), nhead, mask, L)
    _close(got, ref_g32, tol_b)
    # Philox: export the multipliers of the stream, inject them, compare bit for bit
    dr = ops.Drop(p, seed=9, stream_id=1234)
    pm, _ = ops.dropout(None, dr, n=n_seq * nhead * L * L, device=qkv.device)
    pm = pm.view(n_seq * nhead, L, L)
    a1, _ = ops.mha_causal_bf16(qs, offs, nhead, max(lens), prec=prec, want_f32=True, drop=dr)
    a2, _ = ops.mha_causal_bf16(qs, offs, nhead, max(lens), prec=prec, want_f32=True, drop=ops.Drop(p, mask=pm))
    assert torch.equal(a1, a2)
    b1 = ops.mha_causal_bwd(qkv, dout, offs, nhead, max(lens), 1.0, prec=prec, drop=dr)
    b2 = ops.mha_causal_bwd(qkv, dout, offs, nhead, max(lens), 1.0, prec=prec, drop=ops.Drop(p, mask=pm))
    assert torch.equal(b1, b2)
    # and it really drops something
    plain, _ = ops.mha_causal_bf16(qs, offs, nhead, max(lens), prec=prec, want_f32=True)
    assert (plain - a1).abs().max() > 1e-2


def _lstm_layer_ref(gx, w, h0, c0, lengths, rounded=True):
    """fp64 recurrence.  ``rounded``: on the bf16-rounded W_hh, h re-rounded to bf16 as the operand of the next step
    (the kernel's bf16 mode); state kept in full precision either way."""
    T, B, H4 = gx.shape
    rnd = (lambda x: x.to(torch.bfloat16).double()) if rounded else (lambda x: x.double())
    w = rnd(w)
    h, c = h0.double().clone(), c0.double().clone()
    outs = torch.zeros(T, B, H4 // 4, dtype=torch.float64, device=gx.device)
    for t in range(T):
        a = gx[t].double() + rnd(h.float()) @ w.t()
        i, f, g, o = a.chunk(4, dim=1)
        cn = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        hn = torch.sigmoid(o) * torch.tanh(cn)
        live = (t < lengths).view(-1, 1)
        c, h = torch.where(live, cn, c), torch.where(live, hn, h)
        outs[t] = torch.where(live, hn, torch.zeros_like(hn))
    return outs, h, c


@pytest.mark.parametrize("prec", ["bf16", "bf16x3"])
@pytest.mark.parametrize("T,B,H", [(5, 384, 64), (7, 700, 128), (4, 1500, 256), (3, 2048, 1024), (6, 1024, 1024)])
def test_lstm_pair_kernel_matches_the_single_cta_kernel(ops, monkeypatch, T, B, H, prec):
    """lstm_pair_kernel (cta_group::2, the default from three 128-row tiles up; 16 units per CTA in bf16 mode, 8 with
    hi + lo slices in precise mode) against lstm_layer_kernel (BLM_LSTM_NO_PAIR=1) on the same inputs -- same K order,
    same fp32 accumulation: bit-identical -- and both against the fp64 recurrence.  Ragged batches (B not a multiple of
    256), rows of every length incl. 0."""
    g = torch.Generator(device=DEV).manual_seed(T * 1000 + B + H)
    gx = torch.randn(T, B, 4 * H, device=DEV, generator=g)
    w = torch.randn(4 * H, H, device=DEV, generator=g) / H ** 0.5
    h0 = torch.randn(B, H, device=DEV, generator=g) * 0.5
    c0 = torch.randn(B, H, device=DEV, generator=g) * 0.5
    lengths = torch.randint(0, T + 1, (B,), device=DEV, generator=g).to(torch.int32)
    lengths[:3] = T
    ws = ops.split(w, prec)

    def run(rows32=False):
        cs = torch.zeros(T * B, H, device=DEV)
        g2 = gx.view(T * B, 4 * H)
        if rows32:      # the 32-row-block layout of blm_gemm_desc.f32_rows32
            pad = ops.rows32_empty(T * B, 4 * H, DEV).zero_()
            pad[:T * B] = g2
            g2 = pad.view(-1, 32, H, 4).permute(0, 2, 1, 3).contiguous().view(-1, 4 * H)
            assert torch.equal(ops.rows32_to_dense(g2, T * B), gx.view(T * B, 4 * H))
        o32, o, hT, cT = ops.lstm_layer(g2, ws, h0, c0, lengths, T, B, H, prec=prec, want_f32=True, c_seq=cs,
                                        gx_rows32=rows32)
        torch.cuda.synchronize()
        return o32.clone(), o.float().clone(), hT.clone(), cT.clone(), cs

    pair = run()
    pair32 = run(rows32=True)
    monkeypatch.setenv("BLM_LSTM_NO_PAIR", "1")
    single = run()
    single32 = run(rows32=True)
    for other in (pair32, single, single32):
        for a, b in zip(pair, other):
            assert torch.equal(a, b)
    ro, rh, rc = _lstm_layer_ref(gx, w, h0, c0, lengths, rounded=prec == "bf16")
    tol = 2e-3 if prec == "bf16" else 3e-5       # bf16: MUFU gate math on rounded operands; precise: ~16 mantissa bits
    _close(pair[0].view(T, B, H), ro, tol)
    _close(pair[2], rh, tol)
    _close(pair[3], rc, tol)


@pytest.mark.parametrize("M,N,K,prec", [(300, 512, 200, "bf16"), (4096, 4096, 1024, "bf16"), (1000, 1024, 256, "bf16x3"),
                                        (77, 64, 64, "bf16x3"), (33, 40, 64, "bf16")])
def test_gemm_fp32_output_in_32_row_blocks(ops, M, N, K, prec):
    """blm_gemm_desc.f32_rows32: the same values as the row-major output, bit for bit, in [M/32][N/4][32][4] order
    (ragged M, N not a multiple of 32)."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = ops.split(torch.randn(M, K, device=DEV, generator=g), prec)
    b = ops.split(torch.randn(N, K, device=DEV, generator=g), prec)
    bias = torch.randn(N, device=DEV, generator=g)
    dense = torch.empty(M, N, device=DEV)
    ops.gemm(a, b, prec=prec, bias=bias, out_f32=dense)
    blocked = ops.rows32_empty(M, N, DEV)
    ops.gemm(a, b, prec=prec, bias=bias, out_f32=blocked, f32_rows32=True)
    assert torch.equal(ops.rows32_to_dense(blocked, M), dense)


@pytest.mark.parametrize("M,N,K,prec", [(3200, 512, 512, "bf16"), (3200, 4096, 512, "bf16"), (300, 256, 200, "bf16x3"),
                                        (1000, 512, 4096, "bf16x3")])
def test_gemm_epilogue_dropout_equals_the_dropout_kernel(ops, M, N, K, prec):
    """blm_gemm_desc.drop: element (m, col) takes multiplier m * N + col of the site.  Forward: out = resid + mask *
    act(acc + bias), out_pre unmasked.  *_GRAD: out = (mask * acc) * act'(aux), out_pre = mask * acc.  Checked against
    the unfused composition (same GEMM, then blm_dropout) for a Philox site incl. the device seed word, and for the
    exported multipliers re-injected as an explicit mask."""
    from bayeslms_b200.ops import ACT_GELU, ACT_GELU_GRAD, ACT_NONE
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = ops.split(torch.randn(M, K, device=DEV, generator=g) / K ** 0.5, prec)
    b = ops.split(torch.randn(N, K, device=DEV, generator=g), prec)
    bias = torch.randn(N, device=DEV, generator=g)
    resid = torch.randn(M, N, device=DEV, generator=g)
    aux = torch.randn(M, N, device=DEV, generator=g)
    word = torch.tensor([5], dtype=torch.int64, device=DEV)
    site = ops.Drop(0.2, seed=77, stream_id=(64 << 32) | 3, seed_dev=word)
    mult, _ = ops.dropout(None, site, n=M * N, device=torch.device(DEV))
    mult = mult.view(M, N)
    for drop in (site, ops.Drop(0.2, mask=mult.contiguous())):
        # forward, no activation: resid + mask * (acc + bias)
        plain = torch.empty(M, N, device=DEV)
        ops.gemm(a, b, prec=prec, bias=bias, out_f32=plain)
        got = torch.empty(M, N, device=DEV)
        ops.gemm(a, b, prec=prec, bias=bias, resid=resid, out_f32=got, drop=drop)
        assert torch.equal(got, plain * mult + resid)
        # forward, GELU: the saved pre-activation stays unmasked
        h, pre = ops.empty_split(M, N, prec, DEV), torch.empty(M, N, device=DEV)
        h0, pre0 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
        ops.gemm(a, b, prec=prec, bias=bias, act=ACT_GELU, out_f32=h0, out_pre=pre0)
        ops.gemm(a, b, prec=prec, bias=bias, act=ACT_GELU, out=h, out_pre=pre, drop=drop)
        assert torch.equal(pre, pre0)
        want = ops.split(h0 * mult, prec)
        assert torch.equal(h.hi, want.hi) and (h.lo is None or torch.equal(h.lo, want.lo))
        # backward: (mask * acc) * gelu'(aux)
        d0 = torch.empty(M, N, device=DEV)
        ops.gemm(a, b, prec=prec, act=ACT_NONE, out_f32=d0)
        dz = torch.empty(M, N, device=DEV)
        ops.gemm(a, b, prec=prec, act=ACT_GELU_GRAD, aux=aux, out_f32=dz, drop=drop)
        z = aux.double()
        gp = 0.5 * (1 + torch.erf(z / 2 ** 0.5)) + z * torch.exp(-0.5 * z * z) / (2 * np.pi) ** 0.5
        _close(dz, (d0 * mult).double() * gp, 1e-5)
        assert torch.equal(dz == 0, ((d0 * mult) == 0) | (dz == 0))
