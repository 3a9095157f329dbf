/*
 * bayeslm_b200 -- C ABI of the B200 (sm_100a) kernels behind the BayesLMs
 * n-best rescoring / fine-tune hot path.
 *
 * The reference (AmourWaltz/BayesLMs) has no FFI of its own: its seam is the
 * Python module `steps/pytorchnn/model.py`, whose nn.Module classes call into
 * PyTorch (cuBLAS/cuDNN/ATen).  Each entry point below replaces one group of
 * those library calls; the citation after "replaces:" is the reference call
 * site (path:line under the reference checkout).  The Python classes in
 * bayeslms_b200/model.py keep the reference constructors / forward signatures
 * and call these functions through ctypes (bayeslms_b200/_lib.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library
 *     never allocates device memory, never synchronises and only touches the
 *     stream it is given;
 *   - matrices are row-major; "ld*" are leading dimensions in ELEMENTS;
 *   - bf16 buffers are raw uint16_t storage (`blm_bf16`);
 *   - return value: BLM_OK or a negative BLM_ERR_*; blm_last_error() returns a
 *     thread-local message for the last failure;
 *   - there is no CPU fallback and no other architecture: blm_init() fails
 *     with BLM_ERR_ARCH unless the device is compute capability 10.0.
 */
#ifndef BAYESLM_B200_H_
#define BAYESLM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t blm_bf16;
typedef void* blm_stream; /* cudaStream_t */

enum {
  BLM_OK = 0,
  BLM_ERR_SHAPE = -1, /* unsupported / inconsistent sizes            */
  BLM_ERR_ALIGN = -2, /* pointer or leading dimension not 16-B aligned */
  BLM_ERR_ARCH = -3,  /* device is not sm_100                          */
  BLM_ERR_CUDA = -4,  /* a CUDA runtime / driver call failed           */
  BLM_ERR_ARG = -5    /* null pointer / bad enum                       */
};

/* activation fused into the GEMM epilogue */
enum {
  BLM_ACT_NONE = 0,
  BLM_ACT_GELU = 1,  /* exact erf GELU, model.py:1035                          */
  BLM_ACT_GPMIX = 2, /* sum_i coef[i,n]*act_i(z), acts tanh,sigmoid,relu,gelu;
                        model.py:1893-1899 with act_set of model.py:2263       */
  BLM_ACT_SOFTMAX_GRAD = 3, /* (exp(z - lse[m]) - [n == target[m]]) * grad_scale: the gradient of
                        the mean cross-entropy w.r.t. the logits (train.py:332,412),
                        written as bf16 (hi, lo) so the logits themselves never exist */
  BLM_ACT_GELU_GRAD = 4,  /* z * gelu'(aux[m,n]): backward of BLM_ACT_GELU at the saved
                        pre-activation aux                                       */
  BLM_ACT_GPMIX_GRAD = 5, /* z * sum_i coef[i,n] act_i'(aux[m,n]): backward of BLM_ACT_GPMIX */
  BLM_ACT_GELU_FAST = 6,  /* the same erf GELU evaluated in packed fp16 (two elements per
                        instruction, <= 5e-4 relative): for bf16-hi-only outputs (fast mode),
                        where the fp32 evaluation bounds the FFN1 epilogue            */
  BLM_ACT_GPMIX_FAST = 7  /* BLM_ACT_GPMIX in packed fp16 (tanh.approx.f16x2 for tanh and sigmoid):
                        bf16-hi-only outputs of the fast mode, N even                  */
};

/* where the N(0,1) noise of a reparameterised tensor comes from */
enum {
  BLM_EPS_NONE = 0,   /* posterior mean: W = mu                                */
  BLM_EPS_PTR = 1,    /* explicit eps tensor (parity with injected noise)      */
  BLM_EPS_PHILOX = 2  /* Philox4x32-10, key=(seed), counter=(element, stream)  */
};

int blm_version(void);
const char* blm_last_error(void);
/* one-time per-device setup: checks cc 10.0, raises dynamic-smem limits. */
int blm_init(int device);
int blm_num_sms(void);

/* ------------------------------------------------------------------ GEMM
 * C[M,N] = epilogue( sum over segments s of  A_s[M,K_s] * B_s[N,K_s]^T )
 *
 * Every operand is bf16, K-major.  fp32 tensors are carried as a (hi, lo)
 * bf16 pair (x ~= hi + lo, see blm_split_bf16); "precise" products
 * hi*hi + hi*lo + lo*hi are expressed as three segments, a plain bf16 GEMM as
 * one.  Up to BLM_MAX_SEG segments let two models share one accumulation
 * (logit interpolation, compute_sentence_scores_bayes_jianwei.py:163).
 *
 * epilogue:  z = acc + bias[n];  z *= col_scale for n < col_scale_cols
 *            (the q-scaling of model.py:877);  z = act(z);  z += resid[m,n];
 *            then stored to whichever of out_f32 / out_hi / out_lo is non-null
 *            (out_lo = bf16(z - float(out_hi))).
 *
 * replaces: F.linear / nn.Linear at model.py:850,855,876,921,1026,1028,1043,
 *           1129,1290,1303,1885 (cuBLAS behind ATen in the reference).
 * needs:    lda/ldb/ldc % 8 == 0, ldr % 4 == 0, 16-B aligned pointers
 *           (M, N and K_s are arbitrary: ragged edges are zero-filled / masked).
 */
#define BLM_MAX_SEG 6

struct blm_dropout_desc;

typedef struct blm_gemm_desc {
  int64_t M, N;
  int32_t nseg;
  int32_t act;                    /* BLM_ACT_*                                 */
  const blm_bf16* A[BLM_MAX_SEG]; /* [M, K_s], leading dimension lda[s]        */
  const blm_bf16* B[BLM_MAX_SEG]; /* [N, K_s], leading dimension ldb[s]        */
  int64_t K[BLM_MAX_SEG];
  int64_t lda[BLM_MAX_SEG];
  int64_t ldb[BLM_MAX_SEG];
  const float* bias;   /* [N] or null                                          */
  const float* coef;   /* [4, N] (BLM_ACT_GPMIX only)                          */
  float col_scale;     /* applied to columns [0, col_scale_cols)               */
  int32_t col_scale_cols;
  const float* resid;  /* [M, N] fp32 or null, leading dimension ldr           */
  int64_t ldr;
  float* out_f32;      /* any of the three may be null                         */
  blm_bf16* out_hi;
  blm_bf16* out_lo;
  int64_t ldc;         /* shared by the three outputs                          */
  const float* lse;    /* [M] log-sum-exp per row      (BLM_ACT_SOFTMAX_GRAD)  */
  const int32_t* targets; /* [M] target column per row (BLM_ACT_SOFTMAX_GRAD)  */
  float grad_scale;    /* e.g. 1 / M                   (BLM_ACT_SOFTMAX_GRAD)  */
  int32_t k_chunk;     /* > 0: close the tensor-core accumulator every k_chunk K elements
                          (multiple of 64, counted over the concatenated segments) and
                          sum the chunks in fp32 registers with round-to-nearest.  The
                          tensor core truncates on every accumulate, a bias proportional
                          to K; the precise (bf16x3) mode uses 128.  0: one accumulation. */
  float* out_pre;      /* optional fp32 [M, N] (ldc): the value before the activation,
                          i.e. after bias and q-scale (saved for the backward pass)   */
  const float* aux;    /* [M, N] fp32, leading dimension ldaux (BLM_ACT_*_GRAD)  */
  int64_t ldaux;
  int32_t a_f16;       /* 1: A AND B operands hold IEEE fp16 bits instead of bf16 (tcgen05 kind::f16; mixing
                          A = f16 with B = bf16 is an illegal instruction on sm_100a).  One segment, no k_chunk. */
  int32_t a_mn;        /* 1: A[s] is given MN-major, i.e. as the row-major [K_s, M] tensor (leading dimension lda[s]):  */
  int32_t b_mn;        /* 1: B[s] is the row-major [K_s, N] tensor.  Weight gradients dW = dY^T X take dY [tokens, N]
                          and X [tokens, K] as they are (a_mn = b_mn = 1), input gradients dX = dY W take W [N, K]
                          as it is (b_mn = 1): no transposed copies (tcgen05 MN-major shared-memory descriptors)  */
  int32_t fast_act;    /* 1: BLM_ACT_GELU_GRAD evaluates gelu' in packed fp16 (fast mode, <= 1e-3 absolute)            */
  int32_t f32_rows32;  /* 1: out_f32 is laid out in 32-row blocks, [ceil(M/32)][N/4][32 rows][4 floats] (buffer of
                          ceil(M/32)*32*N floats; ldc == N, N % 4 == 0; no out_pre).  A consumer that owns one row per
                          thread -- the LSTM gate math, gx_rows32 of blm_lstm_layer_seq -- then reads 512 contiguous
                          bytes per warp instruction instead of 32 separate rows.                                     */
  const struct blm_dropout_desc* drop;
                       /* optional (null: none): dropout fused into the epilogue.  Element (m, col) takes multiplier
                          m * N + col of the site (N % 32 == 0).  Forward epilogues: out = resid + mask . act(acc + bias),
                          out_pre = the pre-activation, unmasked -- `x + dropout(sublayer(x))`, model.py:1039-1046 in one
                          launch.  BLM_ACT_*_GRAD: out = (mask . acc) . act'(aux), out_pre = mask . acc -- the backward of
                          `dropout(act(z))`.  Same multipliers as blm_dropout on the dense [M, N] tensor.               */
} blm_gemm_desc;

int blm_gemm(const blm_gemm_desc* d, blm_stream stream);

/* --------------------------------------- GEMM + bias + residual + LayerNorm
 * out[M,N] = LayerNorm( resid[M,N] + A[M,K] * B[N,K]^T + bias[N] ) * gamma + beta
 *
 * The post-LN sublayer tail of the Transformer block in ONE kernel (fast bf16 mode): a CTA owns a
 * full 128 x N row block (N <= 512 = all 512 TMEM columns), so the row statistics are local to
 * the epilogue thread that owns the row; the pre-LayerNorm sum never reaches HBM.  The residual
 * arrives and both outputs (fp32 residual stream + bf16 operand copy) leave through TMA.
 * Statistics: shifted one-pass mean / M2 per half row, merged with Chan's formula; eps inside the
 * sqrt (nn.LayerNorm).
 * replaces: `src = norm1(src + dropout1(self_attn(...)))` / `src = norm2(src + dropout2(linear2(...)))`
 *           model.py:1040-1046, 1167-1175, 2283-2286 (eval mode: dropout is the identity).
 * needs:    N in {128, 256, 384, 512}; lda / ldb % 8 == 0; resid / out_f32 / out_hi dense enough for
 *           ldr % 4 == 0, ldc % 8 == 0; 16-B aligned pointers.  resid may alias out_f32.
 * returns BLM_ERR_SHAPE for other N: the caller then runs blm_gemm + blm_layernorm.            */
typedef struct blm_gemm_ln_desc {
  int64_t M, N, K;
  const blm_bf16* A;   /* [M, K] bf16, leading dimension lda                   */
  int64_t lda;
  const blm_bf16* B;   /* [N, K] bf16, leading dimension ldb                   */
  int64_t ldb;
  const float* bias;   /* [N] or null                                          */
  const float* resid;  /* [M, N] fp32, leading dimension ldr                   */
  int64_t ldr;
  const float* gamma;  /* [N]                                                  */
  const float* beta;   /* [N]                                                  */
  float eps;
  int32_t reserved;
  float* out_f32;      /* [M, N] fp32, leading dimension ldc (may be null)     */
  blm_bf16* out_hi;    /* [M, N] bf16, leading dimension ldc (may be null)     */
  int64_t ldc;
} blm_gemm_ln_desc;

int blm_gemm_ln(const blm_gemm_ln_desc* d, blm_stream stream);

/* ------------------------------------------------- tile-fused sampled GEMM (a)
 * C[M,N] = epilogue( A[M,K] * W~[N,K]^T ),  W~ = bf16( mu + sigma * eps ),  sigma = exp(lgstd)
 * built tile by tile in shared memory: TMA stages the bf16 mu and sigma tiles, generator warps
 * draw eps (Philox4x32-10 in registers, or an explicit tensor), overwrite the mu tile with W~
 * in place and hand it to tcgen05.mma; W~ never reaches HBM.  Noise indexing is identical to
 * blm_reparam(seed, stream_id) on the dense [N, K] tensor.  bf16 operands, fp32 accumulation.
 * BLM_EPS_NONE degenerates to a GEMM on mu.  Epilogue as blm_gemm (no q-scale).
 * mu / sigma are the bf16 copies the caller caches once per checkpoint (blm_split_bf16,
 * blm_sigma_bf16): both are sample-independent.
 * replaces: BayesLinear.forward model.py:1098-1129; GPNN weights model.py:1882-1885;
 *           Bayesian embedding model.py:1286-1290.
 * needs:    K % 8 == 0, lda / ldmu / ldc % 8 == 0, sigma dense [N, K], 16-B aligned pointers. */
typedef struct blm_gemm_sampled_desc {
  int64_t M, N, K;
  const blm_bf16* A;     /* [M, K] bf16, leading dimension lda                  */
  int64_t lda;
  const blm_bf16* mu;    /* [N, K] bf16 mean, leading dimension ldmu            */
  int64_t ldmu;
  const blm_bf16* sigma; /* [N, K] bf16 exp(lgstd), dense (null with EPS_NONE)  */
  const float* eps;      /* [N, K] fp32 dense (BLM_EPS_PTR)                     */
  int32_t eps_mode;      /* BLM_EPS_*                                           */
  int32_t act;           /* BLM_ACT_*                                           */
  uint64_t seed, stream_id;
  const float* bias;     /* [N] or null                                         */
  const float* coef;     /* [4, N] (BLM_ACT_GPMIX)                              */
  const float* resid;    /* [M, N] fp32 or null                                 */
  int64_t ldr;
  float* out_f32;
  blm_bf16* out_hi;
  blm_bf16* out_lo;
  int64_t ldc;
  void* workspace;        /* null: tile-stationary kernel (W~ regenerated per group of 4 M tiles, never stored).
                             blm_gemm_sampled_workspace_bytes(N, K) bytes, ZEROED ONCE by the caller and then
                             reusable by any number of launches on one stream: GENERATE-ONCE mode -- each CTA
                             of the persistent grid draws 1/grid of W~ exactly once at kernel entry into this
                             L2-resident scratch (4 MB for the FFN weight), a grid-wide arrival counter gates the
                             first weight-tile TMA load, and the rest of the launch is the pipelined tcgen05
                             GEMM of blm_gemm (all of its epilogues).  Bit-identical results.          */
  int64_t workspace_bytes;
  const float* mu_f32;    /* generate-once mode only, optional: fp32 mean [N, K] (leading dimension ldmu_f32) and  */
  int64_t ldmu_f32;       /* dense fp32 lgstd [N, K]; W~ = bf16(mu + exp(lgstd) eps) is then formed from the fp32  */
  const float* lgstd_f32; /* parameters, bit-identical to blm_reparam's bf16 output (mu / sigma may be null)       */
} blm_gemm_sampled_desc;

int64_t blm_gemm_sampled_workspace_bytes(int64_t N, int64_t K);

/* sigma[i] = bf16(exp(lgstd[i])): the sample-independent scale the fused GEMM multiplies eps by. */
int blm_sigma_bf16(const float* lgstd, blm_bf16* sigma, int64_t n, blm_stream stream);
int blm_gemm_sampled(const blm_gemm_sampled_desc* d, blm_stream stream);

/* ------------------------------------------- output projection + NLL (d)
 * nll[m] = logsumexp_v( h[m,:] . E[v,:] + b[v] ) - ( h[m,:] . E[t_m,:] + b[t_m] )
 * with the [M, V] logits living only in tensor memory: the vocabulary is
 * streamed tile by tile through the tensor cores and folded into a running
 * (max, sum-exp, target-logit) per row.  Segments as in blm_gemm (3 for the
 * precise product, 2x3 for two interpolated models whose activations were
 * pre-scaled by alpha / 1-alpha).
 *
 * workspace: blm_vocab_nll_workspace_bytes(M, V) bytes, 16-B aligned.
 * replaces: decoder nn.Linear (model.py:201,1220,1306) + nn.CrossEntropyLoss
 *           (compute_sentence_scores_bayes_jianwei.py:168,474).
 */
typedef struct blm_vocab_nll_desc {
  int64_t M, V;
  int32_t nseg;
  int32_t reserved;
  const blm_bf16* H[BLM_MAX_SEG]; /* [M, K_s] hidden states                    */
  const blm_bf16* E[BLM_MAX_SEG]; /* [V, K_s] output embedding                 */
  int64_t K[BLM_MAX_SEG];
  int64_t ldh[BLM_MAX_SEG];
  int64_t lde[BLM_MAX_SEG];
  const float* bias;      /* [V] or null                                       */
  const int32_t* targets; /* [M]                                               */
  float* nll;             /* [M] out                                           */
  void* workspace;
  int64_t workspace_bytes;
  float* lse;             /* [M] optional out: log-sum-exp of the row (backward) */
} blm_vocab_nll_desc;

int64_t blm_vocab_nll_workspace_bytes(int64_t M, int64_t V);
int blm_vocab_nll(const blm_vocab_nll_desc* d, blm_stream stream);

/* out[i] = sum of x[seg_offsets[i] .. seg_offsets[i+1])  (per-hypothesis score,
 * compute_sentence_scores_bayes_jianwei.py:170: length * mean CE).            */
int blm_segment_sum(const float* x, const int32_t* seg_offsets, int64_t nseg, float* out,
                    blm_stream stream);

/* Monte-Carlo predictive over K posterior samples (the K-sample score the
 * harness defines):  out[m] = -log( 1/K sum_k exp(-nll[k*M + m]) ). */
int blm_mc_combine(const float* nll, int64_t K, int64_t M, float* out, blm_stream stream);

/* ------------------------------------------------------------ elementwise */
/* fp32 -> bf16 hi (+ optional lo = bf16(x - hi)); n elements. */
int blm_split_bf16(const float* x, blm_bf16* hi, blm_bf16* lo, int64_t n, blm_stream stream);

/* x[m,:] = emb[tok[m],:] * scale + pe[pos[m],:]      (model.py:1284,116)
 * pe may be null (LSTM: plain lookup, model.py:218).                          */
int blm_embed(const int32_t* tokens, const int32_t* pos, const float* emb, const float* pe,
              float scale, int64_t M, int32_t d, float* out_f32, blm_bf16* out_hi,
              blm_bf16* out_lo, blm_stream stream);

/* y = LayerNorm(x) * gamma + beta, eps inside the sqrt (nn.LayerNorm,
 * model.py:1030-1031,1042,1045).                                              */
int blm_layernorm(const float* x, const float* gamma, const float* beta, float eps, int64_t M,
                  int32_t d, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo,
                  blm_stream stream);

/* Reparameterised tensor  w = mu + exp(lgstd) * eps   (model.py:670-725,
 * 1086-1102, 1876-1883), written as fp32 and/or a bf16 (hi, lo) pair.
 * mu is [rows, cols] with leading dimension ldmu (row slices of a larger
 * tensor, model.py:718); lgstd/eps/out are dense [rows, cols].
 * eps_mode BLM_EPS_PHILOX draws eps[i] from Philox4x32-10 with key `seed`,
 * counter (i/4, stream_id), Box-Muller; every rank / launch shape sees the
 * same noise for the same (seed, stream_id).                                  */
int blm_reparam(const float* mu, int64_t ldmu, const float* lgstd, const float* eps,
                int32_t eps_mode, uint64_t seed, uint64_t stream_id, int64_t rows, int64_t cols,
                float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream);
/* The N(0,1) stream blm_reparam uses, exposed so tests can inject it into the
 * oracle. */
int blm_philox_normal(uint64_t seed, uint64_t stream_id, int64_t n, float* out, blm_stream stream);
/* out[i] = scale * that stream (the N(0, 0.1^2) hidden noise of model.py:2786 uses scale 0.1). */
int blm_philox_normal_scaled(uint64_t seed, uint64_t stream_id, int64_t n, float scale, float* out,
                             blm_stream stream);

/* ------------------------------------------------------------ attention (c)
 * Causal multi-head self-attention over packed variable-length hypotheses.
 * qkv is [M, 3*d] fp32 (q already scaled), hypothesis i owns rows
 * [seq_offsets[i], seq_offsets[i+1]); heads are contiguous head_dim-wide
 * column groups; out[M, d] = softmax(q k^T + causal mask) v.
 * replaces: torch.bmm + mask add + F.softmax + torch.bmm, model.py:889-920.
 * needs:    head_dim <= 128, sequence length <= 128.                          */
int blm_mha_causal(const float* qkv, const int32_t* seq_offsets, int64_t nseq, int32_t nhead,
                   int32_t head_dim, int32_t max_len, float* out_f32, blm_bf16* out_hi,
                   blm_bf16* out_lo, blm_stream stream);

/* Same attention on bf16 operands and tensor cores (the production path): qkv is the [M, 3*d]
 * output of the QKV projection written as bf16 hi (+ optional lo: the precise hi*hi + hi*lo + lo*hi
 * product), leading dimension ld; q, k, v blocks are staged with cp.async and multiplied with
 * mma.sync m16n8k16 (fp32 accumulate), online softmax in registers.  out_* have leading dimension ldo.
 * needs:    head_dim == 64, sequence length <= 128.                            */
int blm_mha_causal_bf16(const blm_bf16* qkv_hi, const blm_bf16* qkv_lo, int64_t ld,
                        const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                        int32_t max_len, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo,
                        int64_t ldo, blm_stream stream);

/* ------------------------------------------------------------------ KL (e)
 * out[0] (+)= scale * 0.5 * mean_{rows x cols}( mu^2 - 2 lgstd + exp(2 lgstd) [- 1] )
 * (model.py:762-765 without the -1; 1115; 1255; 1821-1825 with it).
 * mu is a [rows, cols] view with leading dimension ldmu, lgstd is dense.
 * `accumulate` adds to out[0] instead of overwriting it.
 * workspace: blm_kl_workspace_bytes() bytes, zero-initialised once by the
 * caller (the kernel leaves it zeroed again).                                 */
int64_t blm_kl_workspace_bytes(void);
int blm_kl_gauss(const float* mu, int64_t ldmu, const float* lgstd, int64_t rows, int64_t cols,
                 int32_t minus_one, float scale, int32_t accumulate, float* out, void* workspace,
                 blm_stream stream);

/* ---------------------------------------------------------------- LSTM (b)
 * One layer of the LSTM recurrence (gate order i,f,g,o; model.py:812 ->
 * torch LSTM) over a batch of B independent sequences advanced in lock step.
 *   gates_x [T, B, 4H] fp32 : x_t W_ih^T + b_ih + b_hh, hoisted (blm_gemm)
 *   w_hh_hi/lo [4H, H] bf16 : recurrent weight (hi, lo may be null)
 *   h0, c0 [B, H] fp32      : initial state
 *   lengths [B] int32       : steps t >= lengths[b] leave (h, c) of row b
 *                             untouched (right padding)
 *   out_* [T, B, H]         : h_t (fp32 and/or bf16 hi/lo), zero where padded
 *   hT, cT [B, H] fp32      : state after each row's last valid step
 * workspace: blm_lstm_workspace_bytes(B, H) bytes (the call resets its barrier
 * counter itself with a stream-ordered memset).  H % 8 == 0, H <= 1024,
 * B <= 2048 rows per launch.                                                  */
int64_t blm_lstm_workspace_bytes(int64_t B, int64_t H);
int blm_lstm_layer(const float* gates_x, const blm_bf16* w_hh_hi, const blm_bf16* w_hh_lo,
                   const float* h0, const float* c0, const int32_t* lengths, int64_t T, int64_t B,
                   int64_t H, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, float* hT,
                   float* cT, void* workspace, blm_stream stream);
/* The same launch that also records the cell state after every live step (c_seq [T, B, H] fp32; out_f32 holds h):
 * one launch can then carry a whole chain of utterances -- hypothesis #0 of utterance after utterance, the hidden
 * carry of score.py:261-274 -- and the state at every utterance boundary is read back from (out_f32, c_seq).
 * gx_rows32 != 0: gates_x was written by blm_gemm with f32_rows32 = 1 over M = T * B rows (32-row blocks), which
 * makes the gate math's one-row-per-thread reads contiguous; 0: row-major [T, B, 4H].                          */
int blm_lstm_layer_seq(const float* gates_x, int32_t gx_rows32, const blm_bf16* w_hh_hi, const blm_bf16* w_hh_lo,
                       const float* h0, const float* c0, const int32_t* lengths, int64_t T, int64_t B,
                       int64_t H, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo, float* hT,
                       float* cT, float* c_seq, void* workspace, blm_stream stream);

/* GP-LSTM cell update for one timestep (GPLSTMCell.Gplstm, model.py:1743-1777; gpnn_type <= 3,
 * gate_type 1..4 = i, f, g, o).  acc5 [B, 5H] (ld): the four gate pre-activations followed by the GP
 * unit's pre-activation z = W_g [x; h] + b_g, all produced by one blm_gemm on the concatenated weights
 * [W_hh; W_g(h part)] with the hoisted input part as residual.  The chosen gate is replaced by
 * sum_i coef[i, u] act_i(z), acts (sigmoid, tanh, relu)[:n_act] (model.py:1692-1697).  Updates c, h (fp32,
 * in place, rows with t >= lengths[b] untouched), writes the bf16 operand copy of h for the next step and
 * this step's output rows (zero where padded).                                                    */
int blm_gp_lstm_cell(const float* acc5, int64_t ld, const float* coef, int32_t n_act, int32_t gate_type,
                     const int32_t* lengths, int32_t t, int64_t B, int32_t H, float* c, float* h,
                     blm_bf16* h_hi, blm_bf16* h_lo, float* out_f32, blm_bf16* out_hi, blm_bf16* out_lo,
                     blm_stream stream);
/* Plain cell update for one timestep from the four gate pre-activations acc4 [B, 4H] (ld): the step of the GP-LSTM
 * gate types 5-7, whose GP unit sits outside the gate nonlinearities (model.py:1745-1750, 1763-1764) and is a
 * blm_gemm with the BLM_ACT_GPMIX epilogue.  c_in (may be null): the cell state the update starts from when it is
 * not c itself -- gate type 5, c passed through the GP unit first; c is written for live rows only.          */
int blm_lstm_cell_step(const float* acc4, int64_t ld, const float* c_in, const int32_t* lengths, int32_t t, int64_t B,
                       int32_t H, float* c, float* h, blm_bf16* h_hi, blm_bf16* h_lo, float* out_f32,
                       blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream);

/* ------------------------------------------------- fine-tune step (train.py:306-438)
 * Backward twins of the kernels above and the optimiser.  The contractions of the backward pass
 * (dgrad dX = dY W, wgrad dW = dY^T X) are blm_gemm calls on transposed bf16 operand copies
 * (blm_transpose_*), with the activation derivative fused as BLM_ACT_*_GRAD and the softmax
 * gradient as BLM_ACT_SOFTMAX_GRAD; what follows is the HBM-bound rest.  Gradient buffers are
 * fp32; "accumulate" adds to the destination instead of overwriting it.                        */

/* out[c, r] = bf16 hi (+ lo) of x[r, c]: x fp32 [R, C] (ldx), out [C, R] (ldo, multiple of 8).  */
int blm_transpose_split(const float* x, int64_t ldx, int64_t R, int64_t C, blm_bf16* out_hi,
                        blm_bf16* out_lo, int64_t ldo, blm_stream stream);
/* both bf16 copies of an fp32 weight in one pass: hi/lo [R, C] (ld) and its transpose t_hi/t_lo
 * [C, R] (ldt) -- the B operands of the forward product and of its dgrad.                       */
int blm_split_transpose(const float* x, int64_t ldx, int64_t R, int64_t C, blm_bf16* hi, blm_bf16* lo,
                        int64_t ld, blm_bf16* t_hi, blm_bf16* t_lo, int64_t ldt, blm_stream stream);
/* same from a bf16 (hi[, lo]) source.                                                          */
int blm_transpose_bf16(const blm_bf16* hi, const blm_bf16* lo, int64_t ld, int64_t R, int64_t C,
                       blm_bf16* out_hi, blm_bf16* out_lo, int64_t ldo, blm_stream stream);
/* out[n] (+)= scale * sum_m x[m, n]   (bias gradients); fixed summation order.
 * workspace: blm_colsum_workspace_bytes(M, N) bytes, zeroed once by the caller.                 */
int64_t blm_colsum_workspace_bytes(int64_t M, int64_t N);
int blm_colsum(const float* x, int64_t ldx, int64_t M, int64_t N, float scale, int32_t accumulate,
               float* out, void* workspace, blm_stream stream);
int blm_colsum_bf16(const blm_bf16* hi, const blm_bf16* lo, int64_t ld, int64_t M, int64_t N,
                    float scale, int32_t accumulate, float* out, void* workspace, blm_stream stream);

/* LayerNorm backward (nn.LayerNorm, model.py:1030-1031): x is the LayerNorm INPUT.
 * dx = rstd (g dy - mean(g dy) - xhat mean(g dy xhat)); dgamma (+)= sum dy xhat; dbeta (+)= sum dy. */
int64_t blm_layernorm_bwd_workspace_bytes(int64_t M, int32_t d);
int blm_layernorm_bwd(const float* dy, const float* x, const float* gamma, float eps, int64_t M,
                      int32_t d, float* dx, float* dgamma, float* dbeta, int32_t accumulate,
                      void* workspace, blm_stream stream);
/* The same pass with two by-products of the step it sits in: dx_hi / dx_lo (optional), the bf16 operand copies of dx
 * for the dgrad / wgrad GEMMs that consume it, and dxsum [d] (optional) (+)= sum_m dx[m, :] -- dx is the gradient of
 * the projection output that was added to the residual stream, so this is that projection's bias gradient
 * (`linear2.bias` / `o_net.bias`, model.py:1039-1046).  One split and one column-sum launch less per LayerNorm. */
int blm_layernorm_bwd_ex(const float* dy, const float* x, const float* gamma, float eps, int64_t M, int32_t d,
                         float* dx, blm_bf16* dx_hi, blm_bf16* dx_lo, float* dgamma, float* dbeta,
                         int32_t accumulate, float* dxsum, int32_t dxsum_accumulate, void* workspace,
                         blm_stream stream);

/* Backward of blm_mha_causal: qkv fp32 [M, 3d] (q already scaled), dout [M, d] -> dqkv [M, 3d];
 * the q block of dqkv is multiplied by q_scale (gradient w.r.t. the unscaled projection,
 * model.py:877).  Softmax is recomputed, nothing of size T x T is stored.
 * needs: head_dim == 64, sequence length <= 128.                                               */
int blm_mha_causal_bwd(const float* qkv, int64_t ld, const float* dout, int64_t ldo,
                       const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                       int32_t max_len, float q_scale, float* dqkv, int64_t ldd, blm_stream stream);
/* The same gradient on tensor cores (mma.sync m16n8k16 bf16; precise = 1: hi/lo parts, three products per term):
 * S, dP are rebuilt per block in registers, dS / P are re-packed from accumulator to A fragments, nothing T x T
 * is stored.  head_dim 64, max_len <= 128.                                                                    */
int blm_mha_causal_bwd_tc(const float* qkv, int64_t ld, const float* dout, int64_t ldo,
                          const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                          int32_t max_len, float q_scale, int32_t precise, float* dqkv, int64_t ldd,
                          blm_stream stream);

/* ------------------------------------------------- dropout of the fine-tune step
 * The reference trains with nn.Dropout at: the embedding + positional output (model.py:116), the attention
 * probabilities (model.py:912-913), the attention and FFN residual branches and the FFN activation
 * (dropout1 / dropout2 / dropout, model.py:1039-1045, 1163-1174, 2275-2285, 2793-2803), and around the LSTM
 * (model.py:218-221).  A kept element is scaled by 1 / (1 - p).  The multiplier of element i is either read from an
 * explicit fp32 tensor (`mask`, values 0 or 1/(1-p): parity with masks injected into the oracle) or derived from
 * Philox4x32-10: word (i & 3) of counter (i >> 2) on stream `stream_id` under key seed + *seed_dev; kept iff
 * word >= floor(p * 2^32).  `seed_dev` (may be null) is read on the device, so a CUDA graph that captured the call
 * replays with fresh masks after a 8-byte write.  p == 0 and mask == null: identity.                           */
typedef struct blm_dropout_desc {
  const float* mask;         /* explicit multipliers, or null for Philox                  */
  float p;                   /* drop probability in [0, 1)                                */
  int32_t reserved;
  uint64_t seed;             /* Philox key (host part)                                    */
  const uint64_t* seed_dev;  /* device word added to the key, or null                     */
  uint64_t stream_id;        /* Philox stream: one per (site, layer)                      */
} blm_dropout_desc;

/* out = x * m (+ resid), elementwise over n values (n % 4 == 0, 16-B aligned); x null = ones (exports the
 * multipliers).  Outputs: fp32 and / or bf16 (hi[, lo]); out_f32 may alias x.
 * replaces: nn.Dropout at model.py:116, 1039, 1043, 1045 and 218-221, and its autograd twin (same call on the
 * gradient).                                                                                                   */
int blm_dropout(const float* x, int64_t n, const blm_dropout_desc* drop, const float* resid, float* out_f32,
                blm_bf16* out_hi, blm_bf16* out_lo, blm_stream stream);

/* blm_mha_causal_bf16 / blm_mha_causal_bwd_tc with dropout on the attention probabilities (model.py:912-913):
 * O = (m . P) V.  The multiplier of (sequence s, head h, query i, key j) is element
 * ((s * nhead + h) * L + i) * L + j of the mask / Philox stream, L = max_len rounded up to a multiple of 4.  */
int blm_mha_causal_bf16_dropout(const blm_bf16* qkv_hi, const blm_bf16* qkv_lo, int64_t ld,
                                const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                                int32_t max_len, const blm_dropout_desc* drop, float* out_f32, blm_bf16* out_hi,
                                blm_bf16* out_lo, int64_t ldo, blm_stream stream);
int blm_mha_causal_bwd_tc_dropout(const float* qkv, int64_t ld, const float* dout, int64_t ldo,
                                  const int32_t* seq_offsets, int64_t nseq, int32_t nhead, int32_t head_dim,
                                  int32_t max_len, float q_scale, int32_t precise, const blm_dropout_desc* drop,
                                  float* dqkv, int64_t ldd, blm_stream stream);

/* GP mixture (model.py:1893-1899): dcoef[i, n] (+)= sum_m dh[m, n] act_i(z[m, n]).              */
int blm_gpmix_dcoef(const float* z, const float* dh, int64_t ld, int64_t M, int64_t N,
                    int32_t accumulate, float* dcoef, blm_stream stream);

/* Variational hidden noise of VTransformerEncoderLayer (model.py:2785-2801, training, T == 100):
 *   fp = f + e exp(f rho[t]),  e = eps (BLM_EPS_PTR, already N(0, noise_std^2)) or
 *   noise_std * Philox normal.  Rows are sequence-major: row = b*T + t; rho, mean_p are [T, d].
 * Backward also differentiates the layer's KL (model.py:2770-2781)
 *   KL = 0.5 mean_{T,B,d}( ((1 - mean_p) fp)^2 - 2 rho + exp(2 rho) )
 * scaled by kl_scale: dfp (may be null) is the downstream gradient w.r.t. fp; outputs df,
 * drho, dmean_p and klpart[T, d] with KL = 0.5 sum(klpart) / (T B d).                          */
int blm_vnoise_fwd(const float* f, const float* rho, const float* eps, int32_t eps_mode, uint64_t seed,
                   uint64_t stream_id, float noise_std, const float* resid /* optional: out = resid + fp */,
                   int64_t B, int32_t T, int32_t d, float* fp, blm_stream stream);
int blm_vnoise_bwd(const float* dfp, const float* f, const float* rho, const float* mean_p,
                   const float* eps, int32_t eps_mode, uint64_t seed, uint64_t stream_id,
                   float noise_std, int64_t B, int32_t T, int32_t d, float kl_scale, float* df,
                   float* drho, float* dmean_p, float* klpart, blm_stream stream);

/* dE[tokens[m], :] += scale * dx[m, :]   (nn.Embedding backward, fp32 atomics).                 */
int blm_embed_bwd(const float* dx, const int32_t* tokens, float scale, int64_t M, int32_t d, float* dE,
                  blm_stream stream);

/* Gradient of scale * blm_kl_gauss: dmu += c mu, dlgstd += c (exp(2 lgstd) - 1), c = scale/(rows cols). */
int blm_kl_gauss_bwd(const float* mu, int64_t ldmu, const float* lgstd, int64_t rows, int64_t cols,
                     float scale, float* dmu, int64_t lddmu, float* dlgstd, blm_stream stream);
/* Backward of blm_reparam: G = dL/dw [rows, cols] (ldg): dmu (+)= G, dlgstd (+)= G eps exp(lgstd). */
int blm_reparam_bwd(const float* G, int64_t ldg, const float* lgstd, const float* eps, int32_t eps_mode,
                    uint64_t seed, uint64_t stream_id, int64_t rows, int64_t cols, int32_t accumulate,
                    float* dmu, int64_t lddmu, float* dlgstd, blm_stream stream);

/* LSTM backward through time (the LSTM families' share of train.py:306-438; torch's LSTM backward
 * behind model.py:812).  Row layout is time-major: row = t * B + b.
 * blm_lstm_gates_act: gates [T, B, 4H] holds the pre-activations Z = x W_ih^T + b + h_{t-1} W_hh^T
 *   (rebuilt by one blm_gemm over all steps); on return it holds sigmoid/tanh of them (i,f,g,o) and
 *   c_all [T, B, H] the cell states, starting from c0 [B, H].
 * blm_lstm_bwd_step: one step of the backward recurrence for rows [B] at time t:
 *   dgates_t from (dout_t + dh_rec, dc); dc is updated in place (dc_is_zero: treat the carry as 0,
 *   first step of the backward walk).  dh_rec (nullable) = dgates_{t+1} W_hh, a blm_gemm between steps.
 *   dgates leave as fp32 and bf16 hi[, lo], all with leading dimension 4H.                      */
int blm_lstm_gates_act(float* gates, const float* c0, int64_t T, int64_t B, int64_t H, float* c_all,
                       blm_stream stream);
int blm_lstm_bwd_step(const float* gates_t, const float* c_prev, const float* c_t, const float* dout_t,
                      const float* dh_rec, float* dc, int32_t dc_is_zero, int64_t B, int64_t H,
                      float* dgates_f32, blm_bf16* dgates_hi, blm_bf16* dgates_lo, blm_stream stream);

/* out[0] (+)= scale * sum x (squares == 0) or scale * sum x^2 (squares == 1); deterministic.
 * workspace: blm_reduce_workspace_bytes(), zeroed once by the caller.                           */
int64_t blm_reduce_workspace_bytes(void);
int blm_reduce(const float* x, int64_t n, int32_t squares, float scale, int32_t accumulate, float* out,
               void* workspace, blm_stream stream);
/* clip_grad_norm_(max_norm) + SGD momentum (train.py:419,466) over one flat buffer:
 *   c = min(1, max_norm / (sqrt(norm_sq[0]) grad_scale + 1e-6)) (1 if norm_sq is null);
 *   v = momentum v + c grad_scale g;  p -= lr v.                                               */
int blm_sgd_momentum(float* p, const float* g, float* v, int64_t n, float lr, float momentum,
                     const float* norm_sq, float max_norm, float grad_scale, blm_stream stream);
/* The same update that also writes the bf16 (hi[, lo]) operand copies of the updated parameters (out_lo may be
 * null): the next step's GEMMs read them directly instead of re-splitting every weight.          */
int blm_sgd_momentum_split(float* p, const float* g, float* v, int64_t n, float lr, float momentum,
                           const float* norm_sq, float max_norm, float grad_scale, blm_bf16* out_hi,
                           blm_bf16* out_lo, blm_stream stream);

/* ------------------------------------------------- training-mode Variational / GP LSTM cells (row a20)
 * Variational cell (VLSTMCell + VNN, model.py:2470-2579): in training, after every step h <- h + n_t,
 * n_t = e_t exp(hidden_lgstd), e_t ~ N(0, 0.1^2) of shape (1, H) shared by the batch rows.  Additive noise commutes
 * with the recurrent product, so W_hh n_{t-1} is folded into the hoisted input gates as a per-step bias row
 * (blm_rowgroup_add) and blm_lstm_layer runs unchanged on the pure hidden state.
 *   blm_vnn_noise:    n[t, j] = e[t, j] exp(rho[j])                                          (model.py:2557-2563)
 *   blm_rowgroup_add: out[g B + b, :] = x[g B + b, :] + r[g, :], fp32 and / or bf16 (hi, lo); out_f32 may alias x
 *   blm_rowgroup_sum: out[g, :] = sum_b x[g B + b, :]  (gradient of a per-step row; deterministic)
 *   blm_vnn_kl:       kl_out[0] += mean_{B,H}(h^2 - 2 rho + exp(2 h) - 1) / 2 on the pure last-step hidden
 *                     (VNN.kl_divergence, model.py:2545-2551, as written); dh += kl_scale dKL/dh, drho += kl_scale dKL/drho
 *   blm_vnn_drho:     drho[j] += exp(rho[j]) sum_t dn[t, j] e[t, j]
 * GP cell (GPLSTMCell.Gplstm, model.py:1743-1777): blm_gp_lstm_bwd_step is the backward twin of blm_gp_lstm_cell for
 * one timestep: from acc5 (pre-activations i, f, g, o, z), the cell states and dh = dout + dh_rec it writes the
 * gradient of the five pre-activation blocks (the replaced gate's own block gets 0) as fp32 and bf16 (hi, lo),
 * updates the carried dL/dc in place and accumulates dcoef[k, u] += sum_b dgate act_k(z).                        */
int blm_vnn_noise(const float* e, const float* rho, int64_t T, int32_t H, float* n, blm_stream stream);
int blm_rowgroup_add(const float* x, const float* r, int64_t G, int64_t B, int32_t W, float* out_f32, blm_bf16* out_hi,
                     blm_bf16* out_lo, blm_stream stream);
int blm_rowgroup_sum(const float* x, int64_t G, int64_t B, int32_t W, float* out, blm_stream stream);
/* Backward of a stand-alone GP unit of the LSTM cells (gate types 5-7, model.py:1745-1750): h = sum_i coef[i, n]
 * act_i(z) with acts (sigmoid, tanh, relu); dz = dh . dh/dz (fp32 and optional bf16 hi[, lo]), dcoef[3, N] += sum_m dh
 * act_i(z).  dh, z, dz share the leading dimension ld.                                                          */
int blm_gp3_bwd(const float* dh, const float* z, const float* coef, int64_t ld, int64_t M, int32_t N, float* dz,
                blm_bf16* dz_hi, blm_bf16* dz_lo, float* dcoef, blm_stream stream);
int blm_vnn_kl(const float* h, const float* rho, int64_t B, int32_t H, float kl_scale, float* kl_out, float* dh,
               float* drho, blm_stream stream);
int blm_vnn_drho(const float* dn, const float* e, const float* rho, int64_t T, int32_t H, float* drho, blm_stream stream);
int blm_gp_lstm_bwd_step(const float* acc5, int64_t ld, const float* coef, int32_t n_act, int32_t gate_type,
                         const float* c_prev, const float* c_t, const float* dout, const float* dh_rec, float* dc,
                         int32_t dc_is_zero, int64_t B, int32_t H, float* dacc, blm_bf16* dacc_hi, blm_bf16* dacc_lo,
                         int64_t ldd, float* dcoef, blm_stream stream);

/* ------------------------------------------------- scorer file formats (host code, no GPU work)
 * The reference scorer tokenises one hypothesis at a time in Python (load_nbest score.py:20-51, read_vocab :63-84,
 * get_input_and_target :87-120) and formats one score at a time (write_scores :283-303).  These entry points do the
 * same text -> id -> text work in one pass over the file bytes, writing the packed int32 ids straight into the
 * caller's (pinned) staging buffers -- the layout blm_embed / blm_vocab_nll consume.
 *
 * blm_vocab_from_text: `words.txt` bytes ("word idx" per line; any other field count is an error, score.py:79) ->
 *   word -> id map, id = order of first occurrence.  Null on error (blm_last_error).
 * blm_nbest_scan: line starts [n_lines + 1] and scored positions per line (words + 1) of a `words_text` file; pass
 *   null arrays to count only; n_threads host threads split the text at line boundaries.  flags bit 0: the text holds a Unicode-only whitespace character that Python's
 *   str.split() would split on (the caller then takes its own slow path).
 * blm_nbest_tokenize: lines [l0, l1) -> tok = <s> w1..wL, tgt = w1..wL <s>, pos = 0..L at offsets offs[i] - offs[l0]
 *   (offs = prefix sum of the per-line counts); OOV -> <unk>; n_threads host threads split the range.
 * blm_nbest_group: utterance of every line (key before the last '-', dense ids in order of first appearance -- the
 *   reference's dict order), 1-based index inside the utterance, key location in the text.
 * blm_scores_format: "<utt>-<idx> %.4f\n" for the lines in `order`; returns bytes written (or needed, if > cap). */
typedef struct blm_vocab blm_vocab;
blm_vocab* blm_vocab_from_text(const char* text, int64_t nbytes);
int64_t blm_vocab_size(const blm_vocab* vocab);
int32_t blm_vocab_id(const blm_vocab* vocab, const char* word, int64_t len);
void blm_vocab_free(blm_vocab* vocab);
int blm_nbest_scan(const char* text, int64_t nbytes, int64_t cap_lines, int64_t* line_begin, int32_t* line_tokens,
                   int64_t* n_lines, int64_t* n_tokens, int32_t* flags, int32_t n_threads);
int blm_nbest_tokenize(const blm_vocab* vocab, const char* text, const int64_t* line_begin, const int64_t* offs,
                       int64_t l0, int64_t l1, int32_t* tok, int32_t* tgt, int32_t* pos, int32_t n_threads);
int blm_nbest_group(const char* text, const int64_t* line_begin, int64_t n_lines, int32_t* utt_of_line,
                    int32_t* idx_in_utt, int64_t* key_begin, int32_t* key_len, int64_t* n_utts);
int64_t blm_scores_format(const char* text, const int64_t* key_begin, const int32_t* key_len, const int32_t* idx_in_utt,
                          const int64_t* order, int64_t n, const float* scores, char* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BAYESLM_B200_H_ */
