/* ct_b200.h — C ABI of libct_b200.so: the B200 (sm_100a) implementation of CleanTransformer's
 * dense forward/backward + optimizer + gradient all-reduce hot path.
 *
 * The reference (firechecking/CleanTransformer) has no FFI layer: the path sits behind Python
 * classes that call PyTorch ops. Each entry point below replaces the PyTorch-op sequence of the
 * cited reference lines (paths relative to the reference root) and is what a ctypes binding in the
 * reference's own modules would call.
 *
 * Conventions
 *   - every tensor argument is a raw DEVICE pointer owned by the caller (PyTorch); the library never
 *     frees or retains it past the call. Only the ct_comm_* buffers are library-owned.
 *   - dtype enum: 0 = f32, 1 = bf16, 2 = f16.   activation enum: 0 none, 1 relu, 2 gelu_erf,
 *     3 gelu_tanh, 4 tanh.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *     nothing touches the legacy default stream; calls are CUDA-graph capturable unless noted.
 *   - return 0 on success, >0 = cudaError_t, <0 = library code (-1 bad argument/shape,
 *     -2 unsupported dtype/arch, -3 workspace too small, -4 comm not initialised). The message is
 *     available per calling thread through ct_last_error(). Re-entrant: callable concurrently from
 *     the Python main thread and PyTorch's autograd thread.
 */
#ifndef CT_B200_H_
#define CT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CT_B200_VERSION 100

#define CT_F32 0
#define CT_BF16 1
#define CT_F16 2

#define CT_ACT_NONE 0
#define CT_ACT_RELU 1      /* transformer.py:100 */
#define CT_ACT_GELU_ERF 2  /* modeling_bert.py:229 (torch.nn.GELU) */
#define CT_ACT_GELU_TANH 3 /* modeling_bloom.py:335-345, modeling_gpt.py:112-122 */
#define CT_ACT_TANH 4      /* modeling_bert.py:283-286 (pooler) */
/* GEMM epilogues only. CT_ACT_GELU_TANH_SAVE_GRAD as `act`: y = gelu_tanh(t) and the `preact` output receives
 * gelu_tanh'(t) — what GeLUFunction.backward (modeling_bloom.py:288-306) derives from the saved input — so that
 * the backward epilogue only multiplies. CT_ACT_GRAD_PRECOMPUTED as `actgrad_act`: `actgrad_src` holds act'(pre). */
#define CT_ACT_GELU_TANH_SAVE_GRAD 5
#define CT_ACT_GRAD_PRECOMPUTED 6

/* ---- library ------------------------------------------------------------------------------- */
int ct_version(void);
int ct_last_error(char* buf, size_t n);
/* 0 only if `device` is an sm_100 part (B200). */
int ct_device_check(int device);
/* Kernel-variant knobs for A/B measurements and tests (not needed for normal use). Names:
 * "LN_BWD_IMPL", "ATTN_FWD_IMPL", "ATTN_BWD_IMPL", "GEMM_EPI_IMPL", "GEMM_2CTA", "CE_IMPL", "GEMM_SPLITK"; 0 = default.
 * Initial values are read from the environment variables CT_<NAME>. */
int ct_set_option(const char* name, int value);
int ct_get_option(const char* name, int* value);

/* ---- LayerNorm: CleanTransformer/transformer.py:61-89 (LayerNorm._mean / forward) ----------- *
 * y = gamma * (x - mean) / sqrt(mean((x-mean)^2 + eps)) + beta over the last `cols` elements.
 * x: [rows, cols] f32 or bf16. Optional second output y2 (e.g. a bf16 copy that feeds the next
 * GEMM while y stays f32). mean/rstd ([rows] f32) may be NULL for inference. */
int ct_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                     int y_dtype, void* y2, int y2_dtype, float* mean, float* rstd, int64_t rows,
                     int64_t cols, float eps, void* stream);
/* Backward of the above (what autograd derives from transformer.py:86-89).
 * dx = (dx_add ? dx_add : 0) + dLN/dx ; dgamma/dbeta are ACCUMULATED (+=) into f32 buffers when
 * dgb_accumulate != 0, else overwritten. dy may be NULL only if dy2 is given; when both dy and dy2
 * are non-NULL the incoming gradient is their sum (two consumers of the LN output).
 * workspace (nullable): f32 scratch of >= 2 * (2 * #SMs) * cols floats enables a deterministic two-stage
 * dgamma/dbeta reduction; without it the per-CTA partials are combined with atomics. */
int ct_layernorm_bwd(const void* dy, int dy_dtype, const void* dy2, int dy2_dtype, const void* x,
                     int x_dtype, const float* gamma, const float* mean, const float* rstd,
                     const void* dx_add, int dx_add_dtype, void* dx, int dx_dtype, float* dgamma,
                     float* dbeta, int dgb_accumulate, float* workspace, size_t workspace_bytes,
                     int64_t rows, int64_t cols, void* stream);

/* Extended backward: same arithmetic, plus two fused by-products of the dx pass that the pre-LN block
 * backward (modeling_bloom.py:142-159 as differentiated by autograd) otherwise pays separate kernels for:
 *   dx2   - a second copy of dx in another dtype (the bf16 operand of the next dgrad/wgrad GEMM);
 *   dxsum - column sums of dx ([cols] f32): when x was produced by `residual + Linear(...)`
 *           (modeling_bloom.py:121-122, 267-269) this is that Linear's bias gradient.
 * workspace: >= 3 * (2 * #SMs) * cols floats for the deterministic two-stage reductions. */
typedef struct ct_ln_bwd_args {
  int64_t rows, cols;
  const void* dy;  int32_t dy_dtype;      /* nullable when dy2 is given */
  const void* dy2; int32_t dy2_dtype;     /* nullable; gradient = dy + dy2 */
  const void* x;   int32_t x_dtype;
  const float* gamma; const float* mean; const float* rstd;
  const void* dx_add; int32_t dx_add_dtype; /* nullable: dx += dx_add (residual-path gradient) */
  void* dx;  int32_t dx_dtype;
  void* dx2; int32_t dx2_dtype;           /* nullable */
  float* dgamma; float* dbeta; int32_t dgb_accumulate;
  float* dxsum; int32_t dxsum_accumulate; /* nullable */
  float* workspace; size_t workspace_bytes;
} ct_ln_bwd_args;
int ct_layernorm_bwd_ex(const ct_ln_bwd_args* args, void* stream);

/* ---- optimizers: CleanTransformer/optimizer.py ---------------------------------------------- *
 * One vectorised kernel over a flat f32 arena (p, g, m, v all [n]).
 * mode 0 = decoupled weight decay, the arithmetic of torch.optim.AdamW which the examples call
 *          (examples/ft_bloom.py:19,70,90);
 * mode 1 = the reference's own class, optimizer.py:71-97: coupled L2 (g += wd*p written back to g),
 *          m_hat/(sqrt(v_hat)+eps), `step` is the reference's counter which starts at 1.
 * Hyper-parameters are doubles (Python floats); derived constants (1-beta, 1-beta^t, lr/(1-beta1^t))
 * are formed in double and rounded to fp32 once, like the reference's Python-scalar arithmetic.
 * grad_scale multiplies g on load (1/world for a summed all-reduce; 1.0 otherwise).
 * p_shadow (nullable): bf16 copy of the updated parameters for the next forward's GEMMs. */
int ct_adamw_step(float* p, float* g, float* m, float* v, void* p_shadow_bf16, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int64_t step, int mode,
                  float grad_scale, void* stream);
/* Same arithmetic over `ntensors` separate tensors (host arrays of device pointers). */
int ct_adamw_multi(int ntensors, float* const* p, float* const* g, float* const* m, float* const* v,
                   void* const* p_shadow_bf16, const int64_t* sizes, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int64_t step, int mode,
                   float grad_scale, void* stream);
/* optimizer.py:28-50 (SGD.step): g += wd*p; buf = first ? g : momentum*buf + (1-dampening)*g;
 * g = buf; p -= lr*g. `buf` may be NULL when momentum == 0. g is rewritten like the reference. */
int ct_sgd_step(float* p, float* g, float* buf, int64_t n, float lr, float momentum,
                float dampening, float weight_decay, int first_step, void* stream);
/* Same arithmetic over `ntensors` separate tensors (host arrays of device pointers; buf may be NULL). */
int ct_sgd_multi(int ntensors, float* const* p, float* const* g, float* const* buf, const int64_t* sizes,
                 float lr, float momentum, float dampening, float weight_decay, int first_step, void* stream);

/* ---- elementwise helpers --------------------------------------------------------------------- */
/* dst[i] = (dst_dtype) src[i]  (autocast's parameter/activation casts) */
int ct_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* out[j] (+)= sum_i x[i, j]  — bias gradients (rows x cols, row stride ld elements) */
int ct_colsum(const void* x, int x_dtype, int64_t ld, float* out, int accumulate, int64_t rows,
              int64_t cols, void* stream);
/* y = act(x) and dx = dy * act'(x): modeling_bloom.py:335-363, modeling_gpt.py:112-122 */
int ct_act_fwd(const void* x, int x_dtype, void* y, int y_dtype, int act, int64_t n, void* stream);
int ct_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx, int dx_dtype,
               int act, int64_t n, void* stream);

/* ---- GEMM family (tcgen05 / TMEM / TMA): every nn.Linear / Conv1D on the path ---------------- *
 * C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *   a_mn_major = 0: A stored [M rows][K] with K contiguous (row stride lda elements)
 *   a_mn_major = 1: A stored [K rows][M] with M contiguous (row stride lda)      — "A transposed"
 *   b_mn_major likewise for B over (N, K).
 * epilogue, in order:  t = alpha*acc + bias[n];  if (preact) preact[m,n] = t;  t = act(t);
 *   if (actgrad_src) t *= act'(actgrad_src[m,n]);  if (residual) t += residual[m,n];
 *   C[m,n] = t + beta * C_old[m,n]   (beta in {0,1}; beta=1 needs c_dtype f32).
 * Covers: nn.Linear fwd (transformer.py:37,98-102; modeling_bloom.py:79,121,256,267;
 * modeling_bert.py:238-247), Conv1D fwd with its [in,out] weight via b_mn_major=1
 * (modeling_gpt.py:32-46), and the dgrad / wgrad GEMMs autograd derives from them. */
typedef struct {
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  int32_t ab_dtype; /* CT_BF16 or CT_F16 */
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  void* C;
  int32_t c_dtype;
  int64_t ldc;
  float alpha, beta;
  const float* bias; /* [N] or NULL */
  int32_t act;
  void* preact; /* nullable, [M,N] */
  int32_t preact_dtype;
  int64_t ldp;
  const void* actgrad_src; /* nullable, [M,N]: multiply by act'(src) with `actgrad_act` */
  int32_t actgrad_dtype;
  int32_t actgrad_act;
  int64_t ldg;
  const void* residual; /* nullable, [M,N] */
  int32_t res_dtype;
  int64_t ldr;
  int32_t impl; /* 0 auto, 1 force 1-CTA tcgen05, 2 force the SIMT kernel (small/unaligned shapes),
                   3 force the 2-CTA (cta_group::2, 256x256 tile) tcgen05 kernel,
                   4 force the skinny weight-streaming kernel (M <= 32, K-major A and B, K % 32 == 0: the q_len = 1
                   decode step, generation_util.py:57-119); auto picks it whenever that layout holds */
  float* row_stats; /* nullable. LM head (modeling_bloom.py:220-230): softmax statistics of every stored bf16 row,
                       [2*ceil(N/256)][M] float2 = (max * log2e, sum of 2^(v*log2e - max*log2e)) per (slot, row), a
                       slot being the four 32-column chunks of one parity of a 256-column tile; a slot without any
                       column holds (-inf, 0). Consumed by ct_cross_entropy_fwd_stats. Plain bf16 epilogue of the
                       2-CTA kernel only (CT_ERR_UNSUPPORTED otherwise). */
} ct_gemm_args;
int ct_gemm(const ct_gemm_args* args, void* stream);

/* Named wrappers matching SURVEY.md §8 b3.
 * y[M,N] = act(x[M,K] @ w^T + bias) (+ residual); w is [N,K] (nn.Linear) or, with w_in_out = 1,
 * [K,N] (Conv1D). */
int ct_gemm_bias_act(const void* x, const void* w, int w_in_out, const float* bias,
                     const void* residual, int res_dtype, void* y, int y_dtype, void* preact,
                     int act, int64_t M, int64_t N, int64_t K, int ab_dtype, void* stream);
/* dx[M,K] = dy[M,N] @ w   (w [N,K], or [K,N] when w_in_out) ; optional dx *= act'(actgrad_src) */
int ct_gemm_dgrad(const void* dy, const void* w, int w_in_out, void* dx, int dx_dtype,
                  const void* actgrad_src, int actgrad_act, int64_t M, int64_t N, int64_t K,
                  int ab_dtype, void* stream);
/* dw (+)= dy^T @ x  as f32 [N,K] (or [K,N] when w_in_out);  db (+)= colsum(dy) when db != NULL */
int ct_gemm_wgrad_bias(const void* dy, const void* x, int w_in_out, float* dw, float* db,
                       int accumulate, int64_t M, int64_t N, int64_t K, int ab_dtype, void* stream);


/* ---- attention: QK^T -> (+ALiBi / mask) -> softmax -> PV, flash-style -------------------------- *
 * One kernel replaces the batched GEMM + mask + softmax + batched GEMM + re-layout sequences of
 *   CleanTransformer/models/modeling_bloom.py:99-116  (alibi.baddbmm, masked_fill(finfo.min), softmax, PV)
 *   CleanTransformer/models/modeling_gpt.py:83-103    (w*b + -1e4*(1-b), + additive mask, softmax, PV)
 *   CleanTransformer/transformer.py:41-57             (scores/sqrt(d) + additive mask, softmax, PV)
 * without materialising the [B,H,Sq,Sk] score tensor.
 *
 * q/k/v element (b,h,s,d) lives at base + b*sb + h*sh + s*ss + d  (d contiguous), which covers the
 * per-head interleaved fused QKV of Bloom (bloom:81-82), the blocked [Q|K|V] of GPT (gpt:72), three
 * separate projections (transformer.py:37) and [b,h,t,d] KV caches (bloom:88-92, gpt:76-80).
 * o has the merged-head layout the reference produces after transpose+view ([B,Sq,H*D]).
 *
 * score2(i,j) = log2e * ( scale * q_i.k_j )  + kbias2[b,h,j]            (log2 domain)
 *   if causal and j > i + (Sk - Sq):  score2 = causal_fill*log2e + kbias2[b,h,j]
 *   score2 = max(score2, -FLT_MAX)   (so a fully masked row is uniform, as masked_fill(finfo.min) is)
 * kbias2 = log2e * (alibi_slope[h]*pos[b,j] + key_add[b,j]) comes from ct_attn_mask_prep.
 * lse2[b,h,i] = log2(sum_j 2^score2(i,j)) is kept for the backward pass. */
typedef struct {
  int32_t B, H, Sq, Sk, D;
  int32_t dtype; /* CT_BF16 or CT_F16 */
  const void* q;
  int64_t q_sb, q_sh, q_ss;
  const void* k;
  int64_t k_sb, k_sh, k_ss;
  const void* v;
  int64_t v_sb, v_sh, v_ss;
  void* o;
  int64_t o_sb, o_sh, o_ss;
  float* lse2; /* [B,H,Sq], nullable for inference */
  float scale;
  int32_t causal;
  float causal_fill;          /* -FLT_MAX (Bloom fill) or -1e4 (GPT, modeling_gpt.py:89) */
  const float* kbias2;        /* nullable; element (b,h,j) at kbias2 + b*kb_sb + h*kb_sh + j */
  int64_t kb_sb, kb_sh;
  const int32_t* first_valid; /* [B] nullable: first key that is not hard-masked (left padding) */
  int32_t impl;               /* 0 auto, 1 force tcgen05 (D == 64), 2 force SIMT */
  const int32_t* seq_len_dev; /* nullable; q_len = 1 decode only: the number of cached keys is read from device memory
                                 (Sk is then just the capacity), so a captured decode step serves every position */
  /* attention-probability dropout (transformer.py:47-50, modeling_gpt.py:93-96, modeling_bloom.py:111-113,
   * modeling_bert.py attention dropout): P(i,j) is zeroed with probability dropout_p AFTER the softmax (the row sum
   * keeps every key) and the survivors are scaled by 1/(1-p). The keep decision of element (b,h,i,j) is
   * keep(rng_seed, rng_stream, hi = b*H + h, lo = i*Sk + j) as defined under "dropout" below; the backward takes the same
   * three values. */
  float dropout_p;     /* 0: none */
  uint32_t rng_stream; /* distinguishes the dropout sites / calls that share one seed */
  uint64_t rng_seed;
  /* q_len = 1 decode only, nullable: the new token's key / value rows, element (b,h,d) at k_new + b*kn_sb + h*kn_sh + d.
   * The kernel first stores them into cache row (key count - 1) of k / v (modeling_gpt.py:76-80, modeling_bloom.py:88-92:
   * the torch.concat of the reference) and then attends — one launch instead of two appends and the attention. */
  const void* k_new;
  const void* v_new;
  int64_t kn_sb, kn_sh, vn_sb, vn_sh;
} ct_attn_args;
int ct_attn_fwd(const ct_attn_args* args, void* stream);

/* ---- dropout ------------------------------------------------------------------------------------------------------
 * One counter-based generator serves every dropout site, so a mask never has to be stored and a test (or the oracle,
 * oracle/ct_oracle.py: dropout_keep) can restate it bit for bit:
 *   key0 = (uint32)seed + stream * 0x632BE5AB,  key1 = (uint32)(seed >> 32) ^ (stream * 0x2545F491)
 *   h = lo * 0x9E3779B1 + key0;  h ^= hi * 0x85EBCA77 + key1;
 *   h ^= h >> 16; h *= 0x7FEB352D; h ^= h >> 15; h *= 0x846CA68B; h ^= h >> 16;      (all mod 2^32)
 *   keep(seed, stream, hi, lo)  <=>  (h >> 8) >= round(p * 2^24)
 * Elementwise sites (torch.nn.Dropout on hidden states: transformer.py:109-116, modeling_gpt.py:136,
 * modeling_bert.py:253-262, modeling_bloom.py dropout_add): element e of the flattened tensor has hi = e >> 32,
 * lo = (uint32)e.   out = (residual ? residual : 0) + (keep ? x / (1-p) : 0); the backward of the site is the same call
 * on the incoming gradient without a residual. x / residual / out: f32, bf16 or f16 (ct_dtype), n elements. */
int ct_dropout(const void* x, int x_dtype, const void* residual, int res_dtype, void* out, int out_dtype, int64_t n,
               float p, uint64_t seed, uint32_t rng_stream, void* stream);

/* Backward: recomputes P from q,k,lse2; dq/dk/dv use the same (sb,sh,ss) addressing as q/k/v.
 * delta ([B,H,Sq] f32) and dq_accum (B*H*ceil(Sq/128)*128*D f32, layout private to the library) are
 * caller-allocated workspaces. */
typedef struct {
  ct_attn_args f; /* forward arguments (o = forward output, lse2 = saved statistics) */
  const void* dout; /* same layout as o */
  void* dq;
  int64_t dq_sb, dq_sh, dq_ss;
  void* dk;
  int64_t dk_sb, dk_sh, dk_ss;
  void* dv;
  int64_t dv_sb, dv_sh, dv_ss;
  float* delta;
  float* dq_accum;
} ct_attn_bwd_args;
int ct_attn_bwd(const ct_attn_bwd_args* args, void* stream);
/* diagnostic: resident CTAs per SM of the default tcgen05 forward / backward kernels (occupancy API, no launch);
 * detail: NULL or 8 ints (forward kernel: registers, static smem, dynamic smem, occupancy at dynamic smem
 * - {0, 1, 2, 4, 16} KB) */
int ct_attn_occupancy(int* fwd_ctas_per_sm, int* bwd_ctas_per_sm, int* detail);

/* Build kbias2 / first_valid from the caller's attention_mask [B,Sk] (1 = attend).
 * mask_dtype: CT_F32, 3 = int64, 4 = int32.
 * mode 0 (Bloom, modeling_bloom.py:176-185,309-331): key_add = mask ? 0 : -FLT_MAX and ALiBi
 *        pos = (cumsum(mask)-1)*mask times slopes[h]  -> kbias2 [B,H,Sk]
 * mode 1 (GPT, modeling_gpt.py:176-179): key_add = (1-mask)*finfo(f32).min       -> kbias2 [B,1,Sk]
 * mode 2 (BERT, modeling_bert.py:303-304): key_add = (1-mask)*-10000.0           -> kbias2 [B,1,Sk]
 * slopes ([H] f32) only for mode 0. */
int ct_attn_mask_prep(const void* attention_mask, int mask_dtype, int64_t B, int64_t Sk, int64_t H,
                      int mode, const float* slopes, float* kbias2, int32_t* first_valid,
                      void* stream);


/* ---- embedding: modeling_bloom.py:190, modeling_gpt.py:169,184, modeling_bert.py:297-300 ------- *
 * out[t,:] (+)= weight[ids[t],:]  (f32 weights, int64 ids); accumulate adds into `out` (position /
 * segment embeddings). Backward scatters dout rows into dweight with red.add (dweight must hold the
 * running gradient or zeros); rows equal to padding_idx receive no gradient (modeling_bert.py:273). */
int ct_embedding_fwd(const int64_t* ids, const float* weight, float* out, int64_t T, int64_t H,
                     int64_t V, int accumulate, void* stream);
int ct_embedding_bwd(const int64_t* ids, const float* dout, float* dweight, int64_t T, int64_t H,
                     int64_t V, int64_t padding_idx, void* stream);
/* Gather + LayerNorm in one pass (SURVEY §8 f N3): y = LN(table0[ids0] (+ table1[ids1]) (+ table2[ids2])) —
 * modeling_bloom.py:190-191 (word_embeddings -> word_embeddings_layernorm), modeling_bert.py:297-301 (word +
 * segment + position tables -> embedding_post LayerNorm). ids* [rows] int64 (a broadcast position / segment row is
 * expanded by the caller), tables f32 [vocab*, cols]; tables 1 and 2 optional (NULL, in order). `emb` (nullable,
 * f32 [rows, cols]) receives the sum — the LayerNorm input that ct_layernorm_bwd needs; y / y2 / mean / rstd as
 * ct_layernorm_fwd. The backward is ct_layernorm_bwd followed by ct_embedding_bwd per table. Needs
 * cols % 128 == 0, cols <= 1024 and 16-byte aligned pointers (CT_ERR_UNSUPPORTED otherwise: call the two
 * kernels). An id outside its table gives a NaN row, like ct_embedding_fwd. */
int ct_embedding_layernorm_fwd(const int64_t* ids0, const float* table0, int64_t vocab0,
                               const int64_t* ids1, const float* table1, int64_t vocab1,
                               const int64_t* ids2, const float* table2, int64_t vocab2,
                               const float* gamma, const float* beta, float* emb, void* y, int y_dtype,
                               void* y2, int y2_dtype, float* mean, float* rstd, int64_t rows,
                               int64_t cols, float eps, void* stream);

/* ---- cross entropy: modeling_bloom.py:224-230 (shift + torch CrossEntropyLoss, mean) ---------- *
 * logits [rows, V] (bf16 or f32). shift != 0: row r = (b, s) is scored against labels[r + 1] and the
 * last position of every length-S sequence is skipped; shift == 0: labels[r]. Rows whose target is
 * ignore_index do not count. Writes loss (f32 scalar, mean over counted rows) and, when dlogits is
 * non-NULL, dlogits = (softmax - onehot) / count in the logits dtype (zero on skipped rows).
 * workspace: f32 [rows + 4]. */
int ct_cross_entropy_fwd(const void* logits, int dtype, int64_t ld, const int64_t* labels,
                         void* dlogits, int64_t ldd, float* loss, float* workspace, int64_t rows,
                         int64_t V, int64_t S, int shift, int64_t ignore_index, void* stream);
/* Same contract for bf16 logits whose per-row softmax statistics were produced by the logits GEMM itself
 * (ct_gemm_args.row_stats, n_slots = 2*ceil(V/256)): one streaming pass, the rows are not read for the log-sum-exp. */
int ct_cross_entropy_fwd_stats(const void* logits, int64_t ld, const int64_t* labels, void* dlogits, int64_t ldd,
                               float* loss, float* workspace, const float* row_stats, int64_t n_slots,
                               int64_t rows, int64_t V, int64_t S, int shift, int64_t ignore_index, void* stream);
/* x *= *device_scalar (no-op when the scalar is 1.0): applies an upstream dloss to dlogits. */
int ct_scale_by_scalar(void* x, int dtype, int64_t n, const float* device_scalar, void* stream);


/* ---- DDP gradient all-reduce over NVLink/NVSwitch peer memory ---------------------------------- *
 * Replaces the ncclAllReduce calls of torch's DDP reducer behind `DDP(model, device_ids=[rank])`
 * (examples/ft_bloom_DDP.py:99,126,135; README.md:46-52). One process per GPU, one node.
 * Two ways to set the symmetric buffer up (csrc/comm.cu):
 *   VMM + multicast (preferred):
 *     ct_comm_vmm_supported  does the device offer cuMemCreate / POSIX-fd export (vmm_ok) and NVLink multicast?
 *     ct_comm_vmm_init   allocate data_bytes (+ a signal page) with cuMemCreate on `device`, map it, return the local
 *                        pointer and a POSIX file descriptor of the allocation for the peers (the caller passes it
 *                        over a Unix socket with SCM_RIGHTS and closes it afterwards)
 *     ct_comm_vmm_connect  peer_fds[world] (entry [rank] ignored): import and map every peer's allocation
 *     ct_comm_mc_create  (rank 0) create the multicast object, return its descriptor for the peers
 *     ct_comm_mc_import  (other ranks) import it
 *     ct_comm_mc_add_device  (all ranks; then a host barrier) join the multicast team
 *     ct_comm_mc_bind    (all ranks; then a host barrier) bind the local allocation and map the multicast address:
 *                        from here ct_allreduce_bucket adds inside the NVSwitch (multimem.ld_reduce / multimem.st)
 *   cudaMalloc + IPC handles (fallback):
 *     ct_comm_init     cudaMalloc a symmetric f32 buffer of data_bytes (+ a signal buffer) on `device`,
 *                      return the local pointer and two 64-byte cudaIpcMemHandle blobs to publish
 *     ct_comm_connect  data_handles / sig_handles: world x 64 bytes, rank-ordered, as gathered by the
 *                      caller over its bootstrap channel (torch.distributed); maps every peer buffer
 *   ct_comm_info       flags[0] = VMM buffers, flags[1] = multicast mapping present
 * Collectives (same order on all ranks, each on the caller's stream; epochs are kept in device memory, so the
 * kernels can be captured in a CUDA graph and replayed):
 *   ct_allreduce_bucket  in place over elements [offset, offset+count) of the symmetric buffer on
 *                    every rank: x = scale * sum_ranks x. mode 0 auto (NVLS when mapped, else two-shot unicast),
 *                    1 one-shot (needs `count` spare floats after the range), 2 NVLS, 3 two-shot unicast
 *                    (reduce-scatter + all-gather by the slice owner). max_ctas bounds the SMs used so
 *                    backward compute keeps running (0 = default: 16 NVLS, 64 otherwise).
 *   ct_broadcast     copy [offset, offset+count) from root's buffer to every rank's buffer.
 *   ct_comm_barrier  all ranks have reached this point of their stream.
 *   ct_embedding_bwd_allranks  sparse half of a TIED embedding gradient (modeling_bloom.py:215-216,
 *                    modeling_gpt.py:205: lm_head.weight is the token table). Every rank has staged, inside
 *                    its symmetric buffer, an int64 token count T at hdr_offset, T x H f32 token gradients
 *                    at rows_offset and T int64 ids at ids_offset (offsets in floats, same on all ranks);
 *                    every rank then adds scale * rows of ALL ranks into its own table gradient [V,H] at
 *                    grad_offset (rows with id == padding_idx or out of range are skipped). Replaces
 *                    "scatter locally, then all-reduce the whole V x H table" — see csrc/comm.cu.
 * The buffers are library-owned (freed by ct_comm_finalize); PyTorch sees them as non-owning tensors. */
int ct_comm_vmm_supported(int device, int* vmm_ok, int* multicast_ok);
int ct_comm_vmm_init(int rank, int world, int device, size_t data_bytes, void** local_data, int* local_fd_out);
int ct_comm_vmm_connect(const int* peer_fds);
int ct_comm_mc_create(int* mc_fd_out);
int ct_comm_mc_import(int mc_fd);
int ct_comm_mc_add_device(void);
int ct_comm_mc_bind(void);
int ct_comm_info(int* flags);
int ct_comm_init(int rank, int world, int device, size_t data_bytes, void** local_data,
                 void* data_handle_out, void* sig_handle_out);
int ct_comm_connect(const void* data_handles, const void* sig_handles);
int ct_allreduce_bucket(int64_t offset, int64_t count, float scale, int mode, int max_ctas,
                        void* stream);
int ct_broadcast(int64_t offset, int64_t count, int root, void* stream);
int ct_comm_barrier(void* stream);
int ct_embedding_bwd_allranks(int64_t hdr_offset, int64_t rows_offset, int64_t ids_offset, int64_t grad_offset,
                              int64_t H, int64_t V, int64_t padding_idx, float scale, int max_ctas,
                              void* stream);
int ct_comm_finalize(void);


/* ---- decode with a preallocated KV cache (generation_util.py:57-119 calls the models with
 * k_v_pasts; modeling_bloom.py:88-92 / modeling_gpt.py:76-80 grow the cache with torch.concat) --- *
 * ct_kv_append: cache[b,h,pos+s,:] = src[b,h,s,:] for s < S_new (16-bit elements; strides in
 * elements as in ct_attn_args); fails if pos + S_new > t_max.
 * ct_attn_decode: ct_attn_fwd's contract specialised for q_len = 1 (or a few) rows against the
 * cache: warp-per-query-row kernel, any head_dim <= 128, no lse2. */
int ct_kv_append(const void* src, int64_t s_sb, int64_t s_sh, int64_t s_ss, void* cache, int64_t c_sb,
                 int64_t c_sh, int64_t c_ss, int B, int H, int S_new, int D, int pos, int t_max,
                 void* stream);
int ct_attn_decode(const ct_attn_args* args, void* stream);
/* ct_kv_append with the destination read from device memory: rows [*len_dev - S_new, *len_dev) (a captured decode
 * step: *len_dev = cache length after this append). Nothing is written when the rows do not fit t_max. */
int ct_kv_append_dev(const void* src, int64_t s_sb, int64_t s_sh, int64_t s_ss, void* cache, int64_t c_sb,
                     int64_t c_sh, int64_t c_ss, int B, int H, int S_new, int D, const int32_t* len_dev, int t_max,
                     void* stream);
/* One step of GenerationMixin._greedy_search after the model call (generation_util.py:86-101, do_sample = False) on
 * the device: next = argmax(logits[b,:]) (first maximum); next = next*alive + pad*(1-alive); alive &= next not in
 * end_ids; ids_out[b, out_pos] = next; cur_ids[b] = next; pos_ids[b] += 1 (nullable); then once: seq_len += 1,
 * out_pos += 1, done_at = out_pos when no row is alive. state = int32[5] on the device: {seq_len, out_pos, alive rows,
 * done_at (-1), 0}. logits [B, V] with row stride ld, dtype CT_F32 / CT_BF16 / CT_F16.
 * sampled (nullable, int64 [B]): the tokens the caller drew itself (do_sample = True, generation_util.py:78-84:
 * temperature / top-k / top-p / multinomial) — they replace the argmax, the bookkeeping is the same. */
int ct_greedy_step(const void* logits, int logits_dtype, int64_t ld, int64_t B, int64_t V, int64_t* alive,
                   const int64_t* end_ids, int n_end, int64_t pad_id, int64_t* ids_out, int64_t out_stride,
                   int64_t* cur_ids, int64_t* pos_ids, int32_t* state, const int64_t* sampled, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CT_B200_H_ */
