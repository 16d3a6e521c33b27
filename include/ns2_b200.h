/* ns2_b200.h — C ABI of libns2b200.so: the sm_100a kernels behind the NaturalSpeech2 denoiser hot path.
 *
 * The reference (lucidrains/naturalspeech2-pytorch @ 659bec7) has no FFI of its own; its boundary for this
 * path is Python (nn.Module classes).  Each entry point below replaces the PyTorch library calls the
 * reference makes at the cited lines (paths relative to the reference repo; ns2.py =
 * naturalspeech2_pytorch/naturalspeech2_pytorch.py).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates device
 *    memory, never synchronises, and is re-entrant;
 *  - return value 0 = success, negative = error (ns2_last_error() gives a thread-local message);
 *  - activations are token-major (batch, position, channel) with the channel contiguous;
 *  - "bf16" buffers hold IEEE bfloat16, "f32" IEEE binary32, codes are int64 (as in the reference).
 */
#ifndef NS2_B200_H_
#define NS2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NS2_ABI_VERSION 3

typedef void* ns2_stream_t; /* cudaStream_t */

const char* ns2_last_error(void);
int ns2_abi_version(void);
/* Size the persistent grids of every kernel for at most `sms` SMs (rounded down to an even count; 0 = all SMs, the
 * default).  The GEMM / attention kernels run one CTA (pair) per SM for the whole launch, so a concurrent kernel that
 * holds a few SMs - NCCL's all-reduce of the gradients during the backward pass (ns2.py:1723-1726, 1886) - would delay
 * whole CTAs by a full tile loop; leaving those SMs out of the grid avoids that.  Returns the previous limit. */
int ns2_set_sm_limit(int sms);

/* ------------------------------------------------------------------------------------------------
 * 1. Segmented tcgen05 GEMM with fused epilogues.
 *
 * Computes, for every group g, batch b, position n and output channel j
 *     acc_a[b,n,j] = sum over segments s with segs[s].acc == a, k < segs[s].k_len of
 *                    A[b, n - segs[s].shift_units*dil[g], g*a_group_col_stride + segs[s].a_col_off + k]
 *                  * B[g*b_group_row_stride + j, segs[s].b_col_off + k]
 * where rows of A with a negative (or >= a_rows) position read as zero — this is the causal left padding
 * of CausalConv1d (ns2.py:583-595): a k=3 dilated causal conv is three segments with shift_units 2,1,0.
 * A Linear layer (nn.Linear: ns2.py:1021,1024,1051-1053,783) is one segment with shift 0.  A "same"-padded Conv1d
 * (kernel 2p+1, padding p: SpeechPromptEncoder ns2.py:316-320) is 2p+1 segments with shift_units p, p-1, ..., -p.
 *
 * Epilogues (all accumulate in fp32):
 *   NS2_EPI_BF16     out_bf16 = acc0 + bias
 *   NS2_EPI_F32      out_f32  = acc0 + bias (+ resid_f32)           residual add of ns2.py:799,805,809
 *   NS2_EPI_GEGLU    out_bf16[j] = gelu_erf(acc0[gate j] + bias) * (acc0[val j] + bias)   ns2.py:1004-1007;
 *                    B rows must be packed so that each 256-row tile holds 128 value rows followed by
 *                    the 128 matching gate rows (see pack_geglu_weight in the Python host code);
 *                    `n` counts packed B rows (2x the number of output channels)
 *   NS2_EPI_WAVENET  y = (acc0 + bias)*gamma[b,j] + beta[b,j]; out_bf16 = tanh(y)*sigmoid(y) + acc1 + bias1
 *                    the WavenetResBlock body ns2.py:619-636 (acc0 = dilated conv, acc1 = res_conv);
 *                    gamma = film[b*film_batch_stride + g*film_group_stride + j], beta = gamma + n;
 *                    bias1 = bias + bias1_off.
 * All k_len must be multiples of 64 unless the segment ends at the last column of A and B.
 * ------------------------------------------------------------------------------------------------ */
enum { NS2_EPI_BF16 = 0, NS2_EPI_F32 = 1, NS2_EPI_GEGLU = 2, NS2_EPI_WAVENET = 3 };
#define NS2_GEMM_MAX_SEGS 12 /* a k=9 convolution (SpeechPromptEncoder, ns2.py:316-320) is 9 segments */
#define NS2_GEMM_MAX_GROUPS 8

typedef struct ns2_gemm_seg {
  int32_t a_col_off;   /* first A column of this segment (within the group's column window) */
  int32_t b_col_off;   /* first B column (K index in the packed weight) */
  int32_t k_len;       /* reduction length */
  int32_t shift_units; /* row shift = shift_units * dil[group] */
  int32_t acc;         /* accumulator id, 0 or 1 */
} ns2_gemm_seg;

typedef struct ns2_gemm_args {
  const void* A;           /* bf16 (a_batches, a_rows, a_cols) */
  int64_t a_row_stride;    /* elements */
  int64_t a_batch_stride;  /* elements */
  int32_t a_batches, a_rows, a_cols;
  const void* B;           /* bf16 packed weights (b_rows, b_cols), K contiguous */
  int64_t b_row_stride;    /* elements */
  int32_t b_rows, b_cols;
  int32_t n;               /* B rows (accumulator columns) per group */
  int32_t groups;          /* >= 1; independent problems sharing the launch */
  int32_t a_group_col_stride;
  int32_t b_group_row_stride;
  int32_t out_group_col_stride;
  int32_t dil[NS2_GEMM_MAX_GROUPS];
  int32_t num_segs;
  ns2_gemm_seg segs[NS2_GEMM_MAX_SEGS];
  int32_t epilogue;
  const float* bias;       /* indexed like B rows (group offset = g*b_group_row_stride); may be NULL */
  int32_t bias1_off;       /* WAVENET: offset of the res_conv bias inside `bias` */
  void* out;               /* bf16 or f32, row (b*a_rows + n) */
  int64_t out_row_stride;  /* elements */
  const float* resid;      /* F32 epilogue: optional residual, same row indexing as out */
  int64_t resid_row_stride;
  const float* film;       /* WAVENET: FiLM table */
  int64_t film_batch_stride;
  int32_t film_group_stride;
  int32_t flags;           /* 0, or NS2_GEMM_FLAG_* */
  void* debug_timeline;    /* bring-up aid, normally NULL: device buffer of 64*8 int64 receiving clock64 stamps of CTA
                              pair 0 of the CTA-pair kernel (tools/gemm_timeline.py) */
} ns2_gemm_args;

#define NS2_GEMM_FLAG_SKIP_EPILOGUE 1  /* measurement aid: run the TMA/MMA mainloop only, write nothing (CTA-pair kernel) */
#define NS2_GEMM_FLAG_SILU 4 /* BF16 / F32 epilogues: out = silu(acc + bias) (+ resid) — Conv1d + nn.SiLU of the prompt
                               encoder (ns2.py:316-320) and CausalConv1d + SiLU of the phoneme encoder (ns2.py:255-257) */
#define NS2_GEMM_FLAG_NARROW_LAST 8 /* tuning / A-B tests (CTA-pair kernel, n % 256 != 0): schedule the partial-width n-tiles
                                      after all full-width ones instead of n-fastest */
#define NS2_GEMM_FLAG_WAVENET_ONE_PASS 2 /* tuning / A-B tests: WAVENET with two 256-column accumulators and a single epilogue pass */

int ns2_gemm(const ns2_gemm_args* args, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 1b. Weight gradient of a Linear / CausalConv1d tap (autograd's grad_weight = grad_output^T @ input; the reference
 *     reaches it through loss.backward(), README.md:63, ns2.py:1886):
 *        dW[g][n, k] += sum_{b, m} dY[b, m, g*dy_group_col_stride + n] * X[b, m - shift_units*dil[g], g*x_group_col_stride + x_col_off + k]
 *     dY / X: bf16 token-major activations (batches, rows, cols); dW: fp32, ACCUMULATED into (reduce-add), row stride
 *     dw_row_stride, group g at row g*dw_group_row_stride.  n, k multiples of 32 (n a multiple of 128 when groups > 1).
 *     Positions outside [0, rows) of a shifted tap contribute zero (the conv's causal padding).  splits = 0: automatic.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ns2_wgrad_args {
  const void* dY; int64_t dy_row_stride, dy_batch_stride; int32_t dy_cols;
  const void* X;  int64_t x_row_stride, x_batch_stride;  int32_t x_cols;
  int32_t batches, rows;
  int32_t n, k;
  int32_t groups, dy_group_col_stride, x_group_col_stride, x_col_off;
  int32_t dil[NS2_GEMM_MAX_GROUPS];
  int32_t shift_units;
  float* dW; int64_t dw_row_stride; int32_t dw_group_row_stride;
  int32_t splits;
} ns2_wgrad_args;

int ns2_wgrad(const ns2_wgrad_args* args, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 2. Non-causal, unmasked flash attention forward (Attend.forward, attend.py:112-155 with mask=None,
 *    causal=False, dropout=0 — the only configuration the hot path uses).
 *    q/k/v: bf16, head h lives in columns [h*64, h*64+64) of each row; dim_head must be 64.
 *    out[b, i, h*64:(h+1)*64] = softmax_j(q_i . k_j * scale) @ v   (bf16)
 * ------------------------------------------------------------------------------------------------ */
typedef struct ns2_attn_args {
  const void* q; int64_t q_row_stride, q_batch_stride;
  const void* k; int64_t k_row_stride, k_batch_stride;
  const void* v; int64_t v_row_stride, v_batch_stride;
  void* out;     int64_t o_row_stride, o_batch_stride;
  int32_t batches, heads, q_len, kv_len, dim_head;
  float scale;
  int32_t kernel;   /* NS2_ATTN_AUTO, or force one implementation (tests / tuning) */
  void* debug_timeline; /* bring-up aid, normally NULL: device buffer of 64*16 int64 receiving clock64 stamps of CTA 0
                           of the two-tile kernel (tools/attn_timeline.py) */
  float* lse;           /* optional (batches, heads, q_len) f32: log2-domain log-sum-exp of the scaled score rows,
                           saved for ns2_attn_bwd */
} ns2_attn_args;

#define NS2_ATTN_AUTO 0            /* two-tile kernel when q_len > 128 and kv_len > 64, else one-tile */
#define NS2_ATTN_ONE_TILE 1        /* 128 queries x 64-key tiles per CTA, P staged in shared memory */
#define NS2_ATTN_TWO_TILE 2        /* persistent, 2 x 128 queries x 128-key tiles, P and O in tensor memory */
#define NS2_ATTN_TWO_TILE_POLY2 3  /* same, 2 of every 8 exponentials on the FMA pipe (degree-3 polynomial) */
#define NS2_ATTN_TWO_TILE_POLY4 4  /* same, 4 of every 8 */
#define NS2_ATTN_TWO_TILE_LOCKSTEP 5 /* two-tile without the exponential-section turn taking (A/B measurement) */

int ns2_attn_fwd(const ns2_attn_args* args, ns2_stream_t stream);

/* Backward of the above (autograd of F.scaled_dot_product_attention, reached from loss.backward(), ns2.py:1886):
 *   dq_accum (batches, q_len, heads*64) f32, contiguous, must be ZERO on entry (every key tile adds its share);
 *   dk / dv: bf16, same layout conventions as k / v;  lse from ns2_attn_fwd;  delta: scratch (batches, heads, q_len) f32. */
typedef struct ns2_attn_bwd_args {
  const void* q; int64_t q_row_stride, q_batch_stride;
  const void* k; int64_t k_row_stride, k_batch_stride;
  const void* v; int64_t v_row_stride, v_batch_stride;
  const void* o; int64_t o_row_stride, o_batch_stride;
  const void* d_o; int64_t do_row_stride, do_batch_stride;
  const float* lse;
  float* delta;
  float* dq_accum;
  void* dk; int64_t dk_row_stride, dk_batch_stride;
  void* dv; int64_t dv_row_stride, dv_batch_stride;
  int32_t batches, heads, q_len, kv_len, dim_head;
  float scale;
} ns2_attn_bwd_args;

int ns2_attn_bwd(const ns2_attn_bwd_args* args, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3. RMSNorm (+ learned gamma) (+ FiLM) : RMSNorm.forward ns2.py:736-746.
 *    out_bf16[r, :] = x[r,:] / max(||x[r,:]||_2, 1e-12) * sqrt(dim) * gamma * film_gamma[b] + film_beta[b]
 *    gamma may be NULL (=1); film may be NULL (no FiLM); b = r / rows_per_batch;
 *    film_gamma = film + b*film_batch_stride, film_beta = film_gamma + dim.
 * ------------------------------------------------------------------------------------------------ */
int ns2_rmsnorm_film(const float* x, int64_t x_row_stride, int64_t rows, int32_t dim,
                     int32_t rows_per_batch, const float* gamma, const float* film,
                     int64_t film_batch_stride, void* out_bf16, int64_t out_row_stride,
                     ns2_stream_t stream);

/* Same, fp32 output (PerceiverResampler.norm, ns2.py:566,579). */
int ns2_rmsnorm_f32(const float* x, int64_t x_row_stride, int64_t rows, int32_t dim,
                    const float* gamma, float* out, int64_t out_row_stride, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 4. Small dense layers on the conditioning vector (M <= 64 rows), fp32 end to end.
 *    ns2_time_cond : LearnedSinusoidalPosEmb + Linear + SiLU (ns2.py:108-120, 839-843)
 *        out[b, :] = silu(W @ [t_b, sin(2 pi t_b w), cos(2 pi t_b w)] + bias),  W is (n_out, 2*half+1)
 *    ns2_small_linear : out[b,:] = act(W @ x[b,:] + bias), act 0 = none, 1 = SiLU (ns2.py:858-862)
 * ------------------------------------------------------------------------------------------------ */
int ns2_time_cond(const float* times, int32_t batch, const float* freqs, int32_t half_dim,
                  const float* W, const float* bias, int32_t n_out, float* out,
                  int64_t out_row_stride, ns2_stream_t stream);
int ns2_small_linear(const float* x, int64_t x_row_stride, int32_t batch, int32_t k, const float* W,
                     const float* bias, int32_t n_out, int32_t act, float* out,
                     int64_t out_row_stride, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 5. Layout / cast helpers (what `rearrange` + autocast do in the reference, ns2.py:972-997).
 *    ns2_cast_bf16        : out_bf16 = (bf16) (x [+ add]) ; add is broadcast over nothing (same shape) or NULL
 *    ns2_mean_rows        : out[b,:] = mean over n of x[b,n,:]   (Reduce('b n d -> b d','mean'), ns2.py:859)
 *    ns2_transpose_cast   : (B, C, L) f32 channel-first -> (B, L, C) bf16 token-major
 * ------------------------------------------------------------------------------------------------ */
int ns2_cast_bf16(const float* x, const float* add, int64_t count, void* out_bf16,
                  ns2_stream_t stream);
int ns2_mean_rows(const float* x, int32_t batch, int32_t n, int32_t dim, float* out,
                  ns2_stream_t stream);
/*    ns2_cond_inject      : out_bf16[b,n,:] = bf16(x[b,n,:] + c), c = 0 for n >= cond_len (zero padding, ns2.py:70-77),
 *                           null_cond[:] where drop_mask[b] (uint8, may be NULL = keep all), else cproj[b,n,:]
 *                           (cproj: projected aligned condition, token-major (batch, cond_len, dim) f32; ns2.py:978-992)
 *    ns2_select_rows      : out[b,:] = drop_mask[b] ? null_row[:] : src[b,:]  (f32 or bf16 out; ns2.py:954-968) */
int ns2_cond_inject(const float* x, const float* cproj, const uint8_t* drop_mask, const float* null_cond,
                    int32_t batch, int32_t n, int32_t cond_len, int32_t dim, void* out_bf16, ns2_stream_t stream);
int ns2_select_rows(const uint8_t* drop_mask, const float* null_row, const float* src, int64_t src_row_stride,
                    int32_t batch, int32_t row_len, void* out, int64_t out_row_stride, int32_t out_bf16,
                    ns2_stream_t stream);
int ns2_transpose_cast(const float* x, int32_t batch, int32_t channels, int32_t length,
                       void* out_bf16, ns2_stream_t stream);
/*    ns2_embedding_bf16   : out_bf16[r,:] = bf16(table[ids[r] < 0 ? pad_id : ids[r], :]) — nn.Embedding of the phoneme
 *                           encoder with its padding substitution (ns2.py:253, 279-282); ids int64, table f32 */
int ns2_embedding_bf16(const int64_t* ids, int64_t rows, const float* table, int32_t num_rows, int32_t dim,
                       int32_t pad_id, void* out_bf16, ns2_stream_t stream);

/*    ns2_groupnorm_silu   : y = silu(GroupNorm(groups, channels)(x)) (+ resid) on token-major f32 x (batch, rows, channels):
 *                           statistics per (batch element, group) over rows x channels/groups values, biased variance,
 *                           eps inside the square root, per-channel affine (nn.GroupNorm) - Block.forward of the
 *                           duration / pitch predictor (ns2.py:345-365) with the ResnetBlock residual (ns2.py:399-401).
 *                           Writes out_f32 and/or out_bf16 (either may be NULL).
 *    ns2_rowdot           : out[r] = (relu ?) max(0, .) : (.) of dot(x[r,:], w) + bias[0] - Linear(dim, 1) + ReLU heads
 *                           (ns2.py:452-456) */
int ns2_groupnorm_silu(const float* x, int32_t batch, int32_t rows, int32_t channels, int32_t groups,
                       const float* weight, const float* bias, float eps, const float* resid, float* out_f32,
                       void* out_bf16, ns2_stream_t stream);
int ns2_rowdot(const float* x, int64_t rows, int32_t dim, const float* w, const float* bias, int32_t relu, float* out,
               ns2_stream_t stream);
/*    ns2_expand_encodings : length regulation, NaturalSpeech2.expand_encodings (ns2.py:1449-1455) with the hard alignment
 *                           given as one text index per frame: out[b, d, n] = phon[b, m, d] + pitch_table[coarse[b, m], d],
 *                           m = idx[b, n] (int32; negative = frame past the sample's length -> 0).  phon f32 (batch, t_text,
 *                           dim) token-major, coarse int32 (batch, t_text) = f0_to_coarse bins, out f32 (batch, dim, length)
 *                           channel-first (the `cond` layout of Model.forward, ns2.py:929-937). */
int ns2_expand_encodings(const float* phon, const int32_t* coarse, const float* pitch_table, int32_t table_rows,
                         const int32_t* idx, int32_t batch, int32_t t_text, int32_t dim, int32_t length, float* out,
                         ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 6. Diffusion element-wise steps, fp32 (NaturalSpeech2.forward ns2.py:1621-1666; ddim_sample 1392-1429).
 *    All take per-sample scalars as device arrays of length `batch`; `per_sample` = N*D elements.
 *    `objective` selects the parameterisation (ns2.py:1637-1644, 1412-1421): NS2_OBJ_V / NS2_OBJ_EPS / NS2_OBJ_X0.
 *    ns2_q_sample   : x_t = alpha*x0 + sigma*noise ; target = alpha*noise - sigma*x0 (v) | noise (eps) | x0 (x0)
 *    ns2_mse_rows   : out[b] = mean((pred-target)^2) over the sample          (ns2.py:1646-1647);
 *                     deterministic two-level reduction through caller-provided scratch; optionally
 *                     also the batch mean of those per-sample values (one more tiny launch)
 *    ns2_ddim_step  : x0 = alpha*x - sigma*out (v) | (x - sigma*out)/max(alpha,1e-10) (eps) | out (x0) ;
 *                     eps = (x - alpha*x0)/max(sigma,1e-10) ;
 *                     x <- x0*alpha_next + eps*sigma_next                     (ns2.py:1420-1429)
 *    ns2_cfg_combine: out = null + (cond - null)*scale                        (ns2.py:927)
 * ------------------------------------------------------------------------------------------------ */
#define NS2_MSE_SCRATCH_PER_SAMPLE 64
#define NS2_OBJ_V 0
#define NS2_OBJ_EPS 1
#define NS2_OBJ_X0 2
int ns2_q_sample(const float* x0, const float* noise, const float* alpha, const float* sigma,
                 int32_t batch, int64_t per_sample, float* x_t, float* target, int32_t objective,
                 ns2_stream_t stream);
int ns2_mse_rows(const float* pred, const float* target, int32_t batch, int64_t per_sample,
                 float* scratch /* batch * NS2_MSE_SCRATCH_PER_SAMPLE floats */, float* out,
                 float* mean_out /* optional: mean over the batch of out[], ns2.py:1666 */,
                 ns2_stream_t stream);
int ns2_ddim_step(float* x, const float* v, const float* alpha, const float* sigma,
                  const float* alpha_next, const float* sigma_next, int32_t batch,
                  int64_t per_sample, int32_t objective, ns2_stream_t stream);
int ns2_cfg_combine(const float* cond, const float* null_, float scale, int64_t count, float* out,
                    ns2_stream_t stream);
/*    ns2_x_start    : x_start implied by the model output `pred` (ns2.py:1673-1680): alpha*x - sigma*pred (v) |
 *                     (x - sigma*pred)/max(alpha,1e-10) (eps) | pred (x0) */
int ns2_x_start(const float* x, const float* pred, const float* alpha, const float* sigma, int32_t batch,
                int64_t per_sample, float* out, int32_t objective, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 7. Residual vector quantisation (Encodec RVQ encode/decode; third-party code reached from
 *    ns2.py:1445,1611 via audiolm_pytorch.EncodecWrapper -> encodec ResidualVectorQuantizer).
 *    ns2_rvq_prepare : codebooks f32 (Q, K, d) -> cb_f16: NS2_RVQ_PREPARED_HALFS(Q, K, d) fp16 values = the copy
 *                      (Q, K, d) scaled by 2^-e_q followed by the (Q, K, 16) norm blocks (||c||^2 as an fp16 hi/lo pair,
 *                      laid out for the tensor core); ||c||^2 f32 (Q, K); meta f32 (Q, 2) = {max_k ||c_k||, 2^e_q}
 *    ns2_rvq_encode  : frames f32 (F, d) -> codes int64 (F, Q); residual chain in fp32, nearest
 *                      codeword by exact squared L2 distance, ties -> lowest index.  d must be 128,
 *                      K a multiple of 128.
 *    ns2_rvq_decode  : emb f32 (F, d) = sum_q codebooks[q, codes[f,q], :]  (summed in order q = 0..Q-1)
 * ------------------------------------------------------------------------------------------------ */
#define NS2_RVQ_PREPARED_HALFS(q, k, d) ((long long)(q) * (k) * ((d) + 16))
#define NS2_RVQ_STATS_LEN 260  /* 4 counters + 32 stages x 8 clock64 stamps of CTA 0 (bring-up timeline) */
int ns2_rvq_prepare(const float* codebooks, int32_t q, int32_t k, int32_t d, void* cb_f16,
                    float* cb_norm2, float* cb_meta, ns2_stream_t stream);
int ns2_rvq_encode(const float* frames, int64_t num_frames, int32_t d, const float* codebooks,
                   const void* cb_f16, const float* cb_norm2, const float* cb_meta, int32_t q,
                   int32_t k, int64_t* codes,
                   int64_t* stats /* optional NS2_RVQ_STATS_LEN int64 counters, accumulated with atomics:
                                     {lookups, near-ties re-scored, full scans, sub-chunk scans} */,
                   ns2_stream_t stream);
int ns2_rvq_decode(const int64_t* codes, int64_t num_frames, int32_t q, int32_t k, int32_t d,
                   const float* codebooks, float* emb, ns2_stream_t stream);
/*    ns2_rvq_ce      : cross-entropy head of the residual VQ, `codec.rq(x_start, codes)` (ns2.py:1670-1684;
 *                      vector-quantize-pytorch ResidualVQ.forward(x, indices)).  Per stage the logits are the negative
 *                      Euclidean distances -||r_q - c_k||; loss = sum_q mean_{f: target != -1} CE(logits, target[f,q]);
 *                      the residual chain follows `own_codes` (the codec's own nearest codewords, from ns2_rvq_encode).
 *                      ce_scratch: num_frames * q floats; loss: 1 float. */
int ns2_rvq_ce(const float* frames, int64_t num_frames, int32_t d, const float* codebooks,
               const float* cb_norm2, int32_t q, int32_t k, const int64_t* own_codes,
               const int64_t* target_codes, float* ce_scratch, float* loss, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 8. Backward pass (what autograd runs for the reference's loss.backward(), README.md:63, ns2.py:1886).
 *    Matrix products: ns2_gemm with transposed weight packs (dgrad; negative shift_units = anti-causal taps) and
 *    ns2_wgrad.  Between them:
 *    ns2_rmsnorm_film_bwd : backward of RMSNorm (+learned gamma | +FiLM), ns2.py:736-746.  dxr (fp32 residual-stream
 *                           gradient) += dx IN PLACE, its bf16 copy is written to dxr_bf16; dfilm[b, :dim] += d(gamma_b),
 *                           dfilm[b, dim:2dim] += d(beta_b) (atomics); dgamma[:] += d(learned gamma)
 *    ns2_geglu_bwd        : pre (rows, 2*dp) bf16 in the packed [128 value | 128 gate] tile layout is OVERWRITTEN by its
 *                           gradient given dg (rows, dp) bf16   (GEGLU, ns2.py:1004-1007)
 *    ns2_wavenet_gate_bwd : dz of y = tanh(z) sigmoid(z) + res, z = c*gamma_b + beta_b (ns2.py:625-630): dc = dz*gamma_b,
 *                           dfilm += [sum dz*c | sum dz] per batch / group
 *    ns2_colsum_bf16      : out[c] += sum_r t[r, c]            (bias gradients)
 *    ns2_group_sum_bf16   : out[r, c] = sum_g t[r, g*dim + c]  (gradient of an input shared by all dilation columns)
 *    ns2_mse_bwd          : out = coef[b] * (pred - target), bf16 and/or f32   (seed of the backward pass, ns2.py:1646-1666)
 *    ns2_film_wgrad       : dw[r, c] (+)= sum_b dfilm[b, r] * t[b, c]  (FiLM projection weights; batch <= 32 per call)
 *    ns2_attn_bwd         : flash-attention backward (dq, dk, dv) from (q, k, v, o, lse, do)
 * ------------------------------------------------------------------------------------------------ */
int ns2_rmsnorm_film_bwd(const float* x, const void* dh_bf16, int64_t rows, int32_t dim, int32_t rows_per_batch,
                         const float* gamma, const float* film, int64_t film_batch_stride, float* dfilm,
                         int64_t dfilm_batch_stride, float* dgamma, float* dxr, void* dxr_bf16, ns2_stream_t stream);
int ns2_geglu_bwd(void* pre_bf16, const void* dg_bf16, int64_t rows, int32_t dp, ns2_stream_t stream);
int ns2_wavenet_gate_bwd(const void* c_bf16, int64_t c_row_stride, const void* dy_bf16, int64_t dy_row_stride,
                         void* dc_bf16, int64_t dc_row_stride, int32_t batches, int32_t rows_per_batch, int32_t dim,
                         int32_t groups, const float* film, int64_t film_batch_stride, int32_t film_group_stride,
                         float* dfilm, int64_t dfilm_batch_stride, ns2_stream_t stream);
int ns2_colsum_bf16(const void* t_bf16, int64_t rows, int32_t cols, int64_t row_stride, float* out, ns2_stream_t stream);
int ns2_group_sum_bf16(const void* t_bf16, int64_t rows, int32_t dim, int32_t groups, void* out_bf16, ns2_stream_t stream);
int ns2_mse_bwd(const float* pred, const float* target, const float* coef, int32_t batch, int64_t per_sample,
                void* out_bf16 /* optional */, float* out_f32 /* optional */, ns2_stream_t stream);
int ns2_film_wgrad(const float* dfilm, int64_t dfilm_batch_stride /* elements between batch rows of dfilm (>= rows): a
                   column window of the stacked FiLM gradient can be reduced as soon as its layer is final */,
                   const float* t, int32_t batch, int64_t rows, int32_t cols, float* dw,
                   int32_t accumulate /* 0: dw = ..., dw need not be initialised; 1: dw += ... */, ns2_stream_t stream);
/*    ns2_accum_bf16       : acc (f32) += t (bf16); acc_bf16 (optional) = bf16(acc)   (joins a branch gradient) */
int ns2_accum_bf16(float* acc, const void* t_bf16, int64_t count, void* acc_bf16, ns2_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 9. Monotonic alignment search: `maximum_path(value, mask)` of naturalspeech2_pytorch/aligner.py:88-122 (called from
 *    Aligner.forward aligner.py:214, reached from NaturalSpeech2.forward ns2.py:1578 in conditional training).
 *    value, mask: f32 (batch, t_x, t_y) contiguous (t_x text positions <= 1024, t_y mel frames); mask holds 0/1.
 *    Viterbi recursion over the frames with the reference's exact fp32 operations and tie rule, then the backtrack:
 *      idx[b, j]     (int32, batch x t_y)  = text position aligned to frame j
 *      path[b, i, j] (f32, optional)       = (idx[b, j] == i) * mask[b, i, j]      — bit-identical to the reference
 *    neg_const = the reference's `const` (default -inf).  workspace: ns2_maximum_path_workspace_bytes() bytes of
 *    scratch (1-bit decisions, batch x t_y x 128 B).
 * ------------------------------------------------------------------------------------------------ */
int64_t ns2_maximum_path_workspace_bytes(int32_t batch, int32_t t_x, int32_t t_y);
int ns2_maximum_path(const float* value, const float* mask, int32_t batch, int32_t t_x, int32_t t_y, float neg_const,
                     void* workspace, int64_t workspace_bytes, int32_t* idx, float* path, ns2_stream_t stream);

/* Number of kernel launches issued through this library since load (for bench.py's gpu_launches). */
int64_t ns2_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NS2_B200_H_ */
