/* rtts_b200.h - C ABI of libreformer_b200.so: the sm_100a kernels behind Reformer-TTS's LSH-attention /
 * reversible / chunked-FFN hot path.
 *
 * The reference has no native boundary for this path (pure Python); the functions
 * below are what a binding for it would call - each one replaces the run of ATen library launches the
 * cited reference lines issue.  `ref:` = kowaalczyk/reformer-tts, `rp:` = reformer-pytorch 0.19.1
 * (rows R1-R11 of SURVEY.md 8(a)), `hf:` = transformers modeling_reformer.py (5.5.0 line numbers).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the library never allocates,
 *    frees or retains memory: outputs and workspaces are caller-owned.
 *  - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it.
 *  - return 0 on success, negative on error; rtts_last_error() gives a thread-local message.
 *  - activations: bf16 row-major [B, T, H*dh] ("token-major", heads side by side, leading
 *    dimension ld in elements); integer tensors int32; accumulators / statistics fp32.
 *  - no global mutable state; re-entrant; one CUDA context per process.
 */
#ifndef RTTS_B200_H
#define RTTS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTTS_KEYNORM_L2 0  /* rp R5: x / max(|x|_2, 1e-12)                     */
#define RTTS_KEYNORM_RMS 1 /* hf:1042-1056: x * rsqrt(mean(x^2)+1e-6) / sqrt(dh) */
#define RTTS_MASK_QUERY_AND_KEY 0 /* rp R8 */
#define RTTS_MASK_KEY_ONLY 1      /* hf:914-922 */

/* Constants that differ between the two LSH implementations the reference can select
 * (ref:reformer_tts/model/reformer.py:198-213). */
typedef struct rtts_lsh_spec {
  float score_scale; /* rp R7: dh^-0.5 ; hf: 1 */
  int32_t key_norm;  /* RTTS_KEYNORM_* */
  float mask_value;  /* rp R8: -FLT_MAX ; hf:431: -1e9 */
  float self_value;  /* rp R8: -5e4     ; hf:430: -1e5 */
  int32_t mask_mode; /* RTTS_MASK_* */
  int32_t causal;    /* ref:reformer_tts/model/reformer.py:68,117 */
} rtts_lsh_spec;

const char* rtts_last_error(void);
int rtts_abi_version(void);
/* Hash of the sources this binary was built from (csrc/build.py source_hash); "unknown" for a build outside build.py. */
const char* rtts_build_id(void);

/* ---- LSH bucketing --------------------------------------------------------------------------- */

/* Random-rotation hash + argmax + round offsets (rp R2; hf:717-758; replaces randn-einsum-cat-argmax).
 * qk bf16 [B,T,H*dh] (ld elements per token), rot fp32 [rot_heads, dh, R, n_buckets/2] with rot_heads
 * = 1 (rp: shared) or H (hf: per head).  pad_mask uint8 [B,T] (1 = real token) or NULL; when
 * use_pad_bucket != 0 padded tokens get bucket n_buckets and the round stride is n_buckets+1 (hf:740-747).
 * buckets int32 [B,H,R*T] (value = round*stride + id).  Arithmetic: fp32 FMA on the bf16 inputs.
 * sumsq (nullable) fp32 [B,H,T] receives |qk row|^2, the input of the key normalisation (rp R5 / hf:1042-1056). */
int rtts_lsh_hash(const void* qk, int64_t ld, const float* rot, int rot_heads, const uint8_t* pad_mask,
                  int use_pad_bucket, int32_t* buckets, float* sumsq, int B, int T, int H, int dh, int R,
                  int n_buckets, void* stream);

/* The same hash on the tensor pipe (tcgen05): each fp32 rotation is split into three bf16 parts whose sum is the fp32 value, so every
 * product is formed exactly and only the fp32 accumulation rounds (as in the reference's fp32 einsum).  Same inputs / outputs as
 * rtts_lsh_hash plus `workspace` (device, rtts_lsh_hash_tc_workspace_bytes bytes, 16-byte aligned, overwritten on every call).
 * Supported when rtts_lsh_hash_tc_supported() != 0: dh = 64, T % 128 == 0, n_buckets / 2 <= 256, and the projections of one launch
 * (all R rounds when R * n_buckets / 2 <= 256, else the largest group of whole rounds that fits: the 16k-token sweep hashes its 4
 * rounds of 128 projections in two launches) a multiple of 16; callers fall back to rtts_lsh_hash otherwise. */
int rtts_lsh_hash_tc(const void* qk, int64_t ld, const float* rot, int rot_heads, const uint8_t* pad_mask,
                     int use_pad_bucket, int32_t* buckets, float* sumsq, void* workspace, int B, int T, int H, int dh,
                     int R, int n_buckets, void* stream);
int64_t rtts_lsh_hash_tc_workspace_bytes(int rot_heads, int R, int n_buckets);
int rtts_lsh_hash_tc_supported(int T, int dh, int R, int n_buckets);

/* sumsq fp32 [B,H,T] = |qk[b,t,h,:]|^2 on its own (same values rtts_lsh_hash emits). */
int rtts_lsh_sumsq(const void* qk, int64_t ld, float* sumsq, int B, int T, int H, int dh, void* stream);

/* Stable sort by (bucket, position) per (batch*head) row, and its inverse (rp R3; hf:150-156,762-779).
 * sticker[i] = round*T + pos sitting at sorted slot i; undo[sticker[i]] = i.  ids_per_round = round
 * stride used by rtts_lsh_hash (n_buckets or n_buckets+1).  Bit-exact with torch.sort of T*bucket+pos. */
int rtts_lsh_sort(const int32_t* buckets, int32_t* sticker, int32_t* undo, int rows, int T, int R,
                  int ids_per_round, void* stream);

/* ---- chunked shared-QK attention ------------------------------------------------------------- */

/* Gather by sticker, key normalisation, look-one-back, QK^T, masks, softmax, PV and un-sort in one
 * kernel (rp R4-R10; hf:563-599,801-906,1067-1096).  bucket in {64,128}, dh = 64, T % 128 == 0.
 * sumsq fp32 [B,H,T] = |qk row|^2.  mask uint8 [B,T] (1 = real token) or NULL.  Outputs, already UNSORTED:
 *   o_rounds bf16 [B,H,R,T,dh], lse_rounds fp32 [B,H,R,T]. */
int rtts_lsh_attn_fwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const float* sumsq, const uint8_t* mask,
                      const rtts_lsh_spec* spec_host, void* o_rounds, float* lse_rounds, int B, int T, int H,
                      int dh, int R, int bucket, void* stream);

/* Combine hash rounds (rp R11; hf:626-645): out bf16 [B,T,H*dh] (ld_out), lse fp32 [B,H,T] = logsumexp_r. */
int rtts_lsh_merge_fwd(const void* o_rounds, const float* lse_rounds, void* out, int64_t ld_out, float* lse,
                       int B, int T, int H, int dh, int R, void* stream);

/* delta[b,h,t] = <dout[b,t,h,:], out[b,t,h,:]> (fp32), the per-query term of the attention backward. */
int rtts_lsh_delta(const void* dout, const void* out, int64_t ld, float* delta, int B, int T, int H, int dh,
                   void* stream);

/* Backward of rtts_lsh_attn_fwd + rtts_lsh_merge_fwd with in-kernel recompute of the scores
 * (autograd of rp R4-R11; the per-round o / lse of the forward are not needed).  Inputs: qk, v, sticker, sumsq,
 * mask as in forward; dout bf16 [B,T,H*dh] (ld_dout) = gradient of the merged output; lse [B,H,T] from
 * rtts_lsh_merge_fwd; delta [B,H,T] from rtts_lsh_delta.  Outputs, bf16 [B,H,R,T,dh], scattered to the
 * UNSORTED slot like the forward:
 *   dqk_main  gradient w.r.t. the qk row of the slot from the CTA that owns it as a key: its key-role gradient
 *             (key-normalisation Jacobian applied) plus the query-role gradient from that CTA's keys,
 *   dq_b      query-role gradient from the previous CTA (slot seen as look-ahead chunk); written for every
 *             slot when bucket == 128 and only for slots in even sorted chunks when bucket == 64,
 *   dv_rounds value gradient. */
int rtts_lsh_attn_bwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const float* sumsq, const uint8_t* mask,
                      const rtts_lsh_spec* spec_host, const void* dout, int64_t ld_dout, const float* lse,
                      const float* delta, void* dqk_main, void* dq_b, void* dv_rounds, int B, int T, int H, int dh,
                      int R, int bucket, void* stream);

/* Sum the per-round gradients over the R rounds: dqk = sum_r (dqk_main + dq_b), dv = sum_r dv_rounds,
 * both bf16 [B,T,H*dh] (ld).  undo (from rtts_lsh_sort) tells which slots have a dq_b entry; may be NULL
 * when bucket == 128. */
int rtts_lsh_grad_reduce(const void* dqk_main, const void* dq_b, const void* dv_rounds, const int32_t* undo,
                         void* dqk, void* dv, int64_t ld, int B, int T, int H, int dh, int R, int bucket,
                         void* stream);

/* ---- LayerNorm (ref:reformer_tts/model/reformer.py:25-33, eps 1e-5, affine) -------------------- */

/* x fp32 [rows, dim] -> y bf16 [rows, dim]; saves mean / rstd fp32 [rows] when non-NULL. */
int rtts_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int rows, int dim, float eps, void* stream);
/* dy fp32 [rows, dim] (gradient w.r.t. the normalised, affine output) -> dx fp32; dgamma/dbeta fp32 [dim]
 * are ACCUMULATED (+=) so they can point at parameter .grad buffers. */
int rtts_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       float* dx, float* dgamma, float* dbeta, int rows, int dim, void* stream);
/* Same, accumulated onto another gradient: dx = dx_add + (LayerNorm backward); dx_add fp32 [rows, dim], may be NULL or alias dx
 * (the `dx2 += df` of the reversible blocks without a separate pass over the stream). */
int rtts_layernorm_bwd_acc(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                           const float* dx_add, float* dx, float* dgamma, float* dbeta, int rows, int dim, void* stream);

/* ---- GEMM with fused epilogue (projections rp R1, FeedForward ref:reformer_tts/model/modules.py:195-207) */

#define RTTS_EPI_BIAS 1      /* + bias[n] (fp32)                         */
#define RTTS_EPI_RELU 2      /* max(., 0)                                */
#define RTTS_EPI_GATE 4      /* * (gate[m,n] > 0), gate bf16 (ReLU bwd)  */
#define RTTS_EPI_OUT_BF16 8  /* C is bf16 (else fp32)                    */
#define RTTS_EPI_ATOMIC 16   /* C (fp32) += result, split-K allowed      */
#define RTTS_EPI_COLSUM 32   /* also accumulate column sums into colsum  */
#define RTTS_EPI_RESID_ADD 64   /* C (fp32) = resid + result: `gate` points at the fp32 residual [M,N], ldgate its row stride   */
#define RTTS_EPI_RESID_SUB 128  /* C (fp32) = resid - result (reversible reconstruction x1 = y1 - f(x2)); C may alias resid */

/* C[M,N] = epilogue(A . B^T).  Operands bf16, fp32 accumulate in TMEM (tcgen05).
 * a_mn_major = 0: A is [M,K] row-major (lda = K stride); 1: A is stored [K,M] row-major.
 * b_mn_major = 0: B is [N,K] row-major (an nn.Linear weight); 1: B is stored [K,N] row-major.
 * M % 128 == 0, N % 128 == 0, K % 64 == 0.  split_k >= 1 requires RTTS_EPI_ATOMIC when > 1. */
int rtts_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
                   int64_t ldc, const float* bias, const void* gate, int64_t ldgate, float* colsum, int M, int N,
                   int K, int epilogue, int split_k, void* stream);
/* rtts_gemm_bf16 with inverted dropout on the result: C = [resid +/-] keep * keep_scale * (A . B^T + bias) - nn.Dropout after the
 * output projection of the LSH layer (ref: reformer-pytorch post_attn_dropout) - keep_mask uint8 [M,N] (1 = keep, row stride
 * ldkeep, multiple of 4), keep_scale = 1 / (1 - p).  fp32 output only; combines with RTTS_EPI_BIAS and RTTS_EPI_RESID_*.
 * keep_mask == NULL is plain rtts_gemm_bf16. */
int rtts_gemm_bf16_dropout(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
                           int64_t ldc, const float* bias, const void* gate, int64_t ldgate, float* colsum,
                           const uint8_t* keep_mask, int64_t ldkeep, float keep_scale, int M, int N, int K, int epilogue,
                           int split_k, void* stream);

/* ---- dense decoder -> encoder attention core (nn.MultiheadAttention's softmax(QK^T / sqrt(dh)) V) ----------
 * Replaces the attention core of ref:reformer_tts/model/reformer.py:161-186 (torch.nn.MultiheadAttention.forward with
 * key_padding_mask and attention dropout).  q [B,T,H*dh], k / v [B,S,..] (head h at column h*dh, row pitch ldkv), all bf16,
 * token-major; keep uint8 [B,S] (1 = valid key, NULL = all valid); T a multiple of 128, S a multiple of 32 up to 256, dh = 64.
 * Dropout (p_drop > 0) drops normalised probabilities by a counter-based hash of (*seed, element): the same mask is
 * regenerated by every call given the same device seed word.  out bf16 [B,T,H*dh]; lse fp32 [B,H,T] (log2 domain: max * scale *
 * log2(e) + log2(row sum)), consumed by rtts_xattn_bwd. */
int rtts_xattn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* keep, float scale, float p_drop,
                   const uint64_t* seed, void* out, int64_t ldo, float* lse, int B, int T, int S, int H, int dh, void* stream);

/* Backward of rtts_xattn_fwd (scores recomputed in-kernel from q, k and lse).  dout bf16 [B,T,H*dh]; delta fp32 [B,H,T] =
 * <dout, out> per (b,h,t) (rtts_lsh_delta).  dq bf16 [B,T,H*dh] is written; dk / dv fp32 [B,S,..] (head h at column h*dh, row
 * pitch lddkv) are ACCUMULATED into (zero them first): every 128-query tile adds its partial sums with vector reductions. */
int rtts_xattn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* keep, float scale, float p_drop,
                   const uint64_t* seed, const void* dout, int64_t lddo, const float* lse, const float* delta, void* dq, int64_t lddq,
                   float* dk, float* dv, int64_t lddkv, int B, int T, int S, int H, int dh, void* stream);

/* ---- small fused element-wise helpers used by the host mirror -------------------------------- */

/* Column sums of a bf16 matrix (row pitch ld elements): colsum fp32 [cols] += sum_rows x[r, :].  Bias gradients of gradient matrices
 * that are already bf16 (the autograd sum over the batch of nn.Linear / nn.MultiheadAttention biases, ref:reformer_tts/model/reformer.py:161-186). */
int rtts_colsum_bf16(const void* x, int64_t ld, float* colsum, int rows, int cols, void* stream);

/* fp32 -> bf16 cast with optional column-sum accumulation (bias gradients): colsum fp32 [cols] += sum_rows. */
int rtts_cast_bf16_colsum(const float* x, void* y, float* colsum, int rows, int cols, void* stream);
/* Same with inverted dropout applied first: y = bf16(keep * scale * x), colsum += column sums of the masked values
 * (keep uint8 [rows, cols], 1 = keep; the backward of a layer whose output went through rtts_gemm_bf16_dropout). */
int rtts_cast_bf16_colsum_dropout(const float* x, const uint8_t* keep_mask, float keep_scale, void* y, float* colsum, int rows,
                                  int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RTTS_B200_H */
