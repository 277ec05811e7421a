/* lamp_b200 -- C ABI of the B200-native label-graph attention path of LaMP.
 *
 * The reference (QData/LaMP) is pure Python/PyTorch and defines no FFI; its seam for this path is the Python
 * class API.  This header is the native boundary underneath that seam: every entry
 * point cites the reference function (file:line under the reference checkout) whose arithmetic it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller, row-major, 16-byte aligned; the library never
 *    allocates or frees device memory and keeps no state besides per-process one-time kernel attributes;
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *  - return value: 0 on success, a negative LAMP_E* code otherwise; lamp_last_error() gives a thread-local
 *    human-readable message.  Nothing throws, nothing calls exit();
 *  - re-entrant / thread-safe: one process per GPU or several host threads may call concurrently;
 *  - results are bit-reproducible run to run except for the two accumulating backward entry points
 *    (lamp_gemm_tn_acc, lamp_layernorm_bwd's dgamma / dbeta), which sum partial results with fp32 atomics.
 *
 * Number formats
 *  - LAMP_PREC_FP32 : fp32 in / fp32 out; contractions run on the tcgen05 tensor cores as 3-term split-bf16
 *                     products (hi*hi + hi*lo + lo*hi, fp32 accumulate), ~2^-16 relative per product;
 *  - LAMP_PREC_BF16 : operands rounded to bf16 once (1-term), fp32 accumulate / softmax / LayerNorm.
 *  - "planes": an fp32 matrix carried as two bf16 matrices (hi = bf16(x), lo = bf16(x - hi)) with a common
 *    leading dimension; `lo` may be NULL for LAMP_PREC_BF16.  Intermediate activations stay in this form between
 *    the kernels of a layer so that every HBM byte moved is a byte a tensor core consumes.
 */
#ifndef LAMP_B200_H_
#define LAMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAMP_OK 0
#define LAMP_EINVAL (-1)     /* bad shape, stride or alignment                       */
#define LAMP_ECUDA (-2)      /* CUDA runtime / driver error                          */
#define LAMP_EARCH (-3)      /* current device is not compute capability 10.x        */
#define LAMP_EWORKSPACE (-4) /* workspace too small                                  */

#define LAMP_PREC_FP32 0
#define LAMP_PREC_BF16 1

int lamp_version(void);
const char* lamp_last_error(void);
/* 0 if the CURRENT device can run the library (sm_100), LAMP_EARCH otherwise. */
int lamp_device_check(void);
int lamp_sm_count(void);
/* Process-wide tuning knobs (benchmarking aid; results are identical for every setting). */
#define LAMP_TUNE_GEMM_BLOCK_K 1  /* 0 (default: automatic), 32 (64B swizzle, deeper TMA ring) or 64 (128B swizzle) */
#define LAMP_TUNE_ATTN_COMPACT 3  /* 1 (default: tile rows follow L, deepest K/V staging that fits) or 0 (128-row tiles) */
#define LAMP_TUNE_ATTN_STAGE 4    /* 1 (default: attention output planes leave through smem staging + TMA stores) or 0 */
#define LAMP_TUNE_ATTN_PV_SPLIT 5  /* 0 (default) or 1: O += P V as two interleaved N = 64 accumulation chains (d == 128) */
#define LAMP_TUNE_GEMM_TN_TC 6     /* 1 (default): weight gradient dY^T X on tcgen05 (MN-major operands), 0: warp-MMA version */
#define LAMP_TUNE_ATTN_BWD_TC 7    /* 1 (default): attention backward as batched tcgen05 products, 0: warp-MMA kernels */
#define LAMP_TUNE_PDL 8            /* programmatic dependent launch: 2 (default: launches of <= 16384 rows), 1 (always), 0 (never) */
#define LAMP_TUNE_ATTN_KV128_MIN_LK 9 /* L (default 512): 128-key tiles (one K, one V slot) for d > 64 when Lk >= L; 0: always 64-key tiles there */
#define LAMP_TUNE_GEMM_CTA_PAIR 2 /* 1 (default: tcgen05 cta_group::2 pairs for 256-wide tiles) or 0 (single CTAs) */
int lamp_set_tuning(int key, int value);

/* ---------------------------------------------------------------- level 1: kernels ------------------------ */

/* x[rows, cols] fp32 (leading dim ld) -> planes (leading dim ldp).  cols % 4 == 0. */
int lamp_split_planes(const float* x, int64_t rows, int cols, int64_t ld, void* hi, void* lo, int64_t ldp,
                      void* stream);

/* Many small splits in one launch (training: the planes of every projection weight, W and W^T, once per step).
 * Job: fp32 src [rows, cols] (leading dim ld) -> planes [rows, cols], or [cols, rows] with transpose != 0 (leading dim
 * ldp; hi / lo may point into a wider matrix, e.g. the Wq | Wk | Wv row blocks of one [3*H*d, D] operand).  lo may be
 * NULL.  No alignment requirements beyond the element types. */
typedef struct LampSplitJob {
  const float* src;
  void* hi;
  void* lo;
  int rows, cols;
  int64_t ld, ldp;
  int transpose;
} LampSplitJob;
int lamp_split_planes_multi(const LampSplitJob* jobs, int n_jobs, void* stream);

/* C[M,N] = A[M,K] * W[N,K]^T, then (+bias[N]) (ReLU) (+residual[row % resid_mod or row, :]) and store as fp32
 * and/or planes.  Replaces the nn.Linear / Conv1d(k=1) contractions of lamp/SubLayers.py:91-93 (w_qs,w_ks,w_vs),
 * :110 (fc) and :133 (w_1, w_2).  K % 8 == 0, N % 8 == 0.  m_dev (nullable, device int32): process only
 * min(M, *m_dev) rows -- the row count of a padding-aware (packed) batch lives on the device, so launching needs no
 * host synchronisation. */
int lamp_gemm_planes(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                     int64_t ldw, int M, int N, int K, int precision, const float* bias, int relu,
                     const float* residual, int64_t ldr, int resid_mod, float* out_f32, int64_t ldo, void* out_hi,
                     void* out_lo, int64_t ldp, const int32_t* m_dev, void* stream);

/* Training variant of lamp_gemm_planes (3-term products): out_f32 = dropout(A W^T + bias) + residual, the dropout
 * being the counter hash of lamp_dropout_add / lamp_dropout_split (keep(row, col) is a pure function of
 * (seed + *seed_dev, row, col); seed_dev nullable) -- fc / w_2 followed by nn.Dropout and the residual add
 * (lamp/SubLayers.py:113-117, 136-141) without a separate element-wise pass.  p_drop == 0 degenerates to
 * lamp_gemm_planes with a residual. */
int lamp_gemm_planes_drop(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                          int64_t ldw, int M, int N, int K, const float* bias, float p_drop, uint64_t seed,
                          const uint64_t* seed_dev, const float* residual, int64_t ldr, int resid_mod, float* out_f32,
                          int64_t ldo, void* stream);

/* lamp_gemm_planes with the residual given as split-bf16 planes (hi + lo reconstructs it to 2^-17 relative): lets a
 * layer keep its activations in operand form only, without an fp32 copy in HBM.  res_lo may be NULL. */
int lamp_gemm_planes_pres(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                          int64_t ldw, int M, int N, int K, int precision, const float* bias, const void* res_hi,
                          const void* res_lo, int64_t ldr, int resid_mod, float* out_f32, int64_t ldo, void* out_hi,
                          void* out_lo, int64_t ldp, const int32_t* m_dev, void* stream);

/* Same contraction with the residual add AND the LayerNorm of lamp/SubLayers.py:117 / :141 fused into the epilogue:
 * out = LayerNorm(A W^T (+bias) (+residual)) * gamma + beta, written as fp32 and/or planes.  The whole output row
 * must fit the on-chip accumulator: 256 < N <= 512 (otherwise LAMP_EINVAL: use lamp_gemm_planes + lamp_layernorm). */
int lamp_gemm_ln_planes(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                        int64_t ldw, int M, int N, int K, int precision, const float* bias, const float* residual,
                        int64_t ldr, int resid_mod, const float* gamma, const float* beta, float eps, float* out_f32,
                        int64_t ldo, void* out_hi, void* out_lo, int64_t ldp, void* stream);

/* ---- deferred LayerNorm.  Between two GEMMs the LayerNorm of lamp/SubLayers.py:117 / :141 need not run as a pass of
 * its own: the producing GEMM writes the PRE-norm tensor y as planes together with per-row partial sums
 * {sum y, sum y^2} ("row stats", float2 [rows][nparts], nparts = lamp_gemm_stats_parts(N)), and every consumer applies
 * the normalisation itself.  Results equal LayerNorm(y)*gamma+beta up to fp32 rounding (variance as E[y^2]-mean^2). */
int lamp_gemm_stats_parts(int N);

/* Consumer, A operand:  C = LayerNorm_gamma,beta(y) W^T (+bias) (ReLU) -> planes, computed as
 * rstd*(y Wg^T - mean*colsum) + biasf with Wg = W*diag(gamma) given as the weight planes, colsum[n] = sum_k Wg[n,k]
 * and biasf[n] = bias[n] + sum_k beta[k] W[n,k] (both folded by the caller once per weight version).  The LayerNorm
 * width is K. */
int lamp_gemm_planes_dln(const void* y_hi, const void* y_lo, int64_t lda, const float* a_stats, int a_nparts,
                         float a_eps, const void* wg_hi, const void* wg_lo, int64_t ldw, const float* colsum,
                         const float* biasf, int M, int N, int K, int precision, int relu, void* out_hi, void* out_lo,
                         int64_t ldp, const int32_t* m_dev, void* stream);

/* Producer:  y = A W^T (+bias) + residual -> planes + row stats of y (stats_out: float2 [M][lamp_gemm_stats_parts(N)]).
 * The residual is fp32 (`residual`, rows modulo resid_mod when > 0), plain planes (res_hi/res_lo), or -- with r_stats
 * -- itself a deferred LayerNorm of width N: planes of the pre-norm tensor, normalised element-wise with r_gamma /
 * r_beta / r_eps while it is added. */
int lamp_gemm_planes_rstats(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                            int64_t ldw, int M, int N, int K, int precision, const float* bias, const float* residual,
                            const void* res_hi, const void* res_lo, int64_t ldr, int resid_mod, const float* r_stats,
                            int r_nparts, float r_eps, const float* r_gamma, const float* r_beta, void* out_hi,
                            void* out_lo, int64_t ldp, float* stats_out, const int32_t* m_dev, void* stream);

/* Materialise a deferred LayerNorm: out[r,:] = LayerNorm(y[index ? index[r] : r, :])*gamma+beta as fp32 and/or planes
 * (with `index` also the un-packing gather of the encoder output). */
int lamp_ln_apply(const void* y_hi, const void* y_lo, const float* stats, int nparts, const float* gamma,
                  const float* beta, float eps, int64_t rows, int D, const int64_t* index, float* out, void* out_hi,
                  void* out_lo, const int32_t* m_dev, void* stream);

/* lamp_diag_proj on a deferred LayerNorm: logits[b,l] = <LayerNorm(y[b,l,:])*gamma+beta, W[l,:]> (+bias[l]). */
int lamp_diag_proj_ln(const void* y_hi, const void* y_lo, const float* stats, int nparts, const float* gamma,
                      const float* beta, float eps, const float* W, const float* bias, int64_t B, int L, int D,
                      float* logits, void* stream);

/* Masked softmax attention over label nodes for B samples x H heads (lamp/SubLayers.py:27-43 with the head
 * split/merge of :96-107 folded into the addressing).  Q planes: [B*Lq (or Lq if q_bcast), ldq], head h at
 * columns q_col0 + h*d; K/V planes: [B*Lk, ldkv] at k_col0 / v_col0 + h*d.  mask: NULL or bytes (non-zero =
 * masked) addressed as mask[b*msb + i*msq + j*msk] (strides may be 0: [L,L] label mask -> msb = 0; [B,Lk]
 * key padding -> msq = 0).  Outputs: planes and/or fp32 [B*Lq, ld], head h at columns h*d.  If `probs` is not
 * NULL it receives the attention probabilities [H*B, Lq, Lk] (head-major batch index h*B + b, as
 * lamp/SubLayers.py:96-98,121) and row_max/row_sum ([H*B*Lq] floats each) must be provided as scratch.
 * d % 16 == 0, d <= 128.  A fully masked row yields NaN exactly like the reference's softmax over all -inf.
 * Padding-aware keys: with kv_start / kv_len (device int32 [B], both or neither) the K/V plane matrix holds only the
 * non-PAD tokens of the batch, packed (kv_rows rows in total); sample b attends to rows [kv_start[b], +kv_len[b]) and
 * Lk is the upper bound of kv_len.  Equivalent to the key-padding mask of lamp/utils.py:26-34 without touching PAD
 * keys at all.  In this mode `mask` (optional) is one byte per PACKED key row (msb = msq = 0; row kv_start[b] + j),
 * and `probs` must be NULL. */
int lamp_attn_core_planes(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                          const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                          int Lq, int Lk, int d, float temperature, int precision, const uint8_t* mask,
                          int64_t msb, int64_t msq, int64_t msk, void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                          int64_t ldof, float* row_max, float* row_sum, float* probs, const int32_t* kv_start,
                          const int32_t* kv_len, int64_t kv_rows, void* stream);

/* Bit-packed form of a [Bm, Lq, Lk] byte mask (strides msb/msq/msk as above; Bm = 1 for a mask shared by the batch):
 * words[(b*Lq + q)*W + (k >> 5)] bit (k & 31) = mask[b,q,k] != 0, W = ceil(Lk/32).  The label-graph mask of
 * lamp/Decoders.py:113 is packed once per model and read as one 32-bit word per thread and KV tile. */
int lamp_pack_mask_bits(const uint8_t* mask, int64_t msb, int64_t msq, int64_t msk, int64_t Bm, int Lq, int Lk,
                        uint32_t* words, void* stream);

/* lamp_attn_core_planes with the mask given in that packed form (word strides: mbb per sample -- 0 when shared --
 * and mbq per query row, mbq >= ceil(Lk/32)); no probability output, no packed keys. */
int lamp_attn_core_planes_mbits(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                                const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                                int Lq, int Lk, int d, float temperature, int precision, const uint32_t* mask_bits,
                                int64_t mbb, int64_t mbq, void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                                int64_t ldof, void* stream);

/* Backward of the attention core (training path; lamp/SubLayers.py:27-43 differentiated):
 * given q [N,Lq,d], k,v [N,Lk,d], the forward's output O and probabilities P [N,Lq,Lk] (before dropout) and A (after
 * dropout; NULL = no dropout, A == P), and dO, computes dq, dk, dv (fp32, same shapes as q, k, v).  p_drop is the
 * dropout rate the forward used (scale 1/(1-p)); the kept set is read off A.  Masked entries need no mask here: their
 * P is 0.  d % 16 == 0, d <= 128.  Deterministic (no atomics).  Runs as four batched tcgen05 products
 * (dA = dO V^T, dQ = dS K, dV = A^T dO, dK = dS^T Q; transposed operands are read MN-major in place) around one
 * element-wise kernel; LAMP_TUNE_ATTN_BWD_TC = 0 selects the warp-MMA version. */
size_t lamp_attn_core_bwd_workspace_bytes(int N, int Lq, int Lk, int d);
int lamp_attn_core_bwd(const float* q, const float* k, const float* v, const float* dO, const float* O, const float* P,
                       const float* A, float* dq, float* dk, float* dv, int N, int Lq, int Lk, int d, float temperature,
                       float p_drop, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of out = LayerNorm(x)*gamma+beta: dx (fp32), and dgamma / dbeta ACCUMULATED into the given [D] buffers
 * (zero them first).  x is the LayerNorm input saved by the forward. */
int lamp_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int D, float* dx,
                       float* dgamma, float* dbeta, void* stream);

/* lamp_layernorm_bwd that ALSO writes planes(dropout-backward(dx)) = planes(keep(row, col) * dx / (1 - p)) with the
 * counter hash of lamp_dropout_add / lamp_gemm_planes_drop: the gradient of the sub-layer's dropped branch, directly as
 * the operand of its dW / input-gradient products (replaces a lamp_dropout_split pass over dx).  p_drop == 0: a plain
 * split of dx. */
int lamp_layernorm_bwd_drop(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int D, float* dx,
                            float* dgamma, float* dbeta, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* dx_hi,
                            void* dx_lo, void* stream);

/* Weight / bias gradient of Y = X W^T (+ b): dW[N,K] += dY[M,N]^T X[M,K], db[N] += column sums of dY (db may be
 * NULL).  dY and X are given as split-bf16 planes (lo planes NULL in bf16 mode).  Accumulating (fp32 atomics over row
 * chunks): zero or pre-load dW / db.  N % 8 == 0, K % 8 == 0.  The input gradient dX = dY W is lamp_gemm_planes on
 * the transposed weight planes. */
int lamp_gemm_tn_acc(const void* dy_hi, const void* dy_lo, int64_t ldy, const void* x_hi, const void* x_lo, int64_t ldx,
                     int64_t M, int N, int K, float* dW, float* db, void* stream);

/* out = LayerNorm(y (+ add[row % add_mod or row])) * gamma + beta  (torch.nn.LayerNorm semantics, eps inside the
 * sqrt; lamp/SubLayers.py:117,141).  Writes fp32 and/or planes (any may be NULL).  D % 4 == 0, D <= 4096. */
int lamp_layernorm(const float* y, const float* add, int add_mod, const float* gamma, const float* beta, float eps,
                   int64_t rows, int D, float* out, void* out_hi, void* out_lo, const int32_t* m_dev, void* stream);

/* out[r,:] = word_emb[seq[i],:] (+ pos_emb[pos[i],:]), i = row_index ? row_index[r] : r  (lamp/Encoders.py:66,75).
 * seq/pos/row_index are int64; with row_index + m_dev only the first *m_dev listed tokens are embedded (packed batch). */
int lamp_embed(const int64_t* seq, const int64_t* pos, const float* word_emb, const float* pos_emb, int64_t rows,
               int D, float* out, void* out_hi, void* out_lo, const int64_t* row_index, const int32_t* m_dev,
               void* stream);

/* Backward of lamp_embed without row_index (autograd of lamp/Encoders.py:66,75; torch.nn.Embedding(padding_idx)):
 * dword[seq[r],:] += g[r,:] unless seq[r] == pad_word, dpos[pos[r],:] += g[r,:] unless pos[r] == pad_pos (pass -1
 * for "no padding row").  Accumulates with vector reductions into the caller-zeroed tables; either table may be NULL.
 * The order of the additions is not fixed. */
int lamp_embed_bwd(const float* g, const int64_t* seq, const int64_t* pos, int64_t rows, int D, int64_t pad_word,
                   int64_t pad_pos, float* dword, float* dpos, void* stream);

/* out[r,:] = src[index[r],:] (fp32): un-packs a packed activation into the dense [B*T, D] API tensor. */
int lamp_gather_rows(const float* src, const int64_t* index, int64_t rows, int D, float* out, void* stream);

/* Zero rows [*m_dev, *m_dev + nguard) of a packed plane matrix (rows >= max_rows are skipped): keeps the few rows a KV
 * tile may read past the packed data finite. */
int lamp_zero_guard_rows(void* hi, void* lo, int64_t ld, int cols, const int32_t* m_dev, int64_t max_rows, int nguard,
                         void* stream);

/* logits[b,l] = <x[b,l,:], W[l,:]> (+bias[l]) : the diagonal of the [B,L,L] projection, lamp/Models.py:124-126. */
int lamp_diag_proj(const float* x, const float* W, const float* bias, int64_t B, int L, int D, float* logits,
                   void* stream);

/* Backward of lamp_diag_proj: dx[b,l,:] = g[b,l] W[l,:] (dx may be NULL), dW[l,:] = sum_b g[b,l] x[b,l,:] and
 * dbias[l] = sum_b g[b,l] (dW / dbias may be NULL; they are WRITTEN, not accumulated). */
int lamp_diag_proj_bwd(const float* g, const float* x, const float* W, int64_t B, int L, int D, float* dx, float* dW,
                       float* dbias, void* stream);

/* ---- fused training sub-layers (second pass): everything stays in the projection GEMMs' layouts ---- */

/* Training forward of the attention core (lamp/SubLayers.py:27-43 with the dropout of :40 inside the kernel) on operand
 * planes: like lamp_attn_core_planes plus p_drop / seed (+ optional device-side counter added to the seed, so CUDA-graph
 * replays draw fresh masks), and optionally the probability tensors: attn [H*B, Lq, Lk] (after dropout; the reference's
 * return value, head-major) and probs_pre (before dropout; NULL iff p_drop == 0).  attn == NULL: nothing of size Lq*Lk
 * is written -- the backward recomputes P (lamp_attn_bwd_planes, recompute form) from row_max / row_sum [H*B*Lq]. */
int lamp_attn_core_planes_train(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast, const void* kv_hi,
                                const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H, int Lq, int Lk, int d,
                                float temperature, int precision, const uint8_t* mask, int64_t msb, int64_t msq,
                                int64_t msk, void* o_hi, void* o_lo, int64_t ldo, float* row_max, float* row_sum,
                                float* attn, float* probs_pre, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                void* stream);

/* lamp_attn_core_planes_train with the mask given as packed bits (lamp_pack_mask_bits; one 4-byte load per thread and
 * key tile instead of 32 byte loads + ballots -- the label-graph mask of a multi-tile problem, L > 128) and without
 * the probability outputs (recompute-form training: the backward rebuilds P from the row statistics and the BYTE
 * mask). */
int lamp_attn_core_planes_train_mbits(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                                      const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B,
                                      int H, int Lq, int Lk, int d, float temperature, int precision,
                                      const uint32_t* mask_bits, int64_t mbb, int64_t mbq, void* o_hi, void* o_lo,
                                      int64_t ldo, float* row_max, float* row_sum, float p_drop, uint64_t seed,
                                      const uint64_t* seed_dev, void* stream);

/* Backward of the attention core with every operand in place: Q / K / V / dO / O are split-bf16 planes, head h = column
 * slice [col0 + h*d, col0 + (h+1)*d) of a [B*L, ld] matrix (dO and O: [B*Lq, ldo], col0 = 0); P / A: the fp32 head-major
 * [H*B, Lq, Lk] tensors of the training forward (A may be NULL or == P without dropout).  dQ -> planes [B*Lq, lddq] at
 * dq_col0 + h*d; dK / dV -> planes [B*Lk, lddkv] at dk_col0 / dv_col0 + h*d.  Four batched tcgen05 products around one
 * element-wise kernel; no permute / contiguous copies, no fp32 round trips.
 * RECOMPUTE FORM (P == NULL): the forward kept no probability tensor at all (lamp_attn_core_planes_train with attn ==
 * NULL); P is rebuilt from one more batched product S = Q K^T, the forward's row_max / row_sum, the mask (same pointer /
 * strides as the forward) and the dropout counter hash of (seed [+ *seed_dev]). */
size_t lamp_attn_bwd_planes_workspace_bytes(int B, int H, int Lq, int Lk);
int lamp_attn_bwd_planes(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, const void* kv_hi, const void* kv_lo,
                         int64_t ldkv, int k_col0, int v_col0, const void* do_hi, const void* do_lo, const void* o_hi,
                         const void* o_lo, int64_t ldo, const float* P, const float* A, void* dq_hi, void* dq_lo,
                         int64_t lddq, int dq_col0, void* dkv_hi, void* dkv_lo, int64_t lddkv, int dk_col0, int dv_col0,
                         int B, int H, int Lq, int Lk, int d, float temperature, float p_drop, const float* row_max,
                         const float* row_sum, const uint8_t* mask, int64_t msb, int64_t msq, int64_t msk, uint64_t seed,
                         const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, void* stream);

/* y = dropout(y0) + x (lamp/SubLayers.py:113-119 / :139-141): keep(row, col) is a counter hash of (seed [+ *seed_dev],
 * row, col), kept values are scaled by 1/(1-p); x row index is row % x_mod when x_mod > 0 (broadcast residual). */
int lamp_dropout_add(const float* y0, const float* x, int64_t rows, int D, int x_mod, float p_drop, uint64_t seed,
                     const uint64_t* seed_dev, float* y, void* stream);

/* planes(dropout-backward(dy)): the same mask recomputed, result written as split-bf16 planes [rows, D] (p_drop == 0: a
 * plain split). */
int lamp_dropout_split(const float* dy, int64_t rows, int D, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                       void* hi, void* lo, void* stream);

/* ReLU backward on operand planes, in place: g = 0 where the forward activation (its hi plane) is <= 0.  n % 8 == 0. */
int lamp_relu_mask_planes(void* g_hi, void* g_lo, const void* h_hi, int64_t n, void* stream);

/* Device-side training targets (replaces the per-row CPU loop of utils/utils.py:205-216 `get_gold_binary`, called from
 * train.py:34 and test.py:47): gold [B, W] int64 label ids (+`skip` = 4 special tokens), EOS-terminated, PAD = 0 padded.
 * Per row: entries > 0 minus the LAST of them (the EOS) -> out[b, id - skip] = 1, everything else 0.  out [B, L] fp32 is
 * written completely. */
int lamp_gold_binary(const int64_t* gold, int64_t B, int W, int L, int skip, float* out, void* stream);

/* F.binary_cross_entropy_with_logits(logits, target, reduction='mean') (train.py:38) and its gradient in ONE pass:
 * *loss = mean(max(x,0) - x y + log1p(exp(-|x|))), dlogits = (sigmoid(x) - y) / n (dlogits may be NULL).  Deterministic
 * (fixed-order block partials).  workspace: lamp_bce_logits_workspace_bytes() bytes, zeroed ONCE by the caller before
 * the first use (the kernel re-arms it, so CUDA-graph replays can reuse it). */
#define LAMP_BCE_MAX_BLOCKS 256
size_t lamp_bce_logits_workspace_bytes(void);
int lamp_bce_logits(const float* logits, const float* target, int64_t n, float* loss, float* dlogits, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- level 2: reference-shaped ops ----------- */

/* ScaledDotProductAttention.forward (lamp/SubLayers.py:27-43), eval mode.
 * q [N,Lq,d], k,v [N,Lk,d] fp32; mask as above with N in place of B; out [N,Lq,d]; attn [N,Lq,Lk] or NULL. */
size_t lamp_sdpa_workspace_bytes(int N, int Lq, int Lk, int d);
int lamp_sdpa_fwd(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                  int64_t msk, float* out, float* attn, int N, int Lq, int Lk, int d, float temperature,
                  int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Training forward of the same op: dropout with rate p_drop on the probabilities (lamp/SubLayers.py:40) inside the
 * kernel.  The kept set is a pure function of (seed, row, key) -- a counter-based hash, no state -- so `attn` (after
 * dropout, what the reference returns) and `probs_pre` (before dropout, saved for lamp_attn_core_bwd; may be NULL when
 * p_drop == 0) describe exactly what the PV product consumed.  seed_dev (nullable, device uint64): a counter added to
 * `seed` on the device -- a CUDA-graph replay of a training step bakes `seed` into the graph and advances the counter. */
int lamp_sdpa_fwd_train(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                        int64_t msk, float* out, float* attn, float* probs_pre, int N, int Lq, int Lk, int d,
                        float temperature, int precision, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                        void* workspace, size_t workspace_bytes, void* stream);

/* MultiHeadAttention.forward (lamp/SubLayers.py:77-121), eval mode (dropout = identity).
 * q [B,Lq,D]; kv [B,Lk,D] or NULL for self-attention (k = v = q); Wq,Wk,Wv [H*d, D]; Wfc [D, H*d] or NULL iff
 * H == 1 (:72-74); out = LayerNorm(fc(concat heads) + q) [B,Lq,D]; attn [H*B,Lq,Lk] head-major or NULL. */
size_t lamp_mha_workspace_bytes(int B, int Lq, int Lk, int D, int H, int d, int self_attn, int want_attn);
int lamp_mha_fwd(const float* q, const float* kv, const float* Wq, const float* Wk, const float* Wv,
                 const float* Wfc, const float* ln_w, const float* ln_b, const uint8_t* mask, int64_t msb,
                 int64_t msq, int64_t msk, float* out, float* attn, int B, int Lq, int Lk, int D, int H, int d,
                 int precision, float ln_eps, void* workspace, size_t workspace_bytes, void* stream);

/* PositionwiseFeedForward.forward (lamp/SubLayers.py:135-142), eval mode.
 * x [rows, D]; W1 [d_inner, D] (Conv1d weight [d_inner, D, 1]); b1 [d_inner]; W2 [D, d_inner]; b2 [D]. */
size_t lamp_ffn_workspace_bytes(int64_t rows, int D, int d_inner);
int lamp_ffn_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                 const float* ln_w, const float* ln_b, float* out, int64_t rows, int D, int d_inner, int precision,
                 float ln_eps, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAMP_B200_H_ */
