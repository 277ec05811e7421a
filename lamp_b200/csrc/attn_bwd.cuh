// Backward of the masked attention core (SURVEY.md 8f, N4 -- first native piece of the training path):
//   forward (lamp/SubLayers.py:27-43):  S = q k^T / temperature;  P = softmax(mask(S));  A = dropout(P);  O = A v
//   backward:  dV = A^T dO;   dA = dO V^T;   dP = dA o keep / (1 - p);   delta_i = sum_j P_ij dP_ij = <dO_i, O_i>;
//              dS = P o (dP - delta) / temperature;   dQ = dS K;   dK = dS^T Q.
// The training forward keeps P (and A when dropout is active) in HBM exactly like the reference does (it returns
// `attn`), so the backward is four batched contractions plus element-wise work and needs neither the mask nor the
// dropout generator: keep == (A != 0) wherever P != 0, and P == 0 gives dS == 0 regardless.
//
// Two kernels over fp32 head-major tensors q [N, Lq, d], k/v [N, Lk, d], dO/O [N, Lq, d], P/A [N, Lq, Lk]:
//   attn_bwd_dq_kernel : one CTA per (n, 64-row q tile), loop over 64-key tiles: dA tile -> dS tile (written to a
//                        [N, Lq, Lk] fp32 scratch for the second kernel) -> dQ accumulation in registers.
//   attn_bwd_dkv_kernel: one CTA per (n, 64-key tile), loop over q tiles: dV += A^T dO, dK += dS^T Q in registers.
// No atomics, deterministic.  Contractions run on the tensor cores through warp-level mma (nvcuda::wmma, bf16
// 16x16x16, fp32 accumulate) with the same 3-term split-bf16 operands as the forward (hi*hi + hi*lo + lo*hi), staged
// in shared memory; transposed operands are free (col-major fragment loads).  This is a correctness-first version:
// it does not use tcgen05/TMA yet.
#pragma once
#include <mma.h>

#include <type_traits>

#include "sm100_primitives.cuh"

namespace lamp {

constexpr int BWD_TILE = 64;        // q rows / keys per tile
constexpr int BWD_THREADS = 256;    // 8 warps: warp w -> output rows 16*(w&3), column half (w>>2)
constexpr int BWD_DMAX = 128;

struct AttnBwdParams {
  int N, Lq, Lk, d;        // d % 16 == 0, d <= 128
  float inv_temp;          // 1 / temperature
  float drop_scale;        // 1 / (1 - p), 1 when dropout is off
  const float *q, *k, *v, *dO, *O;   // [N, L, d]
  const float* P;          // [N, Lq, Lk] softmax probabilities before dropout
  const float* A;          // [N, Lq, Lk] after dropout (== P when dropout is off; may alias P)
  float* dS;               // [N, Lq, Lk] scratch (dq kernel writes, dkv kernel reads)
  float *dq, *dk, *dv;     // [N, L, d]
};

// fp32 tile [rows x cols] (leading dim ld, rows/cols beyond the valid range read as 0) -> hi / lo bf16 planes in smem
// (leading dim lds), optionally scaled.
__device__ __forceinline__ void bwd_stage_tile(const float* __restrict__ src, long long ld, int rows_valid, int cols_valid,
                                               int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, int lds,
                                               float scale = 1.0f) {
  for (int idx = threadIdx.x; idx < rows * (cols >> 1); idx += BWD_THREADS) {
    const int r = idx / (cols >> 1), c = (idx % (cols >> 1)) << 1;
    float x0 = 0.f, x1 = 0.f;
    if (r < rows_valid) {
      if (c < cols_valid) x0 = src[r * ld + c] * scale;
      if (c + 1 < cols_valid) x1 = src[r * ld + c + 1] * scale;
    }
    uint32_t h, l;
    split_bf16x2(x0, x1, h, l);
    *reinterpret_cast<uint32_t*>(hi + r * lds + c) = h;
    *reinterpret_cast<uint32_t*>(lo + r * lds + c) = l;
  }
}

using FragA_R = nvcuda::wmma::fragment<nvcuda::wmma::matrix_a, 16, 16, 16, __nv_bfloat16, nvcuda::wmma::row_major>;
using FragA_C = nvcuda::wmma::fragment<nvcuda::wmma::matrix_a, 16, 16, 16, __nv_bfloat16, nvcuda::wmma::col_major>;
using FragB_R = nvcuda::wmma::fragment<nvcuda::wmma::matrix_b, 16, 16, 16, __nv_bfloat16, nvcuda::wmma::row_major>;
using FragB_C = nvcuda::wmma::fragment<nvcuda::wmma::matrix_b, 16, 16, 16, __nv_bfloat16, nvcuda::wmma::col_major>;
using FragC = nvcuda::wmma::fragment<nvcuda::wmma::accumulator, 16, 16, 16, float>;

// acc[f] (+)= A[16 rows at a_row0, K] * B[K, 16 cols at b_col0 + 16 f], f < NF, 3-term split products.
//   A_TRANS == false: A element (m, k) at a[(a_row0 + m) * lda + k]         (row-major [M, K])
//   A_TRANS == true : A element (m, k) at a[k * lda + a_row0 + m]           (stored as [K, M]: transposed operand)
//   B_TRANS == false: B element (k, n) at b[k * ldb + b_col0 + n]           (row-major [K, N])
//   B_TRANS == true : B element (k, n) at b[(b_col0 + n) * ldb + k]         (stored as [N, K])
template <int NF, bool A_TRANS, bool B_TRANS>
__device__ __forceinline__ void bwd_mma(FragC (&acc)[NF], const __nv_bfloat16* a_hi, const __nv_bfloat16* a_lo, int lda,
                                        int a_row0, const __nv_bfloat16* b_hi, const __nv_bfloat16* b_lo, int ldb,
                                        int b_col0, int K) {
  using namespace nvcuda;
  for (int k0 = 0; k0 < K; k0 += 16) {
    typename std::conditional<A_TRANS, FragA_C, FragA_R>::type ah, al;
    const int aoff = A_TRANS ? (k0 * lda + a_row0) : (a_row0 * lda + k0);
    wmma::load_matrix_sync(ah, a_hi + aoff, lda);
    wmma::load_matrix_sync(al, a_lo + aoff, lda);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      typename std::conditional<B_TRANS, FragB_C, FragB_R>::type bh, bl;
      const int boff = B_TRANS ? ((b_col0 + 16 * f) * ldb + k0) : (k0 * ldb + b_col0 + 16 * f);
      wmma::load_matrix_sync(bh, b_hi + boff, ldb);
      wmma::load_matrix_sync(bl, b_lo + boff, ldb);
      wmma::mma_sync(acc[f], ah, bh, acc[f]);
      wmma::mma_sync(acc[f], ah, bl, acc[f]);
      wmma::mma_sync(acc[f], al, bh, acc[f]);
    }
  }
}

// Shared-memory plan of both kernels (bytes): four [64 x DP] plane pairs + two [64 x 64] plane pairs + one fp32
// [64 x 64] tile + 64 floats, DP = d padded by 8 elements (bank spread).
__host__ __device__ constexpr int bwd_dp(int d) { return d + 8; }
__host__ __device__ constexpr size_t attn_bwd_smem_bytes(int d) {
  return static_cast<size_t>(3) * 2 * BWD_TILE * bwd_dp(d) * 2 + 2 * 2 * BWD_TILE * (BWD_TILE + 8) * 2 +
         BWD_TILE * (BWD_TILE + 4) * 4 + BWD_TILE * 4 + 128;
}

// ---------------------------------------------------------------------------------------------------- dS and dQ
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_dq_kernel(const AttnBwdParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(128) uint8_t bsm[];
  const int DP = bwd_dp(p.d), TP = BWD_TILE + 8, FP = BWD_TILE + 4;
  __nv_bfloat16* dO_hi = reinterpret_cast<__nv_bfloat16*>(bsm);
  __nv_bfloat16* dO_lo = dO_hi + BWD_TILE * DP;
  __nv_bfloat16* V_hi = dO_lo + BWD_TILE * DP;
  __nv_bfloat16* V_lo = V_hi + BWD_TILE * DP;
  __nv_bfloat16* K_hi = V_lo + BWD_TILE * DP;
  __nv_bfloat16* K_lo = K_hi + BWD_TILE * DP;
  __nv_bfloat16* dS_hi = K_lo + BWD_TILE * DP;
  __nv_bfloat16* dS_lo = dS_hi + BWD_TILE * TP;
  __nv_bfloat16* unused = dS_lo + BWD_TILE * TP;      // (second [64 x 64] plane pair: used by the dkv kernel only)
  float* dA = reinterpret_cast<float*>(unused + 2 * BWD_TILE * TP);
  float* delta = dA + BWD_TILE * FP;

  const int num_qt = (p.Lq + BWD_TILE - 1) / BWD_TILE;
  const int n = blockIdx.x / num_qt, qt = blockIdx.x % num_qt;
  const int q0 = qt * BWD_TILE;
  const int qrows = min(BWD_TILE, p.Lq - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp & 3) * 16, wh = warp >> 2;  // output rows / column half of this warp
  const float* dO = p.dO + (static_cast<long long>(n) * p.Lq + q0) * p.d;
  const float* O = p.O + (static_cast<long long>(n) * p.Lq + q0) * p.d;

  bwd_stage_tile(dO, p.d, qrows, p.d, BWD_TILE, p.d, dO_hi, dO_lo, DP);
  // delta_i = <dO_i, O_i>: one warp per 8 rows
  for (int r = warp * 8; r < warp * 8 + 8; ++r) {
    float s = 0.f;
    if (r < qrows)
      for (int c = lane; c < p.d; c += 32) s += dO[r * p.d + c] * O[r * p.d + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) delta[r] = s;
  }
  constexpr int NFQ = BWD_DMAX / 32;  // dQ fragments per warp at d = 128 (column half = d / 2)
  FragC dq_acc[NFQ];
#pragma unroll
  for (int f = 0; f < NFQ; ++f) wmma::fill_fragment(dq_acc[f], 0.0f);
  const int nfq = p.d / 32;           // fragments actually used (d % 32 == 0 -> d/2 is a multiple of 16)
  const bool odd16 = (p.d % 32) != 0; // d = 16, 48, 80, 112: the second half is 16 columns shorter

  const int num_kt = (p.Lk + BWD_TILE - 1) / BWD_TILE;
  for (int kt = 0; kt < num_kt; ++kt) {
    const int k0 = kt * BWD_TILE;
    const int krows = min(BWD_TILE, p.Lk - k0);
    __syncthreads();  // previous iteration's readers of V / K / dS are done
    bwd_stage_tile(p.v + (static_cast<long long>(n) * p.Lk + k0) * p.d, p.d, krows, p.d, BWD_TILE, p.d, V_hi, V_lo, DP);
    bwd_stage_tile(p.k + (static_cast<long long>(n) * p.Lk + k0) * p.d, p.d, krows, p.d, BWD_TILE, p.d, K_hi, K_lo, DP);
    __syncthreads();
    // dA tile [64 q x 64 keys] = dO_i V_j^T ; warp: rows wr, key columns 32 * wh .. +32
    {
      FragC acc[2];
      wmma::fill_fragment(acc[0], 0.0f);
      wmma::fill_fragment(acc[1], 0.0f);
      bwd_mma<2, false, true>(acc, dO_hi, dO_lo, DP, wr, V_hi, V_lo, DP, 32 * wh, p.d);
      wmma::store_matrix_sync(dA + wr * FP + 32 * wh, acc[0], FP, wmma::mem_row_major);
      wmma::store_matrix_sync(dA + wr * FP + 32 * wh + 16, acc[1], FP, wmma::mem_row_major);
    }
    __syncthreads();
    // dS = P o (dA * keep * drop_scale - delta) / temperature  -> global scratch + planes in smem
    for (int idx = threadIdx.x; idx < BWD_TILE * (BWD_TILE >> 1); idx += BWD_THREADS) {
      const int r = idx / (BWD_TILE >> 1), c = (idx % (BWD_TILE >> 1)) << 1;
      float s0 = 0.f, s1 = 0.f;
      if (r < qrows) {
        const long long g = (static_cast<long long>(n) * p.Lq + q0 + r) * p.Lk + k0 + c;
        if (c < krows) {
          const float pr = p.P[g];
          const float keep = (p.A[g] != 0.0f) ? p.drop_scale : 0.0f;
          s0 = pr * (dA[r * FP + c] * keep - delta[r]) * p.inv_temp;
          p.dS[g] = s0;
        }
        if (c + 1 < krows) {
          const float pr = p.P[g + 1];
          const float keep = (p.A[g + 1] != 0.0f) ? p.drop_scale : 0.0f;
          s1 = pr * (dA[r * FP + c + 1] * keep - delta[r]) * p.inv_temp;
          p.dS[g + 1] = s1;
        }
      }
      uint32_t h, l;
      split_bf16x2(s0, s1, h, l);
      *reinterpret_cast<uint32_t*>(dS_hi + r * TP + c) = h;
      *reinterpret_cast<uint32_t*>(dS_lo + r * TP + c) = l;
    }
    __syncthreads();
    // dQ_i += dS K_j ; warp: rows wr, d columns (d/2) * wh ..
    {
      const int c0 = odd16 ? (wh ? (p.d / 2 + 8) : 0) : (p.d / 2) * wh;  // split d into two 16-aligned halves
      const int nf = odd16 ? (wh ? (p.d - (p.d / 2 + 8)) / 16 : (p.d / 2 + 8) / 16) : nfq;
      switch (nf) {
        case 4: bwd_mma<4, false, false>(dq_acc, dS_hi, dS_lo, TP, wr, K_hi, K_lo, DP, c0, BWD_TILE); break;
        case 3: { FragC(&a3)[3] = reinterpret_cast<FragC(&)[3]>(dq_acc); bwd_mma<3, false, false>(a3, dS_hi, dS_lo, TP, wr, K_hi, K_lo, DP, c0, BWD_TILE); break; }
        case 2: { FragC(&a2)[2] = reinterpret_cast<FragC(&)[2]>(dq_acc); bwd_mma<2, false, false>(a2, dS_hi, dS_lo, TP, wr, K_hi, K_lo, DP, c0, BWD_TILE); break; }
        case 1: { FragC(&a1)[1] = reinterpret_cast<FragC(&)[1]>(dq_acc); bwd_mma<1, false, false>(a1, dS_hi, dS_lo, TP, wr, K_hi, K_lo, DP, c0, BWD_TILE); break; }
        default: break;
      }
    }
  }
  // dQ tile -> global (through the fp32 smem tile, 16 rows x 16 columns per fragment)
  __syncthreads();
  {
    const int c0 = odd16 ? (wh ? (p.d / 2 + 8) : 0) : (p.d / 2) * wh;
    const int nf = odd16 ? (wh ? (p.d - (p.d / 2 + 8)) / 16 : (p.d / 2 + 8) / 16) : nfq;
    float* stage = dA + warp * (16 * 20);  // per-warp 16 x 16 staging (ld 20), inside the [64 x 68] fp32 tile
    float* dq = p.dq + (static_cast<long long>(n) * p.Lq + q0) * p.d;
    for (int f = 0; f < nf; ++f) {
      wmma::store_matrix_sync(stage, dq_acc[f], 20, wmma::mem_row_major);
      __syncwarp();
      for (int e = lane; e < 256; e += 32) {
        const int r = e >> 4, c = e & 15;
        if (wr + r < qrows) dq[(wr + r) * p.d + c0 + 16 * f + c] = stage[r * 20 + c];
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------- dK and dV
__global__ void __launch_bounds__(BWD_THREADS, 1) attn_bwd_dkv_kernel(const AttnBwdParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(128) uint8_t bsm[];
  const int DP = bwd_dp(p.d), TP = BWD_TILE + 8, FP = BWD_TILE + 4;
  __nv_bfloat16* dO_hi = reinterpret_cast<__nv_bfloat16*>(bsm);
  __nv_bfloat16* dO_lo = dO_hi + BWD_TILE * DP;
  __nv_bfloat16* Q_hi = dO_lo + BWD_TILE * DP;
  __nv_bfloat16* Q_lo = Q_hi + BWD_TILE * DP;
  __nv_bfloat16* spare = Q_lo + BWD_TILE * DP;         // third [64 x DP] pair: unused here
  __nv_bfloat16* dS_hi = spare + 2 * BWD_TILE * DP;
  __nv_bfloat16* dS_lo = dS_hi + BWD_TILE * TP;
  __nv_bfloat16* A_hi = dS_lo + BWD_TILE * TP;
  __nv_bfloat16* A_lo = A_hi + BWD_TILE * TP;
  float* stage_all = reinterpret_cast<float*>(A_lo + BWD_TILE * TP);
  (void)FP;

  const int num_kt = (p.Lk + BWD_TILE - 1) / BWD_TILE;
  const int n = blockIdx.x / num_kt, kt = blockIdx.x % num_kt;
  const int k0 = kt * BWD_TILE;
  const int krows = min(BWD_TILE, p.Lk - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp & 3) * 16, wh = warp >> 2;
  const bool odd16 = (p.d % 32) != 0;
  const int c0 = odd16 ? (wh ? (p.d / 2 + 8) : 0) : (p.d / 2) * wh;
  const int nf = odd16 ? (wh ? (p.d - (p.d / 2 + 8)) / 16 : (p.d / 2 + 8) / 16) : p.d / 32;

  constexpr int NFQ = BWD_DMAX / 32;
  FragC dv_acc[NFQ], dk_acc[NFQ];
#pragma unroll
  for (int f = 0; f < NFQ; ++f) {
    wmma::fill_fragment(dv_acc[f], 0.0f);
    wmma::fill_fragment(dk_acc[f], 0.0f);
  }
  const int num_qt = (p.Lq + BWD_TILE - 1) / BWD_TILE;
  for (int qt = 0; qt < num_qt; ++qt) {
    const int q0 = qt * BWD_TILE;
    const int qrows = min(BWD_TILE, p.Lq - q0);
    __syncthreads();
    bwd_stage_tile(p.dO + (static_cast<long long>(n) * p.Lq + q0) * p.d, p.d, qrows, p.d, BWD_TILE, p.d, dO_hi, dO_lo, DP);
    bwd_stage_tile(p.q + (static_cast<long long>(n) * p.Lq + q0) * p.d, p.d, qrows, p.d, BWD_TILE, p.d, Q_hi, Q_lo, DP);
    // A (after dropout) and dS tiles [64 q x 64 keys], row-major as stored; consumed TRANSPOSED below
    const long long g0 = (static_cast<long long>(n) * p.Lq + q0) * p.Lk + k0;
    bwd_stage_tile(p.A + g0, p.Lk, qrows, krows, BWD_TILE, BWD_TILE, A_hi, A_lo, TP);
    bwd_stage_tile(p.dS + g0, p.Lk, qrows, krows, BWD_TILE, BWD_TILE, dS_hi, dS_lo, TP);
    __syncthreads();
    // dV_j += A^T dO_i ; dK_j += dS^T Q_i : output rows = keys (wr), contraction over the 64 q rows
#define LAMP_BWD_DKV(NF_)                                                                                         \
  {                                                                                                               \
    FragC(&av)[NF_] = reinterpret_cast<FragC(&)[NF_]>(dv_acc);                                                    \
    FragC(&ak)[NF_] = reinterpret_cast<FragC(&)[NF_]>(dk_acc);                                                    \
    bwd_mma<NF_, true, false>(av, A_hi, A_lo, TP, wr, dO_hi, dO_lo, DP, c0, BWD_TILE);                            \
    bwd_mma<NF_, true, false>(ak, dS_hi, dS_lo, TP, wr, Q_hi, Q_lo, DP, c0, BWD_TILE);                            \
  }
    switch (nf) {
      case 4: LAMP_BWD_DKV(4) break;
      case 3: LAMP_BWD_DKV(3) break;
      case 2: LAMP_BWD_DKV(2) break;
      case 1: LAMP_BWD_DKV(1) break;
      default: break;
    }
#undef LAMP_BWD_DKV
  }
  __syncthreads();
  float* stage = stage_all + warp * (16 * 20);
  float* dv = p.dv + (static_cast<long long>(n) * p.Lk + k0) * p.d;
  float* dk = p.dk + (static_cast<long long>(n) * p.Lk + k0) * p.d;
  for (int which = 0; which < 2; ++which) {
    for (int f = 0; f < nf; ++f) {
      wmma::store_matrix_sync(stage, which ? dk_acc[f] : dv_acc[f], 20, wmma::mem_row_major);
      __syncwarp();
      float* dst = which ? dk : dv;
      for (int e = lane; e < 256; e += 32) {
        const int r = e >> 4, c = e & 15;
        if (wr + r < krows) dst[(wr + r) * p.d + c0 + 16 * f + c] = stage[r * 20 + c];
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------- probabilities
// Attention probabilities of the training forward (and of `return_attns=True`): P = exp2(q.k * scale - max) / sum from
// the operand planes and the row statistics the attention core saved, on the tensor cores (the warp-per-row
// attn_probs_kernel recomputes q.k with scalar FMAs and took 80 % of a training step).  One CTA per (h, b, 64-row q
// tile), loop over 64-key tiles; writes `probs` (after dropout, what the reference returns) and optionally
// `probs_pre` (before dropout, saved for the backward), both [H*B, Lq, Lk] head-major.
struct ProbsMmaParams {
  int B, H, Lq, Lk, d;
  float scale_log2;
  const __nv_bfloat16 *q_hi, *q_lo, *kv_hi, *kv_lo;  // lo nullable (bf16 mode)
  int ldq, ldkv, q_col0, k_col0, q_bcast;
  const uint8_t* mask;
  long long msb, msq, msk;
  const float *row_max, *row_sum;
  float* probs;
  float* probs_pre;
  uint32_t drop_thresh;
  float drop_scale;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;  // see AttnParams
};

// bf16 plane tile [rows x cols] (cols % 2 == 0, rows beyond rows_valid read as 0) -> smem (leading dim lds)
__device__ __forceinline__ void bwd_stage_planes(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                 long long ld, int rows_valid, int rows, int cols, __nv_bfloat16* dhi,
                                                 __nv_bfloat16* dlo, int lds) {
  for (int idx = threadIdx.x; idx < rows * (cols >> 1); idx += BWD_THREADS) {
    const int r = idx / (cols >> 1), c = (idx % (cols >> 1)) << 1;
    uint32_t h = 0u, l = 0u;
    if (r < rows_valid) {
      h = *reinterpret_cast<const uint32_t*>(hi + r * ld + c);
      if (lo != nullptr) l = *reinterpret_cast<const uint32_t*>(lo + r * ld + c);
    }
    *reinterpret_cast<uint32_t*>(dhi + r * lds + c) = h;
    *reinterpret_cast<uint32_t*>(dlo + r * lds + c) = l;
  }
}

__host__ __device__ constexpr size_t attn_probs_smem_bytes(int d) {
  return static_cast<size_t>(2) * 2 * BWD_TILE * bwd_dp(d) * 2 + BWD_TILE * (BWD_TILE + 4) * 4 + 2 * BWD_TILE * 4 + 128;
}

__global__ void __launch_bounds__(BWD_THREADS) attn_probs_mma_kernel(const ProbsMmaParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(128) uint8_t bsm[];
  const int DP = bwd_dp(p.d), FP = BWD_TILE + 4;
  __nv_bfloat16* Q_hi = reinterpret_cast<__nv_bfloat16*>(bsm);
  __nv_bfloat16* Q_lo = Q_hi + BWD_TILE * DP;
  __nv_bfloat16* K_hi = Q_lo + BWD_TILE * DP;
  __nv_bfloat16* K_lo = K_hi + BWD_TILE * DP;
  float* S = reinterpret_cast<float*>(K_lo + BWD_TILE * DP);
  float* rmx = S + BWD_TILE * FP;
  float* rinv = rmx + BWD_TILE;

  const int num_qt = (p.Lq + BWD_TILE - 1) / BWD_TILE;
  const int qt = blockIdx.x % num_qt;
  const int hb = blockIdx.x / num_qt;          // h * B + b (head-major)
  const int b = hb % p.B, h = hb / p.B;
  const int q0 = qt * BWD_TILE;
  const int qrows = min(BWD_TILE, p.Lq - q0);
  const int warp = threadIdx.x >> 5;
  const int wr = (warp & 3) * 16, wh = warp >> 2;
  const long long qrow0 = (p.q_bcast ? 0LL : static_cast<long long>(b) * p.Lq) + q0;
  bwd_stage_planes(p.q_hi + qrow0 * p.ldq + p.q_col0 + h * p.d, p.q_lo ? p.q_lo + qrow0 * p.ldq + p.q_col0 + h * p.d : nullptr,
                   p.ldq, qrows, BWD_TILE, p.d, Q_hi, Q_lo, DP);
  const long long stat0 = static_cast<long long>(hb) * p.Lq + q0;
  for (int r = threadIdx.x; r < BWD_TILE; r += BWD_THREADS) {
    rmx[r] = r < qrows ? p.row_max[stat0 + r] : 0.0f;
    rinv[r] = r < qrows ? 1.0f / p.row_sum[stat0 + r] : 0.0f;
  }
  const int num_kt = (p.Lk + BWD_TILE - 1) / BWD_TILE;
  for (int kt = 0; kt < num_kt; ++kt) {
    const int k0 = kt * BWD_TILE;
    const int krows = min(BWD_TILE, p.Lk - k0);
    __syncthreads();
    const long long krow0 = static_cast<long long>(b) * p.Lk + k0;
    bwd_stage_planes(p.kv_hi + krow0 * p.ldkv + p.k_col0 + h * p.d,
                     p.kv_lo ? p.kv_lo + krow0 * p.ldkv + p.k_col0 + h * p.d : nullptr, p.ldkv, krows, BWD_TILE, p.d, K_hi,
                     K_lo, DP);
    __syncthreads();
    FragC acc[2];
    wmma::fill_fragment(acc[0], 0.0f);
    wmma::fill_fragment(acc[1], 0.0f);
    bwd_mma<2, false, true>(acc, Q_hi, Q_lo, DP, wr, K_hi, K_lo, DP, 32 * wh, p.d);
    wmma::store_matrix_sync(S + wr * FP + 32 * wh, acc[0], FP, wmma::mem_row_major);
    wmma::store_matrix_sync(S + wr * FP + 32 * wh + 16, acc[1], FP, wmma::mem_row_major);
    __syncthreads();
    for (int idx = threadIdx.x; idx < BWD_TILE * BWD_TILE; idx += BWD_THREADS) {
      const int r = idx >> 6, c = idx & 63;
      if (r < qrows && c < krows) {
        const int i = q0 + r, j = k0 + c;
        bool masked = false;
        if (p.mask != nullptr)
          masked = p.mask[static_cast<long long>(b) * p.msb + static_cast<long long>(i) * p.msq +
                          static_cast<long long>(j) * p.msk] != 0;
        float pr = masked ? (0.0f * rinv[r]) : exp2f(S[r * FP + c] * p.scale_log2 - rmx[r]) * rinv[r];
        const long long g = (stat0 + r) * p.Lk + j;
        if (p.probs_pre != nullptr) p.probs_pre[g] = pr;
        if (p.drop_thresh) {
          const uint32_t rh = drop_rowhash(p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0ull),
                                           static_cast<unsigned long long>(stat0 + r));
          pr = drop_keep(rh, static_cast<uint32_t>(j), p.drop_thresh) ? pr * p.drop_scale : 0.0f;
        }
        p.probs[g] = pr;
      }
    }
  }
}

}  // namespace lamp
