// Masked attention core over label nodes:   O = softmax(mask(Q K^T / temperature)) V      (per sample, per head)
// Reference: lamp/SubLayers.py:27-43 (ScaledDotProductAttention.forward), called from :104.
//
// One persistent CTA per SM loops over work items (b, h, q-tile of 128 label rows) and, inside, over KV tiles:
//   warp 0   : TMA producer -- 3D tensor maps {cols, L, B}: rows >= L are zero-filled by the hardware, so ragged
//              label counts (L = 103, 159, 983 ...) need no padding in HBM.  Q/K/V arrive as 128B-swizzled
//              [rows x 64] bf16 boxes of the split-bf16 planes written by the projection GEMM.
//   warp 1   : MMA issuer   -- S = Q K^T  (A, B K-major) and O (+)= P V  (B = V is MN-major: V is consumed in its
//              natural [keys x d] layout, no transpose pass); 3-term split-bf16 products, fp32 accumulators in TMEM.
//   warps 2-5: softmax      -- TMEM lane == query row, so each thread owns a whole score row: mask + running
//              max/sum (online softmax across KV tiles) + exp2 + hi/lo split; P goes back to smem as the A operand
//              of the PV product; O is rescaled in TMEM when the running max moves; final 1/sum folded into the
//              epilogue which writes the head's slice of the concatenated output (no permute / contiguous copies).
// The label mask is read from its single [Lq, Lk] (or [B, Lk] key-padding) byte copy with arbitrary strides and
// turned into per-row bit words with warp ballots -- it is never tiled per head or per sample
// (the reference materialises H*B*Lq*Lk bytes, lamp/SubLayers.py:102 / lamp/Decoders.py:141).
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

struct AttnParams {
  int B, H, Lq, Lk, d;  // d = head width, multiple of 16, <= 128
  float scale_log2;     // log2(e) / temperature
  int q_col0, k_col0, v_col0;  // column of head 0 inside the Q / KV plane matrices
  int q_bcast;                 // 1: Q is shared by every sample (batch coordinate forced to 0)
  const uint8_t* mask;         // nullptr or bytes (non-zero = masked), element strides below
  long long msb, msq, msk;
  __nv_bfloat16* o_hi;  // [B*Lq, ldo] planes, head h at column h*d (nullable)
  __nv_bfloat16* o_lo;  // nullable
  int ldo;
  float* o_f32;  // optional fp32 copy, [B*Lq, ldof]
  int ldof;
  float* row_max;  // optional [H*B*Lq] (head-major) running max of the SCALED (log2 domain) scores
  float* row_sum;  // optional [H*B*Lq] softmax denominators
};

constexpr int ATTN_BLOCK_M = 128;
constexpr int ATTN_THREADS = 192;
constexpr uint32_t ATTN_TMEM_COLS = 256;  // S: [0,128)   O: [128,256)

// Shared-memory plan (bytes).  kb64 = ceil(d / 64) column blocks of 64 bf16 (= one 128 B swizzle row each).
template <int BLOCK_KV, bool ALIAS_PQ, int NTERMS>
struct AttnSmem {
  static constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  __host__ __device__ static constexpr uint32_t q_bytes(int kb64) { return NPL * kb64 * ATTN_BLOCK_M * 128; }
  __host__ __device__ static constexpr uint32_t kv_bytes(int kb64) { return NPL * kb64 * BLOCK_KV * 128; }
  __host__ __device__ static constexpr uint32_t p_bytes() { return NPL * (BLOCK_KV / 64) * ATTN_BLOCK_M * 128; }
  __host__ __device__ static constexpr uint32_t total(int kb64) {
    const uint32_t qp = ALIAS_PQ ? (q_bytes(kb64) > p_bytes() ? q_bytes(kb64) : p_bytes()) : q_bytes(kb64) + p_bytes();
    return qp + 2 * kv_bytes(kb64) + 1024 /*align*/ + 256 /*barriers*/;
  }
};

template <int BLOCK_KV, bool ALIAS_PQ, int NTERMS>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                 const __grid_constant__ CUtensorMap tmKV_hi, const __grid_constant__ CUtensorMap tmKV_lo,
                 const AttnParams p) {
  using SM = AttnSmem<BLOCK_KV, ALIAS_PQ, NTERMS>;
  constexpr int NPL = SM::NPL;
  constexpr int NW = BLOCK_KV / 32;  // mask words per row

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb64 = (p.d + 63) >> 6;
  const uint32_t q_bytes = SM::q_bytes(kb64), kv_bytes = SM::kv_bytes(kb64), p_bytes = SM::p_bytes();
  uint8_t* sQ = smem;
  uint8_t* sP = ALIAS_PQ ? smem : smem + q_bytes;
  const uint32_t qp_bytes = ALIAS_PQ ? (q_bytes > p_bytes ? q_bytes : p_bytes) : q_bytes + p_bytes;
  uint8_t* sK = smem + qp_bytes;
  uint8_t* sV = sK + kv_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kv_bytes);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  // plane pl (0 = hi, 1 = lo), 64-column block kb
  auto q_tile = [&](int pl, int kb) { return sQ + (pl * kb64 + kb) * (ATTN_BLOCK_M * 128); };
  auto k_tile = [&](int pl, int kb) { return sK + (pl * kb64 + kb) * (BLOCK_KV * 128); };
  auto v_tile = [&](int pl, int kb) { return sV + (pl * kb64 + kb) * (BLOCK_KV * 128); };
  auto p_tile = [&](int pl, int kb) { return sP + (pl * (BLOCK_KV / 64) + kb) * (ATTN_BLOCK_M * 128); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ_hi);
    tma_prefetch_desc(&tmKV_hi);
    if (NPL == 2) {
      tma_prefetch_desc(&tmQ_lo);
      tma_prefetch_desc(&tmKV_lo);
    }
    for (int i = 0; i < 9; ++i) mbar_init(&bars[i], i == 7 ? 128u : 1u);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATTN_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  const int num_qt = (p.Lq + ATTN_BLOCK_M - 1) / ATTN_BLOCK_M;
  const int num_kv = (p.Lk + BLOCK_KV - 1) / BLOCK_KV;
  const int num_items = p.B * p.H * num_qt;
  const int ksteps_d = p.d >> 4;  // UMMA K-steps over the head width

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0, kvc = 0;  // per-CTA counters -> barrier parities
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qt = item % num_qt;
        const int h = (item / num_qt) % p.H;
        const int b = item / (num_qt * p.H);
        for (int j = 0; j < num_kv; ++j, ++kvc) {
          mbar_wait(k_empty, (kvc & 1) ^ 1);
          mbar_arrive_expect_tx(k_full, kv_bytes);
          for (int kb = 0; kb < kb64; ++kb) {
            tma_load_3d(k_tile(0, kb), &tmKV_hi, k_full, p.k_col0 + h * p.d + kb * 64, j * BLOCK_KV, b);
            if (NPL == 2) tma_load_3d(k_tile(1, kb), &tmKV_lo, k_full, p.k_col0 + h * p.d + kb * 64, j * BLOCK_KV, b);
          }
          if (j == 0) {
            mbar_wait(q_empty, (it & 1) ^ 1);
            mbar_arrive_expect_tx(q_full, q_bytes);
            const int bq = p.q_bcast ? 0 : b;
            for (int kb = 0; kb < kb64; ++kb) {
              tma_load_3d(q_tile(0, kb), &tmQ_hi, q_full, p.q_col0 + h * p.d + kb * 64, qt * ATTN_BLOCK_M, bq);
              if (NPL == 2)
                tma_load_3d(q_tile(1, kb), &tmQ_lo, q_full, p.q_col0 + h * p.d + kb * 64, qt * ATTN_BLOCK_M, bq);
            }
          }
          mbar_wait(v_empty, (kvc & 1) ^ 1);
          mbar_arrive_expect_tx(v_full, kv_bytes);
          for (int kb = 0; kb < kb64; ++kb) {
            tma_load_3d(v_tile(0, kb), &tmKV_hi, v_full, p.v_col0 + h * p.d + kb * 64, j * BLOCK_KV, b);
            if (NPL == 2) tma_load_3d(v_tile(1, kb), &tmKV_lo, v_full, p.v_col0 + h * p.d + kb * 64, j * BLOCK_KV, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(ATTN_BLOCK_M, BLOCK_KV, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(ATTN_BLOCK_M, p.d, 0, 1);  // B = V is MN-major
      uint32_t it = 0, kvc = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        for (int j = 0; j < num_kv; ++j, ++kvc) {
          const bool last = (j == num_kv - 1);
          if (j == 0) mbar_wait(q_full, it & 1);
          mbar_wait(k_full, kvc & 1);
          tcgen05_fence_after();
          // S = Q K^T : both operands K-major (d contiguous), SBO = 8 rows * 128 B
          for (int t = 0; t < ksteps_d; ++t) {
            const int kb = t >> 2;
            const uint32_t koff = (t & 3) * 32;
            const uint64_t dq_hi = umma_smem_desc(smem_u32(q_tile(0, kb)) + koff, 16, 1024);
            const uint64_t dk_hi = umma_smem_desc(smem_u32(k_tile(0, kb)) + koff, 16, 1024);
            umma_bf16_ss(tmem_S, dq_hi, dk_hi, idesc_s, t != 0 ? 1u : 0u);
            if (NTERMS == 3) {
              const uint64_t dq_lo = umma_smem_desc(smem_u32(q_tile(1, kb)) + koff, 16, 1024);
              const uint64_t dk_lo = umma_smem_desc(smem_u32(k_tile(1, kb)) + koff, 16, 1024);
              umma_bf16_ss(tmem_S, dq_hi, dk_lo, idesc_s, 1u);
              umma_bf16_ss(tmem_S, dq_lo, dk_hi, idesc_s, 1u);
            }
          }
          umma_commit(s_full);
          umma_commit(k_empty);
          if (last && !ALIAS_PQ) umma_commit(q_empty);
          // O (+)= P V_j
          mbar_wait(p_full, kvc & 1);
          mbar_wait(v_full, kvc & 1);
          tcgen05_fence_after();
          const int kv_valid = min(BLOCK_KV, p.Lk - j * BLOCK_KV);
          const int ksteps_kv = (kv_valid + 15) >> 4;
          for (int t = 0; t < ksteps_kv; ++t) {
            // A = P: K-major, 64-column blocks; B = V: MN-major, K (= key index) advances by 16 rows of 128 B,
            // LBO = distance between the 64-wide d blocks, SBO = 8 key rows.
            const uint32_t pa = (t & 3) * 32;
            const uint64_t dp_hi = umma_smem_desc(smem_u32(p_tile(0, t >> 2)) + pa, 16, 1024);
            const uint64_t dv_hi = umma_smem_desc(smem_u32(v_tile(0, 0)) + t * 2048, BLOCK_KV * 128, 1024);
            const uint32_t accum = (j != 0 || t != 0) ? 1u : 0u;
            umma_bf16_ss(tmem_O, dp_hi, dv_hi, idesc_o, accum);
            if (NTERMS == 3) {
              const uint64_t dp_lo = umma_smem_desc(smem_u32(p_tile(1, t >> 2)) + pa, 16, 1024);
              const uint64_t dv_lo = umma_smem_desc(smem_u32(v_tile(1, 0)) + t * 2048, BLOCK_KV * 128, 1024);
              umma_bf16_ss(tmem_O, dp_hi, dv_lo, idesc_o, 1u);
              umma_bf16_ss(tmem_O, dp_lo, dv_hi, idesc_o, 1u);
            }
          }
          umma_commit(o_done);
          umma_commit(v_empty);
          if (last && ALIAS_PQ) umma_commit(q_empty);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 2..5)
    const int wq = warp & 3;
    const int row = wq * 32 + lane;  // row inside the q tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
    uint32_t mw[NW];  // mask words of this thread's row (bit set = masked)
    long long mkey = -1;  // (b, qt, j) combination the words were built for
    uint32_t kvc = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int qt = item % num_qt;
      const int h = (item / num_qt) % p.H;
      const int b = item / (num_qt * p.H);
      const int qrow = qt * ATTN_BLOCK_M + row;
      float m_run = -INFINITY, l_run = 0.0f;
      for (int j = 0; j < num_kv; ++j, ++kvc) {
        const int k0 = j * BLOCK_KV;
        // ---- mask words for (row, this KV tile); cached while the addressed mask region is unchanged
        {
          const long long key = (p.mask == nullptr)
                                    ? static_cast<long long>(j)
                                    : ((p.msb ? static_cast<long long>(b) : 0) * num_qt + (p.msq ? qt : 0)) * num_kv + j;
          if (key != mkey) {
            mkey = key;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
              const int rem = p.Lk - (k0 + 32 * w);  // valid columns in this word
              mw[w] = rem >= 32 ? 0u : (rem <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << rem));
            }
            if (p.mask != nullptr) {
              const uint8_t* mb = p.mask + static_cast<long long>(b) * p.msb;
              const int nrows = p.msq ? 32 : 1;
              for (int rr = 0; rr < nrows; ++rr) {
                const int qr = qt * ATTN_BLOCK_M + wq * 32 + rr;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                  const int kc = k0 + 32 * w + lane;
                  uint32_t byte = 0;
                  if (kc < p.Lk && qr < p.Lq) byte = mb[static_cast<long long>(qr) * p.msq + static_cast<long long>(kc) * p.msk];
                  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, byte != 0);
                  if (!p.msq || lane == rr) mw[w] |= bal;
                }
              }
            }
          }
        }
        mbar_wait(s_full, kvc & 1);
        tcgen05_fence_after();
        // ---- pass 1: row maximum of the masked scores (raw, unscaled; scale > 0)
        float mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          uint32_t r[32];
          tmem_ld32(tmem_S + lane_sel + 32 * w, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((mw[w] >> e) & 1u)) mx = fmaxf(mx, __uint_as_float(r[e]));
        }
        const float m_new = fmaxf(m_run, mx);
        const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
        const float alpha = (m_new == -INFINITY) ? 1.0f : exp2f((m_run - m_new) * p.scale_log2);
        // ---- rescale the running output when the maximum moved (needs PV of the previous tile retired)
        if (j > 0) {
          mbar_wait(o_done, (kvc - 1) & 1);
          tcgen05_fence_after();
          for (int c = 0; c < p.d; c += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_O + lane_sel + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
            tmem_st16(tmem_O + lane_sel + c, r);
          }
          tmem_wait_st();
        }
        // ---- pass 2: P = exp2((s - m) * scale), hi/lo split, swizzled K-major store for the PV product
        float sum = 0.0f;
        const float mb2 = m_use * p.scale_log2;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          uint32_t r[32];
          tmem_ld32(tmem_S + lane_sel + 32 * w, r);
          tmem_wait_ld();
          float pv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float x = exp2f(__uint_as_float(r[e]) * p.scale_log2 - mb2);
            pv[e] = ((mw[w] >> e) & 1u) ? 0.0f : x;
            sum += pv[e];
          }
          // columns [32w, 32w+32) -> 64-column block (w>>1), 16-byte chunks (4*(w&1) .. +4), XOR-swizzled by row
          uint8_t* ph = p_tile(0, w >> 1) + row * 128;
          uint8_t* pl = p_tile(1, w >> 1) + row * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 hi, lo;
            split_bf16x2(pv[8 * c + 0], pv[8 * c + 1], hi.x, lo.x);
            split_bf16x2(pv[8 * c + 2], pv[8 * c + 3], hi.y, lo.y);
            split_bf16x2(pv[8 * c + 4], pv[8 * c + 5], hi.z, lo.z);
            split_bf16x2(pv[8 * c + 6], pv[8 * c + 7], hi.w, lo.w);
            const uint32_t chunk = static_cast<uint32_t>((4 * (w & 1) + c) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(ph + chunk) = hi;
            if (NPL == 2) *reinterpret_cast<uint4*>(pl + chunk) = lo;
          }
        }
        l_run = l_run * alpha + sum;
        m_run = m_new;
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core
        tcgen05_fence_before();
        mbar_arrive(p_full);
      }
      // ---- epilogue: O / l -> this head's column slice of the concatenated output
      mbar_wait(o_done, (kvc - 1) & 1);
      tcgen05_fence_after();
      const float inv = 1.0f / l_run;  // l == 0 (row fully masked) -> inf -> NaN, as the reference's softmax of all -inf
      const bool row_ok = qrow < p.Lq;
      const size_t grow = static_cast<size_t>(b) * p.Lq + qrow;
      for (int c = 0; c < p.d; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_O + lane_sel + c, r);
        tmem_wait_ld();
        if (row_ok) {
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) * inv;
          const int col = h * p.d + c;
          if (p.o_f32 != nullptr) {
            float4* o = reinterpret_cast<float4*>(p.o_f32 + grow * p.ldof + col);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          }
          if (p.o_hi != nullptr) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 hi, lo;
              split_bf16x2(v[8 * g + 0], v[8 * g + 1], hi.x, lo.x);
              split_bf16x2(v[8 * g + 2], v[8 * g + 3], hi.y, lo.y);
              split_bf16x2(v[8 * g + 4], v[8 * g + 5], hi.z, lo.z);
              split_bf16x2(v[8 * g + 6], v[8 * g + 7], hi.w, lo.w);
              const size_t off = grow * p.ldo + col + 8 * g;
              *reinterpret_cast<uint4*>(p.o_hi + off) = hi;
              if (p.o_lo != nullptr) *reinterpret_cast<uint4*>(p.o_lo + off) = lo;
            }
          }
        }
      }
      if (row_ok && p.row_sum != nullptr) {
        const size_t si = (static_cast<size_t>(h) * p.B + b) * p.Lq + qrow;
        p.row_max[si] = ((m_run == -INFINITY) ? 0.0f : m_run) * p.scale_log2;
        p.row_sum[si] = l_run;
      }
      tcgen05_fence_before();  // O reads retired before the next item's PV (ordered through p_full)
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, ATTN_TMEM_COLS);
  }
}

// Attention probabilities for `return_attns=True` (diagnostic path; lamp/SubLayers.py:121 returns them to the
// caller).  Recomputes s = q.k from the planes with fp32 FMAs and normalises with the row statistics saved by
// attn_core_kernel:  P[h*B+b, i, j] = exp2(s*scale - max_i) / sum_i  (head-major batch, lamp/SubLayers.py:96-98).
struct ProbsParams {
  int B, H, Lq, Lk, d;
  float scale_log2;
  const __nv_bfloat16 *q_hi, *q_lo, *kv_hi, *kv_lo;  // lo nullable
  int ldq, ldkv, q_col0, k_col0, q_bcast;
  const uint8_t* mask;
  long long msb, msq, msk;
  const float *row_max, *row_sum;
  float* probs;  // [H*B, Lq, Lk]
};

__global__ void attn_probs_kernel(const ProbsParams p) {
  // one warp per (h, b, i) row; lanes stride over keys
  const long long warp_g = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long nrows = static_cast<long long>(p.H) * p.B * p.Lq;
  if (warp_g >= nrows) return;
  const int i = static_cast<int>(warp_g % p.Lq);
  const int b = static_cast<int>((warp_g / p.Lq) % p.B);
  const int h = static_cast<int>(warp_g / (static_cast<long long>(p.Lq) * p.B));
  const size_t qrow = (p.q_bcast ? 0 : static_cast<size_t>(b) * p.Lq) + i;
  const __nv_bfloat16* qh = p.q_hi + qrow * p.ldq + p.q_col0 + h * p.d;
  const __nv_bfloat16* ql = p.q_lo ? p.q_lo + qrow * p.ldq + p.q_col0 + h * p.d : nullptr;
  const float mx = p.row_max[warp_g], inv = 1.0f / p.row_sum[warp_g];
  float* out = p.probs + warp_g * p.Lk;
  for (int j = lane; j < p.Lk; j += 32) {
    const size_t krow = static_cast<size_t>(b) * p.Lk + j;
    const __nv_bfloat16* kh = p.kv_hi + krow * p.ldkv + p.k_col0 + h * p.d;
    const __nv_bfloat16* kl = p.kv_lo ? p.kv_lo + krow * p.ldkv + p.k_col0 + h * p.d : nullptr;
    float s = 0.0f;
    for (int c = 0; c < p.d; ++c) {
      const float qa = __bfloat162float(qh[c]), ka = __bfloat162float(kh[c]);
      float t = qa * ka;
      if (ql != nullptr) t += qa * __bfloat162float(kl[c]) + __bfloat162float(ql[c]) * ka;
      s += t;
    }
    bool masked = false;
    if (p.mask != nullptr)
      masked = p.mask[static_cast<long long>(b) * p.msb + static_cast<long long>(i) * p.msq +
                      static_cast<long long>(j) * p.msk] != 0;
    out[j] = masked ? (0.0f * inv) : exp2f(s * p.scale_log2 - mx) * inv;
  }
}

}  // namespace lamp
