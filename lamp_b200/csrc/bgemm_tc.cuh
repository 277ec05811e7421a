// Batched small GEMM on tcgen05 for the attention backward:  C_b[M, N] = opA(A_b) opB(B_b)  for b < batch, every
// operand given as split-bf16 planes read in place by 4-D TMA maps {cols, rows, head, sample} (rows beyond the
// per-batch bound are zero-filled, so ragged label counts need no padding).  The batch index is two-level,
// batch = sample * H + head, with independent element strides for the two coordinates: the same kernel reads
// head-major [H*B, L, d] tensors (the reference's attention layout) and head COLUMN SLICES of the [B*L, H*d] plane
// matrices the projection GEMMs produce / consume -- so the training path needs no permute / contiguous copies.  Each operand is either K-major (the contraction
// index is the contiguous one: [rows = M or N, cols = Kc]) or MN-major ([rows = Kc, cols = M or N]) -- the four
// products of the backward differ only in that:
//     dA = dO V^T     (A: dO K-major,  B: V  K-major)         dQ = dS K      (A: dS K-major,  B: K  MN-major)
//     dV = A^T dO     (A: A  MN-major, B: dO MN-major)        dK = dS^T Q    (A: dS MN-major, B: Q  MN-major)
// Persistent: one CTA per SM walks (batch, 128-row m tile, TN-column n tile) work items; contraction in 64-wide chunks
// through a TMA ring that keeps running across items; 3-term split-bf16 products into one of two TMEM accumulators,
// so the epilogue of an item overlaps the MMAs of the next one.  Epilogue: tcgen05.ld (lane == output row) -> scale ->
// fp32 or split-bf16 -> the warp's 128B-swizzled staging boxes in shared memory -> 4-D TMA stores through output maps
// with the same {cols, rows, head, sample} indexing (rows / columns outside one (sample, head) matrix are clipped by
// the hardware).  (History: one CTA per item -- TMEM allocation and barrier set-up dominated, tensor pipe 7-8 % active;
// then persistent CTAs with per-row scalar stores -- the four epilogue warps executed ~1000 instructions per tile, most
// of them 2- and 4-byte stores with 64-bit address arithmetic, and were the critical path at ~10 K clocks per tile.)
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

constexpr int BG_THREADS = 192;
constexpr int BG_KC = 64;  // contraction chunk
__host__ __device__ constexpr uint32_t bg_stage_bytes(int npl, int tn) { return npl * (128 + tn) * BG_KC * 2; }
__host__ __device__ constexpr int bg_stages(int npl, int tn) { return (192 * 1024) / bg_stage_bytes(npl, tn); }
constexpr uint32_t BG_EPI_WARP = 8192;                  // per epilogue warp: two [32 rows x 128 B] swizzled boxes
constexpr uint32_t BG_EPI_STAGING = 4 * BG_EPI_WARP;
__host__ __device__ constexpr uint32_t bg_smem_bytes(int npl, int tn) {
  return (bg_stages(npl, tn) > 4 ? 4 : bg_stages(npl, tn)) * bg_stage_bytes(npl, tn) + BG_EPI_STAGING + 1024 + 256;
}

struct BgemmParams {
  int batch, H, M, N, Kc;   // batch = samples * H; item -> (sample = batch / H, head = batch % H)
  float scale;              // C = scale * (A B)
  int out_f32;              // 1: tmC0 is an fp32 map (box {32 cols, 32 rows}); 0: tmC0 / tmC1 are the hi / lo plane maps
  int out_lo;               // planes: the lo plane is written as well
};

template <bool A_MN, bool B_MN, int NTERMS, int TN>
__global__ void __launch_bounds__(BG_THREADS, 1)
bgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1, const BgemmParams p) {
  constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  constexpr int STAGES = bg_stages(NPL, TN) > 4 ? 4 : bg_stages(NPL, TN);
  constexpr uint32_t A_BYTES = 128 * BG_KC * 2, B_BYTES = TN * BG_KC * 2;
  constexpr uint32_t STAGE_BYTES = NPL * (A_BYTES + B_BYTES);
  constexpr uint32_t BOX = 64 * 64 * 2;  // one [64 x 64] box of an MN-major operand
  static_assert(STAGES >= 2, "ring too shallow");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + BG_EPI_STAGING);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * STAGES + 2;  // [2], count 128 (the epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  auto a_tile = [&](int s, int pl) { return smem + s * STAGE_BYTES + pl * (A_BYTES + B_BYTES); };
  auto b_tile = [&](int s, int pl) { return smem + s * STAGE_BYTES + pl * (A_BYTES + B_BYTES) + A_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + 127) / 128, tiles_n = (p.N + TN - 1) / TN;
  const int per_batch = tiles_m * tiles_n;
  const long long num_items = static_cast<long long>(p.batch) * per_batch;
  const int num_it = (p.Kc + BG_KC - 1) / BG_KC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmC0);
    if (NPL == 2) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * TN);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int bh = static_cast<int>(item / per_batch), tile = static_cast<int>(item % per_batch);
        const int b = bh / p.H, h = bh % p.H;
        const int m0 = (tile / tiles_n) * 128, n0 = (tile % tiles_n) * TN;
        for (int it = 0; it < num_it; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int kc0 = it * BG_KC;
          for (int pl = 0; pl < NPL; ++pl) {
            const CUtensorMap* ta = pl ? &tmA_lo : &tmA_hi;
            const CUtensorMap* tb = pl ? &tmB_lo : &tmB_hi;
            if (A_MN) {  // map {M cols, Kc rows, head, sample}, boxes {64, 64}
              for (int bx = 0; bx < 2; ++bx) tma_load_4d(a_tile(stage, pl) + bx * BOX, ta, &full_bar[stage], m0 + 64 * bx, kc0, h, b);
            } else {     // map {Kc cols, M rows, head, sample}, box {64, 128}
              tma_load_4d(a_tile(stage, pl), ta, &full_bar[stage], kc0, m0, h, b);
            }
            if (B_MN) {
              for (int bx = 0; bx < TN / 64; ++bx) tma_load_4d(b_tile(stage, pl) + bx * BOX, tb, &full_bar[stage], n0 + 64 * bx, kc0, h, b);
            } else {
              tma_load_4d(b_tile(stage, pl), tb, &full_bar[stage], kc0, n0, h, b);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, TN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0, t = 0;
      for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
        const uint32_t acc = t & 1, use = t >> 1;
        mbar_wait(&acc_empty[acc], (use & 1) ^ 1);  // the epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TN;
        for (int it = 0; it < num_it; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
#pragma unroll
          for (int k = 0; k < BG_KC / 16; ++k) {
            // K-major: 16 contraction elements = 32 B inside the swizzled 128 B row (SBO = 8 rows);
            // MN-major: 16 contraction rows = 2 KB inside every [64 x 64] box (LBO = box stride, SBO = 8 rows)
            auto desc = [&](uint8_t* base, bool mn) {
              return mn ? umma_smem_desc(smem_u32(base) + k * 2048, BOX, 1024) : umma_smem_desc(smem_u32(base) + k * 32, 16, 1024);
            };
            const uint64_t da_hi = desc(a_tile(stage, 0), A_MN), db_hi = desc(b_tile(stage, 0), B_MN);
            umma_bf16_ss(d_tmem, da_hi, db_hi, idesc, (it | k) != 0 ? 1u : 0u);
            if (NTERMS == 3) {
              const uint64_t da_lo = desc(a_tile(stage, 1), A_MN), db_lo = desc(b_tile(stage, 1), B_MN);
              umma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
              umma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    const int wq = warp & 3;
    const uint32_t stg = smem_u32(epi_stage) + (warp - 2) * BG_EPI_WARP;
    const uint32_t srow = stg + lane * 128, sx = lane & 7;   // this lane's row in a box; chunk c sits at c ^ (row & 7)
    uint32_t t = 0;
    for (long long item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
      const int bh = static_cast<int>(item / per_batch), tile = static_cast<int>(item % per_batch);
      const int b = bh / p.H, h = bh % p.H;
      const int m0 = (tile / tiles_n) * 128, n0 = (tile % tiles_n) * TN;
      const uint32_t acc = t & 1, use = t >> 1;
      mbar_wait(&acc_full[acc], use & 1);
      tcgen05_fence_after();
      const int r0 = m0 + wq * 32;          // first output row of this warp
      const bool store = r0 < p.M;          // (rows of the box beyond M are clipped by the TMA unit)
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 64) {
        if (n0 + c0 >= p.N) break;
        uint32_t ra[32], rb[32];
        const uint32_t taddr = tmem_base + acc * TN + (static_cast<uint32_t>(wq * 32) << 16) + c0;
        tmem_ld32(taddr, ra);
        tmem_ld32(taddr + 32, rb);
        tmem_wait_ld();
        if (c0 + 64 >= TN || n0 + c0 + 64 >= p.N) {   // last read of this accumulator: hand it back to the MMA warp
          tcgen05_fence_before();
          mbar_arrive(&acc_empty[acc]);
        }
        if (lane == 0) tma_store_wait_read0();        // the previous stores have read the staging boxes
        __syncwarp();
        if (p.out_f32) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 v0, v1;
            v0.x = __float_as_uint(__uint_as_float(ra[4 * c]) * p.scale);     v0.y = __float_as_uint(__uint_as_float(ra[4 * c + 1]) * p.scale);
            v0.z = __float_as_uint(__uint_as_float(ra[4 * c + 2]) * p.scale); v0.w = __float_as_uint(__uint_as_float(ra[4 * c + 3]) * p.scale);
            v1.x = __float_as_uint(__uint_as_float(rb[4 * c]) * p.scale);     v1.y = __float_as_uint(__uint_as_float(rb[4 * c + 1]) * p.scale);
            v1.z = __float_as_uint(__uint_as_float(rb[4 * c + 2]) * p.scale); v1.w = __float_as_uint(__uint_as_float(rb[4 * c + 3]) * p.scale);
            sts128(srow + ((c ^ sx) << 4), v0);           // box 0: columns c0 .. c0 + 31
            sts128(srow + 4096 + ((c ^ sx) << 4), v1);    // box 1: columns c0 + 32 .. c0 + 63
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {   // 8 columns = one 16-byte chunk of a plane row; columns 8c .. 8c + 7 of the 64
            const uint32_t* src = c < 4 ? ra + 8 * c : rb + 8 * (c - 4);
            uint4 hv, lv;
            split_bf16x2(__uint_as_float(src[0]) * p.scale, __uint_as_float(src[1]) * p.scale, hv.x, lv.x);
            split_bf16x2(__uint_as_float(src[2]) * p.scale, __uint_as_float(src[3]) * p.scale, hv.y, lv.y);
            split_bf16x2(__uint_as_float(src[4]) * p.scale, __uint_as_float(src[5]) * p.scale, hv.z, lv.z);
            split_bf16x2(__uint_as_float(src[6]) * p.scale, __uint_as_float(src[7]) * p.scale, hv.w, lv.w);
            sts128(srow + ((c ^ sx) << 4), hv);
            if (p.out_lo) sts128(srow + 4096 + ((c ^ sx) << 4), lv);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && store) {
          if (p.out_f32) {
            tma_store_4d(&tmC0, stg, n0 + c0, r0, h, b);
            if (n0 + c0 + 32 < p.N) tma_store_4d(&tmC0, stg + 4096, n0 + c0 + 32, r0, h, b);
          } else {
            tma_store_4d(&tmC0, stg, n0 + c0, r0, h, b);
            if (p.out_lo) tma_store_4d(&tmC1, stg + 4096, n0 + c0, r0, h, b);
          }
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all0();   // global writes performed before the CTA (and its shared memory) goes away
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 2 * TN);
  }
}

// dS = P o (dA * keep / (1 - p) - delta) / temperature  ->  split-bf16 planes [N*Lq, ld] (ld = Lk rounded up to 8; the
// pad columns are written as 0), and A (attn after dropout) -> planes, both operands of the dQ / dK / dV products.
// dA (and S in the recompute form) are workspace tensors with row pitch ld as well (TMA-stored by the batched product:
// the pitch must be a multiple of 16 bytes); P and A are the caller's dense [N, Lq, Lk] tensors.
// delta_i = <dO_i, O_i> is computed by the same block (one warp per row) before the row is swept.
__global__ void attn_bwd_ds_kernel(const float* __restrict__ dA, const float* __restrict__ P, const float* __restrict__ A,
                                   const float* __restrict__ dO, const float* __restrict__ O, long long rows, int Lk,
                                   int d, int ld, float inv_temp, float drop_scale, __nv_bfloat16* __restrict__ dS_hi,
                                   __nv_bfloat16* __restrict__ dS_lo, __nv_bfloat16* __restrict__ A_hi,
                                   __nv_bfloat16* __restrict__ A_lo) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float delta = 0.f;
  for (int c = lane; c < d; c += 32) delta += dO[row * d + c] * O[row * d + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xFFFFFFFFu, delta, o);
  for (int c = lane; c < ld; c += 32) {
    float ds = 0.f, a = 0.f;
    if (c < Lk) {
      const long long g = row * Lk + c;
      a = A[g];
      const float keep = (a != 0.0f) ? drop_scale : 0.0f;
      ds = P[g] * (dA[row * ld + c] * keep - delta) * inv_temp;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(ds);
    dS_hi[row * ld + c] = h;
    dS_lo[row * ld + c] = __float2bfloat16_rn(ds - __bfloat162float(h));
    const __nv_bfloat16 ah = __float2bfloat16_rn(a);
    A_hi[row * ld + c] = ah;
    A_lo[row * ld + c] = __float2bfloat16_rn(a - __bfloat162float(ah));
  }
}

// The same for the training path that keeps everything in the projection GEMMs' layouts: dO and O are split-bf16 planes
// [B*Lq, ld_o] whose head h lives in columns [h*d, (h+1)*d); P / A / dA stay head-major [H*B, Lq, Lk] (row index
// (h*B + b)*Lq + i), which is also the layout of the dS / A planes written here.
__global__ void attn_bwd_ds_planes_kernel(const float* __restrict__ dA, const float* __restrict__ P,
                                          const float* __restrict__ A, const __nv_bfloat16* __restrict__ dO_hi,
                                          const __nv_bfloat16* __restrict__ dO_lo, const __nv_bfloat16* __restrict__ O_hi,
                                          const __nv_bfloat16* __restrict__ O_lo, long long ld_o, int B, int H, int Lq,
                                          int Lk, int d, int ld, float inv_temp, float drop_scale,
                                          __nv_bfloat16* __restrict__ dS_hi, __nv_bfloat16* __restrict__ dS_lo,
                                          __nv_bfloat16* __restrict__ A_hi, __nv_bfloat16* __restrict__ A_lo) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(H) * B * Lq;
  if (row >= rows) return;
  const int i = static_cast<int>(row % Lq);
  const long long n = row / Lq;
  const int b = static_cast<int>(n % B), h = static_cast<int>(n / B);
  const long long orow = (static_cast<long long>(b) * Lq + i) * ld_o + static_cast<long long>(h) * d;
  float delta = 0.f;
  for (int c = lane; c < d; c += 32) {
    float g = __bfloat162float(dO_hi[orow + c]), o = __bfloat162float(O_hi[orow + c]);
    if (dO_lo != nullptr) g += __bfloat162float(dO_lo[orow + c]);
    if (O_lo != nullptr) o += __bfloat162float(O_lo[orow + c]);
    delta = fmaf(g, o, delta);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xFFFFFFFFu, delta, o);
  for (int c = lane; c < ld; c += 32) {
    float ds = 0.f, a = 0.f;
    if (c < Lk) {
      const long long g = row * Lk + c;
      a = A[g];
      const float keep = (a != 0.0f) ? drop_scale : 0.0f;
      ds = P[g] * (dA[row * ld + c] * keep - delta) * inv_temp;
    }
    const __nv_bfloat16 hh = __float2bfloat16_rn(ds);
    dS_hi[row * ld + c] = hh;
    dS_lo[row * ld + c] = __float2bfloat16_rn(ds - __bfloat162float(hh));
    const __nv_bfloat16 ah = __float2bfloat16_rn(a);
    A_hi[row * ld + c] = ah;
    A_lo[row * ld + c] = __float2bfloat16_rn(a - __bfloat162float(ah));
  }
}

// Recompute form (no probability tensors kept by the forward): P is rebuilt from the raw scores S = Q K^T (one more
// batched product), the row statistics the forward saved (reference maximum in the scaled log2 domain, denominator), the
// mask and the dropout counter hash -- exactly the expression of the probability kernel (attn_probs_mma_kernel).
struct DsRecomputeParams {
  const float* S;        // [H*B, Lq, Lk] raw scores
  const float* dA;       // [H*B, Lq, Lk]
  const float* row_max;  // [H*B*Lq]
  const float* row_sum;
  const uint8_t* mask;   // nullable; element (b, i, j) at b*msb + i*msq + j*msk, non-zero = masked
  long long msb, msq, msk;
  float scale_log2, inv_temp, drop_scale;
  uint32_t drop_thresh;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  const __nv_bfloat16 *dO_hi, *dO_lo, *O_hi, *O_lo;
  long long ld_o;
  int B, H, Lq, Lk, d, ld;
  __nv_bfloat16 *dS_hi, *dS_lo, *A_hi, *A_lo;
};
// One warp per (head, sample, query) row, four consecutive keys per lane and step: 16-byte loads of S and dA, 8-byte
// stores of the four plane rows.  (The scalar version executed ~107 thread instructions per element -- 2-byte stores
// and 64-bit address arithmetic per array -- and was instruction-bound at 32 % of DRAM.)
__global__ void __launch_bounds__(256) attn_bwd_ds_recompute_kernel(const DsRecomputeParams p) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = static_cast<long long>(p.H) * p.B * p.Lq;
  if (row >= rows) return;
  const int i = static_cast<int>(row % p.Lq);
  const long long n = row / p.Lq;
  const int b = static_cast<int>(n % p.B), h = static_cast<int>(n / p.B);
  const long long orow = (static_cast<long long>(b) * p.Lq + i) * p.ld_o + static_cast<long long>(h) * p.d;
  float delta = 0.f;
  for (int c = lane * 4; c < p.d; c += 128) {   // d is a multiple of 16, the planes are 16-byte aligned per head slice
    const uint2 gh = __ldg(reinterpret_cast<const uint2*>(p.dO_hi + orow + c)), gl = __ldg(reinterpret_cast<const uint2*>(p.dO_lo + orow + c));
    const uint2 oh = __ldg(reinterpret_cast<const uint2*>(p.O_hi + orow + c)), ol = __ldg(reinterpret_cast<const uint2*>(p.O_lo + orow + c));
    const uint32_t g4[4] = {gh.x, gh.y, gl.x, gl.y}, o4[4] = {oh.x, oh.y, ol.x, ol.y};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float g0 = __uint_as_float(g4[j] << 16) + __uint_as_float(g4[j + 2] << 16);
      const float g1 = __uint_as_float(g4[j] & 0xFFFF0000u) + __uint_as_float(g4[j + 2] & 0xFFFF0000u);
      const float o0 = __uint_as_float(o4[j] << 16) + __uint_as_float(o4[j + 2] << 16);
      const float o1 = __uint_as_float(o4[j] & 0xFFFF0000u) + __uint_as_float(o4[j + 2] & 0xFFFF0000u);
      delta = fmaf(g0, o0, delta);
      delta = fmaf(g1, o1, delta);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xFFFFFFFFu, delta, o);
  const float rmx = p.row_max[row], rinv = 1.0f / p.row_sum[row];
  const uint32_t rh = p.drop_thresh ? drop_rowhash(p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0ull),
                                                   static_cast<unsigned long long>(row))
                                    : 0u;
  const uint8_t* mrow = p.mask ? p.mask + static_cast<long long>(b) * p.msb + static_cast<long long>(i) * p.msq : nullptr;
  const float* Srow = p.S + row * p.ld;
  const float* dArow = p.dA + row * p.ld;
  __nv_bfloat16* o_dsh = p.dS_hi + row * p.ld;
  __nv_bfloat16* o_dsl = p.dS_lo + row * p.ld;
  __nv_bfloat16* o_ah = p.A_hi + row * p.ld;
  __nv_bfloat16* o_al = p.A_lo + row * p.ld;
  for (int c = lane * 4; c < p.ld; c += 128) {
    float ds[4] = {0.f, 0.f, 0.f, 0.f}, a[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < p.Lk) {
      const float4 s4 = *reinterpret_cast<const float4*>(Srow + c), d4 = *reinterpret_cast<const float4*>(dArow + c);
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j < p.Lk) {   // (columns Lk .. ld-1 of S / dA were clipped by the store: never read)
          const bool masked = mrow != nullptr && mrow[static_cast<long long>(c + j) * p.msk] != 0;
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(sv[j], p.scale_log2, -rmx)));
          const float pr = masked ? (0.0f * rinv) : e * rinv;
          const float keep = (!p.drop_thresh || drop_keep(rh, static_cast<uint32_t>(c + j), p.drop_thresh)) ? p.drop_scale : 0.0f;
          a[j] = pr * keep;
          ds[j] = pr * (dv[j] * keep - delta) * p.inv_temp;
        }
      }
    }
    uint2 dh, dl, ah, al;
    split_bf16x2(ds[0], ds[1], dh.x, dl.x);
    split_bf16x2(ds[2], ds[3], dh.y, dl.y);
    split_bf16x2(a[0], a[1], ah.x, al.x);
    split_bf16x2(a[2], a[3], ah.y, al.y);
    *reinterpret_cast<uint2*>(o_dsh + c) = dh;
    *reinterpret_cast<uint2*>(o_dsl + c) = dl;
    *reinterpret_cast<uint2*>(o_ah + c) = ah;
    *reinterpret_cast<uint2*>(o_al + c) = al;
  }
}

}  // namespace lamp
