// Weight gradient on the 5th-gen tensor cores:  dW[N, K] += dY[M, N]^T X[M, K]   (the contraction runs over the ROWS)
//
// Both operands are consumed "MN-major": in dY [M, N] and X [M, K] (row-major planes) the contraction index m is the
// slow dimension and the output indices n / k are contiguous, which is exactly the layout tcgen05 reads with the
// a_major / b_major bits set -- no transposed copies of the activations are ever made.  TMA stages [64 rows x 64
// columns] boxes (TNC_BM rows x 64 columns in general; 128B swizzle) of the hi / lo planes: two boxes form the 128-wide n tile of dY (operand A, M = 128),
// four the 256-wide k tile of X (operand B, N = 256); one UMMA K-step consumes 16 rows.  3-term split-bf16 products.
//   warp 0: TMA producer (4-stage ring of 48 KB stages)      warp 1: MMA issuer (accumulator: 256 TMEM columns)
//   warps 2-5: epilogue -- tcgen05.ld (lane == n row), smem transpose, 128-bit vector red.global.add into dW: the M rows are split across
//              CTAs (split-K of this contraction), so that all SMs work on the few output tiles of a 512 x 512 weight.
// One CTA = one (n tile, k tile, row chunk).
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

constexpr int TNC_BM = 32;         // rows (contraction) per stage: 48 KB stages, 4 in flight (64-row stages left only
                                   // ONE 96 KB stage in flight while the other was consumed: load-latency bound)
constexpr int TNC_TILE_N = 128;    // dW rows per CTA (UMMA M)
constexpr int TNC_TILE_K = 256;    // dW columns per CTA (UMMA N)
constexpr int TNC_THREADS = 192;
constexpr uint32_t TNC_BOX_BYTES = TNC_BM * 128;  // one [64 x 64] bf16 box
__host__ __device__ constexpr uint32_t tnc_stage_bytes(int npl) { return npl * (2 + 4) * TNC_BOX_BYTES; }
__host__ __device__ constexpr int tnc_stages(int npl) { return npl == 2 ? 4 : 8; }
__host__ __device__ constexpr uint32_t tnc_smem_bytes(int npl) { return tnc_stages(npl) * tnc_stage_bytes(npl) + 1024 + 256; }

struct GemmTnParams {
  long long M;
  int N, K;
  long long chunk;  // rows per CTA, multiple of TNC_BM
  float* dW;        // [N, K] fp32, accumulated with atomics
};

template <int NTERMS>
__global__ void __launch_bounds__(TNC_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                  const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                  const GemmTnParams p) {
  constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  constexpr int STAGES = tnc_stages(NPL);
  constexpr uint32_t STAGE_BYTES = tnc_stage_bytes(NPL);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  // stage layout: Y_hi (2 boxes) | X_hi (4 boxes) | Y_lo (2) | X_lo (4)
  auto y_tile = [&](int s, int pl) { return smem + s * STAGE_BYTES + pl * 6 * TNC_BOX_BYTES; };
  auto x_tile = [&](int s, int pl) { return smem + s * STAGE_BYTES + pl * 6 * TNC_BOX_BYTES + 2 * TNC_BOX_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_k = (p.K + TNC_TILE_K - 1) / TNC_TILE_K, tiles_n = (p.N + TNC_TILE_N - 1) / TNC_TILE_N;
  const int tile = blockIdx.x % (tiles_k * tiles_n);
  const long long split = blockIdx.x / (tiles_k * tiles_n);
  const int n0 = (tile / tiles_k) * TNC_TILE_N, k0 = (tile % tiles_k) * TNC_TILE_K;
  const long long m_begin = split * p.chunk;
  const long long m_end = (m_begin + p.chunk < p.M) ? m_begin + p.chunk : p.M;
  const int num_it = static_cast<int>((m_end - m_begin + TNC_BM - 1) / TNC_BM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmY_hi);
    tma_prefetch_desc(&tmX_hi);
    if (NPL == 2) {
      tma_prefetch_desc(&tmY_lo);
      tma_prefetch_desc(&tmX_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TNC_TILE_K);
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
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
        const int m0 = static_cast<int>(m_begin + static_cast<long long>(it) * TNC_BM);
        for (int pl = 0; pl < NPL; ++pl) {
          const CUtensorMap* ty = pl ? &tmY_lo : &tmY_hi;
          const CUtensorMap* tx = pl ? &tmX_lo : &tmX_hi;
          for (int bx = 0; bx < 2; ++bx) tma_load_2d(y_tile(stage, pl) + bx * TNC_BOX_BYTES, ty, &full_bar[stage], n0 + 64 * bx, m0);
          for (int bx = 0; bx < 4; ++bx) tma_load_2d(x_tile(stage, pl) + bx * TNC_BOX_BYTES, tx, &full_bar[stage], k0 + 64 * bx, m0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TNC_TILE_N, TNC_TILE_K, 1, 1);  // A and B MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
#pragma unroll
        for (int t = 0; t < TNC_BM / 16; ++t) {
          // 16 contraction rows = 2 KB inside every 64-column box; LBO = distance between the 64-wide boxes along the
          // MN dimension, SBO = 8 rows
          const uint64_t ya_hi = umma_smem_desc(smem_u32(y_tile(stage, 0)) + t * 2048, TNC_BOX_BYTES, 1024);
          const uint64_t xb_hi = umma_smem_desc(smem_u32(x_tile(stage, 0)) + t * 2048, TNC_BOX_BYTES, 1024);
          umma_bf16_ss(tmem_base, ya_hi, xb_hi, idesc, (it | t) != 0 ? 1u : 0u);
          if (NTERMS == 3) {
            const uint64_t ya_lo = umma_smem_desc(smem_u32(y_tile(stage, 1)) + t * 2048, TNC_BOX_BYTES, 1024);
            const uint64_t xb_lo = umma_smem_desc(smem_u32(x_tile(stage, 1)) + t * 2048, TNC_BOX_BYTES, 1024);
            umma_bf16_ss(tmem_base, ya_hi, xb_lo, idesc, 1u);
            umma_bf16_ss(tmem_base, ya_lo, xb_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    // epilogue: TMEM lane == dW row inside the tile; 32 columns per tcgen05.ld.  The chunk is transposed through a
    // per-warp smem slice (the operand ring is idle by now) so that the reductions into dW leave as 128-bit vector
    // red.global.add over CONTIGUOUS row segments: 8 lanes cover the 32 columns of a row, one instruction 4 rows -- a
    // scalar atomicAdd per (lane == row) element touched 32 different sectors per instruction and made the L2 atomic
    // units the bottleneck of this split-K contraction.
    const int wq = warp & 3;
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    if (num_it > 0) {
      float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);  // [32 rows][36]: 16-byte aligned rows
      const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
      for (int c0 = 0; c0 < TNC_TILE_K; c0 += 32) {
        if (k0 + c0 >= p.K) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(stg + lane * 36 + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                                                                        __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        __syncwarp();
        const bool col_ok = k0 + c0 + sub_c < p.K;  // K % 8 == 0: the 4 columns are all in or all out
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub_r;
          const int n = n0 + wq * 32 + rr;
          if (n < p.N && col_ok) {
            const float4 v = *reinterpret_cast<const float4*>(stg + rr * 36 + sub_c);
            float* dst = p.dW + static_cast<long long>(n) * p.K + k0 + c0 + sub_c;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
          }
        }
        __syncwarp();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TNC_TILE_K);
  }
}

}  // namespace lamp
