// Dense projection GEMM on the 5th-gen tensor cores:   C[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
// Operands are "split-bf16 planes" (see sm100_primitives.cuh): each fp32 matrix is carried as a hi and a lo
// bf16 matrix.  With NTERMS == 3 every K-slice issues three tcgen05.mma (hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM), which reproduces an fp32 GEMM to ~2^-16 relative -- this is the path that meets the
// reference's 1e-3 fp32 tolerance with >30x margin.  NTERMS == 1 uses the hi planes only (plain bf16 GEMM).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor 2D, 128B-swizzled [rows x 64] bf16 boxes, STAGES-deep ring
//   warp 1   : MMA issuer     -- one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a double-buffered
//                                TMEM accumulator (2 x BLOCK_N columns); tcgen05.commit releases smem stages
//   warps 2-5: epilogue       -- tcgen05.ld (lane == output row), bias / ReLU / residual, then fp32 and/or
//                                split-bf16 stores; overlaps with the next tile's main loop
// Used for: Q/K/V projections, fc (+residual), both FFN layers (lamp/SubLayers.py:91-93,110,133).
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] or nullptr
  const float* residual;  // [*, ldr] or nullptr; row index = resid_mod ? row % resid_mod : row
  int ldr;
  int resid_mod;
  int relu;
  float* out_f32;  // [M, ldo] or nullptr
  int ldo;
  __nv_bfloat16* out_hi;  // [M, ldp] or nullptr
  __nv_bfloat16* out_lo;  // [M, ldp] or nullptr (nullptr -> hi only)
  int ldp;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N, int NTERMS>
struct GemmCfg {
  static constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  static constexpr uint32_t A_TILE = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr uint32_t W_TILE = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = NPL * (A_TILE + W_TILE);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(STAGES >= 2, "need at least a double-buffered ring");
};

template <int BLOCK_N, int NTERMS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_planes_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                   const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, NTERMS>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NPL = Cfg::NPL;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmW_hi);
    if (NPL == 2) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmW_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int num_tiles = num_m * num_n;

  auto stage_ptr = [&](int s, int which) -> uint8_t* {
    // which: 0 = A_hi, 1 = W_hi, 2 = A_lo, 3 = W_lo
    uint8_t* base = smem + s * Cfg::STAGE_BYTES;
    switch (which) {
      case 0: return base;
      case 1: return base + Cfg::A_TILE;
      case 2: return base + Cfg::A_TILE + Cfg::W_TILE;
      default: return base + 2 * Cfg::A_TILE + Cfg::W_TILE;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * GEMM_BLOCK_M;
        const int n0 = (tile % num_n) * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * GEMM_BLOCK_K;
          tma_load_2d(stage_ptr(stage, 0), &tmA_hi, &full_bar[stage], k0, m0);
          tma_load_2d(stage_ptr(stage, 1), &tmW_hi, &full_bar[stage], k0, n0);
          if (NPL == 2) {
            tma_load_2d(stage_ptr(stage, 2), &tmA_lo, &full_bar[stage], k0, m0);
            tma_load_2d(stage_ptr(stage, 3), &tmW_lo, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_u32(stage_ptr(stage, 0));
          const uint32_t w_hi = smem_u32(stage_ptr(stage, 1));
          const uint32_t a_lo = smem_u32(stage_ptr(stage, 2));
          const uint32_t w_lo = smem_u32(stage_ptr(stage, 3));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            const uint32_t koff = k * 32;  // 16 bf16 = 32 B inside the 128 B swizzle row
            const uint64_t da_hi = umma_smem_desc(a_hi + koff, 16, 1024);
            const uint64_t dw_hi = umma_smem_desc(w_hi + koff, 16, 1024);
            umma_bf16_ss(d_tmem, da_hi, dw_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            if (NTERMS == 3) {
              const uint64_t da_lo = umma_smem_desc(a_lo + koff, 16, 1024);
              const uint64_t dw_lo = umma_smem_desc(w_lo + koff, 16, 1024);
              umma_bf16_ss(d_tmem, da_hi, dw_lo, idesc, 1u);
              umma_bf16_ss(d_tmem, da_lo, dw_hi, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * GEMM_BLOCK_M;
      const int n0 = (tile % num_n) * BLOCK_N;
      const int row = m0 + wq * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BLOCK_N;
      const float* res_row = nullptr;
      if (p.residual != nullptr && row_ok) {
        const int rr = p.resid_mod ? (row % p.resid_mod) : row;
        res_row = p.residual + static_cast<size_t>(rr) * p.ldr;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = n0 + c0 + j;
            if (col < p.N) {  // N % 8 == 0 is enforced by the host wrapper
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
              }
              if (res_row != nullptr) {
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(res_row + col));
                const float4 q1 = __ldg(reinterpret_cast<const float4*>(res_row + col + 4));
                v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w;
                v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
              }
              if (p.out_f32 != nullptr) {
                float4* o = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(row) * p.ldo + col);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
              if (p.out_hi != nullptr) {
                uint4 hi, lo;
                split_bf16x2(v[0], v[1], hi.x, lo.x);
                split_bf16x2(v[2], v[3], hi.y, lo.y);
                split_bf16x2(v[4], v[5], hi.z, lo.z);
                split_bf16x2(v[6], v[7], hi.w, lo.w);
                const size_t off = static_cast<size_t>(row) * p.ldp + col;
                *reinterpret_cast<uint4*>(p.out_hi + off) = hi;
                if (p.out_lo != nullptr) *reinterpret_cast<uint4*>(p.out_lo + off) = lo;
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace lamp
