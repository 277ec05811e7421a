// Dense projection GEMM on the 5th-gen tensor cores:   C[M,N] = A[M,K] * W[N,K]^T  (+ epilogue)
//
// Operands are "split-bf16 planes" (see sm100_primitives.cuh): each fp32 matrix is carried as a hi and a lo
// bf16 matrix.  With NTERMS == 3 every K-slice issues three tcgen05.mma (hi*hi + hi*lo + lo*hi, fp32
// accumulation in TMEM), which reproduces an fp32 GEMM to ~2^-16 relative -- this is the path that meets the
// reference's 1e-3 fp32 tolerance with >100x margin.  NTERMS == 1 uses the hi planes only (plain bf16 GEMM).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor 2D, swizzled [rows x BLOCK_K] bf16 boxes, STAGES-deep ring.
//                                BLOCK_K = 32 (64 B rows, 64B swizzle) keeps 4 stages of 48 KB in flight for the
//                                3-term 128x256 tile; BLOCK_K = 64 (128B swizzle) is kept for comparison.
//   warp 1   : MMA issuer     -- one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a double-buffered
//                                TMEM accumulator (2 x BLOCK_N columns); tcgen05.commit releases smem stages
//   warps 2-9: epilogue       -- 8 warps (two per TMEM lane quarter, alternating 32-column chunks, i.e. two per
//                                scheduler to hide latencies): tcgen05.ld (lane == output row) -> per-warp
//                                XOR-swizzled smem transpose -> every global access (residual read, fp32 / plane
//                                stores) is a full 128 B row segment per 8 lanes; overlaps the next tile's main loop
// Used for: Q/K/V projections, fc (+residual), both FFN layers (lamp/SubLayers.py:91-93,110,133).
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] or nullptr
  const float* residual;  // [*, ldr] or nullptr; row index = resid_mod ? row % resid_mod : row
  const __nv_bfloat16* res_hi;  // alternative residual source: split-bf16 planes [*, ldr] (hi + lo, 2^-17 relative);
  const __nv_bfloat16* res_lo;  // lets a layer keep its activations in operand form only (no fp32 copy in HBM)
  int ldr;
  int resid_mod;
  int relu;
  float* out_f32;  // [M, ldo] or nullptr
  int ldo;
  __nv_bfloat16* out_hi;  // [M, ldp] or nullptr
  __nv_bfloat16* out_lo;  // [M, ldp] or nullptr (nullptr -> hi only)
  int ldp;
  // EPI_LN only: out = LayerNorm(acc + bias + residual) * gamma + beta over the full row (N <= 2 * BLOCK_N)
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  // optional device-side row count: the kernel processes min(M, *m_dev) rows (padding-aware execution: the number
  // of non-PAD tokens of a batch is only known on the device; no host synchronisation is needed to launch)
  const int* m_dev;
  // ---- deferred LayerNorm (EPI_DLN_A / EPI_RSTATS).  A "deferred" activation is the PRE-norm tensor y (planes) plus
  // per-row partial sums {sum y, sum y^2} ("row stats": [rows][nparts] float2); the LayerNorm itself is applied by
  // whoever consumes it, so no LayerNorm kernel (one HBM read + one write of the activation) runs between two GEMMs:
  //   * as the A operand (EPI_DLN_A):  LN(y) W^T = rstd (y (gamma o W)^T - mean colsum) + W beta.  The host folds gamma
  //     into the weight planes; a_colsum[n] = sum_k gamma_k W[n,k], and `bias` carries bias[n] + sum_k beta_k W[n,k].
  //   * as the residual (EPI_RSTATS): normalised element-wise while it is added.
  // EPI_RSTATS also emits the row stats of ITS output (slot nt*2 + half of [M][2*num_n]).
  const float2* a_stats;  // [M][a_nparts]
  int a_nparts;
  float a_eps;
  const float* a_colsum;  // [N]
  const float2* r_stats;  // [M][r_nparts] or nullptr (plain residual)
  int r_nparts;
  float r_eps;
  const float* r_gamma;   // [N]
  const float* r_beta;    // [N]
  float2* stats_out;      // [M][2 * num_n]
  // ---- EPI_F32_DROP (training): out = dropout(acc + bias) + residual with the counter hash of the element-wise
  // dropout kernels (keep(row, col) is a pure function of (seed, row, col): the backward recomputes the mask)
  uint32_t drop_thresh;
  float drop_scale;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;  // nullable: device-side addend of the seed (advanced inside CUDA graphs)
};

// Residual values of the 8 rows (m_base + 4 i, i = 0..7) x 4 consecutive columns handled by one lane of the coalesced
// epilogue phase: fp32 source, or reconstructed from the hi (+ lo) planes.  All loads of a lane are independent and
// issued before any use (8 or 16 requests in flight per lane).
__device__ __forceinline__ void gemm_load_residual8(const GemmParams& p, int m_base, int col, int rows_left, bool col_ok,
                                                    float4 (&dst)[8]) {
  if (p.residual != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * i < rows_left && col_ok) {
        const int row = m_base + 4 * i;
        const int rres = p.resid_mod ? (row % p.resid_mod) : row;
        dst[i] = __ldg(reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(rres) * p.ldr + col));
      }
    }
    return;
  }
  uint2 h[8], l[8];
  const __nv_bfloat16* lo_src = p.res_lo != nullptr ? p.res_lo : p.res_hi;  // branch-free: weight 0 without a lo plane
  const float lo_w = p.res_lo != nullptr ? 1.0f : 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = make_uint2(0u, 0u);
    l[i] = make_uint2(0u, 0u);
    if (4 * i < rows_left && col_ok) {
      const int row = m_base + 4 * i;
      const int rres = p.resid_mod ? (row % p.resid_mod) : row;
      const size_t off = static_cast<size_t>(rres) * p.ldr + col;
      h[i] = __ldg(reinterpret_cast<const uint2*>(p.res_hi + off));
      l[i] = __ldg(reinterpret_cast<const uint2*>(lo_src + off));
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    dst[i].x = fmaf(lo_w, __uint_as_float(l[i].x << 16), __uint_as_float(h[i].x << 16));
    dst[i].y = fmaf(lo_w, __uint_as_float(l[i].x & 0xFFFF0000u), __uint_as_float(h[i].x & 0xFFFF0000u));
    dst[i].z = fmaf(lo_w, __uint_as_float(l[i].y << 16), __uint_as_float(h[i].y << 16));
    dst[i].w = fmaf(lo_w, __uint_as_float(l[i].y & 0xFFFF0000u), __uint_as_float(h[i].y & 0xFFFF0000u));
  }
}

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr uint32_t GEMM_EPI_STAGING = GEMM_EPI_WARPS * 32 * 128;  // per epilogue warp: 32 rows x 128 B
// epilogue flavours (compile-time, so the per-element code carries no dead branches)
constexpr int EPI_PLANES = 0;  // (+bias)(ReLU) -> hi/lo planes
constexpr int EPI_F32 = 1;     // (+bias)(+residual) -> fp32
constexpr int EPI_ANY = 2;     // everything, selected at run time (tests / rare combinations)
constexpr int EPI_LN = 3;      // (+bias)(+residual) -> LayerNorm over the whole row -> fp32 and/or planes.  The two
                               // 256-column accumulator stages hold the two halves of ONE 512-wide row block, so the
                               // row statistics are complete on chip and the pre-norm tensor never touches HBM.
constexpr int EPI_DLN_A = 4;   // A operand is a deferred LayerNorm: rstd*(acc - mean*colsum) + bias' (ReLU) -> planes
constexpr int EPI_RSTATS = 5;  // (+bias) + residual (fp32 | planes | deferred-LayerNorm planes) -> planes + row stats
constexpr int EPI_F32_DROP = 6;  // dropout(acc + bias) + residual -> fp32 (training: fc / w_2 of lamp/SubLayers.py:113-117,136-141)

template <int BLOCK_N, int NTERMS, int BLOCK_K, int CTA_GROUP = 1>
struct GemmCfg {
  static_assert(BLOCK_K == 32 || BLOCK_K == 64, "BLOCK_K selects the 64B / 128B swizzle");
  static_assert(CTA_GROUP == 1 || CTA_GROUP == 2, "one CTA or a CTA pair");
  static constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  static constexpr uint32_t ROW_BYTES = BLOCK_K * 2;
  static constexpr int W_ROWS = BLOCK_N / CTA_GROUP;  // rows of W staged by ONE CTA (a pair splits N between its CTAs)
  static constexpr uint32_t A_TILE = GEMM_BLOCK_M * ROW_BYTES;
  static constexpr uint32_t W_TILE = W_ROWS * ROW_BYTES;
  static constexpr uint32_t STAGE_BYTES = NPL * (A_TILE + W_TILE);
  static constexpr int STAGES_RAW = (192 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_STAGING + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint64_t LAYOUT = (BLOCK_K == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  static constexpr uint32_t SBO = 8 * ROW_BYTES;  // 8-row core-matrix group
  static_assert(STAGES >= 2, "need at least a double-buffered ring");
};

// CTA_GROUP == 2: the two CTAs of a cluster form an MMA pair (tcgen05 cta_group::2, M = 256): each CTA stages its own
// 128 rows of A and only HALF of the W tile, the leader issues one MMA for both, and each CTA keeps the accumulator
// of its own 128 rows.  Per output tile every SM ingests 1/3 fewer operand bytes than with two independent CTAs --
// the operand stream from L2 (~31 B/clk/SM measured) is what bounds this kernel, not the tensor pipe.
template <int BLOCK_N, int NTERMS, int BLOCK_K, int EPI, int CTA_GROUP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_planes_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                   const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, NTERMS, BLOCK_K, CTA_GROUP>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NPL = Cfg::NPL;
  constexpr bool PAIR = (CTA_GROUP == 2);
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + GEMM_EPI_STAGING);
  uint64_t* full_bar = bars;                    // PAIR: only the leader's copies are used
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // PAIR: only the leader's copies are used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);

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
      mbar_init(&tmem_empty[a], 32 * GEMM_EPI_WARPS * CTA_GROUP);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_cg2(tmem_slot, TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();
  griddep_wait();  // PDL: nothing above touches global memory

  const int M = (p.m_dev != nullptr) ? min(p.M, __ldg(p.m_dev)) : p.M;
  const int num_m = (M + GEMM_BLOCK_M * CTA_GROUP - 1) / (GEMM_BLOCK_M * CTA_GROUP);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int num_tiles = num_m * num_n;
  const int tile0 = static_cast<int>(blockIdx.x) / CTA_GROUP;
  const int tile_step = static_cast<int>(gridDim.x) / CTA_GROUP;
  // j-th tile of this CTA (pair) -> (m tile, n tile).  EPI_LN walks both n halves of an m tile back to back so that
  // accumulator stage == n half; everything else strides over the flat tile index.
  auto get_tile = [&](int j, int& mt, int& nt) -> bool {
    if (EPI == EPI_LN) {
      mt = tile0 + (j >> 1) * tile_step;
      nt = j & 1;
      return mt < num_m;
    }
    const int tile = tile0 + j * tile_step;
    mt = tile / num_n;
    nt = tile % num_n;
    return tile < num_tiles;
  };

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
    // ------------------------------------------------------------ TMA producer (every CTA loads its own slices)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int mt, nt;
      for (int j = 0; get_tile(j, mt, nt); ++j) {
        const int m0 = mt * (GEMM_BLOCK_M * CTA_GROUP) + static_cast<int>(rank) * GEMM_BLOCK_M;
        const int n0 = nt * BLOCK_N + static_cast<int>(rank) * Cfg::W_ROWS;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const int k0 = kb * BLOCK_K;
          if (PAIR) {
            // both CTAs' bytes complete on the leader's barrier; the leader posts the expected total
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t fb = mapa_shared(&full_bar[stage], 0);
            tma_load_2d_cg2(stage_ptr(stage, 0), &tmA_hi, fb, k0, m0);
            tma_load_2d_cg2(stage_ptr(stage, 1), &tmW_hi, fb, k0, n0);
            if (NPL == 2) {
              tma_load_2d_cg2(stage_ptr(stage, 2), &tmA_lo, fb, k0, m0);
              tma_load_2d_cg2(stage_ptr(stage, 3), &tmW_lo, fb, k0, n0);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(stage_ptr(stage, 0), &tmA_hi, &full_bar[stage], k0, m0);
            tma_load_2d(stage_ptr(stage, 1), &tmW_hi, &full_bar[stage], k0, n0);
            if (NPL == 2) {
              tma_load_2d(stage_ptr(stage, 2), &tmA_lo, &full_bar[stage], k0, m0);
              tma_load_2d(stage_ptr(stage, 3), &tmW_lo, &full_bar[stage], k0, n0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only for a pair)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BLOCK_M * CTA_GROUP, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int mt, nt;
      for (int j = 0; get_tile(j, mt, nt); ++j) {
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
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint32_t koff = k * 32;  // 16 bf16 = 32 B inside the swizzled row
            const uint64_t da_hi = umma_smem_desc(a_hi + koff, 16, Cfg::SBO, Cfg::LAYOUT);
            const uint64_t dw_hi = umma_smem_desc(w_hi + koff, 16, Cfg::SBO, Cfg::LAYOUT);
            const uint32_t first = (kb | k) != 0 ? 1u : 0u;
            if (PAIR) umma_bf16_ss_cg2(d_tmem, da_hi, dw_hi, idesc, first); else umma_bf16_ss(d_tmem, da_hi, dw_hi, idesc, first);
            if (NTERMS == 3) {
              const uint64_t da_lo = umma_smem_desc(a_lo + koff, 16, Cfg::SBO, Cfg::LAYOUT);
              const uint64_t dw_lo = umma_smem_desc(w_lo + koff, 16, Cfg::SBO, Cfg::LAYOUT);
              if (PAIR) {
                umma_bf16_ss_cg2(d_tmem, da_hi, dw_lo, idesc, 1u);
                umma_bf16_ss_cg2(d_tmem, da_lo, dw_hi, idesc, 1u);
              } else {
                umma_bf16_ss(d_tmem, da_hi, dw_lo, idesc, 1u);
                umma_bf16_ss(d_tmem, da_lo, dw_hi, idesc, 1u);
              }
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above retire
          if (PAIR) umma_commit_cg2_mc(&empty_bar[stage], 0x3); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if (PAIR) umma_commit_cg2_mc(&tmem_full[acc], 0x3); else umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int wq = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2; // which of the two warps of this quarter: handles chunks half, half+2, ...
    uint8_t* stg = staging + (warp - 2) * (32 * 128);
    const uint32_t stg_s = smem_u32(stg);  // shared-space address of this warp's staging slice (forces LDS / STS)
    const int sub_r = lane >> 3;  // coalesced phase: 8 lanes per row, 4 rows per pass
    const int sub_c = lane & 7;   // 16-byte chunk (4 fp32) inside the 128 B row segment
    constexpr bool DLN_A = (EPI == EPI_DLN_A);
    constexpr bool RSTATS = (EPI == EPI_RSTATS);
    constexpr bool DROP = (EPI == EPI_F32_DROP);
    const bool want_f32 = (EPI == EPI_F32) || DROP || (EPI == EPI_ANY && p.out_f32 != nullptr);
    const bool want_pl = (EPI == EPI_PLANES) || DLN_A || RSTATS || (EPI == EPI_ANY && p.out_hi != nullptr);
    const bool want_lo = want_pl && p.out_lo != nullptr;
    const bool has_res = (EPI != EPI_PLANES) && !DLN_A && (p.residual != nullptr || p.res_hi != nullptr);
    const bool has_bias = p.bias != nullptr;
    const bool relu = (EPI != EPI_F32) && !DROP && !RSTATS && p.relu;
    const bool res_dln = RSTATS && p.r_stats != nullptr;  // the residual is a deferred LayerNorm
    const uint32_t te_addr[2] = {mapa_shared(&tmem_empty[0], 0), mapa_shared(&tmem_empty[1], 0)};  // leader's copies
    if constexpr (EPI == EPI_LN) {
      // ---------------- LayerNorm-fused epilogue: stage 0 / 1 hold columns [0,256) / [256,512) of the same rows.
      //   pass 1 (per stage, as soon as its MMAs retire): v = acc + bias + residual -> per-row sum / sum of squares
      //   exchange between the two warps of a lane quarter, then
      //   pass 2 (stage 0 first, released early so the next row block's MMAs can start): normalise, scale, store.
      // row statistics are exchanged through the (idle between the passes) staging slices of the two sibling warps
      float* red_mine = reinterpret_cast<float*>(stg);
      const float* red_sib = reinterpret_cast<const float*>(staging + ((warp - 2) ^ 4) * (32 * 128));
      const bool want_f32 = p.out_f32 != nullptr, want_pl = p.out_hi != nullptr, want_lo = p.out_lo != nullptr;
      const bool has_res = (p.residual != nullptr || p.res_hi != nullptr), has_bias = p.bias != nullptr;
      const float inv_n = 1.0f / static_cast<float>(p.N);
      uint32_t ph = 0;
      for (int mt2 = tile0; mt2 < num_m; mt2 += tile_step, ph ^= 1) {
        const int m0 = mt2 * (GEMM_BLOCK_M * CTA_GROUP) + static_cast<int>(rank) * GEMM_BLOCK_M + wq * 32;
        const int rows_left = M - m0 - sub_r;
        float rsum[8], rsq[8], mean[8], rstd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { rsum[i] = 0.f; rsq[i] = 0.f; mean[i] = 0.f; rstd[i] = 0.f; }
        // one pass over a 256-column stage; final_pass = false: statistics, true: normalise + store
        auto pass = [&](int a, bool final_pass) {
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + a * BLOCK_N;
          const int n0 = a * BLOCK_N;
          // residual rows of one 32-column chunk (8 x 128-bit per lane); fetched one chunk AHEAD of their use so
          // the global/L2 latency overlaps the TMEM load and the smem transpose of the current chunk
          auto load_res = [&](int c0, float4 (&dst)[8]) {
            const int col = n0 + c0 + sub_c * 4;
            if (has_res && c0 < BLOCK_N) {
              gemm_load_residual8(p, m0 + sub_r, col, rows_left, col < p.N, dst);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          };
          float4 resv[8], resn[8];
          load_res(half * 32, resv);
#pragma unroll 1
          for (int c0 = half * 32; c0 < BLOCK_N; c0 += 64) {
            uint32_t r[32];
            tmem_ld32(t_row + c0, r);
            load_res(c0 + 64, resn);
            tmem_wait_ld();
            const int col = n0 + c0 + sub_c * 4;
            const bool col_ok = col < p.N;
            if (n0 + c0 < p.N) {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                sts128(stg_s + lane * 128 + ((c ^ (lane & 7)) << 4), make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]));
              __syncwarp();
              float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = bias4, b4 = bias4;
              if (has_bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
              if (final_pass && col_ok) {
                g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col));
                b4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col));
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + sub_r;
                const uint4 vb = lds128(stg_s + rr * 128 + ((sub_c ^ (rr & 7)) << 4));
                float4 v = make_float4(__uint_as_float(vb.x), __uint_as_float(vb.y), __uint_as_float(vb.z), __uint_as_float(vb.w));
                if (4 * i < rows_left && col_ok) {
                  v.x += bias4.x + resv[i].x; v.y += bias4.y + resv[i].y;
                  v.z += bias4.z + resv[i].z; v.w += bias4.w + resv[i].w;
                  if (!final_pass) {
                    rsum[i] += (v.x + v.y) + (v.z + v.w);
                    rsq[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                  } else {
                    v.x = (v.x - mean[i]) * rstd[i] * g4.x + b4.x;
                    v.y = (v.y - mean[i]) * rstd[i] * g4.y + b4.y;
                    v.z = (v.z - mean[i]) * rstd[i] * g4.z + b4.z;
                    v.w = (v.w - mean[i]) * rstd[i] * g4.w + b4.w;
                    const size_t row = static_cast<size_t>(m0 + 4 * i + sub_r);
                    if (want_f32) *reinterpret_cast<float4*>(p.out_f32 + row * p.ldo + col) = v;
                    if (want_pl) {
                      uint2 hi, lo;
                      split_bf16x2(v.x, v.y, hi.x, lo.x);
                      split_bf16x2(v.z, v.w, hi.y, lo.y);
                      *reinterpret_cast<uint2*>(p.out_hi + row * p.ldp + col) = hi;
                      if (want_lo) *reinterpret_cast<uint2*>(p.out_lo + row * p.ldp + col) = lo;
                    }
                  }
                }
              }
              __syncwarp();
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) resv[i] = resn[i];
          }
        };
        mbar_wait(&tmem_full[0], ph);
        tcgen05_fence_after();
        pass(0, false);
        mbar_wait(&tmem_full[1], ph);
        tcgen05_fence_after();
        pass(1, false);
        // row statistics: 8 lanes share a row -> butterfly over the low 3 lane bits, then the sibling warp via smem
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            rsum[i] += __shfl_xor_sync(0xFFFFFFFFu, rsum[i], o);
            rsq[i] += __shfl_xor_sync(0xFFFFFFFFu, rsq[i], o);
          }
        }
        if (sub_c == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            red_mine[(i * 4 + sub_r) * 2] = rsum[i];
            red_mine[(i * 4 + sub_r) * 2 + 1] = rsq[i];
          }
        }
        named_bar_sync(2, 32 * GEMM_EPI_WARPS);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + sub_r;
          const float s1 = red_mine[rl * 2] + red_sib[rl * 2];
          const float s2 = red_mine[rl * 2 + 1] + red_sib[rl * 2 + 1];
          mean[i] = s1 * inv_n;
          const float var = fmaxf(s2 * inv_n - mean[i] * mean[i], 0.0f);
          rstd[i] = rsqrtf(var + p.ln_eps);
        }
        named_bar_sync(2, 32 * GEMM_EPI_WARPS);  // the staging slices are reused by pass 2 after this point
        pass(0, true);
        tcgen05_fence_before();
        if (PAIR) mbar_arrive_cluster(te_addr[0]); else mbar_arrive(&tmem_empty[0]);
        pass(1, true);
        tcgen05_fence_before();
        if (PAIR) mbar_arrive_cluster(te_addr[1]); else mbar_arrive(&tmem_empty[1]);
      }
    } else {
    // The epilogue of a 128x256 tile has ~13K SM clocks (the tile's MMAs) to retire 32 chunk passes of 32x32 outputs:
    // at 2 warps per scheduler it is ISSUE bound, so the per-element instruction stream is kept minimal --
    // everything that depends on the row only (residual row offsets incl. the `% resid_mod`, validity, LayerNorm
    // mean / rstd) is computed once per tile, loads are unconditional on clamped addresses (no per-row branches),
    // and the residual of the NEXT chunk is fetched as raw bits into a ping-pong buffer and only decoded at use.
    const bool res_f32 = p.residual != nullptr;
    const uint16_t* rsrc_hi = reinterpret_cast<const uint16_t*>(p.res_hi);
    const uint16_t* rsrc_lo = reinterpret_cast<const uint16_t*>(p.res_lo != nullptr ? p.res_lo : p.res_hi);
    const uint32_t lo_keep = (p.res_lo != nullptr) ? 0xFFFFFFFFu : 0u;  // no lo plane: its bits are masked to +0
    struct ResBuf {
      uint4 raw[8];   // fp32 residual: the 4 floats; planes: {hi pair 0, hi pair 1, lo pair 0, lo pair 1}
      float4 g, b;    // deferred residual: gamma / beta of the chunk's columns
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    int mt, nt;
    for (int j = 0; get_tile(j, mt, nt); ++j) {
      const int m0 = mt * (GEMM_BLOCK_M * CTA_GROUP) + static_cast<int>(rank) * GEMM_BLOCK_M + wq * 32;
      const int n0 = nt * BLOCK_N;
      const int rows_left = M - m0 - sub_r;  // row (m0 + sub_r + 4 i) is valid iff 4 i < rows_left
      // element offset of the residual row of each of this lane's 8 rows (clamped to a valid row; host-checked to
      // fit 32 bits)
      int roff[8];
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = min(m0 + sub_r + 4 * i, M - 1);
          roff[i] = (p.resid_mod ? (row % p.resid_mod) : row) * p.ldr;
        }
      }
      uint32_t rh[8];  // EPI_F32_DROP: dropout row hashes of this lane's 8 rows
      if (DROP) {
        const unsigned long long sd = p.drop_seed + (p.drop_seed_dev != nullptr ? __ldg(p.drop_seed_dev) : 0ull);
#pragma unroll
        for (int i = 0; i < 8; ++i) rh[i] = drop_rowhash(sd, static_cast<unsigned long long>(m0 + sub_r + 4 * i));
      }
      auto load_res = [&](int c0, ResBuf& d) {
        if (!has_res || c0 >= BLOCK_N) return;
        const int col = n0 + c0 + sub_c * 4;
        const int colc = col < p.N ? col : 0;
        if (res_f32) {
#pragma unroll
          for (int i = 0; i < 8; ++i) d.raw[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + roff[i] + colc));
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint2 h = __ldg(reinterpret_cast<const uint2*>(rsrc_hi + roff[i] + colc));
            const uint2 l = __ldg(reinterpret_cast<const uint2*>(rsrc_lo + roff[i] + colc));
            d.raw[i] = make_uint4(h.x, h.y, l.x & lo_keep, l.y & lo_keep);
          }
        }
        if (res_dln) {
          d.g = __ldg(reinterpret_cast<const float4*>(p.r_gamma + colc));
          d.b = __ldg(reinterpret_cast<const float4*>(p.r_beta + colc));
        }
      };
      ResBuf rb0, rb1;
      load_res(half * 32, rb0);
      // deferred LayerNorm: scale / shift of this lane's 8 rows (of the A operand, or of the residual), from the
      // partial sums written by the producing GEMM: LN(y) = (y * ra + rc) * gamma + beta, ra = rstd, rc = -mean*rstd
      float ra[8], rc[8];
      if (DLN_A || res_dln) {
        const float2* st = DLN_A ? p.a_stats : p.r_stats;
        const int np = DLN_A ? p.a_nparts : p.r_nparts;
        const float inv_w = 1.0f / static_cast<float>(DLN_A ? p.K : p.N);
        const float eps = DLN_A ? p.a_eps : p.r_eps;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2* sr = st + static_cast<size_t>(min(m0 + sub_r + 4 * i, M - 1)) * np;
          float s1 = 0.f, s2 = 0.f;
          if (np == 4) {  // the common case (256 < N <= 512): two 128-bit loads
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(sr));
            const float4 e1 = __ldg(reinterpret_cast<const float4*>(sr) + 1);
            s1 = (e0.x + e0.z) + (e1.x + e1.z);
            s2 = (e0.y + e0.w) + (e1.y + e1.w);
          } else {
            for (int q = 0; q < np; ++q) {
              const float2 e = __ldg(sr + q);
              s1 += e.x;
              s2 += e.y;
            }
          }
          const float mean = s1 * inv_w;
          ra[i] = rsqrtf(fmaxf(s2 * inv_w - mean * mean, 0.0f) + eps);
          rc[i] = -mean * ra[i];
        }
      }
      float st1[8], st2[8];  // EPI_RSTATS: running row sums of this lane's columns
#pragma unroll
      for (int i = 0; i < 8; ++i) { st1[i] = 0.f; st2[i] = 0.f; }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BLOCK_N;
      // one 32-column chunk: `cur` holds its residual, `nxt` receives the residual of the chunk 64 columns further
      auto chunk = [&](int c0, const ResBuf& cur, ResBuf& nxt) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tmem_wait_ld();
        if (n0 + c0 >= p.N) return;  // warp-uniform: whole 32-column chunk out of range (so are all later ones)
        // registers (lane == row) -> swizzled staging: chunk c of row `lane` lives at chunk (c ^ (lane & 7))
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(stg_s + lane * 128 + ((c ^ (lane & 7)) << 4), make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]));
        __syncwarp();
        load_res(c0 + 64, nxt);  // issued once the accumulator registers are dead; consumed a whole chunk later
        const int col = n0 + c0 + sub_c * 4;
        const bool col_ok = col < p.N;  // N % 8 == 0 (host-checked) -> the 4 columns are all in or all out
        const int colc = col_ok ? col : 0;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = bias4;
        if (has_bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + colc));
        if (DLN_A) cs4 = __ldg(reinterpret_cast<const float4*>(p.a_colsum + colc));
        float* of = want_f32 ? p.out_f32 + static_cast<size_t>(m0 + sub_r) * p.ldo + col : nullptr;
        const size_t poff = static_cast<size_t>(m0 + sub_r) * p.ldp + col;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub_r;
          const uint4 vb = lds128(stg_s + rr * 128 + ((sub_c ^ (rr & 7)) << 4));
          float4 v = make_float4(__uint_as_float(vb.x), __uint_as_float(vb.y), __uint_as_float(vb.z), __uint_as_float(vb.w));
          if (DLN_A) {
            v.x = fmaf(ra[i], v.x, rc[i] * cs4.x); v.y = fmaf(ra[i], v.y, rc[i] * cs4.y);
            v.z = fmaf(ra[i], v.z, rc[i] * cs4.z); v.w = fmaf(ra[i], v.w, rc[i] * cs4.w);
          }
          if (has_bias) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
          if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          if (DROP) {
            const uint32_t cb = static_cast<uint32_t>(col);
            v.x = drop_keep(rh[i], cb, p.drop_thresh) ? v.x * p.drop_scale : 0.f;
            v.y = drop_keep(rh[i], cb + 1, p.drop_thresh) ? v.y * p.drop_scale : 0.f;
            v.z = drop_keep(rh[i], cb + 2, p.drop_thresh) ? v.z * p.drop_scale : 0.f;
            v.w = drop_keep(rh[i], cb + 3, p.drop_thresh) ? v.w * p.drop_scale : 0.f;
          }
          if (has_res) {
            const uint4 q = cur.raw[i];
            float4 rv;
            if (res_f32) {
              rv = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
            } else {
              rv.x = __uint_as_float(q.x << 16) + __uint_as_float(q.z << 16);
              rv.y = __uint_as_float(q.x & 0xFFFF0000u) + __uint_as_float(q.z & 0xFFFF0000u);
              rv.z = __uint_as_float(q.y << 16) + __uint_as_float(q.w << 16);
              rv.w = __uint_as_float(q.y & 0xFFFF0000u) + __uint_as_float(q.w & 0xFFFF0000u);
            }
            if (res_dln) {
              rv.x = fmaf(fmaf(rv.x, ra[i], rc[i]), cur.g.x, cur.b.x); rv.y = fmaf(fmaf(rv.y, ra[i], rc[i]), cur.g.y, cur.b.y);
              rv.z = fmaf(fmaf(rv.z, ra[i], rc[i]), cur.g.z, cur.b.z); rv.w = fmaf(fmaf(rv.w, ra[i], rc[i]), cur.g.w, cur.b.w);
            }
            v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
          }
          if (4 * i < rows_left && col_ok) {
            if (RSTATS) {
              st1[i] += (v.x + v.y) + (v.z + v.w);
              st2[i] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, st2[i]))));
            }
            if (want_f32) *reinterpret_cast<float4*>(of + static_cast<size_t>(4 * i) * p.ldo) = v;
            if (want_pl) {
              uint2 hi, lo;
              split_bf16x2(v.x, v.y, hi.x, lo.x);
              split_bf16x2(v.z, v.w, hi.y, lo.y);
              const size_t off = poff + static_cast<size_t>(4 * i) * p.ldp;
              *reinterpret_cast<uint2*>(p.out_hi + off) = hi;
              if (want_lo) *reinterpret_cast<uint2*>(p.out_lo + off) = lo;
            }
          }
        }
        __syncwarp();
      };
#pragma unroll 1
      for (int c0 = half * 32; c0 < BLOCK_N; c0 += 128) {
        chunk(c0, rb0, rb1);
        chunk(c0 + 64, rb1, rb0);
      }
      if (RSTATS) {
        // 8 lanes (sub_c) share a row: butterfly over the low 3 lane bits, then one float2 per row and warp
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            st1[i] += __shfl_xor_sync(0xFFFFFFFFu, st1[i], o);
            st2[i] += __shfl_xor_sync(0xFFFFFFFFu, st2[i], o);
          }
        }
        if (sub_c == 0) {
          const int np = 2 * num_n;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (4 * i < rows_left)
              p.stats_out[static_cast<size_t>(m0 + sub_r + 4 * i) * np + nt * 2 + half] = make_float2(st1[i], st2[i]);
        }
      }
      tcgen05_fence_before();
      if (PAIR) mbar_arrive_cluster(te_addr[acc]); else mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  }

  tcgen05_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // a pair must not tear down while the peer may still signal it
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace lamp
