// Backward kernels of the dense stages of the training path: LayerNorm backward and the
// weight-gradient contraction dW = dY^T X (+ bias gradient).  The input gradient dX = dY W runs on the forward's
// tcgen05 GEMM (gemm_planes.cuh) with transposed weight planes.  Correctness-first versions (warp-level MMA, fp32
// atomics for the split-M reduction); same 3-term split-bf16 products as everywhere else.
#pragma once
#include "attn_bwd.cuh"

namespace lamp {

// y = LayerNorm(x) * gamma + beta (torch.nn.LayerNorm; lamp/SubLayers.py:117,141).  Given x and dy:
//   xhat = (x - mean) * rstd;  g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));
//   dgamma += sum_rows dy * xhat;  dbeta += sum_rows dy   (accumulated with fp32 atomics: zero them first).
// One warp per row (grid-stride), the row and the warp's dgamma / dbeta partials live in registers (D <= 128 * MAXV).
// PLANES: dx is ALSO written as the split-bf16 planes of dropout-backward(dx) = keep(row, col) * dx / (1 - p) -- the
// operand of the dW / input-gradient products of the sub-layer's last projection (thresh == 0: a plain split), which
// saves the separate dropout_split pass over dx.
struct LnBwdDrop {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  uint32_t thresh;
  float scale;
  unsigned long long seed;
  const unsigned long long* seed_dev;
};
template <int MAXV, bool PLANES>
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                     const float* __restrict__ gamma, float eps, long long rows, int D,
                                     float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     const LnBwdDrop dp) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int d4 = D >> 2;
  float4 gsum[MAXV], bsum[MAXV], gam[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    gsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    bsum[i] = gsum[i];
    const int idx = lane + 32 * i;
    gam[i] = idx < d4 ? __ldg(reinterpret_cast<const float4*>(gamma) + idx) : gsum[i];
  }
  for (long long row = warp0; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const float4* dr = reinterpret_cast<const float4*>(dy + row * D);
    float4 xv[MAXV], dv[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < d4) {
        xv[i] = xr[idx];
        dv[i] = dr[idx];
        s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    const float mean = s / D;
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < d4) {
        const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
        qv += (a * a + b * b) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qv += __shfl_xor_sync(0xFFFFFFFFu, qv, o);
    const float rstd = rsqrtf(qv / D + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < d4) {
        // xhat overwrites xv, g = dy * gamma overwrites nothing (dv is still needed for dgamma / dbeta)
        xv[i].x = (xv[i].x - mean) * rstd; xv[i].y = (xv[i].y - mean) * rstd;
        xv[i].z = (xv[i].z - mean) * rstd; xv[i].w = (xv[i].w - mean) * rstd;
        const float g0 = dv[i].x * gam[i].x, g1 = dv[i].y * gam[i].y, g2 = dv[i].z * gam[i].z, g3 = dv[i].w * gam[i].w;
        c1 += (g0 + g1) + (g2 + g3);
        c2 += (g0 * xv[i].x + g1 * xv[i].y) + (g2 * xv[i].z + g3 * xv[i].w);
        gsum[i].x += dv[i].x * xv[i].x; gsum[i].y += dv[i].y * xv[i].y;
        gsum[i].z += dv[i].z * xv[i].z; gsum[i].w += dv[i].w * xv[i].w;
        bsum[i].x += dv[i].x; bsum[i].y += dv[i].y; bsum[i].z += dv[i].z; bsum[i].w += dv[i].w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, o);
      c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
    }
    c1 /= D;
    c2 /= D;
    float4* outr = reinterpret_cast<float4*>(dx + row * D);
    uint32_t rh = 0u;
    if (PLANES && dp.thresh)
      rh = drop_rowhash(dp.seed + (dp.seed_dev ? __ldg(dp.seed_dev) : 0ull), static_cast<unsigned long long>(row));
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < d4) {
        float4 o;
        o.x = rstd * (dv[i].x * gam[i].x - c1 - xv[i].x * c2);
        o.y = rstd * (dv[i].y * gam[i].y - c1 - xv[i].y * c2);
        o.z = rstd * (dv[i].z * gam[i].z - c1 - xv[i].z * c2);
        o.w = rstd * (dv[i].w * gam[i].w - c1 - xv[i].w * c2);
        outr[idx] = o;
        if (PLANES) {
          const uint32_t col = 4u * idx;
          float4 v;
          v.x = (dp.thresh && !drop_keep(rh, col, dp.thresh)) ? 0.f : o.x * dp.scale;
          v.y = (dp.thresh && !drop_keep(rh, col + 1, dp.thresh)) ? 0.f : o.y * dp.scale;
          v.z = (dp.thresh && !drop_keep(rh, col + 2, dp.thresh)) ? 0.f : o.z * dp.scale;
          v.w = (dp.thresh && !drop_keep(rh, col + 3, dp.thresh)) ? 0.f : o.w * dp.scale;
          uint2 h, l;
          split_bf16x2(v.x, v.y, h.x, l.x);
          split_bf16x2(v.z, v.w, h.y, l.y);
          reinterpret_cast<uint2*>(dp.hi + row * D)[idx] = h;
          reinterpret_cast<uint2*>(dp.lo + row * D)[idx] = l;
        }
      }
    }
  }
  // block-level reduction in shared memory (the block's 8 warps add into one [2][D] buffer), then one global atomic
  // per column and block
  extern __shared__ float lnb_red[];  // [2 * D]
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) lnb_red[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < d4) {
      atomicAdd(lnb_red + 4 * idx + 0, gsum[i].x); atomicAdd(lnb_red + 4 * idx + 1, gsum[i].y);
      atomicAdd(lnb_red + 4 * idx + 2, gsum[i].z); atomicAdd(lnb_red + 4 * idx + 3, gsum[i].w);
      atomicAdd(lnb_red + D + 4 * idx + 0, bsum[i].x); atomicAdd(lnb_red + D + 4 * idx + 1, bsum[i].y);
      atomicAdd(lnb_red + D + 4 * idx + 2, bsum[i].z); atomicAdd(lnb_red + D + 4 * idx + 3, bsum[i].w);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dgamma + c, lnb_red[c]);
    atomicAdd(dbeta + c, lnb_red[D + c]);
  }
}

// dW[N, K] += dY[M, N]^T X[M, K]  -- the weight gradient of Y = X W^T (nn.Linear / Conv1d(k=1): lamp/SubLayers.py:
// 91-93,110,133).  Operands are the split-bf16 planes that already exist (x: split for the forward GEMM, dy: split for
// the dx GEMM).  CTA = one 128 x 128 tile of dW and one chunk of the M rows (split-M: partial tiles are added with fp32
// atomics); per 64-row step the dY [64 x 128] and X [64 x 128] plane tiles are copied to smem and dY is consumed
// transposed (col-major fragment loads).  8 warps: warp -> 32 rows (n) x 64 columns (k) of the tile.
constexpr int TN_TILE = 128;
__host__ __device__ constexpr size_t gemm_tn_smem_bytes() { return static_cast<size_t>(4) * BWD_TILE * (TN_TILE + 8) * 2 + 8 * 16 * 20 * 4; }

__global__ void __launch_bounds__(BWD_THREADS) gemm_tn_kernel(const __nv_bfloat16* __restrict__ dy_hi,
                                                              const __nv_bfloat16* __restrict__ dy_lo, long long ldy,
                                                              const __nv_bfloat16* __restrict__ x_hi,
                                                              const __nv_bfloat16* __restrict__ x_lo, long long ldx,
                                                              long long M, int N, int K, long long chunk,
                                                              float* __restrict__ dW) {
  using namespace nvcuda;
  constexpr int TP = TN_TILE + 8;
  extern __shared__ __align__(128) uint8_t tsm[];
  __nv_bfloat16* Y_hi = reinterpret_cast<__nv_bfloat16*>(tsm);
  __nv_bfloat16* Y_lo = Y_hi + BWD_TILE * TP;
  __nv_bfloat16* X_hi = Y_lo + BWD_TILE * TP;
  __nv_bfloat16* X_lo = X_hi + BWD_TILE * TP;
  float* stage_all = reinterpret_cast<float*>(X_lo + BWD_TILE * TP);
  const int tiles_k = (K + TN_TILE - 1) / TN_TILE, tiles_n = (N + TN_TILE - 1) / TN_TILE;
  const int tile = blockIdx.x % (tiles_k * tiles_n);
  const long long split = blockIdx.x / (tiles_k * tiles_n);
  const int n0 = (tile / tiles_k) * TN_TILE, k0 = (tile % tiles_k) * TN_TILE;
  const int ncols = min(TN_TILE, N - n0), kcols = min(TN_TILE, K - k0);
  const long long m_begin = split * chunk, m_end = min(M, m_begin + chunk);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (warp & 3) * 32, wc = (warp >> 2) * 64;
  FragC acc0[4], acc1[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    wmma::fill_fragment(acc0[f], 0.0f);
    wmma::fill_fragment(acc1[f], 0.0f);
  }
  // plane tile [64 rows x `cols` valid of 128] -> smem, zero beyond the valid rows / columns (cols are even: N, K % 8 == 0)
  auto stage = [&](const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long ld, int rows_valid, int cols_valid,
                   __nv_bfloat16* dhi, __nv_bfloat16* dlo) {
    for (int idx = threadIdx.x; idx < BWD_TILE * (TN_TILE >> 3); idx += BWD_THREADS) {
      const int r = idx / (TN_TILE >> 3), c = (idx % (TN_TILE >> 3)) << 3;
      uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
      if (r < rows_valid && c < cols_valid) {  // 8-element groups are all in or all out
        h = *reinterpret_cast<const uint4*>(hi + r * ld + c);
        if (lo != nullptr) l = *reinterpret_cast<const uint4*>(lo + r * ld + c);
      }
      *reinterpret_cast<uint4*>(dhi + r * TP + c) = h;
      *reinterpret_cast<uint4*>(dlo + r * TP + c) = l;
    }
  };
  for (long long m0 = m_begin; m0 < m_end; m0 += BWD_TILE) {
    const long long mleft = m_end - m0;
    const int mrows = mleft < BWD_TILE ? static_cast<int>(mleft) : BWD_TILE;
    __syncthreads();
    stage(dy_hi + m0 * ldy + n0, dy_lo ? dy_lo + m0 * ldy + n0 : nullptr, ldy, mrows, ncols, Y_hi, Y_lo);
    stage(x_hi + m0 * ldx + k0, x_lo ? x_lo + m0 * ldx + k0 : nullptr, ldx, mrows, kcols, X_hi, X_lo);
    __syncthreads();
    bwd_mma<4, true, false>(acc0, Y_hi, Y_lo, TP, wr, X_hi, X_lo, TP, wc, BWD_TILE);
    bwd_mma<4, true, false>(acc1, Y_hi, Y_lo, TP, wr + 16, X_hi, X_lo, TP, wc, BWD_TILE);
  }
  float* stg = stage_all + warp * (16 * 20);
  for (int half = 0; half < 2; ++half) {
    for (int f = 0; f < 4; ++f) {
      wmma::store_matrix_sync(stg, half ? acc1[f] : acc0[f], 20, wmma::mem_row_major);
      __syncwarp();
      for (int e = lane; e < 256; e += 32) {
        const int r = e >> 4, c = e & 15;
        const int n = wr + 16 * half + r, k = wc + 16 * f + c;
        if (n < ncols && k < kcols) atomicAdd(dW + static_cast<long long>(n0 + n) * K + k0 + k, stg[r * 20 + c]);
      }
      __syncwarp();
    }
  }
}

// db[n] += sum_m dY[m, n] (dY given as planes; N and ldy multiples of 8, planes 16-byte aligned -- what
// lamp_gemm_tn_acc requires anyway).  HBM-bound: 4 B per element, read once.  256 threads = 8 row lanes (warps) x 32 column groups of 8
// columns: a warp reads 512 contiguous bytes of a row per plane with one 16-byte load per thread, four rows in flight
// per thread.  blockIdx.y selects the 256-column slab, blockIdx.x the `chunk` rows; the 8 row lanes are combined in
// shared memory and each column costs one atomic per block.
constexpr int COLSUM_THREADS = 256;
__device__ __forceinline__ void colsum_acc8(float (&s)[8], const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s[2 * j] += __uint_as_float(w[j] << 16);
    s[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__global__ void __launch_bounds__(COLSUM_THREADS) colsum_planes_kernel(const __nv_bfloat16* __restrict__ hi,
                                                                       const __nv_bfloat16* __restrict__ lo, long long ldy,
                                                                       long long M, int N, long long chunk,
                                                                       float* __restrict__ db) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n0 = (blockIdx.y * 32 + cg) * 8;
  const long long m_begin = blockIdx.x * chunk, m_end = min(M, m_begin + chunk);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (n0 < N) {
    const __nv_bfloat16* ph = hi + n0;
    const __nv_bfloat16* pl = lo != nullptr ? lo + n0 : nullptr;
    long long m = m_begin + rl;
    for (; m + 24 < m_end; m += 32) {   // rows m, m + 8, m + 16, m + 24 of this row lane
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = __ldg(reinterpret_cast<const uint4*>(ph + (m + 8 * u) * ldy));
      if (pl != nullptr) {
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = __ldg(reinterpret_cast<const uint4*>(pl + (m + 8 * u) * ldy));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) colsum_acc8(s, a[u]);
      if (pl != nullptr) {
#pragma unroll
        for (int u = 0; u < 4; ++u) colsum_acc8(s, b[u]);
      }
    }
    for (; m < m_end; m += 8) {
      colsum_acc8(s, __ldg(reinterpret_cast<const uint4*>(ph + m * ldy)));
      if (pl != nullptr) colsum_acc8(s, __ldg(reinterpret_cast<const uint4*>(pl + m * ldy)));
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = s[j];
  __syncthreads();
  const int n = blockIdx.y * 256 + threadIdx.x;
  if (n < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    atomicAdd(db + n, t);
  }
}

// Backward of the diagonal label projection logits[b, l] = <x[b, l, :], W[l, :]> (+ bias[l]) (lamp/Models.py:124-126):
//   dx[b, l, :] = g[b, l] * W[l, :]                       (one warp per (b, l) row)
//   dW[l, :]    = sum_b g[b, l] * x[b, l, :],  dbias[l] = sum_b g[b, l]   (blocks over labels x batch slices)
__global__ void diag_proj_bwd_dx_kernel(const float* __restrict__ g, const float* __restrict__ W, long long rows, int L,
                                        int D, float* __restrict__ dx) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float gv = g[row];
  const float4* wr = reinterpret_cast<const float4*>(W + static_cast<long long>(row % L) * D);
  float4* o = reinterpret_cast<float4*>(dx + row * D);
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    const float4 w = __ldg(wr + idx);
    o[idx] = make_float4(gv * w.x, gv * w.y, gv * w.z, gv * w.w);
  }
}
// dW / dbias: blockIdx.x = label, blockIdx.y = slice of the batch (`bchunk` samples); threads own 4 consecutive columns
// (one 16-byte load per sample row), four samples in flight; partial sums leave with one vector reduction per thread
// into the ZEROED dW (and one atomic per block into the zeroed dbias).
__global__ void diag_proj_bwd_dw_kernel(const float* __restrict__ g, const float* __restrict__ x, long long B, int L,
                                        int D, long long bchunk, float* __restrict__ dW, float* __restrict__ dbias) {
  const int l = blockIdx.x;
  const long long b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  const long long rs = static_cast<long long>(L) * D;   // distance between the rows of label l of consecutive samples
  float gs = 0.f;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xp = x + static_cast<long long>(l) * D + c;
    gs = 0.f;
    long long b = b0;
    for (; b + 3 < b1; b += 4) {
      float4 v[4];
      float gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = __ldg(reinterpret_cast<const float4*>(xp + (b + u) * rs));
        gv[u] = __ldg(g + (b + u) * L + l);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s.x = fmaf(gv[u], v[u].x, s.x); s.y = fmaf(gv[u], v[u].y, s.y);
        s.z = fmaf(gv[u], v[u].z, s.z); s.w = fmaf(gv[u], v[u].w, s.w);
        gs += gv[u];
      }
    }
    for (; b < b1; ++b) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xp + b * rs));
      const float gv = __ldg(g + b * L + l);
      s.x = fmaf(gv, v.x, s.x); s.y = fmaf(gv, v.y, s.y); s.z = fmaf(gv, v.z, s.z); s.w = fmaf(gv, v.w, s.w);
      gs += gv;
    }
    float* o = dW + static_cast<long long>(l) * D + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
  }
  if (dbias != nullptr && threadIdx.x == 0) atomicAdd(dbias + l, gs);  // thread 0 always owns columns 0..3
}

// Backward of the token + position embedding sum (lamp/Encoders.py:66,75; torch.nn.Embedding with padding_idx):
//   dword[seq[r], :] += g[r, :] unless seq[r] == pad_word;   dpos[pos[r], :] += g[r, :] unless pos[r] == pad_pos.
// One warp per gradient row, one 16-byte vector reduction per lane and table: rows of the same token / position meet
// in L2 (order of the additions is not fixed: fp32 sums differ in the last bits from run to run).  dword / dpos are
// accumulated into (the caller zeroes them); either may be NULL.
__global__ void embed_bwd_kernel(const float* __restrict__ g, const long long* __restrict__ seq,
                                 const long long* __restrict__ pos, long long rows, int D, long long pad_word,
                                 long long pad_pos, float* __restrict__ dword, float* __restrict__ dpos) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* w = nullptr;
  float* q = nullptr;
  if (dword != nullptr) {
    const long long id = seq[row];
    if (id != pad_word) w = dword + id * D;
  }
  if (dpos != nullptr) {
    const long long id = pos[row];
    if (id != pad_pos) q = dpos + id * D;
  }
  if (w == nullptr && q == nullptr) return;
  const float4* gr = reinterpret_cast<const float4*>(g + row * D);
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    const float4 v = __ldg(gr + idx);
    if (w != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(w + 4 * idx), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (q != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 4 * idx), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

// ---- device-side training targets and loss (train.py:34-38; utils/utils.py:205-216) --------------------------------
// get_gold_binary: `gold` [B, W] int64 holds each document's label ids (+4 for the special tokens), terminated by EOS
// and padded with PAD = 0.  Per row the reference keeps the entries > 0, DROPS THE LAST ONE of them (the EOS), sets
// out[b, id] = 1 in a [B, L + 4] matrix and cuts its first `skip` = 4 columns.  One warp per row; `out` is fully
// written (zeros included), so the caller does not clear it.  Ids whose column falls outside [0, L) are ignored.
__global__ void gold_binary_kernel(const long long* __restrict__ gold, long long B, int W, int L, int skip,
                                   float* __restrict__ out) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const long long* g = gold + row * W;
  float* o = out + row * L;
  for (int c = lane; c < L; c += 32) o[c] = 0.0f;
  int last = -1;  // position of the last entry > 0
  for (int j = lane; j < W; j += 32)
    if (g[j] > 0) last = j;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, off));
  __syncwarp();  // the zeros above are ordered before the ones below (same warp, same row)
  for (int j = lane; j < W; j += 32) {
    const long long id = g[j];
    if (id > 0 && j != last) {
      const long long c = id - skip;
      if (c >= 0 && c < L) o[c] = 1.0f;
    }
  }
}

// F.binary_cross_entropy_with_logits(x, y, reduction='mean') and its gradient in one pass:
//   loss = mean( max(x, 0) - x y + log(1 + exp(-|x|)) ),  dx = (sigmoid(x) - y) / n.
// Deterministic: every block writes one partial sum, the LAST block to finish (ticket counter) adds them in index order.
constexpr int BCE_THREADS = 256;
__global__ void __launch_bounds__(BCE_THREADS)
bce_logits_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, float inv_n,
                  float* __restrict__ dx, float* __restrict__ partials, unsigned int* __restrict__ ticket,
                  float* __restrict__ loss) {
  __shared__ float red[BCE_THREADS / 32];
  __shared__ bool is_last;
  float s = 0.0f;
  for (long long i = static_cast<long long>(blockIdx.x) * BCE_THREADS + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * BCE_THREADS) {
    const float xv = x[i], yv = y[i];
    const float e = __expf(-fabsf(xv));
    s += fmaxf(xv, 0.0f) - xv * yv + log1pf(e);
    if (dx != nullptr) {
      const float sig = xv >= 0.0f ? 1.0f / (1.0f + e) : e / (1.0f + e);
      dx[i] = (sig - yv) * inv_n;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < BCE_THREADS / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.0f;
    for (unsigned b = 0; b < gridDim.x; ++b) t += reinterpret_cast<volatile float*>(partials)[b];
    *loss = t * inv_n;
    *ticket = 0u;  // ready for the next launch (CUDA-graph replays reuse the workspace)
  }
}

// ---- element-wise pieces of the fused training sub-layers (lamp/SubLayers.py:113-119,136-141) -----------------------
// The element-wise dropout after fc / w_2 uses the same counter hash as the attention dropout: keep(row, col) is a pure
// function of (seed, row, col), so the backward recomputes the mask instead of storing it.
//   y = keep * y0 / (1 - p) + x           (x: residual; row index modulo `x_mod` when the residual is broadcast)
__global__ void dropout_add_kernel(const float* __restrict__ y0, const float* __restrict__ x, long long rows, int D,
                                   int x_mod, uint32_t thresh, float scale, unsigned long long seed,
                                   const unsigned long long* __restrict__ seed_dev, float* __restrict__ y) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint32_t rh = drop_rowhash(seed + (seed_dev ? __ldg(seed_dev) : 0ull), static_cast<unsigned long long>(row));
  const float4* a = reinterpret_cast<const float4*>(y0 + row * D);
  const float4* r = reinterpret_cast<const float4*>(x + (x_mod ? row % x_mod : row) * D);
  float4* o = reinterpret_cast<float4*>(y + row * D);
  for (int c = lane; c < (D >> 2); c += 32) {
    float4 v = a[c];
    const float4 rv = __ldg(r + c);
    const uint32_t col = 4u * c;
    v.x = (thresh && !drop_keep(rh, col, thresh) ? 0.f : v.x * scale) + rv.x;
    v.y = (thresh && !drop_keep(rh, col + 1, thresh) ? 0.f : v.y * scale) + rv.y;
    v.z = (thresh && !drop_keep(rh, col + 2, thresh) ? 0.f : v.z * scale) + rv.z;
    v.w = (thresh && !drop_keep(rh, col + 3, thresh) ? 0.f : v.w * scale) + rv.w;
    o[c] = v;
  }
}
//   planes(keep * dy / (1 - p)): the gradient of the dropped branch, directly as the tensor-core operand of the dW / dx
//   products (thresh == 0: a plain split)
__global__ void dropout_split_kernel(const float* __restrict__ dy, long long rows, int D, uint32_t thresh, float scale,
                                     unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint32_t rh = drop_rowhash(seed + (seed_dev ? __ldg(seed_dev) : 0ull), static_cast<unsigned long long>(row));
  const float4* a = reinterpret_cast<const float4*>(dy + row * D);
  uint2* oh = reinterpret_cast<uint2*>(hi + row * D);
  uint2* ol = reinterpret_cast<uint2*>(lo + row * D);
  for (int c = lane; c < (D >> 2); c += 32) {
    float4 v = a[c];
    const uint32_t col = 4u * c;
    v.x = (thresh && !drop_keep(rh, col, thresh)) ? 0.f : v.x * scale;
    v.y = (thresh && !drop_keep(rh, col + 1, thresh)) ? 0.f : v.y * scale;
    v.z = (thresh && !drop_keep(rh, col + 2, thresh)) ? 0.f : v.z * scale;
    v.w = (thresh && !drop_keep(rh, col + 3, thresh)) ? 0.f : v.w * scale;
    uint2 h, l;
    split_bf16x2(v.x, v.y, h.x, l.x);
    split_bf16x2(v.z, v.w, h.y, l.y);
    oh[c] = h;
    ol[c] = l;
  }
}
//   ReLU backward on operand planes, in place: dh = 0 where the forward activation h (its hi plane decides: h > 0 iff
//   hi > 0, the split keeps the sign) was clipped
__global__ void relu_mask_planes_kernel(__nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                                        const __nv_bfloat16* __restrict__ h_hi, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(h_hi) + i);
  uint4 a = reinterpret_cast<uint4*>(g_hi)[i], b = reinterpret_cast<uint4*>(g_lo)[i];
  auto keep = [](uint32_t hv) -> uint32_t {  // two bf16 per word: > 0 <=> sign clear and magnitude non-zero
    const uint32_t lo = ((hv & 0x8000u) == 0u && (hv & 0x7FFFu) != 0u) ? 0x0000FFFFu : 0u;
    const uint32_t hi = ((hv & 0x80000000u) == 0u && (hv & 0x7FFF0000u) != 0u) ? 0xFFFF0000u : 0u;
    return lo | hi;
  };
  const uint32_t k0 = keep(h.x), k1 = keep(h.y), k2 = keep(h.z), k3 = keep(h.w);
  a.x &= k0; a.y &= k1; a.z &= k2; a.w &= k3;
  b.x &= k0; b.y &= k1; b.z &= k2; b.w &= k3;
  reinterpret_cast<uint4*>(g_hi)[i] = a;
  reinterpret_cast<uint4*>(g_lo)[i] = b;
}

}  // namespace lamp
