// HBM-bound row kernels around the tensor-core stages: plane split, (add +) LayerNorm, embedding gather,
// diagonal label projection.  All are one-pass, 128-bit vectorised, one warp per row where a row reduction exists.
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

// fp32 [rows, cols] (leading dim ld) -> hi/lo bf16 planes (leading dim ldp).  cols % 4 == 0.
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int cols, long long ld,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp) {
  const int c4 = cols >> 2;
  const long long total = rows * c4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4;
    const int c = static_cast<int>(i % c4) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ld + c);
    uint2 h, l;
    split_bf16x2(v.x, v.y, h.x, l.x);
    split_bf16x2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + r * ldp + c) = h;
    if (lo != nullptr) *reinterpret_cast<uint2*>(lo + r * ldp + c) = l;
  }
}

// Many small splits in ONE launch (the weight matrices of a training step: every projection weight is needed as
// planes, W for the forward products and W^T for the input-gradient products, and changes every step).  Job j:
// fp32 src [rows, cols] (leading dim ld) -> planes dst [rows, cols] or, with `transpose`, dst [cols, rows] (leading dim
// ldp; dst may be a row / column slice of a wider matrix: that is how Wq | Wk | Wv land in one [3*H*d, D] operand without
// a concatenation).  blockIdx.y = job, blockIdx.x = 32 x 32 tile; 256 threads, tile transposed through shared memory.
struct SplitJob {
  const float* src;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;   // nullable
  int rows, cols;
  long long ld, ldp;
  int transpose;
};
constexpr int SPLIT_MULTI_MAX_JOBS = 56;   // 56 x 56 B of kernel parameters
struct SplitJobs {
  SplitJob job[SPLIT_MULTI_MAX_JOBS];
};
__global__ void __launch_bounds__(256) split_planes_multi_kernel(const __grid_constant__ SplitJobs jobs) {
  __shared__ float tile[32][33];
  const SplitJob& jb = jobs.job[blockIdx.y];
  const int tiles_c = (jb.cols + 31) >> 5, tiles_r = (jb.rows + 31) >> 5;
  if (static_cast<int>(blockIdx.x) >= tiles_c * tiles_r) return;
  const int r0 = (blockIdx.x / tiles_c) << 5, c0 = (blockIdx.x % tiles_c) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 row lanes
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < jb.rows && c < jb.cols) ? jb.src[static_cast<long long>(r) * jb.ld + c] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = ty + 8 * i;   // row of the OUTPUT tile handled by this thread, column tx
    float v;
    long long off;
    bool ok;
    if (jb.transpose) {
      v = tile[tx][a];          // out[c0 + a][r0 + tx] = src[r0 + tx][c0 + a]
      ok = (c0 + a < jb.cols) && (r0 + tx < jb.rows);
      off = static_cast<long long>(c0 + a) * jb.ldp + r0 + tx;
    } else {
      v = tile[a][tx];
      ok = (r0 + a < jb.rows) && (c0 + tx < jb.cols);
      off = static_cast<long long>(r0 + a) * jb.ldp + c0 + tx;
    }
    if (ok) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      jb.hi[off] = h;
      if (jb.lo != nullptr) jb.lo[off] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// out = LayerNorm(y (+ add)) * gamma + beta  (biased variance, eps inside the sqrt: torch.nn.LayerNorm, used at
// lamp/SubLayers.py:117,141).  One warp per row, the row lives in registers (D <= 32*4*MAXV), two-pass variance.
// `add` rows are indexed modulo add_mod when add_mod > 0 (label embeddings shared by every sample).
template <int MAXV>
__global__ void layernorm_kernel(const float* __restrict__ y, const float* __restrict__ add, int add_mod,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 long long rows, int D, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                 __nv_bfloat16* __restrict__ out_lo, const int* __restrict__ m_dev) {
  griddep_launch_dependents();
  griddep_wait();
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m_dev != nullptr && rows > __ldg(m_dev)) rows = __ldg(m_dev);  // device-side row count (padding-aware runs)
  if (row >= rows) return;
  const int d4 = D >> 2;
  const float4* yr = reinterpret_cast<const float4*>(y + row * D);
  const float4* ar = nullptr;
  if (add != nullptr) ar = reinterpret_cast<const float4*>(add + (add_mod ? row % add_mod : row) * D);
  float4 v[MAXV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < d4) {
      v[i] = yr[idx];
      if (ar != nullptr) {
        const float4 a = ar[idx];
        v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
      }
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  const float mean = s / D;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < d4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
  const float rstd = rsqrtf(q / D + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < d4) {
      const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (out != nullptr) reinterpret_cast<float4*>(out + row * D)[idx] = o;
      if (out_hi != nullptr) {
        uint2 h, l;
        split_bf16x2(o.x, o.y, h.x, l.x);
        split_bf16x2(o.z, o.w, h.y, l.y);
        reinterpret_cast<uint2*>(out_hi + row * D)[idx] = h;
        if (out_lo != nullptr) reinterpret_cast<uint2*>(out_lo + row * D)[idx] = l;
      }
    }
  }
}

// enc_input[r, :] = word_emb[src_seq[i], :] (+ pos_emb[src_pos[i], :]),  i = row_index ? row_index[r] : r
// -- lamp/Encoders.py:66,75.  Writes fp32 and/or planes.  One warp per output row.  With `row_index` / `m_dev` the
// kernel gathers only the first *m_dev rows of an index list (the non-PAD tokens of the batch, packed).
__global__ void embed_kernel(const long long* __restrict__ seq, const long long* __restrict__ pos,
                             const float* __restrict__ word_emb, const float* __restrict__ pos_emb, long long rows,
                             int D, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                             __nv_bfloat16* __restrict__ out_lo, const long long* __restrict__ row_index,
                             const int* __restrict__ m_dev) {
  griddep_launch_dependents();
  griddep_wait();
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m_dev != nullptr && rows > __ldg(m_dev)) rows = __ldg(m_dev);
  if (row >= rows) return;
  const long long src = row_index ? row_index[row] : row;
  const float4* w = reinterpret_cast<const float4*>(word_emb + seq[src] * D);
  const float4* q = pos_emb ? reinterpret_cast<const float4*>(pos_emb + pos[src] * D) : nullptr;
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    float4 v = __ldg(w + idx);
    if (q != nullptr) {
      const float4 a = __ldg(q + idx);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (out != nullptr) reinterpret_cast<float4*>(out + row * D)[idx] = v;
    if (out_hi != nullptr) {
      uint2 h, l;
      split_bf16x2(v.x, v.y, h.x, l.x);
      split_bf16x2(v.z, v.w, h.y, l.y);
      reinterpret_cast<uint2*>(out_hi + row * D)[idx] = h;
      if (out_lo != nullptr) reinterpret_cast<uint2*>(out_lo + row * D)[idx] = l;
    }
  }
}

// out[r, :] = src[index[r], :]  (fp32 rows; un-packs the encoder output into the dense [B, T, D] API tensor, the
// representative PAD row being replicated to every PAD position).  One warp per output row.
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ index, long long rows,
                                   int D, float* __restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + index[row] * D);
  float4* o4 = reinterpret_cast<float4*>(out + row * D);
  for (int idx = lane; idx < (D >> 2); idx += 32) o4[idx] = __ldg(s4 + idx);
}

// Zero the `nguard` rows that follow the *m_dev rows in use of a packed plane matrix.  A KV tile of the last sample
// may extend a few rows past the packed data; those keys carry probability 0, and 0 * (uninitialised NaN) must not
// reach the PV product.
__global__ void zero_guard_rows_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld,
                                       int cols, const int* __restrict__ m_dev, long long max_rows, int nguard) {
  const long long first = __ldg(m_dev);
  const long long total = static_cast<long long>(nguard) * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = first + i / cols;
    if (r < max_rows) {
      hi[r * ld + i % cols] = __float2bfloat16(0.0f);
      if (lo != nullptr) lo[r * ld + i % cols] = __float2bfloat16(0.0f);
    }
  }
}

// logits[b, l] = <x[b, l, :], W[l, :]> (+ bias[l])  -- the diagonal of the reference's [B, L, L] projection
// (lamp/Models.py:124-126) without the L-fold redundant work.  One warp per (b, l), fp32 FMA.
__global__ void diag_proj_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                 const float* __restrict__ bias, long long rows, int L, int D,
                                 float* __restrict__ logits) {
  griddep_launch_dependents();
  griddep_wait();
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int l = static_cast<int>(row % L);
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  const float4* wr = reinterpret_cast<const float4*>(W + static_cast<long long>(l) * D);
  float s = 0.0f;
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    const float4 a = xr[idx], b = __ldg(wr + idx);
    s += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) logits[row] = s + (bias ? bias[l] : 0.0f);
}

// ---- deferred LayerNorm (see GemmParams in gemm_planes.cuh): the activation travels as PRE-norm planes y plus per-row
// partial sums {sum y, sum y^2}; these kernels apply the normalisation where a real tensor has to leave the GEMM chain.
__device__ __forceinline__ void dln_row_stats(const float2* __restrict__ stats, int nparts, long long row, int D,
                                              float eps, float& mean, float& rstd) {
  float s1 = 0.0f, s2 = 0.0f;
  for (int q = 0; q < nparts; ++q) {
    const float2 e = __ldg(stats + row * nparts + q);
    s1 += e.x;
    s2 += e.y;
  }
  mean = s1 / D;
  rstd = rsqrtf(fmaxf(s2 / D - mean * mean, 0.0f) + eps);
}

__device__ __forceinline__ float4 planes_load4(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                               long long off) {
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + off));
  float4 v = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u), __uint_as_float(h.y << 16),
                         __uint_as_float(h.y & 0xFFFF0000u));
  if (lo != nullptr) {
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(lo + off));
    v.x += __uint_as_float(l.x << 16); v.y += __uint_as_float(l.x & 0xFFFF0000u);
    v.z += __uint_as_float(l.y << 16); v.w += __uint_as_float(l.y & 0xFFFF0000u);
  }
  return v;
}

// out[r, :] = LayerNorm(y[index ? index[r] : r, :]) * gamma + beta -> fp32 and/or planes.  One warp per output row.
// With `index` this is also the un-packing gather of the encoder output (dense [B, T, D] API tensor).
__global__ void ln_apply_kernel(const __nv_bfloat16* __restrict__ y_hi, const __nv_bfloat16* __restrict__ y_lo,
                                const float2* __restrict__ stats, int nparts, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, long long rows, int D,
                                const long long* __restrict__ index, float* __restrict__ out,
                                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                const int* __restrict__ m_dev) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m_dev != nullptr && rows > __ldg(m_dev)) rows = __ldg(m_dev);
  if (row >= rows) return;
  const long long src = index ? index[row] : row;
  float mean, rstd;
  dln_row_stats(stats, nparts, src, D, eps, mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    const float4 v = planes_load4(y_hi, y_lo, src * D + 4 * idx);
    const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + b.x;
    o.y = (v.y - mean) * rstd * g.y + b.y;
    o.z = (v.z - mean) * rstd * g.z + b.z;
    o.w = (v.w - mean) * rstd * g.w + b.w;
    if (out != nullptr) reinterpret_cast<float4*>(out + row * D)[idx] = o;
    if (out_hi != nullptr) {
      uint2 h, l;
      split_bf16x2(o.x, o.y, h.x, l.x);
      split_bf16x2(o.z, o.w, h.y, l.y);
      reinterpret_cast<uint2*>(out_hi + row * D)[idx] = h;
      if (out_lo != nullptr) reinterpret_cast<uint2*>(out_lo + row * D)[idx] = l;
    }
  }
}

// logits[b, l] = <LayerNorm(y[b, l, :]) * gamma + beta, W[l, :]> (+ bias[l]): the diagonal label projection
// (lamp/Models.py:124-126) applied directly to the deferred output of the last decoder layer.
__global__ void diag_proj_ln_kernel(const __nv_bfloat16* __restrict__ y_hi, const __nv_bfloat16* __restrict__ y_lo,
                                    const float2* __restrict__ stats, int nparts, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float eps, const float* __restrict__ W,
                                    const float* __restrict__ bias, long long rows, int L, int D,
                                    float* __restrict__ logits) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int l = static_cast<int>(row % L);
  float mean, rstd;
  dln_row_stats(stats, nparts, row, D, eps, mean, rstd);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  const float4* wr = reinterpret_cast<const float4*>(W + static_cast<long long>(l) * D);
  float s = 0.0f;
  for (int idx = lane; idx < (D >> 2); idx += 32) {
    const float4 v = planes_load4(y_hi, y_lo, row * D + 4 * idx);
    const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx), w = __ldg(wr + idx);
    s += (((v.x - mean) * rstd * g.x + b.x) * w.x + ((v.y - mean) * rstd * g.y + b.y) * w.y) +
         (((v.z - mean) * rstd * g.z + b.z) * w.z + ((v.w - mean) * rstd * g.w + b.w) * w.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) logits[row] = s + (bias ? bias[l] : 0.0f);
}

// Byte mask (non-zero = masked, element strides msb/msq/msk) -> bit words [Bm][Lq][W], W = ceil(Lk / 32); bit (k & 31) of
// word k >> 5 belongs to key k, keys >= Lk read as 0.  One warp per output word.
__global__ void pack_mask_kernel(const uint8_t* __restrict__ mask, long long msb, long long msq, long long msk,
                                 long long Bm, int Lq, int Lk, uint32_t* __restrict__ words) {
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int W = (Lk + 31) >> 5;
  if (wid >= Bm * Lq * W) return;
  const int w = static_cast<int>(wid % W);
  const int q = static_cast<int>((wid / W) % Lq);
  const long long b = wid / (static_cast<long long>(W) * Lq);
  const int k = w * 32 + lane;
  const bool m = k < Lk && mask[b * msb + q * msq + k * msk] != 0;
  const uint32_t bits = __ballot_sync(0xFFFFFFFFu, m);
  if (lane == 0) words[wid] = bits;
}

}  // namespace lamp
