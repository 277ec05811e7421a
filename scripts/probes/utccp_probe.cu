// tcgen05.cp (smem -> TMEM) probe (measurement / validation tool, not part of the product).
//
// Question: can the A operand of a tcgen05.mma (a 128B-swizzled K-major [128 x 64] bf16 tile, as TMA writes it) be moved
// into tensor memory with tcgen05.cp.128x256b -- one instruction per K = 16 step, given the SAME shared-memory descriptor
// the SS-form MMA would use -- and then be consumed in TS form?  If yes, the attention core can keep Q in TMEM and free
// its shared-memory tile for the next item's Q right away.
//
// Test: Q [128 x 64] and K [64 x 64] bf16 with random bits of moderate magnitude are written to shared memory in the
// 128B-swizzle layout by the threads themselves; D_ss = Q K^T with both operands from shared memory; Q is copied to
// TMEM (4 x tcgen05.cp.128x256b) and D_ts = Q(TMEM) K^T; both accumulators are read back and compared bit for bit, and
// the TMEM image of Q is compared with the source rows.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lamp_b200/csrc -o scripts/probes/utccp_probe
//        scripts/probes/utccp_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sm100_primitives.cuh"

using namespace lamp;

__device__ __forceinline__ void utccp_128x256b(uint32_t dst_tmem, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(sdesc) : "memory");
}

// out: [0, 128*64) D_ss | [128*64, 2*128*64) D_ts | then 128*32 words: TMEM image of Q (32 columns per lane)
__global__ void __launch_bounds__(128, 1) probe(const uint16_t* q, const uint16_t* k, uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;           // 128 rows x 128 B
  uint8_t* sK = smem + 16384;   // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  // 128B swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sQ + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(q + r * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sK + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(k + r * 64 + c * 8);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;  // D_ss: cols [0,64) | D_ts: [64,128) | Q: [128,160)
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    for (int t = 0; t < 4; ++t) {
      const uint64_t dq = umma_smem_desc(smem_u32(sQ) + t * 32, 16, 1024);
      const uint64_t dk = umma_smem_desc(smem_u32(sK) + t * 32, 16, 1024);
      umma_bf16_ss(tmem_base, dq, dk, idesc, t != 0 ? 1u : 0u);
    }
    for (int t = 0; t < 4; ++t) {
      const uint64_t dq = umma_smem_desc(smem_u32(sQ) + t * 32, 16, 1024);
      utccp_128x256b(tmem_base + 128 + 8 * t, dq);
    }
    for (int t = 0; t < 4; ++t) {
      const uint64_t dk = umma_smem_desc(smem_u32(sK) + t * 32, 16, 1024);
      umma_bf16_ts(tmem_base + 64, tmem_base + 128 + 8 * t, dk, idesc, t != 0 ? 1u : 0u);
    }
    umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;
  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {      // D_ss then D_ts: 64 columns each = 2 x 32
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tmem_base + lane_sel + half * 64 + c0, r);
      tmem_wait_ld();
      for (int e = 0; e < 32; ++e) out[half * 128 * 64 + row * 64 + c0 + e] = r[e];
    }
  }
  tmem_ld32(tmem_base + lane_sel + 128, r);
  tmem_wait_ld();
  for (int e = 0; e < 32; ++e) out[2 * 128 * 64 + row * 32 + e] = r[e];
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int main() {
  std::vector<uint16_t> hq(128 * 64), hk(64 * 64);
  srand(1);
  auto rnd_bf16 = []() -> uint16_t {  // sign, exponent in [120, 130], random mantissa: finite, moderate magnitude
    return static_cast<uint16_t>(((rand() & 1) << 15) | ((120 + rand() % 11) << 7) | (rand() & 0x7F));
  };
  for (auto& v : hq) v = rnd_bf16();
  for (auto& v : hk) v = rnd_bf16();
  uint16_t *dq, *dk;
  uint32_t* dout;
  const size_t nout = 2 * 128 * 64 + 128 * 32;
  cudaMalloc(&dq, hq.size() * 2);
  cudaMalloc(&dk, hk.size() * 2);
  cudaMalloc(&dout, nout * 4);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xFF, nout * 4);
  const size_t smem = 16384 + 8192 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dq, dk, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"probe\": \"utccp\", \"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint32_t> ho(nout);
  cudaMemcpy(ho.data(), dout, nout * 4, cudaMemcpyDeviceToHost);
  long long d_mismatch = 0, q_mismatch = 0;
  for (int i = 0; i < 128 * 64; ++i) d_mismatch += ho[i] != ho[128 * 64 + i];
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 32; ++c) {
      const uint32_t want = static_cast<uint32_t>(hq[r * 64 + 2 * c]) | (static_cast<uint32_t>(hq[r * 64 + 2 * c + 1]) << 16);
      q_mismatch += ho[2 * 128 * 64 + r * 32 + c] != want;
    }
  // reference check of D_ss on the host (fp32 accumulation of exact bf16 products; order differs -> tolerance)
  double max_rel = 0;
  for (int r = 0; r < 128; r += 17)
    for (int c = 0; c < 64; c += 5) {
      double acc = 0, mag = 0;
      for (int kk = 0; kk < 64; ++kk) {
        uint32_t a = static_cast<uint32_t>(hq[r * 64 + kk]) << 16, b = static_cast<uint32_t>(hk[c * 64 + kk]) << 16;
        float fa, fb;
        memcpy(&fa, &a, 4);
        memcpy(&fb, &b, 4);
        acc += static_cast<double>(fa) * fb;
        mag += fabs(static_cast<double>(fa) * fb);
      }
      float got;
      memcpy(&got, &ho[r * 64 + c], 4);
      const double rel = fabs(got - acc) / (mag + 1e-30);
      if (rel > max_rel) max_rel = rel;
    }
  printf("{\"probe\": \"utccp\", \"d_ts_vs_d_ss_mismatching_words\": %lld, \"q_tmem_image_mismatching_words\": %lld, "
         "\"d_ss_vs_host_max_rel\": %.3g}\n", d_mismatch, q_mismatch, max_rel);
  return 0;
}
