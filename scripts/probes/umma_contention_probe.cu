// tcgen05.mma rate under contention (measurement tool, not part of the product).
//
// Inside the attention core a tcgen05.mma costs ~93-105 clocks, alone 55-64 (scripts/probes/umma_probe.cu).  This probe
// runs the S-like instruction stream (SS form, M = 128, N = 64, 3 instructions per K-step on alternating hi / lo operand
// tiles, one accumulator) from one thread and switches on, one at a time, the other activities of that kernel:
//   bit 0  tcgen05.ld traffic   : 8 warps read 32 TMEM columns each in a loop (the softmax warps reading S)
//   bit 1  tcgen05.st traffic   : the same warps write 16 columns back (P written over S)
//   bit 2  shared-memory writes : bulk async copies global -> shared (the K / V ring being refilled by TMA)
//   bit 3  a second issuing thread running the PV-like stream (TS form, N = 128) into another accumulator
//   bit 4  MUFU / ALU load      : the 8 warps also run an ex2 + fma loop (the softmax arithmetic)
// Reported: clocks per S instruction (difference quotient over 96 vs 384 instructions), all SMs busy.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lamp_b200/csrc
//        -o scripts/probes/umma_contention_probe scripts/probes/umma_contention_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sm100_primitives.cuh"

using namespace lamp;

constexpr int CTRL_WARPS = 3;   // 0: S issuer, 1: PV issuer, 2: bulk-copy producer
constexpr int BG_WARPS = 8;
constexpr int THREADS = 32 * (CTRL_WARPS + BG_WARPS);
constexpr uint32_t TILE = 128 * 128;   // one [128 rows x 64 bf16] swizzled tile
constexpr uint32_t RING_BYTES = 4 * 32768;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) probe(int n_instr, int mode, const uint8_t* gsrc, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                    // hi | lo tiles of "Q": 2 x 16 KB
  uint8_t* sK = smem + 2 * TILE;         // hi | lo tiles of "K" (64 rows used)
  uint8_t* sV = smem + 4 * TILE;         // "V": 2 x 16 KB
  uint8_t* ring = smem + 6 * TILE;       // bulk-copy landing zone
  uint64_t* bar = reinterpret_cast<uint64_t*>(ring + RING_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 10);
  for (int i = threadIdx.x; i < (6 * TILE) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    *stop = 0;
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;  // S acc [0,64) | PV acc [128,256) | P-like A operand [256,320) | ld/st area [320,448)
  if (warp >= CTRL_WARPS) {
    // background warps: initialise the TMEM areas they touch, then loop until told to stop
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t v[16];
    for (int e = 0; e < 16; ++e) v[e] = 0x3F803F80u;
    for (int c = 256; c < 448; c += 16) tmem_st16(tmem_base + lane_sel + c, v);
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t dq = umma_smem_desc(smem_u32(sQ), 16, 1024), dk = umma_smem_desc(smem_u32(sK), 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n_instr; i += 3) {
      const uint32_t koff = ((i / 3) & 3) * 2;
      umma_bf16_ss(tmem_base, dq + koff, dk + koff, idesc, i != 0 ? 1u : 0u);
      umma_bf16_ss(tmem_base, dq + koff, dk + koff + (TILE >> 4), idesc, 1u);
      umma_bf16_ss(tmem_base, dq + koff + (TILE >> 4), dk + koff, idesc, 1u);
    }
    umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    out[blockIdx.x] = clock64() - t0;
    *stop = 1;
  } else if (warp == 1 && lane == 0 && (mode & 8)) {
    const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 1);
    const uint64_t dv = umma_smem_desc(smem_u32(sV), 8192, 1024);
    uint32_t it = 0;
    while (!*stop) {
      for (int t = 0; t < 12; ++t) umma_bf16_ts(tmem_base + 128, tmem_base + 256 + 8 * (t & 3), dv + (t & 7) * 128u, idesc, 1u);
      umma_commit(&bar[1]);
      mbar_wait(&bar[1], it & 1);
      ++it;
    }
  } else if (warp == 2 && lane == 0 && (mode & 4)) {
    uint32_t it = 0;
    while (!*stop) {
      mbar_arrive_expect_tx(&bar[2], RING_BYTES);
      for (int s = 0; s < 4; ++s) bulk_g2s(ring + s * 32768, gsrc + (static_cast<size_t>(blockIdx.x) * 64 + ((it * 4 + s) & 63)) * 32768, 32768, &bar[2]);
      mbar_wait(&bar[2], it & 1);
      ++it;
    }
  } else if (warp >= CTRL_WARPS) {
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t col = 320 + 32 * ((warp - CTRL_WARPS) >> 2);
    float acc = 0.f;
    while (!*stop) {
      uint32_t r[32];
      if (mode & 1) {
        tmem_ld32(tmem_base + lane_sel + col, r);
        tmem_wait_ld();
      } else {
        for (int e = 0; e < 32; ++e) r[e] = 0x3F800000u + e;
      }
      if (mode & 16) {
        for (int e = 0; e < 32; ++e) {
          float x;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(__uint_as_float(r[e]) * 0.001f));
          acc = fmaf(x, 1.0001f, acc);
          r[e] = __float_as_uint(x);
        }
      }
      if (mode & 2) {
        uint32_t w[16];
        for (int e = 0; e < 16; ++e) w[e] = r[e] ^ r[e + 16];
        tmem_st16(tmem_base + lane_sel + col, w);
        tmem_wait_st();
      }
      if (!(mode & 19)) __nanosleep(200);
      acc += __uint_as_float(r[lane & 31]);
    }
    if (acc == 123.456f) sink[0] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* d_out;
  float* d_sink;
  uint8_t* d_src;
  cudaMalloc(&d_out, sizeof(long long) * sms);
  cudaMalloc(&d_sink, 64);
  const size_t src_bytes = static_cast<size_t>(sms) * 64 * 32768;
  cudaMalloc(&d_src, src_bytes);
  cudaMemset(d_src, 0x3C, src_bytes);
  const size_t smem = 6 * TILE + RING_BYTES + 1024 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<long long> h(sms);
  const char* names[] = {"alone", "+tmem ld", "+tmem st", "+tmem ld+st", "+smem bulk writes", "+second issuer (PV)", "+ex2/fma only",
                         "+ld+st+ex2 (softmax-like)", "+softmax-like +bulk", "+softmax-like +bulk +second issuer (all)"};
  const int modes[] = {0, 1, 2, 3, 4, 8, 16, 19, 23, 31};
  for (int m = 0; m < 10; ++m) {
    double per[2];
    const int counts[2] = {96, 384};
    for (int c = 0; c < 2; ++c) {
      probe<<<sms, THREADS, smem>>>(counts[c], modes[m], d_src, d_out, d_sink);
      probe<<<sms, THREADS, smem>>>(counts[c], modes[m], d_src, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("{\"probe\": \"umma_contention\", \"mode\": \"%s\", \"error\": \"%s\"}\n", names[m], cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      double s = 0;
      for (int b = 0; b < sms; ++b) s += h[b];
      per[c] = s / sms;
    }
    printf("{\"probe\": \"umma_contention\", \"mode\": \"%s\", \"mode_bits\": %d, \"clk_per_S_mma\": %.1f}\n", names[m], modes[m],
           (per[1] - per[0]) / (counts[1] - counts[0]));
    fflush(stdout);
  }
  return 0;
}
