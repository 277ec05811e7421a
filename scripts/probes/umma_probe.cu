// tcgen05.mma cost probe (measurement tool, not part of the product).
//
// Question behind it: what does ONE tcgen05.mma cost as a function of its N when a single
// thread issues a stream of them -- the situation of the attention core's S = Q K^T (N = 64 / 128 keys, K = head width)
// and of a per-(sample, head) fused projection (N = 128, K = d_model)?  The programming guides give the dispatch floor
// max(M,128) * N / 256 clocks; the attention traces of round 1 suggested a fixed extra cost per instruction.
//
// For M = 128 and N in {16 .. 256}, kind::f16 (bf16 operands, fp32 accumulate):
//   form  SS  : A and B from shared memory (128B-swizzled K-major tiles)          -- QK^T, projections
//   form  TS  : A from tensor memory, B from shared memory                        -- P V
//   chain dep : all instructions accumulate into ONE TMEM accumulator             -- a K loop
//   chain alt : instructions alternate between TWO accumulators                   -- two independent K loops interleaved
// One elected thread per CTA issues `n` instructions back to back and commits once; clock64() around issue + completion
// (mbarrier wait).  Reported: clocks per instruction = (t_done - t_start) / n for n = 64 and n = 512 (the difference
// quotient removes the fixed launch/commit latency), with 1 CTA and with one CTA on every SM.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lamp_b200/csrc -o scripts/probes/umma_probe
//        scripts/probes/umma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sm100_primitives.cuh"

using namespace lamp;

struct Result {
  long long clocks;
};

// FORM 0 = SS, 1 = TS.  ALT: alternate between two accumulators.
template <int FORM, bool ALT>
__global__ void __launch_bounds__(128, 1) probe(int N, int n_instr, Result* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B (64 bf16 of K): 16 KB
  uint8_t* sB = smem + 16384;         // 256 rows x 128 B: 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  // finite operand bytes (bf16 1.0 = 0x3F80) so that nothing denormal/NaN-specific is measured
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (FORM == 1) {
    // A operand in TMEM: columns [448, 512) of every lane, bf16 pairs of 1.0
    uint32_t v[16];
    for (int e = 0; e < 16; ++e) v[e] = 0x3F803F80u;
    const uint32_t lane_sel = static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16;
    for (int c = 0; c < 64; c += 16) tmem_st16(tmem_base + lane_sel + 448 + c, v);
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int i = 0; i < n_instr; ++i) {
      const uint32_t koff = (i & 3) * 32;  // walk the 4 K-steps of the 64-wide swizzled tile like a real K loop
      const uint64_t da = umma_smem_desc(a0 + koff, 16, 1024, UMMA_LAYOUT_SW128);
      const uint64_t db = umma_smem_desc(b0 + koff, 16, 1024, UMMA_LAYOUT_SW128);
      const uint32_t d = tmem_base + ((ALT && (i & 1)) ? (N <= 128 ? 128u : 256u) : 0u);
      if (FORM == 0) umma_bf16_ss(d, da, db, idesc, i > 1 ? 1u : 0u);
      else umma_bf16_ts(d, tmem_base + 448 + (i & 3) * 8, db, idesc, i > 1 ? 1u : 0u);
    }
    umma_commit(&bar[0]);
    const long long t1 = clock64();
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2 + 0].clocks = t1 - t0;   // issue loop only
    out[blockIdx.x * 2 + 1].clocks = t2 - t0;   // until the last instruction retired
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int FORM, bool ALT>
void run(const char* label, int N, int grid, Result* d_out, std::vector<Result>& h) {
  const size_t smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(probe<FORM, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double per[2][2];
  const int counts[2] = {64, 512};
  for (int c = 0; c < 2; ++c) {
    probe<FORM, ALT><<<grid, 128, smem>>>(N, counts[c], d_out);  // warm-up
    probe<FORM, ALT><<<grid, 128, smem>>>(N, counts[c], d_out);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, sizeof(Result) * 2 * grid, cudaMemcpyDeviceToHost);
    double issue = 0, done = 0;
    for (int b = 0; b < grid; ++b) {
      issue += h[2 * b].clocks;
      done += h[2 * b + 1].clocks;
    }
    per[c][0] = issue / grid;
    per[c][1] = done / grid;
  }
  const double slope_issue = (per[1][0] - per[0][0]) / (counts[1] - counts[0]);
  const double slope_done = (per[1][1] - per[0][1]) / (counts[1] - counts[0]);
  const double floor_clk = 128.0 * N / 256.0;
  printf("{\"probe\": \"umma\", \"form\": \"%s\", \"M\": 128, \"N\": %d, \"ctas\": %d, \"clk_per_mma_issue\": %.1f, "
         "\"clk_per_mma_retire\": %.1f, \"dispatch_floor_clk\": %.0f, \"fixed_latency_clk\": %.0f}\n",
         label, N, grid, slope_issue, slope_done, floor_clk, per[0][1] - slope_done * counts[0]);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  Result* d_out;
  cudaMalloc(&d_out, sizeof(Result) * 2 * sms);
  std::vector<Result> h(2 * sms);
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  for (int grid : {1, sms}) {
    for (int N : Ns) {
      run<0, false>("SS dependent chain", N, grid, d_out, h);
      run<0, true>("SS two accumulators", N, grid, d_out, h);
      run<1, false>("TS dependent chain", N, grid, d_out, h);
      if (N <= 192) run<1, true>("TS two accumulators", N, grid, d_out, h);  // (N = 256 x 2 would overlap the A columns)
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
