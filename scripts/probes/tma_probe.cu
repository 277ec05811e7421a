// TMA streaming probe (measurement tool, not part of the product): how fast can one producer thread per SM stream
// [104 x 64] bf16 boxes (the attention core's operand tiles) into shared memory, as a function of layout and of the
// number of boxes in flight?   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lamp_b200/csrc
//                                     -o scripts/probes/tma_probe scripts/probes/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "sm100_primitives.cuh"

using namespace lamp;

constexpr int BOX_ROWS = 104;
constexpr int BOX_BYTES = BOX_ROWS * 128;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mode 0: 3D map {cols, L, B} (in-model layout; item -> (b, head, part, kb)); mode 1: 2D map over contiguous tiles;
// mode 2: 1D bulk copies of contiguous tiles.
template <int MODE>
__global__ void __launch_bounds__(64, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm, const __nv_bfloat16* base, int n_boxes_total, int NS, int heads,
             int parts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16 * 14336);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // boxes are dealt round-robin by ITEM (= 2*parts boxes of one (b, head)) so that neighbouring SMs read
  // neighbouring heads of the same sample, as in the attention kernel
  const int per_item = 2 * parts;
  const int n_items = n_boxes_total / per_item;
  if (warp == 0 && lane == 0) {
    uint32_t u = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int h = item % heads, b = item / heads;
      for (int k = 0; k < per_item; ++k, ++u) {
        const int s = u % NS;
        mbar_wait(&empty[s], ((u / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], BOX_BYTES);
        uint8_t* dst = smem + s * 14336;
        if (MODE == 0) {
          const int part = k >> 1, kb = k & 1;
          tma_load_3d(dst, &tm, &full[s], part * heads * 128 + h * 128 + kb * 64, 0, b);
        } else if (MODE == 1) {
          tma_load_2d(dst, &tm, &full[s], 0, (item * per_item + k) * BOX_ROWS);
        } else {
          bulk_load_1d(dst, base + static_cast<size_t>(item * per_item + k) * (BOX_BYTES / 2), BOX_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t u = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      for (int k = 0; k < per_item; ++k, ++u) {
        const int s = u % NS;
        mbar_wait(&full[s], (u / NS) & 1);
        mbar_arrive(&empty[s]);
      }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 1100, L = 103, H = 4, PARTS = 3;
  const size_t cols = PARTS * H * 128, rows = (size_t)B * L;
  __nv_bfloat16* buf;
  cudaMalloc(&buf, rows * cols * 2 + (1 << 20));
  cudaMemset(buf, 0, rows * cols * 2 + (1 << 20));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  CUtensorMap tm3, tm2;
  {
    cuuint64_t dims[3] = {cols, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {cols * 2, (cuuint64_t)L * cols * 2};
    cuuint32_t box[3] = {64, BOX_ROWS, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("enc3 %d\n", r); return 1; }
  }
  const int n_boxes = B * H * 2 * PARTS;  // one plane: 26400 boxes of 13 KB = 351 MB
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)n_boxes * BOX_ROWS};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, BOX_ROWS}, es[2] = {1, 1};
    CUresult r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("enc2 %d\n", r); return 1; }
  }
  const int smem = 16 * 14336 + 1024 + 512;
  cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[3] = {"3D strided (pitch 3072 B, in-model QKV layout)", "2D contiguous tiles", "1D bulk contiguous"};
  for (int mode = 0; mode < 3; ++mode)
    for (int NS : {1, 2, 4, 8, 12, 16}) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        for (int it = 0; it < 4; ++it) {
          if (mode == 0) probe_kernel<0><<<148, 64, smem>>>(tm3, buf, n_boxes, NS, H, PARTS);
          if (mode == 1) probe_kernel<1><<<148, 64, smem>>>(tm2, buf, n_boxes, NS, H, PARTS);
          if (mode == 2) probe_kernel<2><<<148, 64, smem>>>(tm2, buf, n_boxes, NS, H, PARTS);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms / 4 < best) best = ms / 4;
      }
      cudaError_t err = cudaGetLastError();
      const double bytes = (double)n_boxes * (mode == 0 ? 103 * 128 : BOX_BYTES);
      printf("%-48s boxes in flight %2d: %8.1f us  %7.1f GB/s  (%.1f KB in flight / SM)%s\n", names[mode], NS,
             best * 1e3, bytes / (best * 1e-3) / 1e9, NS * BOX_BYTES / 1024.0, err ? cudaGetErrorString(err) : "");
    }
  return 0;
}
