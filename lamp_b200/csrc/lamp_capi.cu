// C ABI of lamp_b200 (see include/lamp_b200.h).  Host-side validation, TMA descriptor construction and launches.
#include "../../include/lamp_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "attn_bwd.cuh"
#include "attn_core.cuh"
#include "bgemm_tc.cuh"
#include "elementwise.cuh"
#include "gemm_planes.cuh"
#include "gemm_tn_tc.cuh"
#include "train_bwd.cuh"

using namespace lamp;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) return fail(LAMP_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_));    \
  } while (0)

#define REQUIRE(cond, ...)                              \
  do {                                                  \
    if (!(cond)) return fail(LAMP_EINVAL, __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- driver entry point for TMA descriptor encoding (no link-time libcuda dependency)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn f = cached.load(std::memory_order_acquire);
  if (f) return f;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  f = reinterpret_cast<EncodeTiledFn>(sym);
  cached.store(f, std::memory_order_release);
  return f;
}

// bf16 matrix viewed as {cols, rows[, batch]}; box = {64 cols (128 B), box_rows[, 1]}, 128B swizzle, zero OOB fill.
int make_tmap(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t ld_elems,
              uint32_t box_rows, bool three_d, uint32_t box_cols = 64) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LAMP_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (!aligned16(base)) return fail(LAMP_EINVAL, "TMA base pointer not 16-byte aligned");
  if ((ld_elems * 2) % 16 != 0) return fail(LAMP_EINVAL, "leading dimension %llu not a multiple of 8 elements",
                                            (unsigned long long)ld_elems);
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {ld_elems * 2, rows * ld_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  const CUtensorMapSwizzle swz = (box_cols == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, three_d ? 3 : 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LAMP_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return LAMP_OK;
}

// Per-device host state.  One process may drive several GPUs (nn.DataParallel threads as in the reference's
// main.py:106-108, a model moved to cuda:1, tests switching devices): the >48 KB dynamic-shared-memory opt-in
// (cudaFuncSetAttribute) is a per-device attribute and SM counts / architectures are per device, so everything that is
// cached is cached per device ordinal of the calling thread's current device.
constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

// bf16 tensor viewed as {cols, rows, heads, samples} with independent element strides for rows / heads / samples;
// box = {64 cols, box_rows, 1, 1}, 128B swizzle, zero OOB fill.  Used by the batched attention-backward products, whose
// operands are either head-major [H*B, L, d] tensors or head column slices of [B*L, H*d] plane matrices.
int make_tmap4(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t heads, uint64_t samples,
               uint64_t ld_elems, uint64_t head_stride, uint64_t sample_stride, uint32_t box_rows, bool f32 = false) {
  // f32 = true: the same view of an fp32 tensor, box = {32 cols (128 B), box_rows, 1, 1} (the fp32 outputs of the
  // batched products are stored through such maps)
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LAMP_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (!aligned16(base)) return fail(LAMP_EINVAL, "TMA base pointer not 16-byte aligned");
  const uint64_t eb = f32 ? 4 : 2;
  if ((ld_elems * eb) % 16 != 0 || (head_stride * eb) % 16 != 0 || (sample_stride * eb) % 16 != 0)
    return fail(LAMP_EINVAL, "TMA strides (%llu, %llu, %llu elements) must be multiples of 16 bytes",
                (unsigned long long)ld_elems, (unsigned long long)head_stride, (unsigned long long)sample_stride);
  cuuint64_t dims[4] = {cols, rows, heads, samples};
  cuuint64_t strides[3] = {ld_elems * eb, head_stride * eb, sample_stride * eb};
  cuuint32_t box[4] = {f32 ? 32u : 64u, box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LAMP_ECUDA, "cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
  return LAMP_OK;
}

int sm_count_cached() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return 0;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  n[dev].store(v, std::memory_order_relaxed);
  return v;
}

int arch_check() {
  static std::atomic<int> ok[kMaxDevices];  // 0 unknown, 1 ok, -1 bad
  const int dev = current_device();
  if (dev < 0) return fail(LAMP_ECUDA, "no CUDA device");
  int v = ok[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return fail(LAMP_ECUDA, "no CUDA device");
    v = (major == 10) ? 1 : -1;
    ok[dev].store(v, std::memory_order_relaxed);
  }
  if (v < 0) return fail(LAMP_EARCH, "lamp_b200 kernels are built for sm_100a only");
  return LAMP_OK;
}

// "once per device" for the function attributes: `init` may run twice under a race between two threads on the same
// device, which is harmless (cudaFuncSetAttribute is idempotent).
struct PerDeviceOnce {
  std::atomic<int> done[kMaxDevices];
};

template <typename F>
int per_device_once(PerDeviceOnce& o, F&& init) {
  const int dev = current_device();
  if (dev < 0) return fail(LAMP_ECUDA, "no CUDA device");
  if (o.done[dev].load(std::memory_order_acquire)) return LAMP_OK;
  const int rc = init();
  if (rc == LAMP_OK) o.done[dev].store(1, std::memory_order_release);
  return rc;
}

template <typename K>
int set_smem(K kernel, uint32_t bytes) {
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return LAMP_OK;
}

// dropout rate -> 32-bit threshold of the counter hash (an element is dropped iff hash < threshold); one definition for
// the attention kernels, the probability kernel, the element-wise dropouts and the recompute backward
inline uint32_t drop_threshold(float p_drop) {
  if (!(p_drop > 0.0f)) return 0u;
  const uint32_t t = (uint32_t)fmin((double)p_drop * 4294967296.0, 4294967295.0);
  return t == 0u ? 1u : t;
}

int launch_check() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "kernel launch: %s", cudaGetErrorString(e));
  return LAMP_OK;
}

// tuning knob: programmatic dependent launch for the inference-path kernels (their prologues overlap the predecessor's
// tail).  0 = off, 1 = on, 2 (default) = on for small launches only: measured +6.5 % at B = 32 (0.597 -> 0.561 ms per
// forward), -1 % (noise level) at B = 1100 where a kernel's prologue is < 1 % of its run time.
std::atomic<int> g_pdl{2};
constexpr long long kPdlMaxRows = 16384;
inline bool use_pdl(long long rows) {
  const int m = g_pdl.load();
  return m == 1 || (m == 2 && rows <= kPdlMaxRows);
}

template <int BLOCK_N, int NTERMS, int BLOCK_K, int EPI, int CTA_GROUP>
int launch_gemm_epi(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                    const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BLOCK_N, NTERMS, BLOCK_K, CTA_GROUP>;
  auto kernel = gemm_planes_kernel<BLOCK_N, NTERMS, BLOCK_K, EPI, CTA_GROUP>;
  static PerDeviceOnce once;
  if (int once_rc = per_device_once(once, [kernel] { int rc_ = set_smem(kernel, Cfg::SMEM_BYTES); return rc_; })) return once_rc;
  const int m_tiles = (p.M + GEMM_BLOCK_M * CTA_GROUP - 1) / (GEMM_BLOCK_M * CTA_GROUP);
  const int tiles = (EPI == EPI_LN) ? m_tiles : m_tiles * ((p.N + BLOCK_N - 1) / BLOCK_N);
  const int max_groups = sm_count_cached() / CTA_GROUP;
  const int groups = tiles < max_groups ? tiles : max_groups;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * CTA_GROUP);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA_GROUP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl(p.M) ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a_hi, a_lo, w_hi, w_lo, p);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "gemm launch: %s", cudaGetErrorString(e));
  return launch_check();
}

template <int BLOCK_N, int NTERMS, int BLOCK_K, int CTA_GROUP>
int launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi, const CUtensorMap& w_lo,
                const GemmParams& p, cudaStream_t st) {
  if (p.drop_thresh != 0) {
    if constexpr (NTERMS == 3)   // training runs on 3-term products only
      return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_F32_DROP, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
    else
      return fail(LAMP_EINVAL, "gemm: the dropout epilogue exists for LAMP_PREC_FP32 only");
  }
  if (p.a_stats != nullptr)
    return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_DLN_A, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
  if (p.stats_out != nullptr)
    return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_RSTATS, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
  if (p.ln_gamma != nullptr) {
    if constexpr (BLOCK_N == 256)
      return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_LN, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
    else
      return fail(LAMP_EINVAL, "gemm: fused LayerNorm needs the 256-wide tile");
  }
  if (p.out_hi != nullptr && p.out_f32 == nullptr && p.residual == nullptr && p.res_hi == nullptr)
    return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_PLANES, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
  if (p.out_hi == nullptr && !p.relu)
    return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_F32, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
  return launch_gemm_epi<BLOCK_N, NTERMS, BLOCK_K, EPI_ANY, CTA_GROUP>(a_hi, a_lo, w_hi, w_lo, p, st);
}

std::atomic<int> g_gemm_block_k{0};   // tuning knob (lamp_set_tuning): 32 -> 64B swizzle / deep ring, 64 -> 128B swizzle,
                                      // 0 -> automatic (64 for CTA pairs: 3 x 64 KB stages; 32 otherwise: 4 x 48 KB)
std::atomic<int> g_attn_compact{1};   // tuning knob: 1 -> L-dependent tile rows + deepest K/V staging that fits, 0 -> full 128-row tiles, 1 stage
std::atomic<int> g_attn_stage{1};     // tuning knob: 1 -> O planes leave through the smem staging tile + TMA stores when it fits
std::atomic<int> g_attn_pv_split{0};  // tuning knob: 1 -> PV product as two interleaved N = 64 chains when d == 128
std::atomic<int> g_gemm_tn_tc{1};     // tuning knob: 1 -> weight gradient on tcgen05 (gemm_tn_tc.cuh), 0 -> warp-MMA version
std::atomic<int> g_attn_bwd_tc{1};    // tuning knob: 1 -> attention backward as batched tcgen05 products, 0 -> warp-MMA kernels
std::atomic<int> g_attn_kv128_min_lk{512};  // tuning knob: > 0 -> 128-key tiles for d > 64 when Lk >= this (dense keys); 0 -> always 64-key tiles there
std::atomic<int> g_gemm_pair{1};      // tuning knob: 1 -> CTA pairs (cta_group::2) for the 256-wide tiles, 0 -> single CTAs

constexpr uint32_t kMaxDynSmem = 232448;  // 227 KB: the sm_100 per-CTA opt-in maximum

// <<<>>> replacement that can request programmatic dependent launch (only for kernels that call griddep_wait())
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, long long rows,
                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl(rows) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <int BLOCK_KV, int NTERMS, bool DROP>
int launch_attn_k(const CUtensorMap (&tm)[8], const AttnParams& p, uint32_t smem_bytes, cudaStream_t st);

template <int BLOCK_KV, int NTERMS>
int launch_attn(const CUtensorMap (&tm)[8], const AttnParams& p, uint32_t smem_bytes, cudaStream_t st) {
  return p.drop_thresh ? launch_attn_k<BLOCK_KV, NTERMS, true>(tm, p, smem_bytes, st)
                       : launch_attn_k<BLOCK_KV, NTERMS, false>(tm, p, smem_bytes, st);
}

template <int BLOCK_KV, int NTERMS, bool DROP>
int launch_attn_k(const CUtensorMap (&tm)[8], const AttnParams& p, uint32_t smem_bytes, cudaStream_t st) {
  auto kernel = attn_core_kernel<BLOCK_KV, NTERMS, DROP>;
  static PerDeviceOnce once;
  if (int once_rc = per_device_once(once, [kernel] { int rc_ = set_smem(kernel, kMaxDynSmem); return rc_; })) return once_rc;
  const int items = p.B * p.H * ((p.Lq + ATTN_BLOCK_M - 1) / ATTN_BLOCK_M);
  const int grid = items < sm_count_cached() ? items : sm_count_cached();
  cudaError_t e = launch_k(kernel, (unsigned)grid, (unsigned)attn_threads(BLOCK_KV), smem_bytes, st,
                           (long long)p.B * p.Lq, tm[0], tm[1], tm[2], tm[3],
                           tm[4], tm[5], tm[6], tm[7], p);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "attn launch: %s", cudaGetErrorString(e));
  return launch_check();
}

// Shared-memory plan: Q tile (+ an equally shaped O staging tile when the planes leave through TMA stores) + as many
// K/V ring slots as fit.  Staging is used when it still leaves >= 4 slots, or 2 for single-tile problems (one K and
// one V slot in flight is all such an item has).
template <int BLOCK_KV, int NTERMS>
int launch_attn_plan(const CUtensorMap (&tm)[8], AttnParams& p, bool can_stage, cudaStream_t st) {
  const int npl = NTERMS == 3 ? 2 : 1, kb64 = (p.d + 63) / 64;
  const uint32_t q_bytes = attn_tile_bytes(npl, kb64, p.qrows);
  p.slot_bytes = attn_tile_bytes(npl, kb64, p.krows > p.vrows ? p.krows : p.vrows);
  auto slots_for = [&](int staged) {
    const uint32_t fixed = attn_smem_bytes(q_bytes, p.slot_bytes, 0, staged);
    if (fixed >= kMaxDynSmem) return 0;
    const int n = (int)((kMaxDynSmem - fixed) / p.slot_bytes);
    return n > ATTN_MAX_SLOTS ? ATTN_MAX_SLOTS : n;
  };
  const bool single = p.Lk <= BLOCK_KV;
  const int ns = can_stage && g_attn_stage.load() ? slots_for(1) : 0;
  // multi-tile rows need ring depth more than they need the staged store (K and V rings of >= 2 slots each: the
  // load-to-use latency is ~1.5 unit periods), single-tile problems have one K and one V tile in flight anyway
  p.staged = ((!single && ns >= 4) || (single && ns >= 2)) ? 1 : 0;
  p.kv_slots = p.staged ? ns : slots_for(0);
  if (p.kv_slots < 2) return fail(LAMP_EINVAL, "attn: tile does not fit shared memory");
  return launch_attn<BLOCK_KV, NTERMS>(tm, p, attn_smem_bytes(q_bytes, p.slot_bytes, p.kv_slots, p.staged), st);
}

inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  }
};

}  // namespace

extern "C" {

int lamp_version(void) { return 100; }
const char* lamp_last_error(void) { return g_err; }
int lamp_device_check(void) { return arch_check(); }
int lamp_sm_count(void) { return sm_count_cached(); }

int lamp_set_tuning(int key, int value) {
  if (key == LAMP_TUNE_GEMM_BLOCK_K && (value == 0 || value == 32 || value == 64)) {
    g_gemm_block_k.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_ATTN_COMPACT && (value == 0 || value == 1)) {
    g_attn_compact.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_ATTN_STAGE && (value == 0 || value == 1)) {
    g_attn_stage.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_ATTN_PV_SPLIT && (value == 0 || value == 1)) {
    g_attn_pv_split.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_GEMM_TN_TC && (value == 0 || value == 1)) {
    g_gemm_tn_tc.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_ATTN_BWD_TC && (value == 0 || value == 1)) {
    g_attn_bwd_tc.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_PDL && (value == 0 || value == 1 || value == 2)) {
    g_pdl.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_ATTN_KV128_MIN_LK && value >= 0) {
    g_attn_kv128_min_lk.store(value);
    return LAMP_OK;
  }
  if (key == LAMP_TUNE_GEMM_CTA_PAIR && (value == 0 || value == 1)) {
    g_gemm_pair.store(value);
    return LAMP_OK;
  }
  return fail(LAMP_EINVAL, "set_tuning: unknown key/value %d/%d", key, value);
}

int lamp_split_planes(const float* x, int64_t rows, int cols, int64_t ld, void* hi, void* lo, int64_t ldp,
                      void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(x && hi, "split_planes: null pointer");
  REQUIRE(cols % 4 == 0 && ld % 4 == 0 && ldp % 4 == 0, "split_planes: cols/ld/ldp must be multiples of 4");
  REQUIRE(aligned16(x) && (reinterpret_cast<uintptr_t>(hi) % 8 == 0), "split_planes: alignment");
  if (rows == 0 || cols == 0) return LAMP_OK;
  const long long total = rows * (cols / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = 32LL * sm_count_cached();
  if (blocks > cap) blocks = cap;
  split_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      x, rows, cols, ld, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), ldp);
  return launch_check();
}

int lamp_split_planes_multi(const LampSplitJob* jobs, int n_jobs, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(n_jobs >= 0 && (jobs || n_jobs == 0), "split_planes_multi: null job list");
  cudaStream_t st = (cudaStream_t)stream;
  for (int j0 = 0; j0 < n_jobs; j0 += SPLIT_MULTI_MAX_JOBS) {
    SplitJobs pack;
    const int n = n_jobs - j0 < SPLIT_MULTI_MAX_JOBS ? n_jobs - j0 : SPLIT_MULTI_MAX_JOBS;
    long long max_tiles = 0;
    for (int j = 0; j < n; ++j) {
      const LampSplitJob& in = jobs[j0 + j];
      REQUIRE(in.src && in.hi && in.rows > 0 && in.cols > 0, "split_planes_multi: job %d: null pointer or empty shape", j0 + j);
      REQUIRE(in.ld >= in.cols && in.ldp >= (in.transpose ? in.rows : in.cols), "split_planes_multi: job %d: leading dimension too small", j0 + j);
      SplitJob& o = pack.job[j];
      o.src = in.src; o.hi = static_cast<__nv_bfloat16*>(in.hi); o.lo = static_cast<__nv_bfloat16*>(in.lo);
      o.rows = in.rows; o.cols = in.cols; o.ld = in.ld; o.ldp = in.ldp; o.transpose = in.transpose;
      const long long tiles = (long long)((in.rows + 31) / 32) * ((in.cols + 31) / 32);
      if (tiles > max_tiles) max_tiles = tiles;
    }
    for (int j = n; j < SPLIT_MULTI_MAX_JOBS; ++j) pack.job[j] = SplitJob{nullptr, nullptr, nullptr, 0, 0, 0, 0, 0};
    REQUIRE(max_tiles < (1LL << 31), "split_planes_multi: matrix too large");
    split_planes_multi_kernel<<<dim3((unsigned)max_tiles, (unsigned)n), 256, 0, st>>>(pack);
    if (int rc = launch_check()) return rc;
  }
  return LAMP_OK;
}

namespace {
struct DlnArgs {  // deferred-LayerNorm extras of a GEMM launch (all null / 0: none)
  const float* a_stats = nullptr; int a_nparts = 0; float a_eps = 0.f; const float* a_colsum = nullptr;
  const float* r_stats = nullptr; int r_nparts = 0; float r_eps = 0.f; const float* r_gamma = nullptr;
  const float* r_beta = nullptr; float* stats_out = nullptr;
  uint32_t drop_thresh = 0; float drop_scale = 1.0f; unsigned long long drop_seed = 0;   // EPI_F32_DROP
  const unsigned long long* drop_seed_dev = nullptr;
};
}  // namespace

static int gemm_impl(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                     int64_t ldw, int M, int N, int K, int precision, const float* bias, int relu,
                     const float* residual, const void* res_hi, const void* res_lo, int64_t ldr, int resid_mod,
                     float* out_f32, int64_t ldo, void* out_hi,
                     void* out_lo, int64_t ldp, const float* ln_gamma, const float* ln_beta, float ln_eps,
                     const int32_t* m_dev, void* stream, const DlnArgs& dln = DlnArgs()) {
  if (int rc = arch_check()) return rc;
  if (ln_gamma != nullptr) {
    REQUIRE(ln_beta != nullptr, "gemm_ln: beta missing");
    REQUIRE(N > 256 && N <= 512, "gemm_ln: the fused LayerNorm epilogue covers 256 < N <= 512 (got %d)", N);
    REQUIRE(!relu, "gemm_ln: ReLU is not part of this epilogue");
    REQUIRE(aligned16(ln_gamma) && aligned16(ln_beta), "gemm_ln: gamma/beta alignment");
  }
  REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  REQUIRE(K % 8 == 0 && N % 8 == 0, "gemm: K (%d) and N (%d) must be multiples of 8", K, N);
  REQUIRE(a_hi && w_hi, "gemm: null operand");
  const bool three = (precision == LAMP_PREC_FP32);
  REQUIRE(precision == LAMP_PREC_FP32 || precision == LAMP_PREC_BF16, "gemm: unknown precision %d", precision);
  REQUIRE(!three || (a_lo && w_lo), "gemm: LAMP_PREC_FP32 needs lo planes");
  REQUIRE(out_f32 || out_hi, "gemm: no output");
  REQUIRE(!out_f32 || (aligned16(out_f32) && ldo % 4 == 0), "gemm: out_f32 alignment");
  REQUIRE(!out_hi || (aligned16(out_hi) && ldp % 8 == 0 && (!out_lo || aligned16(out_lo))), "gemm: plane alignment");
  REQUIRE(!residual || (aligned16(residual) && ldr % 4 == 0), "gemm: residual alignment");
  REQUIRE(!(residual && res_hi), "gemm: give the residual either as fp32 or as planes, not both");
  REQUIRE(!res_hi || (aligned16(res_hi) && ldr % 8 == 0 && (!res_lo || aligned16(res_lo))), "gemm: residual plane alignment");
  REQUIRE(!bias || aligned16(bias), "gemm: bias alignment");
  REQUIRE(!(residual || res_hi) || ((int64_t)(resid_mod ? resid_mod : M) + 1) * ldr < (int64_t)1 << 31,
          "gemm: residual matrix too large for 32-bit row offsets");
  if (M == 0) return LAMP_OK;
  const bool wide = (N > 128);
  const bool pair = wide && g_gemm_pair.load() != 0 && sm_count_cached() >= 2;
  const uint32_t box_n = wide ? (pair ? 128 : 256) : 128;  // a pair's CTAs each stage half of the 256-row W tile
  CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
  int bk = g_gemm_block_k.load();
  if (bk == 0) bk = pair ? 64 : 32;
  if (int rc = make_tmap(&ta_hi, a_hi, K, M, 1, lda, GEMM_BLOCK_M, false, bk)) return rc;
  if (int rc = make_tmap(&tw_hi, w_hi, K, N, 1, ldw, box_n, false, bk)) return rc;
  if (three) {
    if (int rc = make_tmap(&ta_lo, a_lo, K, M, 1, lda, GEMM_BLOCK_M, false, bk)) return rc;
    if (int rc = make_tmap(&tw_lo, w_lo, K, N, 1, ldw, box_n, false, bk)) return rc;
  } else {
    ta_lo = ta_hi;
    tw_lo = tw_hi;
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.residual = residual; p.ldr = (int)ldr;
  p.res_hi = static_cast<const __nv_bfloat16*>(res_hi);
  p.res_lo = static_cast<const __nv_bfloat16*>(res_lo); p.resid_mod = resid_mod; p.relu = relu;
  p.out_f32 = out_f32; p.ldo = (int)ldo;
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.ldp = (int)ldp;
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_eps = ln_eps;
  p.m_dev = m_dev;
  p.a_stats = reinterpret_cast<const float2*>(dln.a_stats); p.a_nparts = dln.a_nparts; p.a_eps = dln.a_eps;
  p.a_colsum = dln.a_colsum;
  p.r_stats = reinterpret_cast<const float2*>(dln.r_stats); p.r_nparts = dln.r_nparts; p.r_eps = dln.r_eps;
  p.r_gamma = dln.r_gamma; p.r_beta = dln.r_beta;
  p.stats_out = reinterpret_cast<float2*>(dln.stats_out);
  p.drop_thresh = dln.drop_thresh; p.drop_scale = dln.drop_scale; p.drop_seed = dln.drop_seed;
  p.drop_seed_dev = dln.drop_seed_dev;
  cudaStream_t st = (cudaStream_t)stream;
#define LAMP_GEMM_DISPATCH(BK)                                                                                      \
  do {                                                                                                             \
    if (pair) {                                                                                                    \
      if (three) return launch_gemm<256, 3, BK, 2>(ta_hi, ta_lo, tw_hi, tw_lo, p, st);                             \
      return launch_gemm<256, 1, BK, 2>(ta_hi, ta_lo, tw_hi, tw_lo, p, st);                                        \
    }                                                                                                              \
    if (three) return wide ? launch_gemm<256, 3, BK, 1>(ta_hi, ta_lo, tw_hi, tw_lo, p, st)                         \
                           : launch_gemm<128, 3, BK, 1>(ta_hi, ta_lo, tw_hi, tw_lo, p, st);                        \
    return wide ? launch_gemm<256, 1, BK, 1>(ta_hi, ta_lo, tw_hi, tw_lo, p, st)                                    \
                : launch_gemm<128, 1, BK, 1>(ta_hi, ta_lo, tw_hi, tw_lo, p, st);                                   \
  } while (0)
  if (bk == 64) LAMP_GEMM_DISPATCH(64);
  LAMP_GEMM_DISPATCH(32);
#undef LAMP_GEMM_DISPATCH
}

int lamp_gemm_planes(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                     int64_t ldw, int M, int N, int K, int precision, const float* bias, int relu,
                     const float* residual, int64_t ldr, int resid_mod, float* out_f32, int64_t ldo, void* out_hi,
                     void* out_lo, int64_t ldp, const int32_t* m_dev, void* stream) {
  return gemm_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, M, N, K, precision, bias, relu, residual, nullptr, nullptr, ldr,
                   resid_mod, out_f32, ldo, out_hi, out_lo, ldp, nullptr, nullptr, 0.0f, m_dev, stream);
}

int lamp_gemm_planes_drop(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                          int64_t ldw, int M, int N, int K, const float* bias, float p_drop, uint64_t seed,
                          const uint64_t* seed_dev, const float* residual, int64_t ldr, int resid_mod, float* out_f32,
                          int64_t ldo, void* stream) {
  REQUIRE(out_f32 != nullptr, "gemm_planes_drop: fp32 output required");
  REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "gemm_planes_drop: rate outside [0, 1)");
  DlnArgs d;
  d.drop_thresh = drop_threshold(p_drop);
  d.drop_scale = 1.0f / (1.0f - p_drop);
  d.drop_seed = seed;
  d.drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  return gemm_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, M, N, K, LAMP_PREC_FP32, bias, 0, residual, nullptr, nullptr, ldr,
                   resid_mod, out_f32, ldo, nullptr, nullptr, 0, nullptr, nullptr, 0.0f, nullptr, stream, d);
}

int lamp_gemm_planes_pres(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                          int64_t ldw, int M, int N, int K, int precision, const float* bias, const void* res_hi,
                          const void* res_lo, int64_t ldr, int resid_mod, float* out_f32, int64_t ldo, void* out_hi,
                          void* out_lo, int64_t ldp, const int32_t* m_dev, void* stream) {
  REQUIRE(res_hi != nullptr, "gemm_planes_pres: residual planes missing");
  return gemm_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, M, N, K, precision, bias, 0, nullptr, res_hi, res_lo, ldr,
                   resid_mod, out_f32, ldo, out_hi, out_lo, ldp, nullptr, nullptr, 0.0f, m_dev, stream);
}

int lamp_gemm_ln_planes(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                        int64_t ldw, int M, int N, int K, int precision, const float* bias, const float* residual,
                        int64_t ldr, int resid_mod, const float* gamma, const float* beta, float eps, float* out_f32,
                        int64_t ldo, void* out_hi, void* out_lo, int64_t ldp, void* stream) {
  REQUIRE(gamma && beta, "gemm_ln: null gamma/beta");
  return gemm_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, M, N, K, precision, bias, 0, residual, nullptr, nullptr, ldr,
                   resid_mod, out_f32, ldo, out_hi, out_lo, ldp, gamma, beta, eps, nullptr, stream);
}

int lamp_gemm_stats_parts(int N) {
  const int block_n = N > 128 ? 256 : 128;  // tile width chosen by gemm_impl
  return 2 * ((N + block_n - 1) / block_n);
}

int lamp_gemm_planes_dln(const void* y_hi, const void* y_lo, int64_t lda, const float* a_stats, int a_nparts,
                         float a_eps, const void* wg_hi, const void* wg_lo, int64_t ldw, const float* colsum,
                         const float* biasf, int M, int N, int K, int precision, int relu, void* out_hi, void* out_lo,
                         int64_t ldp, const int32_t* m_dev, void* stream) {
  REQUIRE(a_stats && colsum && biasf && out_hi, "gemm_dln: null pointer");
  REQUIRE(a_nparts > 0 && a_eps >= 0.0f, "gemm_dln: bad row-stat layout");
  REQUIRE(aligned16(colsum) && (reinterpret_cast<uintptr_t>(a_stats) % 8 == 0), "gemm_dln: alignment");
  DlnArgs d;
  d.a_stats = a_stats; d.a_nparts = a_nparts; d.a_eps = a_eps; d.a_colsum = colsum;
  return gemm_impl(y_hi, y_lo, lda, wg_hi, wg_lo, ldw, M, N, K, precision, biasf, relu, nullptr, nullptr, nullptr, 0, 0,
                   nullptr, 0, out_hi, out_lo, ldp, nullptr, nullptr, 0.0f, m_dev, stream, d);
}

int lamp_gemm_planes_rstats(const void* a_hi, const void* a_lo, int64_t lda, const void* w_hi, const void* w_lo,
                            int64_t ldw, int M, int N, int K, int precision, const float* bias, const float* residual,
                            const void* res_hi, const void* res_lo, int64_t ldr, int resid_mod, const float* r_stats,
                            int r_nparts, float r_eps, const float* r_gamma, const float* r_beta, void* out_hi,
                            void* out_lo, int64_t ldp, float* stats_out, const int32_t* m_dev, void* stream) {
  REQUIRE(out_hi && stats_out, "gemm_rstats: null output");
  REQUIRE(reinterpret_cast<uintptr_t>(stats_out) % 8 == 0, "gemm_rstats: stats alignment");
  DlnArgs d;
  if (r_stats != nullptr) {
    REQUIRE(res_hi && r_gamma && r_beta && r_nparts > 0 && resid_mod == 0, "gemm_rstats: deferred residual needs planes, gamma, beta");
    REQUIRE(aligned16(r_gamma) && aligned16(r_beta) && (reinterpret_cast<uintptr_t>(r_stats) % 8 == 0), "gemm_rstats: alignment");
    d.r_stats = r_stats; d.r_nparts = r_nparts; d.r_eps = r_eps; d.r_gamma = r_gamma; d.r_beta = r_beta;
  }
  d.stats_out = stats_out;
  return gemm_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, M, N, K, precision, bias, 0, residual, res_hi, res_lo, ldr,
                   resid_mod, nullptr, 0, out_hi, out_lo, ldp, nullptr, nullptr, 0.0f, m_dev, stream, d);
}

int lamp_ln_apply(const void* y_hi, const void* y_lo, const float* stats, int nparts, const float* gamma,
                  const float* beta, float eps, int64_t rows, int D, const int64_t* index, float* out, void* out_hi,
                  void* out_lo, const int32_t* m_dev, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(y_hi && stats && gamma && beta && (out || out_hi), "ln_apply: null pointer");
  REQUIRE(D % 4 == 0 && D > 0 && nparts > 0, "ln_apply: D=%d must be a positive multiple of 4", D);
  REQUIRE(aligned16(gamma) && aligned16(beta) && (!out || aligned16(out)), "ln_apply: alignment");
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  ln_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(y_hi), static_cast<const __nv_bfloat16*>(y_lo),
      reinterpret_cast<const float2*>(stats), nparts, gamma, beta, eps, rows, D,
      reinterpret_cast<const long long*>(index), out, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo), m_dev);
  return launch_check();
}

int lamp_diag_proj_ln(const void* y_hi, const void* y_lo, const float* stats, int nparts, const float* gamma,
                      const float* beta, float eps, const float* W, const float* bias, int64_t B, int L, int D,
                      float* logits, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(y_hi && stats && gamma && beta && W && logits, "diag_proj_ln: null pointer");
  REQUIRE(D % 4 == 0 && nparts > 0 && aligned16(W) && aligned16(gamma) && aligned16(beta), "diag_proj_ln: D multiple of 4 and 16-byte alignment required");
  const long long rows = B * L;
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  diag_proj_ln_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      static_cast<const __nv_bfloat16*>(y_hi), static_cast<const __nv_bfloat16*>(y_lo),
      reinterpret_cast<const float2*>(stats), nparts, gamma, beta, eps, W, bias, rows, L, D, logits);
  return launch_check();
}

int lamp_pack_mask_bits(const uint8_t* mask, int64_t msb, int64_t msq, int64_t msk, int64_t Bm, int Lq, int Lk,
                        uint32_t* words, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(mask && words && Bm >= 0 && Lq > 0 && Lk > 0, "pack_mask_bits: bad arguments");
  const long long nwords = Bm * Lq * ((Lk + 31) / 32);
  if (nwords == 0) return LAMP_OK;
  const long long blocks = (nwords * 32 + 255) / 256;
  pack_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mask, msb, msq, msk, Bm, Lq, Lk, words);
  return launch_check();
}

static int attn_impl(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                     const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                     int Lq, int Lk, int d, float temperature, int precision, const uint8_t* mask,
                     int64_t msb, int64_t msq, int64_t msk, const uint32_t* mask_bits, int64_t mbb, int64_t mbq,
                     void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                     int64_t ldof, float* row_max, float* row_sum, float* probs, const int32_t* kv_start,
                     const int32_t* kv_len, int64_t kv_rows, void* stream, float p_drop = 0.0f, uint64_t seed = 0,
                     float* probs_pre = nullptr, const uint64_t* seed_dev = nullptr);

int lamp_attn_core_planes(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                          const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                          int Lq, int Lk, int d, float temperature, int precision, const uint8_t* mask,
                          int64_t msb, int64_t msq, int64_t msk, void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                          int64_t ldof, float* row_max, float* row_sum, float* probs, const int32_t* kv_start,
                          const int32_t* kv_len, int64_t kv_rows, void* stream) {
  return attn_impl(q_hi, q_lo, ldq, q_col0, q_bcast, kv_hi, kv_lo, ldkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature,
                   precision, mask, msb, msq, msk, nullptr, 0, 0, o_hi, o_lo, ldo, o_f32, ldof, row_max, row_sum, probs,
                   kv_start, kv_len, kv_rows, stream);
}

int lamp_attn_core_planes_mbits(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                                const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                                int Lq, int Lk, int d, float temperature, int precision, const uint32_t* mask_bits,
                                int64_t mbb, int64_t mbq, void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                                int64_t ldof, void* stream) {
  REQUIRE(mask_bits != nullptr && mbq >= (Lk + 31) / 32, "attn_mbits: packed mask missing or row stride too small");
  return attn_impl(q_hi, q_lo, ldq, q_col0, q_bcast, kv_hi, kv_lo, ldkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature,
                   precision, nullptr, 0, 0, 0, mask_bits, mbb, mbq, o_hi, o_lo, ldo, o_f32, ldof, nullptr, nullptr,
                   nullptr, nullptr, nullptr, 0, stream);
}

static int attn_impl(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                     const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H,
                     int Lq, int Lk, int d, float temperature, int precision, const uint8_t* mask,
                     int64_t msb, int64_t msq, int64_t msk, const uint32_t* mask_bits, int64_t mbb, int64_t mbq,
                     void* o_hi, void* o_lo, int64_t ldo, float* o_f32,
                     int64_t ldof, float* row_max, float* row_sum, float* probs, const int32_t* kv_start,
                     const int32_t* kv_len, int64_t kv_rows, void* stream, float p_drop, uint64_t seed,
                     float* probs_pre, const uint64_t* seed_dev) {
  if (int rc = arch_check()) return rc;
  REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "attn: dropout rate %f outside [0, 1)", (double)p_drop);
  REQUIRE(p_drop == 0.0f || !kv_len, "attn: dropout is a training-path feature (dense keys)");
  REQUIRE(!mask_bits || !probs, "attn: the probability output needs the byte mask");
  REQUIRE(!probs_pre || probs, "attn: probs_pre goes with probs");
  REQUIRE((kv_start == nullptr) == (kv_len == nullptr), "attn: kv_start and kv_len go together");
  REQUIRE(!kv_len || (!probs && kv_rows > 0), "attn: packed keys exclude the probability output");
  REQUIRE(!kv_len || !mask || (msb == 0 && msq == 0), "attn: with packed keys the mask is one byte per packed key row");
  REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "attn: bad shape B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  REQUIRE(d % 16 == 0 && d >= 16 && d <= 128, "attn: head width %d must be a multiple of 16 in [16,128]", d);
  REQUIRE(precision == LAMP_PREC_FP32 || precision == LAMP_PREC_BF16, "attn: unknown precision %d", precision);
  const bool three = (precision == LAMP_PREC_FP32);
  REQUIRE(q_hi && kv_hi && (!three || (q_lo && kv_lo)), "attn: null operand planes");
  REQUIRE(o_hi || o_f32, "attn: no output");
  REQUIRE(!o_hi || (aligned16(o_hi) && ldo % 8 == 0 && (!o_lo || aligned16(o_lo))), "attn: output plane alignment");
  REQUIRE(!o_f32 || (aligned16(o_f32) && ldof % 4 == 0), "attn: o_f32 alignment");
  REQUIRE(!probs || (row_max && row_sum), "attn: probs need row_max/row_sum scratch");
  REQUIRE(temperature > 0.0f, "attn: temperature must be positive");
  REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0, "attn: column offsets must be multiples of 8");
  if (B == 0) return LAMP_OK;
  const bool multi = Lk > 128;
  // 64-key tiles for wide heads on multi-tile rows: K and V rings of >= 2 slots each fit next to Q.  Long rows (knob
  // LAMP_TUNE_ATTN_KV128_MIN_LK, default 512) use 128-key tiles with ONE K and ONE V slot: an N = 64 score MMA
  // costs 55 clocks for 32 clocks of work, an N = 128 one 64 for 64, and a 128-key unit is long enough (~3 K clocks of
  // MMAs) to cover the load of the next tile.
  const int kv128_min = g_attn_kv128_min_lk.load();
  const int block_kv = (d > 64 && multi && !(kv128_min > 0 && Lk >= kv128_min && !kv_len)) ? 64 : 128;
  auto round_up = [](int x, int m) { return (x + m - 1) / m * m; };
  // TMA box rows follow the label count: a single-tile problem (Lk <= 128) only stages the rows that exist
  const bool compact = g_attn_compact.load() != 0;
  const int qrows = compact ? round_up(Lq < ATTN_BLOCK_M ? Lq : ATTN_BLOCK_M, 8) : ATTN_BLOCK_M;
  const int krows = (multi || !compact) ? block_kv : round_up(Lk, 8);
  const int vrows = (multi || !compact) ? block_kv : round_up(Lk, 16);  // PV consumes keys in steps of 16; TMA zero-fills rows >= Lk
  CUtensorMap tm[8];
  if (int rc = make_tmap(&tm[0], q_hi, (uint64_t)ldq, Lq, q_bcast ? 1 : B, ldq, qrows, true)) return rc;
  const uint64_t kv_tm_rows = kv_len ? (uint64_t)kv_rows : (uint64_t)Lk;  // packed keys: one long row dimension
  const uint64_t kv_tm_batch = kv_len ? 1 : (uint64_t)B;
  if (int rc = make_tmap(&tm[2], kv_hi, (uint64_t)ldkv, kv_tm_rows, kv_tm_batch, ldkv, krows, true)) return rc;
  if (int rc = make_tmap(&tm[4], kv_hi, (uint64_t)ldkv, kv_tm_rows, kv_tm_batch, ldkv, vrows, true)) return rc;
  if (three) {
    if (int rc = make_tmap(&tm[1], q_lo, (uint64_t)ldq, Lq, q_bcast ? 1 : B, ldq, qrows, true)) return rc;
    if (int rc = make_tmap(&tm[3], kv_lo, (uint64_t)ldkv, kv_tm_rows, kv_tm_batch, ldkv, krows, true)) return rc;
    if (int rc = make_tmap(&tm[5], kv_lo, (uint64_t)ldkv, kv_tm_rows, kv_tm_batch, ldkv, vrows, true)) return rc;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
    tm[5] = tm[4];
  }
  // output planes as TMA-store targets: {ldo cols, Lq, B}; rows >= Lq and nothing else are clipped
  const bool can_stage = o_hi != nullptr && o_f32 == nullptr && d % 64 == 0 && (!three || o_lo != nullptr);
  if (can_stage) {
    if (int rc = make_tmap(&tm[6], o_hi, (uint64_t)ldo, Lq, B, ldo, qrows, true)) return rc;
    if (three) {
      if (int rc = make_tmap(&tm[7], o_lo, (uint64_t)ldo, Lq, B, ldo, qrows, true)) return rc;
    } else {
      tm[7] = tm[6];
    }
  } else {
    tm[6] = tm[0];
    tm[7] = tm[0];
  }
  AttnParams p;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.d = d;
  p.scale_log2 = 1.4426950408889634f / temperature;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.q_bcast = q_bcast;
  p.mask = mask; p.msb = msb; p.msq = msq; p.msk = msk;
  p.mask_bits = mask_bits; p.mbb = mbb; p.mbq = mbq;
  p.o_hi = static_cast<__nv_bfloat16*>(o_hi);
  p.o_lo = static_cast<__nv_bfloat16*>(o_lo);
  p.ldo = (int)ldo; p.o_f32 = o_f32; p.ldof = (int)ldof;
  p.row_max = row_max; p.row_sum = row_sum;
  p.qrows = qrows; p.krows = krows; p.vrows = vrows;
  p.kv_start = kv_start; p.kv_len = kv_len;
  p.pv_split = (d == 128 && g_attn_pv_split.load() != 0) ? 1 : 0;
  p.drop_thresh = drop_threshold(p_drop);
  p.drop_scale = 1.0f / (1.0f - p_drop);
  p.drop_seed = seed;
  p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (block_kv == 128)
    rc = three ? launch_attn_plan<128, 3>(tm, p, can_stage, st) : launch_attn_plan<128, 1>(tm, p, can_stage, st);
  else
    rc = three ? launch_attn_plan<64, 3>(tm, p, can_stage, st) : launch_attn_plan<64, 1>(tm, p, can_stage, st);
  if (rc != LAMP_OK) return rc;
  if (probs) {
    // probabilities from the operand planes and the saved row statistics, on the tensor cores
    ProbsMmaParams pp;
    pp.B = B; pp.H = H; pp.Lq = Lq; pp.Lk = Lk; pp.d = d; pp.scale_log2 = p.scale_log2;
    pp.q_hi = static_cast<const __nv_bfloat16*>(q_hi);
    pp.q_lo = three ? static_cast<const __nv_bfloat16*>(q_lo) : nullptr;
    pp.kv_hi = static_cast<const __nv_bfloat16*>(kv_hi);
    pp.kv_lo = three ? static_cast<const __nv_bfloat16*>(kv_lo) : nullptr;
    pp.ldq = (int)ldq; pp.ldkv = (int)ldkv; pp.q_col0 = q_col0; pp.k_col0 = k_col0; pp.q_bcast = q_bcast;
    pp.mask = mask; pp.msb = msb; pp.msq = msq; pp.msk = msk;
    pp.row_max = row_max; pp.row_sum = row_sum; pp.probs = probs;
    pp.probs_pre = probs_pre; pp.drop_thresh = p.drop_thresh; pp.drop_scale = p.drop_scale; pp.drop_seed = seed;
    pp.drop_seed_dev = p.drop_seed_dev;
    static PerDeviceOnce once;
    if (int once_rc = per_device_once(once, [] { int rc_ = set_smem(attn_probs_mma_kernel, (uint32_t)attn_probs_smem_bytes(BWD_DMAX)); return rc_; })) return once_rc;
    const long long grid = (long long)H * B * ((Lq + BWD_TILE - 1) / BWD_TILE);
    REQUIRE(grid < (1LL << 31), "attn: probability grid too large");
    attn_probs_mma_kernel<<<(unsigned)grid, BWD_THREADS, (uint32_t)attn_probs_smem_bytes(d), st>>>(pp);
    rc = launch_check();
  }
  return rc;
}

extern "C++" {
namespace {
struct AttnBwdPlan {
  size_t qp, kp, vp, dop, dA, dSp, Ap, total;
  int ld;
};
AttnBwdPlan attn_bwd_plan(int N, int Lq, int Lk, int d) {
  AttnBwdPlan pl{};
  pl.ld = (Lk + 7) / 8 * 8;
  Carver c(nullptr);
  const size_t nq = (size_t)N * Lq, nk = (size_t)N * Lk;
  pl.qp = c.off;  c.take(nq * d * 4);
  pl.kp = c.off;  c.take(nk * d * 4);
  pl.vp = c.off;  c.take(nk * d * 4);
  pl.dop = c.off; c.take(nq * d * 4);
  pl.dA = c.off;  c.take(nq * pl.ld * 4);   // fp32, row pitch ld (TMA-stored)
  pl.dSp = c.off; c.take(nq * pl.ld * 4);
  pl.Ap = c.off;  c.take(nq * pl.ld * 4);
  pl.total = c.off;
  return pl;
}

// One operand of a batched product: planes + (cols, rows) of ONE (sample, head) matrix and the element strides of the
// row / head / sample coordinates.
struct BgOperand {
  const void* hi;
  const void* lo;
  uint64_t cols, rows, ld, head_stride, sample_stride;
};
// head-major contiguous [N, rows, ld] tensor: N plain batches (H = 1)
inline BgOperand bg_headmajor(const void* hi, const void* lo, uint64_t cols, uint64_t rows, uint64_t ld) {
  return BgOperand{hi, lo, cols, rows, ld, rows * ld, rows * ld};
}
// head column slice of a [B*rows, ld] plane matrix: head h at column col0 + h*d
inline BgOperand bg_slices(const void* hi, const void* lo, uint64_t col0, uint64_t d, uint64_t rows, uint64_t ld) {
  return BgOperand{static_cast<const __nv_bfloat16*>(hi) + col0,
                   lo ? static_cast<const void*>(static_cast<const __nv_bfloat16*>(lo) + col0) : nullptr, d, rows, ld, d,
                   rows * ld};
}

// C (fp32 and/or planes): element (sample, head, m, n) at [sample*stride_c + head*stride_ch + m*ldc + n]
struct BgOutput {
  float* f32;
  void* hi;
  void* lo;
  long long ldc, stride_c, stride_ch;
};

template <bool A_MN, bool B_MN>
int launch_bgemm(const BgOperand& a, const BgOperand& b, int samples, int H, int M, int N, int Kc, float scale,
                 const BgOutput& c_in, cudaStream_t st) {
  constexpr int TN = 128;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  const uint32_t a_box_rows = A_MN ? 64 : 128, b_box_rows = B_MN ? 64 : TN;
  auto mk = [&](CUtensorMap* m, const void* base, const BgOperand& o, uint32_t box_rows) {
    return make_tmap4(m, base, o.cols, o.rows, (uint64_t)H, (uint64_t)samples, o.ld, o.head_stride, o.sample_stride, box_rows);
  };
  if (int rc = mk(&ta_hi, a.hi, a, a_box_rows)) return rc;
  if (int rc = mk(&ta_lo, a.lo ? a.lo : a.hi, a, a_box_rows)) return rc;
  if (int rc = mk(&tb_hi, b.hi, b, b_box_rows)) return rc;
  if (int rc = mk(&tb_lo, b.lo ? b.lo : b.hi, b, b_box_rows)) return rc;
  // output maps: the same {cols, rows, head, sample} indexing, boxes of 32 rows x 128 B (one per epilogue warp)
  CUtensorMap tc0, tc1;
  BgOutput c = c_in;
  if (H == 1 && c.stride_ch == 0) c.stride_ch = c.stride_c;   // a size-1 dimension still needs a valid stride
  if (c.f32 != nullptr) {
    if (c.hi != nullptr) return fail(LAMP_EINVAL, "bgemm: give an fp32 output or plane outputs, not both");
    if (int rc = make_tmap4(&tc0, c.f32, (uint64_t)N, (uint64_t)M, (uint64_t)H, (uint64_t)samples, (uint64_t)c.ldc,
                            (uint64_t)c.stride_ch, (uint64_t)c.stride_c, 32, true)) return rc;
    tc1 = tc0;
  } else {
    if (int rc = make_tmap4(&tc0, c.hi, (uint64_t)N, (uint64_t)M, (uint64_t)H, (uint64_t)samples, (uint64_t)c.ldc,
                            (uint64_t)c.stride_ch, (uint64_t)c.stride_c, 32)) return rc;
    if (c.lo != nullptr) {
      if (int rc = make_tmap4(&tc1, c.lo, (uint64_t)N, (uint64_t)M, (uint64_t)H, (uint64_t)samples, (uint64_t)c.ldc,
                              (uint64_t)c.stride_ch, (uint64_t)c.stride_c, 32)) return rc;
    } else {
      tc1 = tc0;
    }
  }
  auto kernel = bgemm_tc_kernel<A_MN, B_MN, 3, TN>;
  static PerDeviceOnce once;
  if (int once_rc = per_device_once(once, [kernel] { int rc_ = set_smem(kernel, bg_smem_bytes(2, TN)); return rc_; })) return once_rc;
  BgemmParams p;
  p.batch = samples * H; p.H = H; p.M = M; p.N = N; p.Kc = Kc; p.scale = scale;
  p.out_f32 = c.f32 != nullptr ? 1 : 0; p.out_lo = c.lo != nullptr ? 1 : 0;
  const long long items = (long long)p.batch * ((M + 127) / 128) * ((N + TN - 1) / TN);
  const long long grid = items < sm_count_cached() ? items : sm_count_cached();  // persistent: one CTA per SM
  kernel<<<(unsigned)grid, BG_THREADS, bg_smem_bytes(2, TN), st>>>(ta_hi, ta_lo, tb_hi, tb_lo, tc0, tc1, p);
  return launch_check();
}
}  // namespace
}  // extern "C++"

size_t lamp_attn_core_bwd_workspace_bytes(int N, int Lq, int Lk, int d) { return attn_bwd_plan(N, Lq, Lk, d).total; }

int lamp_attn_core_bwd(const float* q, const float* k, const float* v, const float* dO, const float* O, const float* P,
                       const float* A, float* dq, float* dk, float* dv, int N, int Lq, int Lk, int d, float temperature,
                       float p_drop, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(q && k && v && dO && O && P && dq && dk && dv, "attn_bwd: null pointer");
  REQUIRE(N >= 0 && Lq > 0 && Lk > 0, "attn_bwd: bad shape N=%d Lq=%d Lk=%d", N, Lq, Lk);
  REQUIRE(d % 16 == 0 && d >= 16 && d <= BWD_DMAX, "attn_bwd: head width %d must be a multiple of 16 in [16,128]", d);
  REQUIRE(temperature > 0.0f && p_drop >= 0.0f && p_drop < 1.0f, "attn_bwd: bad temperature / dropout rate");
  REQUIRE((long long)N * Lq * Lk < (1LL << 40), "attn_bwd: probability tensor too large");
  const AttnBwdPlan pl = attn_bwd_plan(N, Lq, Lk, d);
  if (!workspace || workspace_bytes < pl.total) return fail(LAMP_EWORKSPACE, "attn_bwd: workspace too small");
  if (N == 0) return LAMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t nq = (size_t)N * Lq, nk = (size_t)N * Lk;
  if (g_attn_bwd_tc.load() == 0) {
    // warp-MMA version (attn_bwd.cuh): fp32 operands, dS through the dA slot of the workspace
    AttnBwdParams p;
    p.N = N; p.Lq = Lq; p.Lk = Lk; p.d = d;
    p.inv_temp = 1.0f / temperature;
    p.drop_scale = 1.0f / (1.0f - p_drop);
    p.q = q; p.k = k; p.v = v; p.dO = dO; p.O = O; p.P = P; p.A = A ? A : P;
    p.dS = reinterpret_cast<float*>(ws + pl.dA); p.dq = dq; p.dk = dk; p.dv = dv;
    const uint32_t smem = (uint32_t)attn_bwd_smem_bytes(d);
    static PerDeviceOnce once;
    if (int once_rc = per_device_once(once, [] { int rc_ = set_smem(attn_bwd_dq_kernel, (uint32_t)attn_bwd_smem_bytes(BWD_DMAX));
      if (rc_ == LAMP_OK) rc_ = set_smem(attn_bwd_dkv_kernel, (uint32_t)attn_bwd_smem_bytes(BWD_DMAX)); return rc_; })) return once_rc;
    const long long gq = (long long)N * ((Lq + BWD_TILE - 1) / BWD_TILE);
    const long long gk = (long long)N * ((Lk + BWD_TILE - 1) / BWD_TILE);
    REQUIRE(gq < (1LL << 31) && gk < (1LL << 31), "attn_bwd: grid too large");
    attn_bwd_dq_kernel<<<(unsigned)gq, BWD_THREADS, smem, st>>>(p);
    if (int rc = launch_check()) return rc;
    attn_bwd_dkv_kernel<<<(unsigned)gk, BWD_THREADS, smem, st>>>(p);
    return launch_check();
  }
  // tcgen05 version: four batched products on the operand planes + one element-wise kernel
  auto hi = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  __nv_bfloat16 *qh = hi(pl.qp), *ql = qh + nq * d, *kh = hi(pl.kp), *kl = kh + nk * d, *vh = hi(pl.vp), *vl = vh + nk * d;
  __nv_bfloat16 *gh = hi(pl.dop), *gl = gh + nq * d;
  __nv_bfloat16 *sh = hi(pl.dSp), *sl = sh + nq * pl.ld, *ah = hi(pl.Ap), *al = ah + nq * pl.ld;
  float* dA = reinterpret_cast<float*>(ws + pl.dA);
  if (int rc = lamp_split_planes(q, (int64_t)nq, d, d, qh, ql, d, stream)) return rc;
  if (int rc = lamp_split_planes(k, (int64_t)nk, d, d, kh, kl, d, stream)) return rc;
  if (int rc = lamp_split_planes(v, (int64_t)nk, d, d, vh, vl, d, stream)) return rc;
  if (int rc = lamp_split_planes(dO, (int64_t)nq, d, d, gh, gl, d, stream)) return rc;
  // dA = dO V^T  [N, Lq, Lk]
  if (int rc = launch_bgemm<false, false>(bg_headmajor(gh, gl, d, Lq, d), bg_headmajor(vh, vl, d, Lk, d), N, 1, Lq, Lk, d, 1.0f,
                                          BgOutput{dA, nullptr, nullptr, pl.ld, (long long)Lq * pl.ld, 0}, st)) return rc;
  // dS (scaled by 1/temperature) and A as planes [N*Lq, ld]
  {
    const long long blocks = ((long long)nq * 32 + 255) / 256;
    attn_bwd_ds_kernel<<<(unsigned)blocks, 256, 0, st>>>(dA, P, A ? A : P, dO, O, (long long)nq, Lk, d, pl.ld,
                                                         1.0f / temperature, 1.0f / (1.0f - p_drop), sh, sl, ah, al);
    if (int rc = launch_check()) return rc;
  }
  // dQ = dS K   (A: dS K-major, B: K MN-major)
  if (int rc = launch_bgemm<false, true>(bg_headmajor(sh, sl, Lk, Lq, pl.ld), bg_headmajor(kh, kl, d, Lk, d), N, 1, Lq, d, Lk, 1.0f,
                                         BgOutput{dq, nullptr, nullptr, d, (long long)Lq * d, 0}, st)) return rc;
  // dV = A^T dO, dK = dS^T Q   (both operands MN-major: contraction over the q rows)
  if (int rc = launch_bgemm<true, true>(bg_headmajor(ah, al, Lk, Lq, pl.ld), bg_headmajor(gh, gl, d, Lq, d), N, 1, Lk, d, Lq, 1.0f,
                                        BgOutput{dv, nullptr, nullptr, d, (long long)Lk * d, 0}, st)) return rc;
  return launch_bgemm<true, true>(bg_headmajor(sh, sl, Lk, Lq, pl.ld), bg_headmajor(qh, ql, d, Lq, d), N, 1, Lk, d, Lq, 1.0f,
                                  BgOutput{dk, nullptr, nullptr, d, (long long)Lk * d, 0}, st);
}

/* Attention backward in the layouts of the projection GEMMs (training path; no head-major copies, no fp32 round trips):
 * Q / K / V / dO / O are split-bf16 planes whose head h is the column slice [col0 + h*d, col0 + (h+1)*d) of a
 * [B*L, ld] matrix; P (softmax before dropout) and A (after dropout; may equal P) are the head-major fp32
 * [H*B, Lq, Lk] tensors the training forward wrote.  dQ / dK / dV leave as planes in the same slice layout. */
size_t lamp_attn_bwd_planes_workspace_bytes(int B, int H, int Lq, int Lk) {
  const size_t nq = (size_t)B * H * Lq, ld = (size_t)(Lk + 7) / 8 * 8;
  return 4 * align_up(nq * ld * 4);  // dA, S (recompute form; fp32 with row pitch ld), dS planes, A planes
}

int lamp_attn_bwd_planes(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, const void* kv_hi, const void* kv_lo,
                         int64_t ldkv, int k_col0, int v_col0, const void* do_hi, const void* do_lo, const void* o_hi,
                         const void* o_lo, int64_t ldo, const float* P, const float* A, void* dq_hi, void* dq_lo,
                         int64_t lddq, int dq_col0, void* dkv_hi, void* dkv_lo, int64_t lddkv, int dk_col0, int dv_col0,
                         int B, int H, int Lq, int Lk, int d, float temperature, float p_drop, const float* row_max,
                         const float* row_sum, const uint8_t* mask, int64_t msb, int64_t msq, int64_t msk, uint64_t seed,
                         const uint64_t* seed_dev, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(q_hi && q_lo && kv_hi && kv_lo && do_hi && do_lo && o_hi && o_lo && dq_hi && dq_lo && dkv_hi && dkv_lo,
          "attn_bwd_planes: null pointer (3-term planes required)");
  REQUIRE(P != nullptr || (row_max && row_sum), "attn_bwd_planes: give P (and A), or the forward's row statistics to recompute them");
  REQUIRE(B >= 0 && H > 0 && Lq > 0 && Lk > 0, "attn_bwd_planes: bad shape");
  REQUIRE(d % 16 == 0 && d >= 16 && d <= 128, "attn_bwd_planes: head width %d must be a multiple of 16 in [16,128]", d);
  REQUIRE(temperature > 0.0f && p_drop >= 0.0f && p_drop < 1.0f, "attn_bwd_planes: bad temperature / dropout rate");
  REQUIRE(q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0 && dq_col0 % 8 == 0 && dk_col0 % 8 == 0 && dv_col0 % 8 == 0,
          "attn_bwd_planes: column offsets must be multiples of 8");
  if (!workspace || workspace_bytes < lamp_attn_bwd_planes_workspace_bytes(B, H, Lq, Lk))
    return fail(LAMP_EWORKSPACE, "attn_bwd_planes: workspace too small");
  if (B == 0) return LAMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nq = (size_t)B * H * Lq;
  const int ld = (Lk + 7) / 8 * 8;
  Carver cv(workspace);
  float* dA = static_cast<float*>(cv.take(nq * ld * 4));
  float* S = static_cast<float*>(cv.take(nq * ld * 4));
  __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(cv.take(nq * ld * 4));
  __nv_bfloat16* sl = sh + nq * ld;
  __nv_bfloat16* ah = static_cast<__nv_bfloat16*>(cv.take(nq * ld * 4));
  __nv_bfloat16* al = ah + nq * ld;
  const BgOperand Q = bg_slices(q_hi, q_lo, q_col0, d, Lq, ldq), K = bg_slices(kv_hi, kv_lo, k_col0, d, Lk, ldkv),
                  V = bg_slices(kv_hi, kv_lo, v_col0, d, Lk, ldkv), G = bg_slices(do_hi, do_lo, 0, d, Lq, ldo);
  // head-major [H*B, Lq, .] tensors: sample stride Lq*ld, head stride B*Lq*ld
  auto hm = [&](const void* hi, const void* lo, uint64_t cols, uint64_t pitch) {
    return BgOperand{hi, lo, cols, (uint64_t)Lq, pitch, (uint64_t)B * Lq * pitch, (uint64_t)Lq * pitch};
  };
  // dA = dO V^T   [H*B, Lq, Lk] fp32
  if (int rc = launch_bgemm<false, false>(G, V, B, H, Lq, Lk, d, 1.0f,
                                          BgOutput{dA, nullptr, nullptr, ld, (long long)Lq * ld, (long long)B * Lq * ld}, st)) return rc;
  const long long ds_blocks = ((long long)nq * 32 + 255) / 256;
  if (P != nullptr) {
    attn_bwd_ds_planes_kernel<<<(unsigned)ds_blocks, 256, 0, st>>>(
        dA, P, A ? A : P, static_cast<const __nv_bfloat16*>(do_hi), static_cast<const __nv_bfloat16*>(do_lo),
        static_cast<const __nv_bfloat16*>(o_hi), static_cast<const __nv_bfloat16*>(o_lo), ldo, B, H, Lq, Lk, d, ld,
        1.0f / temperature, 1.0f / (1.0f - p_drop), sh, sl, ah, al);
    if (int rc = launch_check()) return rc;
  } else {
    // recompute form: S = Q K^T once more, P rebuilt inside the element-wise kernel
    if (int rc = launch_bgemm<false, false>(Q, K, B, H, Lq, Lk, d, 1.0f,
                                            BgOutput{S, nullptr, nullptr, ld, (long long)Lq * ld, (long long)B * Lq * ld}, st)) return rc;
    DsRecomputeParams rp;
    rp.S = S; rp.dA = dA; rp.row_max = row_max; rp.row_sum = row_sum;
    rp.mask = mask; rp.msb = msb; rp.msq = msq; rp.msk = msk;
    rp.scale_log2 = 1.4426950408889634f / temperature; rp.inv_temp = 1.0f / temperature;
    rp.drop_scale = 1.0f / (1.0f - p_drop);
    rp.drop_thresh = drop_threshold(p_drop);
    rp.drop_seed = seed; rp.drop_seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
    rp.dO_hi = static_cast<const __nv_bfloat16*>(do_hi); rp.dO_lo = static_cast<const __nv_bfloat16*>(do_lo);
    rp.O_hi = static_cast<const __nv_bfloat16*>(o_hi); rp.O_lo = static_cast<const __nv_bfloat16*>(o_lo);
    rp.ld_o = ldo; rp.B = B; rp.H = H; rp.Lq = Lq; rp.Lk = Lk; rp.d = d; rp.ld = ld;
    rp.dS_hi = sh; rp.dS_lo = sl; rp.A_hi = ah; rp.A_lo = al;
    attn_bwd_ds_recompute_kernel<<<(unsigned)ds_blocks, 256, 0, st>>>(rp);
    if (int rc = launch_check()) return rc;
  }
  const BgOperand dS = hm(sh, sl, Lk, ld), Ap = hm(ah, al, Lk, ld);
  auto out = [&](void* hi, void* lo, int col0, int64_t ldc, int rows) {
    return BgOutput{nullptr, static_cast<__nv_bfloat16*>(hi) + col0, static_cast<__nv_bfloat16*>(lo) + col0, (long long)ldc,
                    (long long)rows * ldc, (long long)d};
  };
  // dQ = dS K (K MN-major), dV = A^T dO, dK = dS^T Q (both operands MN-major)
  if (int rc = launch_bgemm<false, true>(dS, K, B, H, Lq, d, Lk, 1.0f, out(dq_hi, dq_lo, dq_col0, lddq, Lq), st)) return rc;
  if (int rc = launch_bgemm<true, true>(Ap, G, B, H, Lk, d, Lq, 1.0f, out(dkv_hi, dkv_lo, dv_col0, lddkv, Lk), st)) return rc;
  return launch_bgemm<true, true>(dS, Q, B, H, Lk, d, Lq, 1.0f, out(dkv_hi, dkv_lo, dk_col0, lddkv, Lk), st);
}

static int layernorm_bwd_impl(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int D, float* dx,
                              float* dgamma, float* dbeta, const LnBwdDrop& dp, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(x && dy && gamma && dx && dgamma && dbeta, "layernorm_bwd: null pointer");
  REQUIRE(D % 4 == 0 && D > 0 && D <= 4096, "layernorm_bwd: D=%d must be a multiple of 4, <= 4096", D);
  REQUIRE(aligned16(x) && aligned16(dy) && aligned16(gamma) && aligned16(dx), "layernorm_bwd: alignment");
  REQUIRE(!dp.hi || (dp.lo && aligned16(dp.hi) && aligned16(dp.lo)), "layernorm_bwd: plane outputs need hi and lo, 16-byte aligned");
  if (rows == 0) return LAMP_OK;
  long long blocks = (rows * 32 + 255) / 256;
  const long long cap = 4LL * sm_count_cached();  // grid-stride: few, long-lived blocks keep the atomic count low
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t red = (size_t)2 * D * sizeof(float);
#define LAMP_LNB(MAXV)                                                                                                  \
  do {                                                                                                                  \
    if (dp.hi != nullptr)                                                                                               \
      layernorm_bwd_kernel<MAXV, true><<<(unsigned)blocks, 256, red, st>>>(x, dy, gamma, eps, rows, D, dx, dgamma, dbeta, dp); \
    else                                                                                                                \
      layernorm_bwd_kernel<MAXV, false><<<(unsigned)blocks, 256, red, st>>>(x, dy, gamma, eps, rows, D, dx, dgamma, dbeta, dp); \
  } while (0)
  if (D <= 512) LAMP_LNB(4);
  else if (D <= 1024) LAMP_LNB(8);
  else LAMP_LNB(32);
#undef LAMP_LNB
  return launch_check();
}

int lamp_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int D, float* dx,
                       float* dgamma, float* dbeta, void* stream) {
  return layernorm_bwd_impl(x, dy, gamma, eps, rows, D, dx, dgamma, dbeta, LnBwdDrop{nullptr, nullptr, 0u, 1.0f, 0ull, nullptr}, stream);
}

int lamp_layernorm_bwd_drop(const float* x, const float* dy, const float* gamma, float eps, int64_t rows, int D, float* dx,
                            float* dgamma, float* dbeta, float p_drop, uint64_t seed, const uint64_t* seed_dev, void* dx_hi,
                            void* dx_lo, void* stream) {
  REQUIRE(dx_hi && dx_lo, "layernorm_bwd_drop: plane outputs required");
  REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "layernorm_bwd_drop: rate outside [0, 1)");
  LnBwdDrop dp{static_cast<__nv_bfloat16*>(dx_hi), static_cast<__nv_bfloat16*>(dx_lo), drop_threshold(p_drop),
               1.0f / (1.0f - p_drop), seed, reinterpret_cast<const unsigned long long*>(seed_dev)};
  return layernorm_bwd_impl(x, dy, gamma, eps, rows, D, dx, dgamma, dbeta, dp, stream);
}

// bias gradient next to a weight gradient: column sums of the dY planes into the (zeroed or accumulating) db
static int launch_colsum(const void* dy_hi, const void* dy_lo, int64_t ldy, int64_t M, int N, float* db, cudaStream_t st) {
  const unsigned slabs = (unsigned)((N + 255) / 256);
  long long rb = (4LL * sm_count_cached() + slabs - 1) / slabs;   // ~4 blocks per SM
  long long chunk = ((M + rb - 1) / rb + 31) / 32 * 32;           // whole 32-row steps of the unrolled loop
  if (chunk < 32) chunk = 32;
  rb = (M + chunk - 1) / chunk;
  colsum_planes_kernel<<<dim3((unsigned)rb, slabs), COLSUM_THREADS, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dy_hi), static_cast<const __nv_bfloat16*>(dy_lo), ldy, M, N, chunk, db);
  return launch_check();
}

int lamp_gemm_tn_acc(const void* dy_hi, const void* dy_lo, int64_t ldy, const void* x_hi, const void* x_lo, int64_t ldx,
                     int64_t M, int N, int K, float* dW, float* db, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(dy_hi && x_hi && dW && M >= 0 && N > 0 && K > 0, "gemm_tn: bad arguments");
  REQUIRE(N % 8 == 0 && K % 8 == 0 && ldy % 8 == 0 && ldx % 8 == 0 && ldy >= N && ldx >= K, "gemm_tn: N, K and the leading dimensions must be multiples of 8");
  REQUIRE(aligned16(dy_hi) && aligned16(x_hi) && (!dy_lo || aligned16(dy_lo)) && (!x_lo || aligned16(x_lo)), "gemm_tn: alignment");
  if (M == 0) return LAMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_gemm_tn_tc.load() != 0) {
    // tcgen05 path: dY and X read MN-major straight from their row-major planes
    const bool three = dy_lo != nullptr && x_lo != nullptr;
    CUtensorMap ty_hi, ty_lo, tx_hi, tx_lo;
    if (int rc = make_tmap(&ty_hi, dy_hi, (uint64_t)N, (uint64_t)M, 1, (uint64_t)ldy, TNC_BM, false)) return rc;
    if (int rc = make_tmap(&tx_hi, x_hi, (uint64_t)K, (uint64_t)M, 1, (uint64_t)ldx, TNC_BM, false)) return rc;
    if (three) {
      if (int rc = make_tmap(&ty_lo, dy_lo, (uint64_t)N, (uint64_t)M, 1, (uint64_t)ldy, TNC_BM, false)) return rc;
      if (int rc = make_tmap(&tx_lo, x_lo, (uint64_t)K, (uint64_t)M, 1, (uint64_t)ldx, TNC_BM, false)) return rc;
    } else {
      ty_lo = ty_hi;
      tx_lo = tx_hi;
    }
    static PerDeviceOnce once_tc;
    if (int once_tc_rc = per_device_once(once_tc, [] { int rc_ = set_smem(gemm_tn_tc_kernel<3>, tnc_smem_bytes(2));
      if (rc_ == LAMP_OK) rc_ = set_smem(gemm_tn_tc_kernel<1>, tnc_smem_bytes(1)); return rc_; })) return once_tc_rc;
    const long long tiles = (long long)((N + TNC_TILE_N - 1) / TNC_TILE_N) * ((K + TNC_TILE_K - 1) / TNC_TILE_K);
    // one CTA per SM (192 KB of shared memory each) and ONE wave: tiles * splits <= #SMs.  (Rounding the split count
    // up gave 152 / 160 / 168 CTAs on 148 SMs: the 4-20 CTAs of the second wave doubled the kernel's duration -- ncu
    // showed the tensor pipe at 74-89 % of the active but only 40-45 % of the elapsed cycles.)
    long long splits = sm_count_cached() / tiles;
    const long long max_splits = (M + TNC_BM - 1) / TNC_BM;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    GemmTnParams gp;
    gp.M = M; gp.N = N; gp.K = K; gp.dW = dW;
    gp.chunk = ((M + splits - 1) / splits + TNC_BM - 1) / TNC_BM * TNC_BM;
    splits = (M + gp.chunk - 1) / gp.chunk;
    REQUIRE(M < (1LL << 31), "gemm_tn: M too large for TMA coordinates");
    if (three)
      gemm_tn_tc_kernel<3><<<(unsigned)(tiles * splits), TNC_THREADS, tnc_smem_bytes(2), st>>>(ty_hi, ty_lo, tx_hi, tx_lo, gp);
    else
      gemm_tn_tc_kernel<1><<<(unsigned)(tiles * splits), TNC_THREADS, tnc_smem_bytes(1), st>>>(ty_hi, ty_lo, tx_hi, tx_lo, gp);
    if (int rc = launch_check()) return rc;
    if (db != nullptr) return launch_colsum(dy_hi, dy_lo, ldy, M, N, db, st);
    return LAMP_OK;
  }
  static PerDeviceOnce once;
  if (int once_rc = per_device_once(once, [] { int rc_ = set_smem(gemm_tn_kernel, (uint32_t)gemm_tn_smem_bytes()); return rc_; })) return once_rc;
  const long long tiles = (long long)((N + TN_TILE - 1) / TN_TILE) * ((K + TN_TILE - 1) / TN_TILE);
  // split the M rows so that ~3 CTAs per SM exist; chunks are multiples of the 64-row staging tile
  long long splits = (3LL * sm_count_cached() + tiles - 1) / tiles;
  const long long max_splits = (M + BWD_TILE - 1) / BWD_TILE;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long chunk = ((M + splits - 1) / splits + BWD_TILE - 1) / BWD_TILE * BWD_TILE;
  splits = (M + chunk - 1) / chunk;
  REQUIRE(tiles * splits < (1LL << 31), "gemm_tn: grid too large");
  gemm_tn_kernel<<<(unsigned)(tiles * splits), BWD_THREADS, (uint32_t)gemm_tn_smem_bytes(), st>>>(
      static_cast<const __nv_bfloat16*>(dy_hi), static_cast<const __nv_bfloat16*>(dy_lo), ldy,
      static_cast<const __nv_bfloat16*>(x_hi), static_cast<const __nv_bfloat16*>(x_lo), ldx, M, N, K, chunk, dW);
  if (int rc = launch_check()) return rc;
  if (db != nullptr) return launch_colsum(dy_hi, dy_lo, ldy, M, N, db, st);
  return LAMP_OK;
}

int lamp_diag_proj_bwd(const float* g, const float* x, const float* W, int64_t B, int L, int D, float* dx, float* dW,
                       float* dbias, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(g && x && W && (dx || dW), "diag_proj_bwd: null pointer");
  REQUIRE(D % 4 == 0 && aligned16(W) && (!dx || aligned16(dx)), "diag_proj_bwd: D multiple of 4 and 16-byte alignment required");
  const long long rows = B * L;
  if (rows == 0) return LAMP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx != nullptr) {
    const long long blocks = (rows * 32 + 255) / 256;
    diag_proj_bwd_dx_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, W, rows, L, D, dx);
    if (int rc = launch_check()) return rc;
  }
  if (dW != nullptr) {
    REQUIRE(aligned16(x) && aligned16(dW), "diag_proj_bwd: x and dW must be 16-byte aligned");
    // ~4 blocks per SM: labels x batch slices, partial sums reduced into the zeroed outputs
    long long slices = (4LL * sm_count_cached() + L - 1) / L;
    if (slices > (B + 7) / 8) slices = (B + 7) / 8;
    if (slices < 1) slices = 1;
    const long long bchunk = (B + slices - 1) / slices;
    slices = (B + bchunk - 1) / bchunk;
    if (cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)L * D, st) != cudaSuccess) return launch_check();
    if (dbias != nullptr && cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)L, st) != cudaSuccess) return launch_check();
    const int threads = D / 4 >= 256 ? 256 : (D / 4 >= 128 ? 128 : 64);
    diag_proj_bwd_dw_kernel<<<dim3((unsigned)L, (unsigned)slices), threads, 0, st>>>(g, x, B, L, D, bchunk, dW, dbias);
    return launch_check();
  }
  return LAMP_OK;
}

/* Training forward of the attention core on operand planes (the layouts of the projection GEMMs): like
 * lamp_attn_core_planes, plus dropout on the probabilities inside the kernel and the two probability tensors the
 * backward consumes (attn: after dropout -- the reference's return value; probs_pre: before dropout, NULL when
 * p_drop == 0).  row_max / row_sum: [H*B*Lq] scratch. */
int lamp_attn_core_planes_train(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast, const void* kv_hi,
                                const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B, int H, int Lq, int Lk, int d,
                                float temperature, int precision, const uint8_t* mask, int64_t msb, int64_t msq,
                                int64_t msk, void* o_hi, void* o_lo, int64_t ldo, float* row_max, float* row_sum,
                                float* attn, float* probs_pre, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                void* stream) {
  REQUIRE(row_max && row_sum, "attn_train: the row statistics buffers are required");
  REQUIRE(attn == nullptr || p_drop == 0.0f || probs_pre != nullptr,
          "attn_train: probs_pre is required next to attn when dropout is active");
  return attn_impl(q_hi, q_lo, ldq, q_col0, q_bcast, kv_hi, kv_lo, ldkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature,
                   precision, mask, msb, msq, msk, nullptr, 0, 0, o_hi, o_lo, ldo, nullptr, 0, row_max, row_sum, attn,
                   nullptr, nullptr, 0, stream, p_drop, seed, probs_pre, seed_dev);
}

int lamp_attn_core_planes_train_mbits(const void* q_hi, const void* q_lo, int64_t ldq, int q_col0, int q_bcast,
                                      const void* kv_hi, const void* kv_lo, int64_t ldkv, int k_col0, int v_col0, int B,
                                      int H, int Lq, int Lk, int d, float temperature, int precision,
                                      const uint32_t* mask_bits, int64_t mbb, int64_t mbq, void* o_hi, void* o_lo,
                                      int64_t ldo, float* row_max, float* row_sum, float p_drop, uint64_t seed,
                                      const uint64_t* seed_dev, void* stream) {
  REQUIRE(row_max && row_sum, "attn_train: the row statistics buffers are required");
  REQUIRE(mask_bits != nullptr && mbq >= (Lk + 31) / 32, "attn_train_mbits: packed mask missing or row stride too small");
  return attn_impl(q_hi, q_lo, ldq, q_col0, q_bcast, kv_hi, kv_lo, ldkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature,
                   precision, nullptr, 0, 0, 0, mask_bits, mbb, mbq, o_hi, o_lo, ldo, nullptr, 0, row_max, row_sum, nullptr,
                   nullptr, nullptr, 0, stream, p_drop, seed, nullptr, seed_dev);
}

int lamp_dropout_add(const float* y0, const float* x, int64_t rows, int D, int x_mod, float p_drop, uint64_t seed,
                     const uint64_t* seed_dev, float* y, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(y0 && x && y && rows >= 0 && D > 0 && D % 4 == 0, "dropout_add: bad arguments");
  REQUIRE(aligned16(y0) && aligned16(x) && aligned16(y), "dropout_add: alignment");
  REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "dropout_add: rate outside [0, 1)");
  if (rows == 0) return LAMP_OK;
  const uint32_t thresh = drop_threshold(p_drop);
  const long long blocks = (rows * 32 + 255) / 256;
  dropout_add_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      y0, x, rows, D, x_mod, thresh, 1.0f / (1.0f - p_drop), seed, reinterpret_cast<const unsigned long long*>(seed_dev), y);
  return launch_check();
}

int lamp_dropout_split(const float* dy, int64_t rows, int D, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                       void* hi, void* lo, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(dy && hi && lo && rows >= 0 && D > 0 && D % 4 == 0, "dropout_split: bad arguments");
  REQUIRE(aligned16(dy) && aligned16(hi) && aligned16(lo), "dropout_split: alignment");
  REQUIRE(p_drop >= 0.0f && p_drop < 1.0f, "dropout_split: rate outside [0, 1)");
  if (rows == 0) return LAMP_OK;
  const uint32_t thresh = drop_threshold(p_drop);
  const long long blocks = (rows * 32 + 255) / 256;
  dropout_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      dy, rows, D, thresh, 1.0f / (1.0f - p_drop), seed, reinterpret_cast<const unsigned long long*>(seed_dev),
      static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo));
  return launch_check();
}

int lamp_relu_mask_planes(void* g_hi, void* g_lo, const void* h_hi, int64_t n, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(g_hi && g_lo && h_hi && n >= 0 && n % 8 == 0, "relu_mask_planes: element count must be a multiple of 8");
  REQUIRE(aligned16(g_hi) && aligned16(g_lo) && aligned16(h_hi), "relu_mask_planes: alignment");
  if (n == 0) return LAMP_OK;
  const long long n8 = n / 8, blocks = (n8 + 255) / 256;
  relu_mask_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      static_cast<__nv_bfloat16*>(g_hi), static_cast<__nv_bfloat16*>(g_lo), static_cast<const __nv_bfloat16*>(h_hi), n8);
  return launch_check();
}

int lamp_gold_binary(const int64_t* gold, int64_t B, int W, int L, int skip, float* out, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(gold && out && B >= 0 && W > 0 && L > 0 && skip >= 0, "gold_binary: bad arguments");
  if (B == 0) return LAMP_OK;
  const long long blocks = (B * 32 + 255) / 256;
  gold_binary_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(gold), B, W, L,
                                                                         skip, out);
  return launch_check();
}

size_t lamp_bce_logits_workspace_bytes(void) { return (LAMP_BCE_MAX_BLOCKS + 4) * sizeof(float); }

int lamp_bce_logits(const float* logits, const float* target, int64_t n, float* loss, float* dlogits, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(logits && target && loss && workspace && n > 0, "bce_logits: bad arguments");
  REQUIRE(workspace_bytes >= lamp_bce_logits_workspace_bytes() && aligned16(workspace), "bce_logits: workspace too small");
  // workspace: [0] ticket counter (zero on first use: the caller clears the workspace ONCE; the kernel re-arms it),
  // [4 ..] per-block partial sums
  long long blocks = (n + BCE_THREADS * 4 - 1) / (BCE_THREADS * 4);
  if (blocks > LAMP_BCE_MAX_BLOCKS) blocks = LAMP_BCE_MAX_BLOCKS;
  float* ws = static_cast<float*>(workspace);
  bce_logits_kernel<<<(unsigned)blocks, BCE_THREADS, 0, (cudaStream_t)stream>>>(
      logits, target, n, 1.0f / static_cast<float>(n), dlogits, ws + 4, reinterpret_cast<unsigned int*>(ws), loss);
  return launch_check();
}

#ifdef LAMP_ATTN_TRACE
/* debug build only: copies the [64][64] clock64() stamps of CTA 0 of the last attention launch to the host */
int lamp_debug_attn_trace(unsigned long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, lamp::g_attn_trace, sizeof(unsigned long long) * 64 * 64) == cudaSuccess ? 0 : -1;
}
#endif

int lamp_layernorm(const float* y, const float* add, int add_mod, const float* gamma, const float* beta, float eps,
                   int64_t rows, int D, float* out, void* out_hi, void* out_lo, const int32_t* m_dev, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(y && gamma && beta && (out || out_hi), "layernorm: null pointer");
  REQUIRE(D % 4 == 0 && D > 0 && D <= 4096, "layernorm: D=%d must be a multiple of 4, <= 4096", D);
  REQUIRE(aligned16(y) && (!add || aligned16(add)) && aligned16(gamma) && aligned16(beta), "layernorm: alignment");
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* lo = static_cast<__nv_bfloat16*>(out_lo);
  const long long rows_ll = rows;
  cudaError_t e;
  if (D <= 512)
    e = launch_k(layernorm_kernel<4>, (unsigned)blocks, 256u, 0, st, rows_ll, y, add, add_mod, gamma, beta, eps, rows_ll, D, out, hi, lo, m_dev);
  else if (D <= 1024)
    e = launch_k(layernorm_kernel<8>, (unsigned)blocks, 256u, 0, st, rows_ll, y, add, add_mod, gamma, beta, eps, rows_ll, D, out, hi, lo, m_dev);
  else
    e = launch_k(layernorm_kernel<32>, (unsigned)blocks, 256u, 0, st, rows_ll, y, add, add_mod, gamma, beta, eps, rows_ll, D, out, hi, lo, m_dev);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "layernorm launch: %s", cudaGetErrorString(e));
  return launch_check();
}

int lamp_embed(const int64_t* seq, const int64_t* pos, const float* word_emb, const float* pos_emb, int64_t rows,
               int D, float* out, void* out_hi, void* out_lo, const int64_t* row_index, const int32_t* m_dev,
               void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(seq && word_emb && (out || out_hi), "embed: null pointer");
  REQUIRE(!pos_emb || pos, "embed: pos_emb without pos ids");
  REQUIRE(D % 4 == 0, "embed: D must be a multiple of 4");
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  cudaError_t e = launch_k(embed_kernel, (unsigned)blocks, 256u, 0, (cudaStream_t)stream, (long long)rows,
                           reinterpret_cast<const long long*>(seq), reinterpret_cast<const long long*>(pos), word_emb, pos_emb,
                           (long long)rows, D, out, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo),
                           reinterpret_cast<const long long*>(row_index), m_dev);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "embed launch: %s", cudaGetErrorString(e));
  return launch_check();
}

int lamp_embed_bwd(const float* g, const int64_t* seq, const int64_t* pos, int64_t rows, int D, int64_t pad_word,
                   int64_t pad_pos, float* dword, float* dpos, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(g && (dword || dpos), "embed_bwd: null pointer");
  REQUIRE((!dword || seq) && (!dpos || pos), "embed_bwd: a gradient table without its ids");
  REQUIRE(D % 4 == 0 && aligned16(g) && (!dword || aligned16(dword)) && (!dpos || aligned16(dpos)),
          "embed_bwd: D multiple of 4 and 16-byte alignment required");
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  embed_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      g, reinterpret_cast<const long long*>(seq), reinterpret_cast<const long long*>(pos), (long long)rows, D,
      (long long)pad_word, (long long)pad_pos, dword, dpos);
  return launch_check();
}

int lamp_gather_rows(const float* src, const int64_t* index, int64_t rows, int D, float* out, void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(src && index && out, "gather_rows: null pointer");
  REQUIRE(D % 4 == 0 && aligned16(src) && aligned16(out), "gather_rows: D multiple of 4 and 16-byte alignment required");
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  cudaError_t e = launch_k(gather_rows_kernel, (unsigned)blocks, 256u, 0, (cudaStream_t)stream, (long long)rows, src,
                           reinterpret_cast<const long long*>(index), (long long)rows, D, out);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "gather_rows launch: %s", cudaGetErrorString(e));
  return launch_check();
}

int lamp_zero_guard_rows(void* hi, void* lo, int64_t ld, int cols, const int32_t* m_dev, int64_t max_rows, int nguard,
                         void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(hi && m_dev && nguard > 0 && cols > 0, "zero_guard_rows: bad arguments");
  zero_guard_rows_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(static_cast<__nv_bfloat16*>(hi),
                                                               static_cast<__nv_bfloat16*>(lo), ld, cols, m_dev,
                                                               max_rows, nguard);
  return launch_check();
}

int lamp_diag_proj(const float* x, const float* W, const float* bias, int64_t B, int L, int D, float* logits,
                   void* stream) {
  if (int rc = arch_check()) return rc;
  REQUIRE(x && W && logits, "diag_proj: null pointer");
  REQUIRE(D % 4 == 0 && aligned16(x) && aligned16(W), "diag_proj: D multiple of 4 and 16-byte alignment required");
  const long long rows = B * L;
  if (rows == 0) return LAMP_OK;
  const long long blocks = (rows * 32 + 255) / 256;
  cudaError_t e = launch_k(diag_proj_kernel, (unsigned)blocks, 256u, 0, (cudaStream_t)stream, rows, x, W, bias, rows, L, D, logits);
  if (e != cudaSuccess) return fail(LAMP_ECUDA, "diag_proj launch: %s", cudaGetErrorString(e));
  return launch_check();
}

// ------------------------------------------------------------------------------------ level 2

size_t lamp_sdpa_workspace_bytes(int N, int Lq, int Lk, int d) {
  Carver c(nullptr);
  c.take((size_t)N * Lq * d * 4);      // Q planes
  c.take((size_t)N * Lk * 2 * d * 4);  // [K | V] planes
  c.take((size_t)N * Lq * 8);          // row statistics
  return c.off;
}

static int sdpa_impl(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                     int64_t msk, float* out, float* attn, float* probs_pre, int N, int Lq, int Lk, int d,
                     float temperature, int precision, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* workspace, size_t workspace_bytes, void* stream);

int lamp_sdpa_fwd(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                  int64_t msk, float* out, float* attn, int N, int Lq, int Lk, int d, float temperature,
                  int precision, void* workspace, size_t workspace_bytes, void* stream) {
  return sdpa_impl(q, k, v, mask, msb, msq, msk, out, attn, nullptr, N, Lq, Lk, d, temperature, precision, 0.0f, 0,
                   nullptr, workspace, workspace_bytes, stream);
}

int lamp_sdpa_fwd_train(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                        int64_t msk, float* out, float* attn, float* probs_pre, int N, int Lq, int Lk, int d,
                        float temperature, int precision, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                        void* workspace, size_t workspace_bytes, void* stream) {
  REQUIRE(attn != nullptr, "sdpa_train: the attention map is part of the training forward");
  REQUIRE(p_drop == 0.0f || probs_pre != nullptr, "sdpa_train: probs_pre is required when dropout is active");
  return sdpa_impl(q, k, v, mask, msb, msq, msk, out, attn, probs_pre, N, Lq, Lk, d, temperature, precision, p_drop,
                   seed, seed_dev, workspace, workspace_bytes, stream);
}

static int sdpa_impl(const float* q, const float* k, const float* v, const uint8_t* mask, int64_t msb, int64_t msq,
                     int64_t msk, float* out, float* attn, float* probs_pre, int N, int Lq, int Lk, int d,
                     float temperature, int precision, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     void* workspace, size_t workspace_bytes, void* stream) {
  REQUIRE(q && k && v && out, "sdpa: null pointer");
  REQUIRE(N >= 0 && Lq > 0 && Lk > 0 && d > 0, "sdpa: bad shape");
  if (!workspace || workspace_bytes < lamp_sdpa_workspace_bytes(N, Lq, Lk, d))
    return fail(LAMP_EWORKSPACE, "sdpa: workspace too small");
  const bool three = precision == LAMP_PREC_FP32;
  Carver c(workspace);
  const size_t qn = (size_t)N * Lq * d, kvn = (size_t)N * Lk * 2 * d;
  __nv_bfloat16* qp = static_cast<__nv_bfloat16*>(c.take(qn * 4));
  __nv_bfloat16* kvp = static_cast<__nv_bfloat16*>(c.take(kvn * 4));
  float* stats = static_cast<float*>(c.take((size_t)N * Lq * 8));
  __nv_bfloat16* qlo = three ? qp + qn : nullptr;
  __nv_bfloat16* kvlo = three ? kvp + kvn : nullptr;
  if (int rc = lamp_split_planes(q, (int64_t)N * Lq, d, d, qp, qlo, d, stream)) return rc;
  // K and V side by side in one [N*Lk, 2d] matrix so that a single tensor map serves both operands
  if (int rc = lamp_split_planes(k, (int64_t)N * Lk, d, d, kvp, kvlo, 2 * d, stream)) return rc;
  if (int rc = lamp_split_planes(v, (int64_t)N * Lk, d, d, kvp + d, three ? kvlo + d : nullptr, 2 * d, stream)) return rc;
  return attn_impl(qp, qlo, d, 0, 0, kvp, kvlo, 2 * d, 0, d, N, 1, Lq, Lk, d, temperature, precision, mask, msb, msq, msk,
                   nullptr, 0, 0, nullptr, nullptr, 0, out, d, stats, stats + (size_t)N * Lq, attn, nullptr, nullptr, 0,
                   stream, p_drop, seed, probs_pre, seed_dev);
}

namespace {
struct MhaPlan {
  size_t xq, xkv, wqkv, wfc, qkv, kv, o, y, stats, total;
};
MhaPlan mha_plan(int B, int Lq, int Lk, int D, int H, int d, int self_attn, int want_attn) {
  const size_t hd = (size_t)H * d, mq = (size_t)B * Lq, mk = (size_t)B * Lk;
  Carver c(nullptr);
  MhaPlan p{};
  p.xq = c.off;   c.take(mq * D * 4);
  p.xkv = c.off;  c.take(self_attn ? 0 : mk * D * 4);
  p.wqkv = c.off; c.take(3 * hd * D * 4);
  p.wfc = c.off;  c.take(H > 1 ? hd * D * 4 : 0);
  p.qkv = c.off;  c.take(self_attn ? mq * 3 * hd * 4 : mq * hd * 4);
  p.kv = c.off;   c.take(self_attn ? 0 : mk * 2 * hd * 4);
  p.o = c.off;    c.take(mq * hd * 4);
  p.y = c.off;    c.take(mq * (size_t)(D > (int)hd ? D : hd) * 4);
  p.stats = c.off; c.take(want_attn ? (size_t)H * mq * 8 : 0);
  p.total = c.off;
  return p;
}
}  // namespace

size_t lamp_mha_workspace_bytes(int B, int Lq, int Lk, int D, int H, int d, int self_attn, int want_attn) {
  return mha_plan(B, Lq, Lk, D, H, d, self_attn, want_attn).total;
}

int lamp_mha_fwd(const float* q, const float* kv, const float* Wq, const float* Wk, const float* Wv,
                 const float* Wfc, const float* ln_w, const float* ln_b, const uint8_t* mask, int64_t msb,
                 int64_t msq, int64_t msk, float* out, float* attn, int B, int Lq, int Lk, int D, int H, int d,
                 int precision, float ln_eps, void* workspace, size_t workspace_bytes, void* stream) {
  REQUIRE(q && Wq && Wk && Wv && ln_w && ln_b && out, "mha: null pointer");
  REQUIRE((H > 1) == (Wfc != nullptr), "mha: fc weight must be given iff n_head > 1 (lamp/SubLayers.py:72-74)");
  REQUIRE(H == 1 ? d == D : true, "mha: with n_head == 1 the head width must equal d_model (no fc)");
  REQUIRE(D % 8 == 0, "mha: d_model must be a multiple of 8");
  const int self_attn = (kv == nullptr || kv == q);
  REQUIRE(!self_attn || Lq == Lk, "mha: self-attention needs Lq == Lk");
  const MhaPlan pl = mha_plan(B, Lq, Lk, D, H, d, self_attn, attn != nullptr);
  if (!workspace || workspace_bytes < pl.total) return fail(LAMP_EWORKSPACE, "mha: workspace too small");
  const bool three = precision == LAMP_PREC_FP32;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t hd = (size_t)H * d, mq = (size_t)B * Lq, mk = (size_t)B * Lk;
  auto hi = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto lo = [&](size_t off, size_t n) { return three ? reinterpret_cast<__nv_bfloat16*>(ws + off) + n : nullptr; };
  // 1. operand planes: activations and (stateless API) weights
  if (int rc = lamp_split_planes(q, mq, D, D, hi(pl.xq), lo(pl.xq, mq * D), D, stream)) return rc;
  if (!self_attn)
    if (int rc = lamp_split_planes(kv, mk, D, D, hi(pl.xkv), lo(pl.xkv, mk * D), D, stream)) return rc;
  const float* ws_in[3] = {Wq, Wk, Wv};
  for (int i = 0; i < 3; ++i)
    if (int rc = lamp_split_planes(ws_in[i], hd, D, D, hi(pl.wqkv) + i * hd * D, three ? lo(pl.wqkv, 3 * hd * D) + i * hd * D : nullptr, D, stream)) return rc;
  if (H > 1)
    if (int rc = lamp_split_planes(Wfc, D, (int)hd, hd, hi(pl.wfc), lo(pl.wfc, hd * D), hd, stream)) return rc;
  // 2. projections (lamp/SubLayers.py:91-93) -> planes
  const __nv_bfloat16 *qh, *ql, *kvh, *kvl;
  int64_t ldq, ldkv;
  int k_col0, v_col0;
  if (self_attn) {
    if (int rc = lamp_gemm_planes(hi(pl.xq), lo(pl.xq, mq * D), D, hi(pl.wqkv), lo(pl.wqkv, 3 * hd * D), D, (int)mq, (int)(3 * hd), D, precision,
                                  nullptr, 0, nullptr, 0, 0, nullptr, 0, hi(pl.qkv), lo(pl.qkv, mq * 3 * hd), 3 * hd, nullptr, stream)) return rc;
    qh = kvh = hi(pl.qkv); ql = kvl = lo(pl.qkv, mq * 3 * hd);
    ldq = ldkv = 3 * hd; k_col0 = (int)hd; v_col0 = (int)(2 * hd);
  } else {
    if (int rc = lamp_gemm_planes(hi(pl.xq), lo(pl.xq, mq * D), D, hi(pl.wqkv), lo(pl.wqkv, 3 * hd * D), D, (int)mq, (int)hd, D, precision,
                                  nullptr, 0, nullptr, 0, 0, nullptr, 0, hi(pl.qkv), lo(pl.qkv, mq * hd), hd, nullptr, stream)) return rc;
    if (int rc = lamp_gemm_planes(hi(pl.xkv), lo(pl.xkv, mk * D), D, hi(pl.wqkv) + hd * D, three ? lo(pl.wqkv, 3 * hd * D) + hd * D : nullptr, D,
                                  (int)mk, (int)(2 * hd), D, precision, nullptr, 0, nullptr, 0, 0, nullptr, 0, hi(pl.kv), lo(pl.kv, mk * 2 * hd), 2 * hd, nullptr, stream)) return rc;
    qh = hi(pl.qkv); ql = lo(pl.qkv, mq * hd); kvh = hi(pl.kv); kvl = lo(pl.kv, mk * 2 * hd);
    ldq = hd; ldkv = 2 * hd; k_col0 = 0; v_col0 = (int)hd;
  }
  // 3. masked softmax attention (lamp/SubLayers.py:104 -> :27-43), temperature sqrt(d_k) (:63/:65)
  float* y = reinterpret_cast<float*>(ws + pl.y);
  float* stats = attn ? reinterpret_cast<float*>(ws + pl.stats) : nullptr;
  const float temperature = sqrtf((float)d);
  if (int rc = lamp_attn_core_planes(qh, ql, ldq, 0, 0, kvh, kvl, ldkv, k_col0, v_col0, B, H, Lq, Lk, d, temperature, precision, mask, msb,
                                     msq, msk, H > 1 ? hi(pl.o) : nullptr, H > 1 ? lo(pl.o, mq * hd) : nullptr, hd, H > 1 ? nullptr : y, hd,
                                     stats, stats ? stats + (size_t)H * mq : nullptr, attn, nullptr, nullptr, 0, stream)) return rc;
  // 4. fc + residual (:110,:117) then LayerNorm
  if (H > 1) {
    if (int rc = lamp_gemm_planes(hi(pl.o), lo(pl.o, mq * hd), hd, hi(pl.wfc), lo(pl.wfc, hd * D), hd, (int)mq, D, (int)hd, precision, nullptr, 0,
                                  q, D, 0, y, D, nullptr, nullptr, 0, nullptr, stream)) return rc;
    return lamp_layernorm(y, nullptr, 0, ln_w, ln_b, ln_eps, mq, D, out, nullptr, nullptr, nullptr, stream);
  }
  return lamp_layernorm(y, q, 0, ln_w, ln_b, ln_eps, mq, D, out, nullptr, nullptr, nullptr, stream);
}

size_t lamp_ffn_workspace_bytes(int64_t rows, int D, int d_inner) {
  Carver c(nullptr);
  c.take((size_t)rows * D * 4);        // x planes
  c.take((size_t)d_inner * D * 4);     // W1 planes
  c.take((size_t)d_inner * D * 4);     // W2 planes
  c.take((size_t)rows * d_inner * 4);  // hidden planes
  c.take((size_t)rows * D * 4);        // pre-LN fp32
  return c.off;
}

int lamp_ffn_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                 const float* ln_w, const float* ln_b, float* out, int64_t rows, int D, int d_inner, int precision,
                 float ln_eps, void* workspace, size_t workspace_bytes, void* stream) {
  REQUIRE(x && W1 && b1 && W2 && b2 && ln_w && ln_b && out, "ffn: null pointer");
  REQUIRE(D % 8 == 0 && d_inner % 8 == 0, "ffn: D and d_inner must be multiples of 8");
  if (!workspace || workspace_bytes < lamp_ffn_workspace_bytes(rows, D, d_inner))
    return fail(LAMP_EWORKSPACE, "ffn: workspace too small");
  const bool three = precision == LAMP_PREC_FP32;
  Carver c(workspace);
  const size_t nx = (size_t)rows * D, nw = (size_t)d_inner * D, nh = (size_t)rows * d_inner;
  __nv_bfloat16* xp = static_cast<__nv_bfloat16*>(c.take(nx * 4));
  __nv_bfloat16* w1p = static_cast<__nv_bfloat16*>(c.take(nw * 4));
  __nv_bfloat16* w2p = static_cast<__nv_bfloat16*>(c.take(nw * 4));
  __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(c.take(nh * 4));
  float* y = static_cast<float*>(c.take(nx * 4));
  auto lo = [&](__nv_bfloat16* p, size_t n) { return three ? p + n : nullptr; };
  if (int rc = lamp_split_planes(x, rows, D, D, xp, lo(xp, nx), D, stream)) return rc;
  if (int rc = lamp_split_planes(W1, d_inner, D, D, w1p, lo(w1p, nw), D, stream)) return rc;
  if (int rc = lamp_split_planes(W2, D, d_inner, d_inner, w2p, lo(w2p, nw), d_inner, stream)) return rc;
  // w_1 + ReLU (lamp/SubLayers.py:138), hidden kept as planes only
  if (int rc = lamp_gemm_planes(xp, lo(xp, nx), D, w1p, lo(w1p, nw), D, (int)rows, d_inner, D, precision, b1, 1, nullptr, 0, 0, nullptr, 0, hp,
                                lo(hp, nh), d_inner, nullptr, stream)) return rc;
  // w_2 + residual (:138-141)
  if (int rc = lamp_gemm_planes(hp, lo(hp, nh), d_inner, w2p, lo(w2p, nw), d_inner, (int)rows, D, d_inner, precision, b2, 0, x, D, 0, y, D, nullptr,
                                nullptr, 0, nullptr, stream)) return rc;
  return lamp_layernorm(y, nullptr, 0, ln_w, ln_b, ln_eps, rows, D, out, nullptr, nullptr, nullptr, stream);
}

}  // extern "C"
