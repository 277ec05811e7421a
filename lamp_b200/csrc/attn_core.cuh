// Masked attention core over label nodes:   O = softmax(mask(Q K^T / temperature)) V      (per sample, per head)
// Reference: lamp/SubLayers.py:27-43 (ScaledDotProductAttention.forward), called from :104.
//
// One persistent CTA per SM walks a sequence of "units" = (work item (b, h, 128-row q tile), KV tile j):
//   warp 0   : TMA producer -- 3D tensor maps {cols, L, B}: rows >= L are zero-filled by the hardware, so ragged
//              label counts (L = 103, 159, 983 ...) need no padding in HBM.  Q/K/V arrive as 128B-swizzled
//              [rows x 64] bf16 boxes of the split-bf16 planes written by the projection GEMM.  The `kv_slots` equal
//              slots are split into a K ring (first kv_slots/2 slots) and a V ring (the rest), each slot with its own
//              full/empty barrier; warp 0 feeds Q and the K ring, warp 3 the V ring.  A K slot is free as soon as its
//              S = Q K^T retires (early in a unit), a V slot only after its PV product (late), so with separate rings
//              and producers the K tiles run ahead by the ring depth instead of queueing behind the V loads -- with
//              multi-tile rows (L = 983) the measured load-to-use latency is ~4400 clocks, i.e. 1.5 unit periods.
//   warp 1   : S issuer     -- one thread issues S(u) = Q K_u^T as soon as its Q / K tiles have landed and a TMEM score
//              buffer is free.
//   warp 2   : PV issuer    -- one thread issues O (+)= P(u) V_u as soon as P(u) is published and V_u has landed.  Two
//              issuing threads, not one: a tcgen05.mma issue blocks its thread until the tensor pipe accepts it, i.e.
//              for about the execution time of the queue ahead; a single event-driven thread therefore serialised
//              S(u+1) -> PV(u) -> S(u+2) ... with its polling latency exposed between any two (tensor pipe 61 % busy
//              in the trace); with two, PV(u) is queued while S(u+1) still executes.  S goes to one of the TMEM score
//              buffers, O to one of two TMEM output buffers (item parity), so neither the next item's PV nor the next S
//              ever waits for the epilogue.  P is read from TENSOR MEMORY (TS form), V is consumed MN-major straight
//              from its natural [keys x d] layout (no transpose pass); 3-term split-bf16 products.
//   warp 3   : TMA producer of the V ring.
//   warps 5.. : softmax     -- 4 * (BLOCK_KV/32) warps: TMEM lane == query row and every warp owns one 32-column
//              chunk of the rows of its lane quarter.  Strided byte mask -> bit words via warp ballots, row max
//              exchanged through smem, online max/sum across KV tiles with a LAZY rescale (the reference maximum only
//              moves when it grew by more than 2^8, so the O accumulator is rarely touched and PV(u) does not
//              serialise behind PV(u-1)), ex2, hi/lo split, P written back over S in TMEM.  The epilogue of the
//              PREVIOUS item (1/sum, hi/lo split) runs after the current unit's P is published; it writes the tile
//              in the 128B-swizzled box layout into a staging buffer from which the first softmax thread stores it
//              with TMA once a named barrier has collected the tile (coalesced, asynchronous, rows >= L clipped by
//              the tensor map); the same thread waits for the store to have READ the tile before the next epilogue
//              overwrites it.  (A dedicated store warp used to do this; its slot now belongs to the PV issuer.)
// TMEM columns: S0/P0 [0,128) | S1/P1 [128,256) | O0 [256,384) | O1 [384,512).  64-key tiles (BLOCK_KV == 64: the
// multi-tile label<-input / L = 983 shapes at d = 128) need only 64 score columns per buffer and use THREE of them,
// S0/P0 [0,64) | S1/P1 [64,128) | S2/P2 [128,192): with two, S(u+2) had to wait until PV(u) had retired (P(u) lives in
// its score buffer), which serialised the tensor pipe into S(u+1) -> PV(u) -> S(u+2) -> ... with the issue / completion
// latencies exposed at every step (period 5.7 K clocks per unit for 3.6 K of tensor work); with three, S runs two
// units ahead and the pipe only ever waits for operands.  P overwrites S in place: the warp
// that owns a 32-column chunk of the scores holds them in registers and writes the chunk's hi (16 columns) and lo
// (16 columns) bf16 pairs back over the same 32 columns.
// Shared memory: Q tile | O staging tile (same shape) | ring of K/V slots; tile rows follow the label count
// (box rows = L rounded up to 8 / 16), e.g. L = 103: Q 52 KB + O 52 KB + 2 x 56 KB.
// The label mask is read from its single [Lq, Lk] (or [B, Lk] key-padding) copy -- it is never tiled per head or
// per sample (the reference materialises H*B*Lq*Lk bytes, lamp/SubLayers.py:102 / lamp/Decoders.py:141).
#pragma once
#include "sm100_primitives.cuh"

namespace lamp {

struct AttnParams {
  int B, H, Lq, Lk, d;  // d = head width, multiple of 16, <= 128
  float scale_log2;     // log2(e) / temperature
  int q_col0, k_col0, v_col0;  // column of head 0 inside the Q / KV plane matrices
  int q_bcast;                 // 1: Q is shared by every sample (batch coordinate forced to 0)
  const uint8_t* mask;         // nullptr or bytes (non-zero = masked), element strides below
  long long msb, msq, msk;
  // alternative mask form: bit-packed words (bit k & 31 of word [b*mbb + q*mbq + (k >> 5)] set = masked), produced by
  // lamp_pack_mask_bits.  One 4-byte load per thread and KV tile instead of 32 byte loads + ballots: the form used
  // whenever a [.., Lq, Lk] mask meets more than one KV tile (L = 983: the byte path cost 4x the tile's MMAs).
  const uint32_t* mask_bits;
  long long mbb, mbq;
  __nv_bfloat16* o_hi;  // [B*Lq, ldo] planes, head h at column h*d (nullable)
  __nv_bfloat16* o_lo;  // nullable
  int ldo;
  float* o_f32;  // optional fp32 copy, [B*Lq, ldof]
  int ldof;
  float* row_max;  // optional [H*B*Lq] (head-major) reference maximum of the SCALED (log2 domain) scores
  float* row_sum;  // optional [H*B*Lq] softmax denominators (relative to row_max)
  int qrows, krows, vrows;  // TMA box rows of the Q / K / V tiles (multiples of 8 / 8 / 16)
  // Padding-aware keys (optional): the K/V matrix holds only the non-PAD tokens of the batch, packed; sample b owns
  // rows [kv_start[b], kv_start[b] + kv_len[b]).  Lk then only bounds the per-sample key count.
  const int* kv_start;
  const int* kv_len;
  int kv_slots;         // K/V ring depth (2 .. ATTN_MAX_SLOTS)
  uint32_t slot_bytes;  // bytes of one ring slot (fits a K tile and a V tile)
  int staged;           // 1: O planes leave through the smem staging tile + TMA store (needs d % 64 == 0)
  int pv_split;         // 1 (d == 128 only): O (+)= P V issued as two interleaved N = 64 accumulation chains
  // training forward: dropout on the probabilities (lamp/SubLayers.py:40).  keep(n, i, j) is a pure function of
  // (drop_seed, head-major row (h*B + b)*Lq + i, key j), so the probability kernel reproduces the same kept set.
  uint32_t drop_thresh;  // 0: off; else an element is dropped iff hash < drop_thresh (= p * 2^32)
  float drop_scale;      // 1 / (1 - p)
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;  // optional device-side counter ADDED to drop_seed (CUDA-graph replays of a
                                            // training step: the host seed is baked into the graph, the counter moves)
};

#ifdef LAMP_ATTN_TRACE
// Debug build only (scripts/attn_trace.py): clock64() stamps of CTA 0's pipeline events, [event][unit].
__device__ unsigned long long g_attn_trace[16 + 48][64];  // rows 16..: per softmax warp (16 warps x {S seen, max bar, P stored})
#define ATTN_TRACE(ev, idx)                                                              \
  do {                                                                                   \
    if (blockIdx.x == 0 && (idx) < 64u) g_attn_trace[ev][idx] = clock64();                \
  } while (0)
#define ATTN_TRACE_W(k, idx)                                                                              \
  do {                                                                                                    \
    if (blockIdx.x == 0 && lane == 0 && (idx) < 64u) g_attn_trace[16 + (warp - LAMP_TRACE_SW0) * 3 + (k)][idx] = clock64(); \
  } while (0)
#else
#define ATTN_TRACE(ev, idx) do { } while (0)
#define ATTN_TRACE_W(k, idx) do { } while (0)
#endif

constexpr int ATTN_BLOCK_M = 128;
constexpr uint32_t ATTN_TMEM_COLS = 512;
constexpr uint32_t ATTN_TMEM_S = 0, ATTN_TMEM_O = 256;
constexpr int ATTN_MAX_SLOTS = 6;
constexpr float ATTN_RESCALE_LOG2 = 8.0f;  // lazy rescale: P stays below 2^8, harmless for fp32 accumulation
// row-statistics exchange between the column-warps of a row: max [2 units][NW][128] + sum [2 items][NW][128] floats
constexpr uint32_t ATTN_RED_BYTES = 2 * 2 * 4 * 128 * 4;
constexpr uint32_t ATTN_BAR_BYTES = 256;
// warps 0..3: Q/K TMA producer, S issuer, PV issuer, V TMA producer; warps 4..: softmax (warp % 4 = TMEM lane quarter)
// 4 control warps (Q/K producer, S issuer, PV issuer, V producer) + the softmax warps.
__host__ __device__ constexpr int attn_ctrl_warps(int) { return 4; }
__host__ __device__ constexpr int attn_threads(int block_kv) { return 32 * attn_ctrl_warps(block_kv) + 128 * (block_kv / 32); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Shared-memory plan (bytes).  kb64 = ceil(d / 64) column blocks of 64 bf16 (= one 128 B swizzle row each).
__host__ __device__ constexpr uint32_t attn_tile_bytes(int npl, int kb64, int rows) { return npl * kb64 * rows * 128; }
__host__ __device__ constexpr uint32_t attn_smem_bytes(uint32_t q_bytes, uint32_t slot_bytes, int slots, int staged) {
  return q_bytes * (staged ? 2 : 1) + slots * slot_bytes + ATTN_RED_BYTES + 1024 /*align*/ + ATTN_BAR_BYTES;
}

// DROP: training forward (dropout on the probabilities); a compile-time switch so that the inference kernel carries
// none of it (as a run-time branch inside the exp loop it cost the d = 64 / bf16 shapes 29 %).
template <int BLOCK_KV, int NTERMS, bool DROP>
__global__ void __launch_bounds__(attn_threads(BLOCK_KV), 1)
attn_core_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                 const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                 const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo,
                 const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo,
                 const AttnParams p) {
  constexpr int NPL = (NTERMS == 3) ? 2 : 1;
  constexpr int NW = BLOCK_KV / 32;        // 32-column score chunks per row == softmax warps per lane quarter
  constexpr int NSW = 4 * NW * 32;         // softmax threads
  constexpr uint32_t NSB = (BLOCK_KV == 64) ? 3u : 2u;     // score / P buffers in TMEM (see the TMEM map above)
  constexpr int SW0 = attn_ctrl_warps(BLOCK_KV);           // first softmax warp
#ifdef LAMP_ATTN_TRACE
  const int LAMP_TRACE_SW0 = SW0;
#endif
  constexpr uint32_t S_STRIDE = (BLOCK_KV == 64) ? 64u : 128u;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb64 = (p.d + 63) >> 6;
  const int R = p.kv_slots;
  const int RK = R >> 1, RV = R - RK;  // K ring: slots [0, RK), V ring: slots [RK, R)
  const uint32_t q_bytes = attn_tile_bytes(NPL, kb64, p.qrows);
  const uint32_t k_bytes = attn_tile_bytes(NPL, kb64, p.krows);
  const uint32_t v_bytes = attn_tile_bytes(NPL, kb64, p.vrows);
  uint8_t* sQ = smem;
  uint8_t* sO = sQ + q_bytes;                          // staging tile, same geometry as Q (absent if !staged)
  uint8_t* sKV = sO + (p.staged ? q_bytes : 0);        // ring of R slots
  float* red_max = reinterpret_cast<float*>(sKV + R * p.slot_bytes);  // [2][4][128]
  float* red_l = red_max + 2 * 4 * 128;                               // [2][4][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red_max) + ATTN_RED_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* s_full = bars + 2;     // [NSB <= 3] S(u) complete (tcgen05.commit), buffer u % NSB
  uint64_t* s_free = bars + 5;     // [NSB] score/P buffer released by the PV product that consumed it
  uint64_t* p_full = bars + 8;     // [NSB] P(u) published by the NSW softmax threads
  uint64_t* pv_done = bars + 11;   // [2] PV(u) complete, indexed by unit parity
  uint64_t* o_free = bars + 13;    // [2] O buffer (item parity) drained by the epilogue, count NSW
  uint64_t* ostage_full = bars + 15;   // staging tile written, count NSW
  uint64_t* ostage_free = bars + 16;   // TMA store has read the staging tile
  uint64_t* kv_full = bars + 17;   // [ATTN_MAX_SLOTS]
  uint64_t* kv_empty = bars + 17 + ATTN_MAX_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17 + 2 * ATTN_MAX_SLOTS);
  static_assert((17 + 2 * ATTN_MAX_SLOTS) * 8 + 4 <= ATTN_BAR_BYTES, "barrier area too small");

  // plane pl (0 = hi, 1 = lo), 64-column block kb
  auto q_tile = [&](int pl, int kb) { return sQ + (pl * kb64 + kb) * (p.qrows * 128); };
  auto o_tile = [&](int pl, int kb) { return sO + (pl * kb64 + kb) * (p.qrows * 128); };
  auto k_tile = [&](int slot, int pl, int kb) { return sKV + slot * p.slot_bytes + (pl * kb64 + kb) * (p.krows * 128); };
  auto v_tile = [&](int slot, int pl, int kb) { return sKV + slot * p.slot_bytes + (pl * kb64 + kb) * (p.vrows * 128); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ_hi);
    tma_prefetch_desc(&tmK_hi);
    tma_prefetch_desc(&tmV_hi);
    if (NPL == 2) {
      tma_prefetch_desc(&tmQ_lo);
      tma_prefetch_desc(&tmK_lo);
      tma_prefetch_desc(&tmV_lo);
    }
    if (p.staged) {
      tma_prefetch_desc(&tmO_hi);
      if (NPL == 2) tma_prefetch_desc(&tmO_lo);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 1);
      mbar_init(&p_full[i], NSW);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_free[i], NSW);
    }
    mbar_init(ostage_full, NSW);
    mbar_init(ostage_free, 1);
    for (int i = 0; i < ATTN_MAX_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATTN_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch_dependents();
  griddep_wait();  // PDL: nothing above touches global memory

  const int num_qt = (p.Lq + ATTN_BLOCK_M - 1) / ATTN_BLOCK_M;
  const int num_kv = (p.Lk + BLOCK_KV - 1) / BLOCK_KV;  // upper bound; per item: item_keys()
  const int num_items = p.B * p.H * num_qt;
  const int ksteps_d = p.d >> 4;  // UMMA K-steps over the head width
  const bool varlen = p.kv_len != nullptr;
  // keys of sample b: (count, first row in the K/V matrix, batch coordinate of the tensor map, KV tiles >= 1)
  auto item_keys = [&](int b, int& lk, int& kbase, int& kbatch, int& nkv) {
    lk = varlen ? min(__ldg(p.kv_len + b), p.Lk) : p.Lk;
    kbase = varlen ? __ldg(p.kv_start + b) : 0;
    kbatch = varlen ? 0 : b;
    nkv = (lk + BLOCK_KV - 1) / BLOCK_KV;
    if (nkv < 1) nkv = 1;  // a sample without keys still produces (NaN) rows, like a fully masked reference row
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0, u = 0;  // per-CTA item / unit counters -> ring positions and barrier parities
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qt = item % num_qt;
        const int h = (item / num_qt) % p.H;
        const int b = item / (num_qt * p.H);
        int lk, kbase, kbatch, nkv;
        item_keys(b, lk, kbase, kbatch, nkv);
        for (int j = 0; j < nkv; ++j, ++u) {
          const int ks = u % RK;
          const int krow = kbase + j * BLOCK_KV;
          mbar_wait(&kv_empty[ks], ((u / RK) & 1) ^ 1);
          ATTN_TRACE(0, u);
          mbar_arrive_expect_tx(&kv_full[ks], k_bytes);
          for (int kb = 0; kb < kb64; ++kb) {
            tma_load_3d(k_tile(ks, 0, kb), &tmK_hi, &kv_full[ks], p.k_col0 + h * p.d + kb * 64, krow, kbatch);
            if (NPL == 2)
              tma_load_3d(k_tile(ks, 1, kb), &tmK_lo, &kv_full[ks], p.k_col0 + h * p.d + kb * 64, krow, kbatch);
          }
          if (j == 0) {
            mbar_wait(q_empty, (it & 1) ^ 1);
            ATTN_TRACE(1, u);
            mbar_arrive_expect_tx(q_full, q_bytes);
            const int bq = p.q_bcast ? 0 : b;
            for (int kb = 0; kb < kb64; ++kb) {
              tma_load_3d(q_tile(0, kb), &tmQ_hi, q_full, p.q_col0 + h * p.d + kb * 64, qt * ATTN_BLOCK_M, bq);
              if (NPL == 2)
                tma_load_3d(q_tile(1, kb), &tmQ_lo, q_full, p.q_col0 + h * p.d + kb * 64, qt * ATTN_BLOCK_M, bq);
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ---------------------------------------------------------------- TMA producer of the V ring
    if (lane == 0) {
      uint32_t u = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int h = (item / num_qt) % p.H;
        const int b = item / (num_qt * p.H);
        int lk, kbase, kbatch, nkv;
        item_keys(b, lk, kbase, kbatch, nkv);
        for (int j = 0; j < nkv; ++j, ++u) {
          const int vs = RK + u % RV;
          const int krow = kbase + j * BLOCK_KV;
          mbar_wait(&kv_empty[vs], ((u / RV) & 1) ^ 1);
          ATTN_TRACE(2, u);
          mbar_arrive_expect_tx(&kv_full[vs], v_bytes);
          for (int kb = 0; kb < kb64; ++kb) {
            tma_load_3d(v_tile(vs, 0, kb), &tmV_hi, &kv_full[vs], p.v_col0 + h * p.d + kb * 64, krow, kbatch);
            if (NPL == 2)
              tma_load_3d(v_tile(vs, 1, kb), &tmV_lo, &kv_full[vs], p.v_col0 + h * p.d + kb * 64, krow, kbatch);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ---------------------------------------------------------------- MMA issuers (warp 1: S, warp 2: PV)
    if (lane == 0) {
      const uint32_t idesc_o = umma_idesc_bf16(ATTN_BLOCK_M, p.d, 0, 1);  // A = P from TMEM, B = V MN-major
      struct UnitIt {
        int item, j, nkv, lk;
        uint32_t it, u;  // per-CTA item / unit counters
      };
      auto load_it = [&](UnitIt& x) {
        if (x.item < num_items) {
          int kbase, kbatch;
          item_keys(x.item / (num_qt * p.H), x.lk, kbase, kbatch, x.nkv);
        }
      };
      auto advance = [&](UnitIt& x) {
        ++x.u;
        if (++x.j == x.nkv) {
          x.j = 0;
          x.item += gridDim.x;
          ++x.it;
          load_it(x);
        }
      };

      // Shared-memory descriptors are built ONCE: inside the issue loops a descriptor is the base plus a small offset in
      // 16-byte units (the start-address field occupies the low 14 bits; all offsets stay below 256 KB).  The issuing
      // thread is the serial resource of this kernel -- 36 tcgen05.mma per 64-key unit -- and rebuilding four
      // descriptors per K-step from pointers cost it ~95 clocks per instruction against the ~55-64 the tensor pipe
      // needs (scripts/probes/umma_probe.cu).
      const uint64_t dQ0 = umma_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dK0 = umma_smem_desc(smem_u32(sKV), 16, 1024);
      const uint64_t dV0 = umma_smem_desc(smem_u32(sKV), p.vrows * 128, 1024);
      const uint32_t q_kb16 = (p.qrows * 128) >> 4, q_pl16 = kb64 * q_kb16;   // 64-column block / plane strides
      const uint32_t k_kb16 = (p.krows * 128) >> 4, k_pl16 = kb64 * k_kb16;
      const uint32_t v_pl16 = kb64 * ((p.vrows * 128) >> 4);
      const uint32_t slot16 = p.slot_bytes >> 4;

      // S(u) = Q K_u^T into score buffer u & 1 (N = the tile's key count rounded up to 16)
      auto issue_s = [&](const UnitIt& x) {
        const uint32_t u = x.u;
        const int ks = u % RK;
        const uint32_t sb = u % NSB;
        int kvn = (min(BLOCK_KV, x.lk - x.j * BLOCK_KV) + 15) & ~15;
        if (kvn < 16) kvn = 16;
        const uint32_t idesc_s = umma_idesc_bf16(ATTN_BLOCK_M, kvn, 0, 0);
        ATTN_TRACE(4, u);
        tcgen05_fence_after();
        const uint32_t tS = tmem_base + ATTN_TMEM_S + sb * S_STRIDE;
        const uint64_t dKs = dK0 + static_cast<uint32_t>(ks) * slot16;
        for (int t = 0; t < ksteps_d; ++t) {
          const uint32_t kb = t >> 2, koff16 = (t & 3) * 2;  // 16 bf16 = 32 B inside the swizzled 128 B row
          const uint64_t dq_hi = dQ0 + (kb * q_kb16 + koff16);
          const uint64_t dk_hi = dKs + (kb * k_kb16 + koff16);
          umma_bf16_ss(tS, dq_hi, dk_hi, idesc_s, t != 0 ? 1u : 0u);
          if (NTERMS == 3) {
            umma_bf16_ss(tS, dq_hi, dk_hi + k_pl16, idesc_s, 1u);
            umma_bf16_ss(tS, dq_hi + q_pl16, dk_hi, idesc_s, 1u);
          }
        }
        umma_commit(&s_full[sb]);
        umma_commit(&kv_empty[ks]);
        if (x.j == x.nkv - 1) umma_commit(q_empty);
        ATTN_TRACE(3, u);
      };

      // O(item parity) (+)= P(u) V_u
      auto issue_pv = [&](const UnitIt& x) {
        const uint32_t u = x.u;
        const int vs = RK + u % RV;
        const uint32_t sb = u % NSB;
        ATTN_TRACE(6, u);
        tcgen05_fence_after();
        const int kv_valid = max(0, min(BLOCK_KV, x.lk - x.j * BLOCK_KV));
        const int ksteps_kv = max(1, (kv_valid + 15) >> 4);
        const uint32_t tO = tmem_base + ATTN_TMEM_O + (x.it & 1) * 128;
        const uint32_t tP = tmem_base + ATTN_TMEM_S + sb * S_STRIDE;  // P lives where S(u) was
        if (p.pv_split) {
          // two independent N = 64 chains (the two 64-wide d blocks of V, O columns [0,64) / [64,128)), interleaved
          const uint32_t idesc_h = umma_idesc_bf16(ATTN_BLOCK_M, 64, 0, 1);
          for (int t = 0; t < ksteps_kv; ++t) {
            const uint32_t pa = tP + 32 * (t >> 1) + 8 * (t & 1);
            const uint32_t accum = (x.j != 0 || t != 0) ? 1u : 0u;
            uint64_t dvh[2], dvl[2];
            for (int kb = 0; kb < 2; ++kb) {
              dvh[kb] = umma_smem_desc(smem_u32(v_tile(vs, 0, kb)) + t * 2048, p.vrows * 128, 1024);
              dvl[kb] = umma_smem_desc(smem_u32(v_tile(vs, 1, kb)) + t * 2048, p.vrows * 128, 1024);
            }
            umma_bf16_ts(tO, pa, dvh[0], idesc_h, accum);
            umma_bf16_ts(tO + 64, pa, dvh[1], idesc_h, accum);
            if (NTERMS == 3) {
              umma_bf16_ts(tO, pa, dvl[0], idesc_h, 1u);
              umma_bf16_ts(tO + 64, pa, dvl[1], idesc_h, 1u);
              umma_bf16_ts(tO, pa + 16, dvh[0], idesc_h, 1u);
              umma_bf16_ts(tO + 64, pa + 16, dvh[1], idesc_h, 1u);
            }
          }
        } else {
          const uint64_t dVs = dV0 + static_cast<uint32_t>(vs) * slot16;
          for (int t = 0; t < ksteps_kv; ++t) {
            // A = P from TMEM: keys 16t..16t+15 sit in 32-column chunk t/2: hi pairs at +8*(t&1), lo pairs 16 further.
            // B = V MN-major: K (= key index) advances by 16 rows of 128 B (= 128 sixteen-byte units), LBO = distance
            // between the 64-wide d blocks, SBO = 8 key rows.
            const uint32_t pa = tP + 32 * (t >> 1) + 8 * (t & 1);
            const uint64_t dv_hi = dVs + static_cast<uint32_t>(t) * 128u;
            const uint32_t accum = (x.j != 0 || t != 0) ? 1u : 0u;
            umma_bf16_ts(tO, pa, dv_hi, idesc_o, accum);
            if (NTERMS == 3) {
              umma_bf16_ts(tO, pa, dv_hi + v_pl16, idesc_o, 1u);
              umma_bf16_ts(tO, pa + 16, dv_hi, idesc_o, 1u);
            }
          }
        }
        umma_commit(&pv_done[u & 1]);
        umma_commit(&kv_empty[vs]);
        umma_commit(&s_free[sb]);  // the score / P buffer may be overwritten by S(u + NSB)
        ATTN_TRACE(5, u);
      };

      UnitIt x{static_cast<int>(blockIdx.x), 0, 1, 0, 0u, 0u};
      load_it(x);
      if (warp == 1) {
        // S(u): Q tile (first unit of an item), K tile, a free score buffer
        for (; x.item < num_items; advance(x)) {
          const uint32_t u = x.u;
          if (x.j == 0) mbar_wait(q_full, x.it & 1);
          mbar_wait(&kv_full[u % RK], (u / RK) & 1);
          ATTN_TRACE(14, u);
          mbar_wait(&s_free[u % NSB], ((u / NSB) & 1) ^ 1);
          ATTN_TRACE(15, u);
          issue_s(x);
        }
      } else {
        // PV(u): P published, V tile, the item's O buffer drained by the epilogue of the item two back
        for (; x.item < num_items; advance(x)) {
          const uint32_t u = x.u;
          mbar_wait(&p_full[u % NSB], (u / NSB) & 1);
          ATTN_TRACE(12, u);
          mbar_wait(&kv_full[RK + u % RV], (u / RV) & 1);
          ATTN_TRACE(13, u);
          if (x.j == 0) mbar_wait(&o_free[x.it & 1], ((x.it >> 1) & 1) ^ 1);
          issue_pv(x);
        }
      }
    }
  } else if (warp >= SW0) {
    // ---------------------------------------------------------------- softmax + epilogue (warps SW0 .. SW0+4*NW)
    const int wq = warp & 3;           // TMEM lane quarter (hardware rule: warp_id % 4)
    const int cw = (warp - SW0) >> 2;  // 32-column chunk of the score tile owned by this warp
    const int row = wq * 32 + lane;    // row inside the q tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
    const int ngroups = p.d >> 4;      // 16-column groups of O; group g belongs to column-warp g % NW
    const bool tracer = (warp == SW0 && lane == 0);  // first softmax thread: trace stamps AND the TMA stores of the staged O tile
    uint32_t mw = 0;                   // mask bits of (row, this warp's 32 columns); bit set = masked
    long long mkey = -1;               // (b, qt, j) combination the word was built for

    // O / l of item `pit` -> this head's column slice of the concatenated output.
    auto epilogue = [&](int b, int h, int qt, uint32_t pit, float m_run) {
      const int par = pit & 1;
      const uint32_t tO = tmem_base + ATTN_TMEM_O + par * 128 + lane_sel;
      float l_tot = 0.0f;
#pragma unroll
      for (int c = 0; c < NW; ++c) l_tot += red_l[(par * 4 + c) * 128 + row];
      const float inv = 1.0f / l_tot;  // l == 0 (row fully masked) -> inf -> NaN, as the reference's softmax of all -inf
      const int qrow = qt * ATTN_BLOCK_M + row;
      const bool row_ok = qrow < p.Lq;
      const size_t grow = static_cast<size_t>(b) * p.Lq + qrow;
      if (p.staged) {
        // the previous item's TMA store must have READ the staging tile before anyone overwrites it
        if (tracer) tma_store_wait_read0();
        named_bar_sync(2, NSW);
      }
      for (int g = cw; g < ngroups; g += NW) {
        uint32_t r[16];
        tmem_ld16(tO + 16 * g, r);
        tmem_wait_ld();
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) * inv;
        uint4 hi[2], lo[2];
        if (p.o_hi != nullptr) {
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            split_bf16x2(v[8 * q2 + 0], v[8 * q2 + 1], hi[q2].x, lo[q2].x);
            split_bf16x2(v[8 * q2 + 2], v[8 * q2 + 3], hi[q2].y, lo[q2].y);
            split_bf16x2(v[8 * q2 + 4], v[8 * q2 + 5], hi[q2].z, lo[q2].z);
            split_bf16x2(v[8 * q2 + 6], v[8 * q2 + 7], hi[q2].w, lo[q2].w);
          }
        }
        if (p.staged) {
          // 128B-swizzled box layout: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
          if (row < p.qrows) {
            const int kb = g >> 2, c0 = (g & 3) * 2, sw = row & 7;
            uint8_t* th = o_tile(0, kb) + row * 128;
            *reinterpret_cast<uint4*>(th + ((c0 ^ sw) << 4)) = hi[0];
            *reinterpret_cast<uint4*>(th + (((c0 + 1) ^ sw) << 4)) = hi[1];
            if (NPL == 2) {
              uint8_t* tl = o_tile(1, kb) + row * 128;
              *reinterpret_cast<uint4*>(tl + ((c0 ^ sw) << 4)) = lo[0];
              *reinterpret_cast<uint4*>(tl + (((c0 + 1) ^ sw) << 4)) = lo[1];
            }
          }
        } else if (row_ok) {
          const int col = h * p.d + 16 * g;
          if (p.o_f32 != nullptr) {
            float4* o = reinterpret_cast<float4*>(p.o_f32 + grow * p.ldof + col);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          }
          if (p.o_hi != nullptr) {
            uint4* oh = reinterpret_cast<uint4*>(p.o_hi + grow * p.ldo + col);
            oh[0] = hi[0];
            oh[1] = hi[1];
            if (p.o_lo != nullptr) {
              uint4* ol = reinterpret_cast<uint4*>(p.o_lo + grow * p.ldo + col);
              ol[0] = lo[0];
              ol[1] = lo[1];
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&o_free[par]);  // this thread's part of the O buffer has been read
      if (p.staged) {
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store
        named_bar_sync(2, NSW);    // the whole tile is staged
        if (tracer) {
          for (int kb = 0; kb < kb64; ++kb) {
            tma_store_3d(&tmO_hi, o_tile(0, kb), h * p.d + kb * 64, qt * ATTN_BLOCK_M, b);
            if (NPL == 2 && p.o_lo != nullptr) tma_store_3d(&tmO_lo, o_tile(1, kb), h * p.d + kb * 64, qt * ATTN_BLOCK_M, b);
          }
          tma_store_commit();
        }
      }
      if (cw == 0 && row_ok && p.row_sum != nullptr) {
        const size_t si = (static_cast<size_t>(h) * p.B + b) * p.Lq + qrow;
        p.row_max[si] = ((m_run == -INFINITY) ? 0.0f : m_run) * p.scale_log2;
        p.row_sum[si] = l_tot;
      }
    };

    uint32_t u = 0, it = 0;
    bool have_prev = false;
    int pb = 0, ph = 0, pqt = 0;
    float pm = 0.0f;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qt = item % num_qt;
      const int h = (item / num_qt) % p.H;
      const int b = item / (num_qt * p.H);
      const uint32_t tO = tmem_base + ATTN_TMEM_O + (it & 1) * 128 + lane_sel;
      float m_run = -INFINITY, l_part = 0.0f;  // m_run: the reference maximum the exponentials are taken against
      int lk, kbase, kbatch_unused, nkv;
      item_keys(b, lk, kbase, kbatch_unused, nkv);
      for (int j = 0; j < nkv; ++j, ++u) {
        const int k0 = j * BLOCK_KV + 32 * cw;  // first key column of this warp's chunk
        const uint32_t sb = u % NSB;
        const uint32_t tS = tmem_base + ATTN_TMEM_S + sb * S_STRIDE + lane_sel + 32 * cw;
        // ---- mask word for (row, this chunk); cached while the addressed mask region is unchanged.  For a
        //      key-padding mask (query stride 0: one byte per key, new every sample) the byte load is issued here and
        //      consumed after the score tile has arrived, so its latency hides behind the barrier wait.
        bool pad_pending = false;
        uint32_t pad_byte = 0;
        if (p.mask_bits != nullptr) {
          // packed mask: issue this thread's word now, OR it in after the score tile has arrived
          const int rem = lk - k0;
          mw = rem >= 32 ? 0u : (rem <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << rem));
          const int qr = qt * ATTN_BLOCK_M + row;
          pad_pending = true;  // (reuses the deferred-OR slot of the key-padding path; see below)
          if (rem > 0 && qr < p.Lq)
            pad_byte = __ldg(p.mask_bits + static_cast<long long>(b) * p.mbb + static_cast<long long>(qr) * p.mbq + (k0 >> 5));
          mkey = -1;
        } else {
          const long long key = (p.mask == nullptr)
                                    ? static_cast<long long>(j)
                                    : ((p.msb ? static_cast<long long>(b) : 0) * num_qt + (p.msq ? qt : 0)) * num_kv + j;
          if (key != mkey || varlen) {  // per-sample key counts: the bounds word changes with every item
            mkey = key;
            const int rem = lk - k0;  // valid columns in this chunk
            mw = rem >= 32 ? 0u : (rem <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << rem));
            if (p.mask != nullptr) {
              const uint8_t* mb = p.mask + static_cast<long long>(b) * p.msb;
              const int kc = k0 + lane;
              if (!p.msq) {
                pad_pending = true;
                if (kc < lk) pad_byte = mb[static_cast<long long>(kbase + kc) * p.msk];  // packed keys: offset by the sample's first row
              } else {
                for (int rr = 0; rr < 32; ++rr) {
                  const int qr = qt * ATTN_BLOCK_M + wq * 32 + rr;
                  uint32_t byte = 0;
                  if (kc < lk && qr < p.Lq)
                    byte = mb[static_cast<long long>(qr) * p.msq + static_cast<long long>(kc) * p.msk];
                  const uint32_t bal = __ballot_sync(0xFFFFFFFFu, byte != 0);
                  if (lane == rr) mw |= bal;
                }
              }
            }
          }
        }
        mbar_wait(&s_full[sb], (u / NSB) & 1);
        if (tracer) ATTN_TRACE(7, u);
        ATTN_TRACE_W(0, u);
        tcgen05_fence_after();
        // ---- pass 1: masked scores of this chunk stay in registers; chunk max -> smem -> row max
        uint32_t r[32];
        tmem_ld32(tS, r);
        if (pad_pending) mw |= (p.mask_bits != nullptr) ? pad_byte : __ballot_sync(0xFFFFFFFFu, pad_byte != 0);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float sv = ((mw >> e) & 1u) ? -INFINITY : __uint_as_float(r[e]);
          r[e] = __float_as_uint(sv);
          mx = fmaxf(mx, sv);
        }
        float* rm = red_max + (u & 1) * (4 * 128);
        rm[cw * 128 + row] = mx;
        named_bar_sync(1, NSW);
        if (tracer) ATTN_TRACE(8, u);
        ATTN_TRACE_W(1, u);
#pragma unroll
        for (int c = 0; c < NW; ++c) mx = fmaxf(mx, rm[c * 128 + row]);
        // ---- lazy reference maximum: move it only when the tile maximum exceeds it by more than 2^8 (or on the
        //      first finite score).  The decision depends on the row only, so the NW column-warps of a row agree.
        const bool move = (j == 0) || ((mx - m_run) * p.scale_log2 > ATTN_RESCALE_LOG2);
        const float m_new = move ? fmaxf(m_run, mx) : m_run;
        const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
        const float alpha = (!move || m_new == -INFINITY) ? 1.0f : ex2_approx((m_run - m_new) * p.scale_log2);
        // ---- multi-tile rows whose reference moved: the previous PV must have retired before the partial output
        //      is rescaled (warp-collective TMEM access -> the whole warp takes part when any row needs it)
        if (j > 0 && __any_sync(0xFFFFFFFFu, alpha != 1.0f)) {
          mbar_wait(&pv_done[(u - 1) & 1], ((u - 1) >> 1) & 1);
          tcgen05_fence_after();
          for (int g = cw; g < ngroups; g += NW) {
            uint32_t o[16];
            tmem_ld16(tO + 16 * g, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st16(tO + 16 * g, o);
          }
        }
        // ---- pass 2: P = 2^((s - m) * scale) -> hi/lo bf16 planes in TMEM (A operand of the PV product)
        float sum = 0.0f;
        const float mb2 = m_use * p.scale_log2;
        uint32_t ph_[16], pl_[16];
        const uint32_t rh = DROP
                                ? drop_rowhash(p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0ull),
                                               (static_cast<unsigned long long>(h) * p.B + b) * p.Lq + qt * ATTN_BLOCK_M + row)
                                : 0u;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float x0 = ex2_approx(fmaf(__uint_as_float(r[e]), p.scale_log2, -mb2));      // 2^(-inf) = 0 for masked
          float x1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), p.scale_log2, -mb2));
          sum += x0 + x1;  // the softmax denominator is taken before dropout
          if (DROP) {
            x0 = drop_keep(rh, static_cast<uint32_t>(k0 + e), p.drop_thresh) ? x0 * p.drop_scale : 0.0f;
            x1 = drop_keep(rh, static_cast<uint32_t>(k0 + e + 1), p.drop_thresh) ? x1 * p.drop_scale : 0.0f;
          }
          split_bf16x2(x0, x1, ph_[e >> 1], pl_[e >> 1]);
        }
        tmem_st16(tS, ph_);                      // P overwrites this warp's own 32 score columns: hi pairs ...
        if (NPL == 2) tmem_st16(tS + 16, pl_);   // ... then lo pairs
        l_part = l_part * alpha + sum;
        m_run = m_new;
        if (j == nkv - 1) red_l[((it & 1) * 4 + cw) * 128 + row] = l_part;  // read after the next unit's barrier
        tmem_wait_st();
        if (tracer) ATTN_TRACE(9, u);
        ATTN_TRACE_W(2, u);
        tcgen05_fence_before();
        mbar_arrive(&p_full[sb]);
        // ---- deferred epilogue of the previous item (its O buffer is the other one): overlaps PV(u)
        if (j == 0 && have_prev) {
          mbar_wait(&pv_done[(u - 1) & 1], ((u - 1) >> 1) & 1);
          if (tracer) ATTN_TRACE(10, u);
          tcgen05_fence_after();
          epilogue(pb, ph, pqt, it - 1, pm);
          if (tracer) ATTN_TRACE(11, u);
        }
      }
      have_prev = true;
      pb = b; ph = h; pqt = qt; pm = m_run;
    }
    if (have_prev) {
      named_bar_sync(1, NSW);  // make the last item's row sums visible
      mbar_wait(&pv_done[(u - 1) & 1], ((u - 1) >> 1) & 1);
      tcgen05_fence_after();
      epilogue(pb, ph, pqt, it - 1, pm);
    }
    if (tracer && p.staged) tma_store_wait_all0();  // global writes complete before the CTA exits
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, ATTN_TMEM_COLS);
  }
}

}  // namespace lamp
