// qserve_b200 -- W4A8 (per-channel / per-group) and W8A8 GEMM for sm_100a.
//
// Replaces  kernels/csrc/qgemm/w4a8_per_chn/gemm_cuda.cu:303-652,
//           kernels/csrc/qgemm/w4a8_per_group/gemm_cuda.cu:328-702,
//           kernels/csrc/qgemm/w8a8/w8a8_gemm_cuda.cu:267-577        (reference: mma.sync + ldmatrix + cp.async)
//
// B200 design (see DESIGN.md "GEMM"):
//   * swap-AB: the 128 output channels of a tile are the UMMA M dimension, the tokens are UMMA N (<= 256),
//     so decode-sized batches (M = 64 tokens) still use full-height 128-row tensor-core instructions.
//   * the packed INT4 weights are consumed in the checkpoint's own "compute-aware reordered" layout
//     (w4a8_linear.py:292-322): one 32x32 tile is 512 contiguous bytes, a 32-channel band is contiguous in K.
//     ONE TMA tensor copy per 128-K block stages 4 bands x 2 KB into shared memory; four unpack warps read one
//     16-byte lane chunk each (LDS.128), split nibbles in registers (per-group: level-2 dequant q*s2+z2 with the
//     reference's exact 32-bit multiply + vadd4), and write INT8 rows straight into TENSOR MEMORY with
//     tcgen05.st.16x128b -- the mma.sync B-fragment order of the checkpoint is exactly the 16x128b TMEM store
//     pattern, so no shuffle and no shared-memory round trip is needed.
//   * two issuer threads (warps 1 and 7: even / odd 256-K stages) issue tcgen05.mma.kind::i8 with A (weights) from TMEM and
//     B (INT8 activations, TMA-loaded with the 128-byte swizzle) from shared memory; INT32 accumulators live in TMEM and are
//     zero-filled up front so that every MMA accumulates (the integer result cannot depend on how the two streams interleave).
//   * weights and activations travel in separate rings: the weight producer never waits for the previous kernel, so with
//     programmatic dependent launch the ring and the unpacked TMEM stages fill while the preceding norm / quant kernel runs.
//   * decode shapes have too few 128-channel tiles to fill 148 SMs, so K is split across a thread-block CLUSTER
//     (2/4/8 CTAs); every CTA pushes its INT32 partials of the channels a peer finishes into that peer's receive buffer
//     through DISTRIBUTED SHARED MEMORY (st.shared::cluster) and each CTA sums and finishes a 1/S slice of the channels --
//     integer adds, so the result is bit-identical for every split.  Split launches run one CTA per SM.
//   * two CTAs per SM (<= 113 KB smem, 256 TMEM columns each) so one CTA's prologue / epilogue overlaps the other's
//     weight stream.
//   * epilogue fused: acc*s1[n]*sa[m] - s1z[n]*asum[m] (per-channel) or acc*(s1[n]*sa[m]) (per-group, W8A8) -> fp16,
//     IEEE fp32 in the reference's source order (bit-exact against the oracle).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "launch.h"

namespace qs {

namespace {

constexpr int kBM = 128;  // output channels per tile (UMMA M)
constexpr int kBK = 128;  // K sub-block (= one g128 group, one 128-byte swizzled activation row)
constexpr int kSub = 2;    // sub-blocks per pipeline stage: per-stage barrier latencies are amortised over 256 K
constexpr int kNumThreads = 256;
#ifndef QS_PAIR_CVT
#define QS_PAIR_CVT 1
#endif
constexpr bool kPairCvt = QS_PAIR_CVT != 0;
#ifndef QS_DECODE_CVT
#define QS_DECODE_CVT 1
#endif
constexpr bool kDecodeCvt = QS_DECODE_CVT != 0;  // the same choice for the epilogue of gemm_kernel (A/B: profiles/r02_notes.md 4b)
constexpr int kPairThreads = 384;  // gemm_pair_kernel: the 8 warps of gemm_kernel + 4 more epilogue warps
// warp roles: 0 weight producer | 1 TMEM owner + MMA issuer (even stages) | 2..5 unpack + TMEM epilogue | 6 activation producer |
// 7 MMA issuer (odd stages).
// All eight warps take part in the final reduce / scale / store phase.

enum { kModeW4Chn = 0, kModeW4Grp = 1, kModeW8 = 2 };

struct GemmParams {
  const uint8_t* s2_scales;  // per-group: [K/128, N] (shuffled per 32 columns, as stored in the checkpoint)
  const uint8_t* s2_zeros;   // per-group: [K/128, N]
  const __half* wscales;     // [N]
  const __half* w_szs;       // [N]  (per-channel only)
  const __half* ascales;     // [M]
  const __half* a_ssums;     // [M]  (per-channel only)
  __half* out;               // [M, N]
  int32_t* acc_out;          // optional: raw INT32 accumulators [M, N] (parity tests)
  int M, N, K;
  int m_tiles, kb_per_tile, split;
  unsigned long long* prof;  // optional: 16 globaltimer stamps per CTA (tools/gemm_timeline.py)
};

// WS = depth of the WEIGHT ring in shared memory.  Weights are static, so the producer streams them before the
// programmatic-dependent-launch wait: everything that fits in the ring (plus the unpacked stages parked in tensor memory)
// is already on chip when the preceding activation kernel finishes.
template <int MODE, int NT, int WS, int AS>
struct Cfg {
  static constexpr int kActStages = AS;  // activation ring (L2-resident operand): deep enough to cover the TMA round trip
  static constexpr int kActSub = NT * kBK;                                       // one swizzled [NT x 128 B] activation sub-tile
  static constexpr int kWSub = (MODE == kModeW8) ? kBM * kBK : kBM * kBK / 2;   // int8 rows or packed int4 tiles of one sub-block
  static constexpr int kS2Sub = (MODE == kModeW4Grp) ? 2 * kBM : 0;             // scales | zeros of one group
  static constexpr int kActBytes = kSub * kActSub;
  static constexpr int kWBytes = kSub * kWSub;
  static constexpr int kS2Bytes = kSub * kS2Sub;
  static constexpr int kWStageTx = kWBytes + kS2Bytes;
  static constexpr int kAStageCols = kSub * (kBK / 4);                          // TMEM columns of one unpacked-A stage
  // unpacked-A ring in tensor memory: as many stages as fit next to the accumulator in 256 (two CTAs / SM) or 512 columns
  static constexpr int kTA = (MODE == kModeW8) ? 0 : (NT <= 64 ? 3 : NT <= 128 ? 2 : 4);
  static constexpr int kTmemNeed = NT + kTA * kAStageCols;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "TMEM overflow");
  // shared memory carve-up (offsets from a 1024-aligned base; activation and W8 weight stages are 1024-byte multiples)
  static constexpr int kOffAct = 0;
  static constexpr int kOffW = kOffAct + kActStages * kActBytes;
  static constexpr int kOffS2 = kOffW + WS * kWBytes;
  static constexpr int kPipeBytes = kOffS2 + WS * kS2Bytes;
  static constexpr int kRedBytes = NT * kBM * 4;  // INT32 partial tile [NT][128], aliases the pipeline buffers
  static constexpr int kOffRow = (kPipeBytes > kRedBytes ? kPipeBytes : kRedBytes);  // float ascales[NT], asums[NT]
  static constexpr int kOffBar = kOffRow + 2 * NT * 4;
  static constexpr int kNumBars = 2 * WS + 2 * kActStages + 2 * (kTA > 0 ? kTA : 1) + 2;
  static constexpr int kOffMisc = kOffBar + kNumBars * 8;  // tmem ptr
  static constexpr int kSmemBytes = kOffMisc + 16;
  // split-K (cluster) launches append a dedicated receive buffer: every CTA pushes its INT32 partials of the channels another
  // CTA of the cluster finishes straight into that CTA's buffer (posted st.shared::cluster from the TMEM-load registers); the
  // owner then sums S local rows.  Such launches run one CTA per SM (the buffer does not fit next to a second CTA).
  static constexpr int kOffRx = (kSmemBytes + 127) / 128 * 128;
  static constexpr int kRxBytes = NT * kBM * 4;  // [sender][token][128 / S channels]
  static constexpr int kSmemBytesSplit = kOffRx + kRxBytes;
  static constexpr bool kSplitFits = kSmemBytesSplit <= 226 * 1024;
  // two co-resident CTAs per SM when both the TMEM columns (<= 256 each) and the shared memory (<= 113 KB each) allow it
  static constexpr int kCtasPerSm = (kTmemCols <= 256 && kSmemBytes <= 113 * 1024) ? 2 : 1;
  static_assert(kSmemBytes <= 226 * 1024, "shared memory overflow");
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
  return t;
}
#define QS_PROF(slot) do { if (p.prof) p.prof[blockIdx.x * 16 + (slot)] = gtime(); } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// INT32 -> FP32, round to nearest even, without the conversion instruction:  v = hi * 65536 + lo, both halves are converted exactly by the
// 1.5 * 2^23 magic-number add, and the single rounding of the fused multiply-add equals cvt.rn.f32.s32(v).  Round 1 used it in every epilogue
// on the assumption that I2F is a quarter-rate XU instruction; the pipe micro-benchmark (profiles/r01_ubench_pipes.txt: 2.5 cycles per warp
// instruction, against 5 x ~2 for this sequence) and the A/B of round 2 say otherwise (decode step 2.797 -> 2.771 ms, prefill tiles +2 %), so the
// kernels now convert with the instruction; this stays as the -DQS_DECODE_CVT=0 / -DQS_PAIR_CVT=0 variant.
__device__ __forceinline__ float int2float_rn_noxu(int32_t v) {
  const int32_t hi = v >> 16, lo = v & 0xFFFF;
  const float fhi = __fsub_rn(__int_as_float(0x4B400000 + hi), 12582912.f);
  const float flo = __fsub_rn(__int_as_float(0x4B400000 + lo), 12582912.f);
  return __fmaf_rn(fhi, 65536.f, flo);
}

template <int V>
struct IntTag { static constexpr int value = V; };

// CVT = true: cvt.rn.f32.s32 itself instead of the five-instruction emulation above -- same value.
template <int MODE, bool CVT = false>
__device__ __forceinline__ float epilogue_one(int32_t acc, float ws, float wsz, float as, float asum) {
  // IEEE fp32, reference source order, no FMA contraction (bit-exact against oracle/w4a8.py)
  float ps = CVT ? __int2float_rn(acc) : int2float_rn_noxu(acc);
  if constexpr (MODE == kModeW4Chn) {
    // w4a8_per_chn/gemm_cuda.cu:586  psum * wscale * ascale - w_sz * a_ssum
    float t = __fmul_rn(__fmul_rn(ps, ws), as);
    float u = __fmul_rn(wsz, asum);
    return __fsub_rn(t, u);
  } else {
    // w4a8_per_group/gemm_cuda.cu:619, w8a8_gemm_cuda.cu:522   psum *= wscale * ascale
    return __fmul_rn(ps, __fmul_rn(ws, as));
  }
}

// grid.x = tiles * split; the `split` CTAs of a cluster share one (n_tile, m_tile) and own disjoint K ranges
template <int MODE, int NT, int WS, int AS, bool ACC>
__global__ void __launch_bounds__(kNumThreads, Cfg<MODE, NT, WS, AS>::kCtasPerSm)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  using C = Cfg<MODE, NT, WS, AS>;
  constexpr int kActStages = AS;
  constexpr int TA = C::kTA > 0 ? C::kTA : 1;
  // The two MMA issuers take alternate iterations, so on a ring of ODD depth each issuer sees only every other phase of a given barrier and a
  // parity wait could be satisfied by the phase before the one it skipped.  W4A8 is safe for any depth: an issuer first waits for the TMEM A stage of
  // its iteration, which the (in-order) unpack warps can only fill after the OTHER issuer's commit for the iteration three back -- and that issuer
  // had waited for the very activation / weight phase in question.  W8A8 has no such guard (operands come straight from TMA-completed barriers,
  // which may complete out of order): its rings must have even depth, so that every barrier belongs to one issuer and is observed phase by phase.
  // (Round 2: a 3-deep W8A8 weight ring produced a launch failure in the first test that used it.)
  static_assert(MODE != kModeW8 || (WS % 2 == 0 && AS % 2 == 0), "W8A8: ring depths must be even (two alternating MMA issuers)");
  extern __shared__ __align__(1024) uint8_t smem[];  // no static shared memory in this kernel: the window starts 1024-aligned
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_act = smem + C::kOffAct;
  uint8_t* s_w = smem + C::kOffW;
  uint8_t* s_s2 = smem + C::kOffS2;
  int32_t* s_red = reinterpret_cast<int32_t*>(smem);  // aliases the pipeline buffers once the mainloop has drained
  float* s_asc = reinterpret_cast<float*>(smem + C::kOffRow);
  float* s_asum = s_asc + NT;
  uint64_t* bar_wfull = reinterpret_cast<uint64_t*>(smem + C::kOffBar);  // TMA -> unpack warps (W4) / MMA (W8)
  uint64_t* bar_wempty = bar_wfull + WS;                                   // unpack warps (W4) / MMA commit (W8) -> weight producer
  uint64_t* bar_xfull = bar_wempty + WS;                                   // activation TMA -> MMA
  uint64_t* bar_xempty = bar_xfull + kActStages;                           // MMA commit -> activation producer
  uint64_t* bar_afull = bar_xempty + kActStages;                           // unpack warps -> MMA (TMEM A stage written)
  uint64_t* bar_aempty = bar_afull + TA;                                   // MMA commit -> unpack warps
  uint64_t* bar_dfull = bar_aempty + TA;
  uint64_t* bar_zero = bar_dfull + 1;                                        // accumulator zero-filled (4 epilogue warps)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kOffMisc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.split;
  const int rank = (S > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile = blockIdx.x / S;
  const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;
  // balanced K split: rank r owns pipeline stages (256 K each) [kb_begin, kb_end)
  const int KB = p.kb_per_tile;
  const int kb_begin = (KB * rank) / S, kb_end = (KB * (rank + 1)) / S;
  const int n_kb = kb_end - kb_begin;
  if (threadIdx.x == 0) QS_PROF(0);
  qs_trace(QS_K_GEMM, 0);
  // final-phase geometry: this thread finishes channel pair `pr` of the tile; its (static) weight scales are fetched now
  const int npairs = 64 / S;
  const int pr = (64 * rank) / S + static_cast<int>(threadIdx.x) % npairs;
  const __half2 ws_h2 = __ldg(reinterpret_cast<const __half2*>(p.wscales + n_tile * kBM + 2 * pr));
  __half2 wz_h2 = __half2half2(__ushort_as_half(0));
  if constexpr (MODE == kModeW4Chn) wz_h2 = __ldg(reinterpret_cast<const __half2*>(p.w_szs + n_tile * kBM + 2 * pr));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_act);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < WS; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], MODE == kModeW8 ? 1 : 4);
    }
    for (int i = 0; i < kActStages; ++i) {
      mbar_init(&bar_xfull[i], 1);
      mbar_init(&bar_xempty[i], 1);
    }
    for (int i = 0; i < TA; ++i) {
      mbar_init(&bar_afull[i], 4);
      mbar_init(&bar_aempty[i], 1);
    }
    mbar_init(bar_dfull, 2);  // two MMA issuers
    mbar_init(bar_zero, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) { pdl_launch_dependents(); QS_PROF(1); }

  if (warp == 0) {
    // ===================================== weight producer: never waits for the previous kernel =====================================
    if (lane == 0) {
      const int w_row = (MODE == kModeW8) ? n_tile * kBM : n_tile * 4;  // W8: row of [N,K]; W4: band index
      const uint8_t* s2s = p.s2_scales + static_cast<size_t>(n_tile) * kBM;
      const uint8_t* s2z = p.s2_zeros + static_cast<size_t>(n_tile) * kBM;
      // one stage = kSub sub-blocks of 128 K.  A sub-block past the end of K (K % 256 == 128) is zero-filled by the
      // tensor maps (weights and activations), so it contributes nothing; its g128 params are re-read from the last group.
      int s = 0;
      uint32_t ph = 0;  // parity of the (it / WS - 1)-th completion of wempty[s]
      for (int it = 0; it < n_kb; ++it) {
        if (it >= WS) mbar_wait(&bar_wempty[s], ph);
        mbar_expect_tx(&bar_wfull[s], C::kWStageTx);
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const int kb = (kb_begin + it) * kSub + u;
          // W4: u64 elements, 256 per 128-K block of one band (4 tiles x 512 B); W8: bytes
          tma_load_2d(s_w + s * C::kWBytes + u * C::kWSub, &tmap_w, (MODE == kModeW8) ? kb * kBK : kb * 256, w_row, &bar_wfull[s]);
          if constexpr (MODE == kModeW4Grp) {
            const int kg = kb < p.K / kBK ? kb : p.K / kBK - 1;
            bulk_copy_g2s(s_s2 + s * C::kS2Bytes + u * C::kS2Sub, s2s + static_cast<size_t>(kg) * p.N, kBM, &bar_wfull[s]);
            bulk_copy_g2s(s_s2 + s * C::kS2Bytes + u * C::kS2Sub + kBM, s2z + static_cast<size_t>(kg) * p.N, kBM, &bar_wfull[s]);
          }
        }
        if (++s == WS) { s = 0; if (it >= WS) ph ^= 1; }
      }
      QS_PROF(3);
    }
  } else if (warp == 6) {
    // ===================================== activation producer =====================================
    if (lane == 0) {
      const int a_row = m_tile * NT;
      pdl_wait();  // the activations are the previous kernel's output
      QS_PROF(2);
      qs_trace(QS_K_GEMM, 1, 6 * 32);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_kb; ++it) {
        if (it >= kActStages) mbar_wait(&bar_xempty[s], ph);
        mbar_expect_tx(&bar_xfull[s], C::kActBytes);
#pragma unroll
        for (int u = 0; u < kSub; ++u)
          tma_load_2d(s_act + s * C::kActBytes + u * C::kActSub, &tmap_act, ((kb_begin + it) * kSub + u) * kBK, a_row, &bar_xfull[s]);
        if (++s == kActStages) { s = 0; if (it >= kActStages) ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 7) {
    // ===================================== two MMA issuers: warp 1 takes the even stages, warp 7 the odd ones =====================
    // One issuer spends ~1060 cycles per stage: two mbarrier waits (~240 cycles each even on a completed phase), 8 MMAs
    // (~48 cycles of tensor pipe each) and two commits -- the tensor pipe idles half of the time.  Interleaving two issuers hides
    // one's waits behind the other's MMAs.  The accumulator is zero-filled up front, so every MMA accumulates and the (integer)
    // result does not depend on the order in which the two streams reach the tensor pipe.
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_i8(kBM, NT, 1u, 1u);
      constexpr int RW = (MODE == kModeW8) ? WS : TA;
      const int first = (warp == 1) ? 0 : 1;
      int sw = first % RW, sx = first % kActStages;
      uint32_t phw = (first / RW) & 1, phx = (first / kActStages) & 1;
      mbar_wait(bar_zero, 0);
      tc_fence_after();
      for (int it = first; it < n_kb; it += 2) {
        // W4: the unpack warps observed wfull before arriving on afull, so afull alone orders the weight data
        if constexpr (MODE == kModeW8) mbar_wait(&bar_wfull[sw], phw); else mbar_wait(&bar_afull[sw], phw);
        mbar_wait(&bar_xfull[sx], phx);
        if (it == 0) QS_PROF(4);
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const uint64_t bdesc = umma_desc_sw128(smem_u32(s_act + sx * C::kActBytes + u * C::kActSub));
#pragma unroll
          for (int t = 0; t < kBK / 32; ++t) {
            if constexpr (MODE == kModeW8) {
              const uint64_t adesc = umma_desc_sw128(smem_u32(s_w + sw * C::kWBytes + u * C::kWSub));
              umma_i8_ss(tmem_base, adesc + t * 2, bdesc + t * 2, idesc, 1u);
            } else {
              umma_i8_ts(tmem_base, tmem_base + NT + sw * C::kAStageCols + u * (kBK / 4) + t * 8, bdesc + t * 2, idesc, 1u);
            }
          }
        }
        if constexpr (MODE == kModeW8) umma_commit(&bar_wempty[sw]); else umma_commit(&bar_aempty[sw]);
        umma_commit(&bar_xempty[sx]);
        sw += 2; if (sw >= RW) { sw -= RW; phw ^= 1; }
        sx += 2; if (sx >= kActStages) { sx -= kActStages; phx ^= 1; }
      }
      if (n_kb > first) umma_commit(bar_dfull); else mbar_arrive(bar_dfull);
      if (first == 0) QS_PROF(6);
    }
  } else if (warp >= 2 && warp <= 5) {
    // ===================================== unpack + TMEM epilogue warps =====================================
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int epi_tid = quad * 32 + lane; // 0..127 == channel row inside the tile
    {
      // zero-fill this quadrant of the accumulator (all MMAs accumulate: see the issuers)
      const uint32_t tz = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
      for (int c = 0; c < NT; c += 8) {
        tmem_st_16x128b_x2(tz + c, 0u, 0u, 0u, 0u);
        tmem_st_16x128b_x2(tz + c + (16u << 16), 0u, 0u, 0u, 0u);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_zero);
    }
    if constexpr (MODE != kModeW8) {
      int s = 0, ta = 0;
      uint32_t ph = 0, pha = 0;  // pha: parity of the (it / TA - 1)-th completion of aempty[ta]
      for (int it = 0; it < n_kb; ++it) {
        mbar_wait(&bar_wfull[s], ph);
        if (it >= TA) {
          mbar_wait(&bar_aempty[ta], pha);
          tc_fence_after();
        }
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const uint8_t* wsrc = s_w + s * C::kWBytes + u * C::kWSub + quad * 2048 + lane * 16;
          const uint32_t tdst = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + NT + ta * C::kAStageCols + u * (kBK / 4);
          uint32_t sc4 = 0, zp4 = 0;
          if constexpr (MODE == kModeW4Grp) {
            const uint8_t* s2 = s_s2 + s * C::kS2Bytes + u * C::kS2Sub + quad * 32 + (lane >> 2) * 4;
            sc4 = *reinterpret_cast<const uint32_t*>(s2);
            zp4 = *reinterpret_cast<const uint32_t*>(s2 + kBM);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint4 v = *reinterpret_cast<const uint4*>(wsrc + t * 512);
            uint32_t xl = v.x & 0x0F0F0F0Fu, xh = (v.x >> 4) & 0x0F0F0F0Fu;
            uint32_t yl = v.y & 0x0F0F0F0Fu, yh = (v.y >> 4) & 0x0F0F0F0Fu;
            uint32_t zl = v.z & 0x0F0F0F0Fu, zh = (v.z >> 4) & 0x0F0F0F0Fu;
            uint32_t wl = v.w & 0x0F0F0F0Fu, wh = (v.w >> 4) & 0x0F0F0F0Fu;
            if constexpr (MODE == kModeW4Grp) {
              // w4a8_per_group/gemm_cuda.cu:298-324: 32-bit multiply of four nibble-bytes, then vadd4 with the s8 zero
              const uint32_t s0 = sc4 & 0xFF, s1 = (sc4 >> 8) & 0xFF, s2 = (sc4 >> 16) & 0xFF, s3 = sc4 >> 24;
              const uint32_t z0 = __byte_perm(zp4, 0, 0x0000), z1 = __byte_perm(zp4, 0, 0x1111);
              const uint32_t z2 = __byte_perm(zp4, 0, 0x2222), z3 = __byte_perm(zp4, 0, 0x3333);
              xl = __vadd4(xl * s0, z0); zl = __vadd4(zl * s0, z0);   // channel c
              yl = __vadd4(yl * s1, z1); wl = __vadd4(wl * s1, z1);   // channel c + 8
              xh = __vadd4(xh * s2, z2); zh = __vadd4(zh * s2, z2);   // channel c + 16
              yh = __vadd4(yh * s3, z3); wh = __vadd4(wh * s3, z3);   // channel c + 24
            }
            // lanes 0..15 of the quadrant <- channels c, c+8 ; lanes 16..31 <- channels c+16, c+24 ; 8 columns = 32 k
            tmem_st_16x128b_x2(tdst + t * 8, xl, yl, zl, wl);
            tmem_st_16x128b_x2(tdst + t * 8 + (16u << 16), xh, yh, zh, wh);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bar_afull[ta]);
          mbar_arrive(&bar_wempty[s]);  // the shared-memory stage has been consumed into registers / tensor memory
        }
        if (++s == WS) { s = 0; ph ^= 1; }
        if (++ta == TA) { ta = 0; if (it >= TA) pha ^= 1; }
      }
    }

    // ------------------------------------------ epilogue ------------------------------------------
    pdl_wait();  // ascales / a_ssums / out belong to the dependency chain
    const int m0 = m_tile * NT;
    for (int j = epi_tid; j < NT; j += 128) {
      const bool ok = (m0 + j) < p.M;
      s_asc[j] = ok ? __half2float(p.ascales[m0 + j]) : 0.f;
      if constexpr (MODE == kModeW4Chn) s_asum[j] = ok ? __half2float(p.a_ssums[m0 + j]) : 0.f;
    }
    if (epi_tid == 0) QS_PROF(7);
    mbar_wait(bar_dfull, 0);   // all MMAs retired: accumulators complete, pipeline buffers free
    tc_fence_after();
    if (epi_tid == 0) QS_PROF(8);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    if (S == 1) {
      // TMEM -> shared memory, transposed to [token][channel] so that channel pairs are contiguous
#pragma unroll 1
      for (int c = 0; c < NT / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) s_red[(c * 32 + i) * kBM + epi_tid] = static_cast<int32_t>(r[i]);
      }
    } else {
      // TMEM -> the receive buffer of the CTA that finishes this channel: rx[sender = rank][token][channel % (128 / S)].
      // Posted remote stores (>= 64 contiguous bytes per warp instruction); the cluster barrier below publishes them.
      const int cps = kBM / S;  // channels finished per CTA
      const uint32_t owner = static_cast<uint32_t>(epi_tid / cps);
      const uint32_t rx_peer = map_to_cta(smem_u32(smem + C::kOffRx), owner) + static_cast<uint32_t>(((rank * NT) * cps + (epi_tid % cps)) * 4);
#pragma unroll 1
      for (int c = 0; c < NT / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) st_dsmem_u32(rx_peer + static_cast<uint32_t>((c * 32 + i) * cps * 4), r[i]);
      }
    }
    if (epi_tid == 0) QS_PROF(9);
  } else {
    pdl_wait();
  }

  // ---------------- cross-CTA (cluster) reduction of the INT32 partial tiles through distributed shared memory ----------------
  tc_fence_before();
  if (S > 1) cluster_sync_all(); else __syncthreads();
  {
    // all 256 threads: this CTA finishes channel pairs [pair0, pair1) of the tile for all NT tokens
    pdl_wait();  // already resolved; makes every storing thread an observer of the dependency
    const int tid = threadIdx.x;
    if (tid == 0) QS_PROF(10);
    const int m0 = m_tile * NT;
    // npairs divides 256, so a thread keeps one channel pair and walks the tokens
    const int tstep = kNumThreads / npairs;    // 4 * S tokens are finished per pass of the CTA
    const int n = n_tile * kBM + 2 * pr;
    float wz0 = 0.f, wz1 = 0.f;
    if constexpr (MODE == kModeW4Chn) { wz0 = __low2float(wz_h2); wz1 = __high2float(wz_h2); }
    const float ws0 = __low2float(ws_h2), ws1 = __high2float(ws_h2);
    const int tok_end = min(NT, p.M - m0);  // tokens of this tile that exist
    const int tok0 = tid / npairs;
    const uint32_t off0 = static_cast<uint32_t>((tok0 * kBM + 2 * pr) * 4), off_step = static_cast<uint32_t>(tstep * kBM * 4);
    __half* optr = p.out + static_cast<size_t>(m0 + tok0) * p.N + n;
    const size_t ostep = static_cast<size_t>(tstep) * p.N;
    auto finish = [&](int tok, int2 acc, __half* dst) {
      const float as = s_asc[tok], asum = s_asum[tok];
      const float o0 = epilogue_one<MODE, kDecodeCvt>(acc.x, ws0, wz0, as, asum);
      const float o1 = epilogue_one<MODE, kDecodeCvt>(acc.y, ws1, wz1, as, asum);
      *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(o0, o1);  // one packed conversion, each half rounded to nearest even
      if constexpr (ACC) *reinterpret_cast<int2*>(p.acc_out + (dst - p.out)) = acc;
    };
    // tokens per thread: NT / (4 S); every round has 16 (distributed-)shared-memory loads in flight per thread
    if (S == 1) {
#pragma unroll 1
      for (int i0 = 0; i0 * tstep + tok0 < NT; i0 += 16) {
        int2 v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (NT / 4 > e) v[e] = *reinterpret_cast<const int2*>(reinterpret_cast<const uint8_t*>(s_red) + off0 + (i0 + e) * off_step);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int tok = tok0 + (i0 + e) * tstep;
          if (NT / 4 > e && tok < tok_end) finish(tok, v[e], optr + (i0 + e) * ostep);
        }
      }
    } else {
      // S > 1: the partials of all S senders sit in this CTA's receive buffer; sum them locally
      const int cps = kBM / S;
      const int32_t* rx = reinterpret_cast<const int32_t*>(smem + C::kOffRx);
      const int prl = tid % npairs;  // channel pair inside this CTA's slice
      // tokens of this thread: NT / (4 S); S is a launch constant of the cluster: specialise so that all S x tokens shared-memory loads of a
      // thread are in flight together (the rolled loop serialised 4 dependent LDS -> add -> convert -> scale -> store chains)
      auto reduce_rows = [&](auto s_tag) {
        constexpr int SS = decltype(s_tag)::value;
        constexpr int TPT = NT / (4 * SS) > 0 ? NT / (4 * SS) : 1;
        int2 v[TPT][SS];
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
          const int tok = tok0 + i * tstep;
#pragma unroll
          for (int r = 0; r < SS; ++r)
            v[i][r] = (tok < NT) ? *reinterpret_cast<const int2*>(rx + (r * NT + tok) * cps + 2 * prl) : make_int2(0, 0);
        }
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
          const int tok = tok0 + i * tstep;
          int2 acc = v[i][0];
#pragma unroll
          for (int r = 1; r < SS; ++r) { acc.x += v[i][r].x; acc.y += v[i][r].y; }
          if (tok < tok_end) finish(tok, acc, optr + i * ostep);
        }
      };
      if (S == 2) reduce_rows(IntTag<2>{});
      else if (S == 4) reduce_rows(IntTag<4>{});
      else reduce_rows(IntTag<8>{});
    }
    if (tid == 0) QS_PROF(11);
  }
  // S > 1: every remote store into this CTA's receive buffer was ordered before the cluster barrier above; none follows
  __syncthreads();
  if (threadIdx.x == 0) QS_PROF(12);
  qs_trace(QS_K_GEMM, 2);
  if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// =============================================================================================
// Prefill tiles (M >= 512, W4A8): a PAIR of CTAs on the two SMs of a TPC computes 256-channel x 256-token tiles with cta_group::2 UMMAs,
// persistently (one pair per TPC walks the tile list).
//
// Why: with 128-token tiles every CTA pulls its whole 128 x K weight band AND its whole 128 x K activation band out of L2: M = 4096 x N = 28672 x
// K = 4096 moves 5.5 GB L2 -> SM in 413 us = 13 TB/s, which IS the chip's L2 throughput (~6300 B/clk, B300_MICROARCH.md) -- the tensor pipe idles
// at 52 %.  In a CTA pair each SM unpacks its own 128 channels into its own tensor memory (A operand) but stages only HALF of the 256 tokens of
// the activation tile (B operand); the 2-CTA MMA reads the other half from the peer SM.  L2 -> SM traffic halves (2.7 GB for the shape above).
// The first version of this kernel (one tile per launch-time CTA pair, 1 pair per TPC because of its 512 TMEM columns) lost all of that again to
// un-overlapped per-tile fill / drain (24 us per tile for 8.3 us of MMA work; profiles/r02_notes.md), hence the persistent form: operand rings run
// across tile boundaries, the only exposed per-tile cost is the accumulator drain (TMEM -> registers -> global, no shared-memory staging).
//
// Roles per CTA (384 threads): warp 0 weight producer | warp 1 TMEM allocator, leader: MMA issuer (even stages) | warps 2..5 unpack + epilogue |
// warp 6 activation producer (both CTAs signal the LEADER's "full" barrier: cta_group::2 TMA) | warp 7 leader: MMA issuer (odd stages) |
// warps 8..11 epilogue only (a lone warp per scheduler issues one instruction per clock: four warps needed 9 us to drain a tile, eight 2 x fewer).
// Only the even CTA of the pair (the leader) issues MMAs; its commits are multicast to the "empty" barriers of both CTAs.  All rings have even
// depth (two alternating issuers, see gemm_kernel).  The accumulator is zero-filled by the epilogue warps (tcgen05.st) and every MMA accumulates:
// the two issuers' first instructions of a tile are not ordered against each other.
// =============================================================================================
template <int MODE>
struct PairCfg {
  static constexpr int NT = 256;                       // tokens per pair tile (UMMA N)
  static constexpr int NH = NT / 2;                    // tokens staged per CTA
  static constexpr int WS = 4, AS = 4, TA = 4;         // powers of two (stage = iteration & (depth - 1))
  static constexpr int kActSub = NH * kBK;             // 16 KB: one swizzled [128 x 128 B] activation sub-tile (this CTA's half)
  static constexpr int kActBytes = kSub * kActSub;     // 32 KB per stage
  static constexpr int kWSub = kBM * kBK / 2;          // 8 KB packed INT4 per 128-K sub-block
  static constexpr int kWBytes = kSub * kWSub;         // 16 KB per stage
  static constexpr int kS2Sub = (MODE == kModeW4Grp) ? 2 * kBM : 0;
  static constexpr int kS2Bytes = kSub * kS2Sub;
  static constexpr int kWStageTx = kWBytes + kS2Bytes;
  static constexpr int kAStageCols = kSub * (kBK / 4);  // 64 TMEM columns per unpacked-A stage
  static constexpr int kTmemCols = 512;                 // 256 accumulator columns + 4 x 64 A-ring columns
  static constexpr int kOffAct = 0;
  static constexpr int kOffW = AS * kActBytes;          // 128 KB
  static constexpr int kOffS2 = kOffW + WS * kWBytes;
  static constexpr int kPipeBytes = kOffS2 + WS * kS2Bytes;
  static constexpr int kOffRow = kPipeBytes;            // float ascales[2][NT], asums[2][NT] (double-buffered by tile parity)
  static constexpr int kOffBar = kOffRow + 4 * NT * 4;
  static constexpr int kNumBars = 2 * WS + 2 * AS + 2 * TA + 2;
  static constexpr int kOffMisc = kOffBar + kNumBars * 8;
  static constexpr int kSmemBytes = kOffMisc + 16;
  static_assert(kSmemBytes <= 226 * 1024, "shared memory overflow");
};

template <int MODE, bool ACC>
__global__ void __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  using C = PairCfg<MODE>;
  constexpr int NT = C::NT, WS = C::WS, AS = C::AS, TA = C::TA;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_act = smem + C::kOffAct;
  uint8_t* s_w = smem + C::kOffW;
  uint8_t* s_s2 = smem + C::kOffS2;
  float* s_row = reinterpret_cast<float*>(smem + C::kOffRow);
  uint64_t* bar_wfull = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* bar_wempty = bar_wfull + WS;
  uint64_t* bar_xfull = bar_wempty + WS;    // leader only: BOTH halves of the activation stage have landed (the peer's TMA signals it remotely)
  uint64_t* bar_xempty = bar_xfull + AS;    // MMA commit (multicast) -> this CTA's activation producer
  uint64_t* bar_afull = bar_xempty + AS;    // leader only: 16 arrivals = the eight unpack warps of both CTAs
  uint64_t* bar_aempty = bar_afull + TA;    // MMA commit (multicast) -> this CTA's unpack warps
  uint64_t* bar_dfull = bar_aempty + TA;    // both MMA issuers' last commits of a tile (multicast): accumulators complete
  uint64_t* bar_zero = bar_dfull + 1;       // leader only: accumulators of both CTAs drained and zero-filled (16 arrivals: 8 epilogue warps each)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kOffMisc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int n_pairs = static_cast<int>(gridDim.x) >> 1;
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int total = p.m_tiles * ((p.N / kBM) >> 1);
  const int my_tiles = (total - pair + n_pairs - 1) / n_pairs;  // >= 1: the host launches at most `total` pairs
  const int n_kb = p.kb_per_tile;
  qs_trace(QS_K_GEMM, 0);
  if (threadIdx.x == 0) QS_PROF(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_act);
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < WS; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 8); }
    for (int i = 0; i < AS; ++i) { mbar_init(&bar_xfull[i], 1); mbar_init(&bar_xempty[i], 1); }
    for (int i = 0; i < TA; ++i) { mbar_init(&bar_afull[i], 16); mbar_init(&bar_aempty[i], 1); }
    mbar_init(bar_dfull, 2);
    mbar_init(bar_zero, 16);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<C::kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) { pdl_launch_dependents(); QS_PROF(1); }

  // tile t of this pair: the m index runs fastest, so the pairs working at the same time share a few weight bands and the whole activation matrix in L2
  auto tile_m = [&](int lt) { return (pair + lt * n_pairs) % p.m_tiles; };
  auto tile_n = [&](int lt) { return 2 * ((pair + lt * n_pairs) / p.m_tiles) + static_cast<int>(rank); };

  if (warp == 0) {
    // ===================================== weight producer (static data: never waits for the previous kernel) =====================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int n_tile = tile_n(lt);
        const int w_row = n_tile * 4;
        const uint8_t* s2s = p.s2_scales + static_cast<size_t>(n_tile) * kBM;
        const uint8_t* s2z = p.s2_zeros + static_cast<size_t>(n_tile) * kBM;
        for (int it = 0; it < n_kb; ++it) {
          mbar_wait(&bar_wempty[s], ph ^ 1);  // a fresh barrier passes the wait on the "previous" phase
          mbar_expect_tx(&bar_wfull[s], C::kWStageTx);
#pragma unroll
          for (int u = 0; u < kSub; ++u) {
            const int kb = it * kSub + u;
            tma_load_2d(s_w + s * C::kWBytes + u * C::kWSub, &tmap_w, kb * 256, w_row, &bar_wfull[s]);
            if constexpr (MODE == kModeW4Grp) {
              const int kg = kb < p.K / kBK ? kb : p.K / kBK - 1;
              bulk_copy_g2s(s_s2 + s * C::kS2Bytes + u * C::kS2Sub, s2s + static_cast<size_t>(kg) * p.N, kBM, &bar_wfull[s]);
              bulk_copy_g2s(s_s2 + s * C::kS2Bytes + u * C::kS2Sub + kBM, s2z + static_cast<size_t>(kg) * p.N, kBM, &bar_wfull[s]);
            }
          }
          if (++s == WS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================================== activation producer: this CTA's 128 of the tile's 256 tokens =====================================
    if (lane == 0) {
      pdl_wait();
      qs_trace(QS_K_GEMM, 1, 6 * 32);
      int s = 0;
      uint32_t ph = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int a_row = tile_m(lt) * NT + static_cast<int>(rank) * C::NH;
        for (int it = 0; it < n_kb; ++it) {
          mbar_wait(&bar_xempty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&bar_xfull[s], 2 * C::kActBytes);  // my half + the peer's half
#pragma unroll
          for (int u = 0; u < kSub; ++u)
            tma_load_2d_pair(s_act + s * C::kActBytes + u * C::kActSub, &tmap_act, (it * kSub + u) * kBK, a_row, &bar_xfull[s]);
          if (++s == AS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (leader && (warp == 1 || warp == 7)) {
    // ===================================== leader: two MMA issuers (even / odd iterations), 256 x 256 x 32 INT8 UMMAs over both SMs ==========
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_i8(2 * kBM, NT, 1u, 1u);
      const int g_end = my_tiles * n_kb;
      int it = (warp == 1) ? 0 : 1;  // iteration inside the current tile
      int lt = 0;
      bool fresh = true;
      for (int g = it; g < g_end; g += 2) {
        if (fresh) {
          mbar_wait(bar_zero, lt & 1);  // accumulators of both CTAs drained (previous tile) and zero-filled
          fresh = false;
          if (g == 0) QS_PROF(2);
          if (g == n_kb) QS_PROF(6);
        }
        const int sa = g & (TA - 1), sx = g & (AS - 1);
        mbar_wait(&bar_afull[sa], (g / TA) & 1);  // both CTAs' unpack warps have written their A stage
        if (g == 0) QS_PROF(3);
        mbar_wait(&bar_xfull[sx], (g / AS) & 1);          // both activation halves
        if (g == 0) QS_PROF(4);
        if (g == 8) QS_PROF(5);
        if (g == 14) QS_PROF(13);
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < kSub; ++u) {
          const uint64_t bdesc = umma_desc_sw128(smem_u32(s_act + sx * C::kActBytes + u * C::kActSub));
#pragma unroll
          for (int t = 0; t < kBK / 32; ++t)
            umma_i8_ts_2cta(tmem_base, tmem_base + NT + sa * C::kAStageCols + u * (kBK / 4) + t * 8, bdesc + t * 2, idesc, 1u);
        }
        umma_commit_2cta(&bar_aempty[sa]);
        umma_commit_2cta(&bar_xempty[sx]);
        it += 2;
        if (it >= n_kb) {  // that was this issuer's last iteration of the tile (n_kb >= 2: both issuers own at least one)
          umma_commit_2cta(bar_dfull);
          if (lt == 0 && warp == 1) QS_PROF(7);
          it -= n_kb;
          ++lt;
          fresh = true;
        }
      }
    }
  } else if ((warp >= 2 && warp <= 5) || warp >= 8) {
    // ===================================== unpack + epilogue: warps 2..5 and 8..11, both CTAs =====================================
    // A warp reaches the TMEM lanes 32 * (warp % 4) ..: the two warps of a lane quadrant split every stage's two 128-K sub-blocks (unpack) and the
    // 256 token columns (epilogue).  (Per-group mode is bound by the integer work of the level-2 dequantisation in this loop: 4 -> 8 warps.)
    const bool unpacker = warp < 8;           // first / second warp of its lane quadrant
    const int quad = warp & 3;
    const int usel = unpacker ? 0 : 1;        // the 128-K sub-block of a stage this warp unpacks
    const int col0 = unpacker ? 0 : NT / 2;   // first accumulator column (token) of this warp's half
    const int epi_tid = quad * 32 + lane;     // channel inside the tile
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    auto zero_fill = [&]() {
#pragma unroll
      for (int c = 0; c < NT / 2; c += 8) {
        tmem_st_16x128b_x2(trow + col0 + c, 0u, 0u, 0u, 0u);
        tmem_st_16x128b_x2(trow + col0 + c + (16u << 16), 0u, 0u, 0u, 0u);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_zero, 0);
    };
    zero_fill();
    int s = 0, ta = 0;
    uint32_t ph = 0, pha = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int m_tile = tile_m(lt), n_tile = tile_n(lt);
      {
        for (int it = 0; it < n_kb; ++it) {
          mbar_wait(&bar_wfull[s], ph);
          mbar_wait(&bar_aempty[ta], pha ^ 1);
          tc_fence_after();
          static_assert(kSub == 2, "two warps per lane quadrant take one sub-block each");
          {
            const int u = usel;
            const uint8_t* wsrc = s_w + s * C::kWBytes + u * C::kWSub + quad * 2048 + lane * 16;
            const uint32_t tdst = trow + NT + ta * C::kAStageCols + u * (kBK / 4);
            uint32_t sc4 = 0, zp4 = 0;
            if constexpr (MODE == kModeW4Grp) {
              const uint8_t* s2 = s_s2 + s * C::kS2Bytes + u * C::kS2Sub + quad * 32 + (lane >> 2) * 4;
              sc4 = *reinterpret_cast<const uint32_t*>(s2);
              zp4 = *reinterpret_cast<const uint32_t*>(s2 + kBM);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint4 v = *reinterpret_cast<const uint4*>(wsrc + t * 512);
              uint32_t xl = v.x & 0x0F0F0F0Fu, xh = (v.x >> 4) & 0x0F0F0F0Fu;
              uint32_t yl = v.y & 0x0F0F0F0Fu, yh = (v.y >> 4) & 0x0F0F0F0Fu;
              uint32_t zl = v.z & 0x0F0F0F0Fu, zh = (v.z >> 4) & 0x0F0F0F0Fu;
              uint32_t wl = v.w & 0x0F0F0F0Fu, wh = (v.w >> 4) & 0x0F0F0F0Fu;
              if constexpr (MODE == kModeW4Grp) {
                const uint32_t s0 = sc4 & 0xFF, s1 = (sc4 >> 8) & 0xFF, s2 = (sc4 >> 16) & 0xFF, s3 = sc4 >> 24;
                const uint32_t z0 = __byte_perm(zp4, 0, 0x0000), z1 = __byte_perm(zp4, 0, 0x1111);
                const uint32_t z2 = __byte_perm(zp4, 0, 0x2222), z3 = __byte_perm(zp4, 0, 0x3333);
                xl = __vadd4(xl * s0, z0); zl = __vadd4(zl * s0, z0);
                yl = __vadd4(yl * s1, z1); wl = __vadd4(wl * s1, z1);
                xh = __vadd4(xh * s2, z2); zh = __vadd4(zh * s2, z2);
                yh = __vadd4(yh * s3, z3); wh = __vadd4(wh * s3, z3);
              }
              tmem_st_16x128b_x2(tdst + t * 8, xl, yl, zl, wl);
              tmem_st_16x128b_x2(tdst + t * 8 + (16u << 16), xh, yh, zh, wh);
            }
          }
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(&bar_afull[ta], 0);  // the leader's barrier counts the unpack warps of both CTAs
            mbar_arrive(&bar_wempty[s]);
          }
          if (++s == WS) { s = 0; ph ^= 1; }
          if (++ta == TA) { ta = 0; pha ^= 1; }
          if (epi_tid == 0 && lt == 0 && it == 8) QS_PROF(14);
        }
        if (epi_tid == 0 && lt == 0) QS_PROF(8);
      }
      // ------------------------------ epilogue: this CTA's 128 channels x 256 tokens, TMEM -> registers -> global ------------------------------
      // thread = one channel (TMEM lane) x 128 tokens; a warp's store covers 32 consecutive channels of one token = 64 B (two full sectors)
      pdl_wait();
      const int m0 = m_tile * NT;
      float* s_asc = s_row + (lt & 1) * 2 * NT;  // double-buffered by tile parity: a fast warp may already fill the next tile's values
      float* s_asum = s_asc + NT;
      {
        const int j = col0 + epi_tid;  // the eight epilogue warps fetch one token's scales each
        const bool ok = (m0 + j) < p.M;
        s_asc[j] = ok ? __half2float(p.ascales[m0 + j]) : 0.f;
        if constexpr (MODE == kModeW4Chn) s_asum[j] = ok ? __half2float(p.a_ssums[m0 + j]) : 0.f;
      }
      const int n = n_tile * kBM + epi_tid;
      const float ws = __half2float(__ldg(p.wscales + n));
      float wz = 0.f;
      if constexpr (MODE == kModeW4Chn) wz = __half2float(__ldg(p.w_szs + n));
      named_bar_sync(1, 256);  // the eight epilogue warps: per-token scales visible
      mbar_wait(bar_dfull, lt & 1);  // every MMA of the tile has retired
      if (epi_tid == 0 && lt == 0 && unpacker) QS_PROF(9);
      tc_fence_after();
      const int tok_end = min(NT, p.M - m0);
#pragma unroll 1
      for (int c = 0; c < NT / 64; ++c) {
        const int tok0 = col0 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(trow + tok0, r);
        tmem_wait_ld();
        __half* out = p.out + static_cast<size_t>(m0 + tok0) * p.N + n;
        if (tok0 + 32 <= tok_end) {
          // whole chunk inside M: no per-token branches, the scales come in as vectors, 32 independent chains
          float as[32], am[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(s_asc + tok0 + i);
            as[i] = a4.x; as[i + 1] = a4.y; as[i + 2] = a4.z; as[i + 3] = a4.w;
            if constexpr (MODE == kModeW4Chn) {
              const float4 m4 = *reinterpret_cast<const float4*>(s_asum + tok0 + i);
              am[i] = m4.x; am[i + 1] = m4.y; am[i + 2] = m4.z; am[i + 3] = m4.w;
            } else {
              am[i] = am[i + 1] = am[i + 2] = am[i + 3] = 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float o = epilogue_one<MODE, kPairCvt>(static_cast<int32_t>(r[i]), ws, wz, as[i], am[i]);
            out[static_cast<size_t>(i) * p.N] = __float2half_rn(o);
            if constexpr (ACC) p.acc_out[static_cast<size_t>(m0 + tok0 + i) * p.N + n] = static_cast<int32_t>(r[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {  // (fully unrolled: a dynamic index would push r[] into local memory for the fast path as well)
            if (tok0 + i < tok_end) {
              const float o = epilogue_one<MODE, kPairCvt>(static_cast<int32_t>(r[i]), ws, wz, s_asc[tok0 + i], s_asum[tok0 + i]);
              out[static_cast<size_t>(i) * p.N] = __float2half_rn(o);
              if constexpr (ACC) p.acc_out[static_cast<size_t>(m0 + tok0 + i) * p.N + n] = static_cast<int32_t>(r[i]);
            }
          }
        }
      }
      if (epi_tid == 0 && lt == 0 && unpacker) QS_PROF(10);
      if (lt + 1 < my_tiles) zero_fill();  // hands the accumulator back to the MMA issuers
      if (epi_tid == 0 && lt == 0 && unpacker) QS_PROF(11);
    }
  } else {
    pdl_wait();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still read this CTA's activation half / arrive on its barriers until its own last MMA has retired
  qs_trace(QS_K_GEMM, 2);
  if (threadIdx.x == 0) QS_PROF(12);
  if (warp == 1) tmem_dealloc_2cta<C::kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Tensor maps are pure functions of (address, shape, box): cuTensorMapEncodeTiled costs ~1 us of host time and an eager decode step issues
// ~260 of them, so they are memoised (weights: one entry per layer and projection; activations: a handful of buffers).
struct TmapKey {
  const void* ptr; uint64_t a, b; uint32_t box, kind;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && a == o.a && b == o.b && box == o.box && kind == o.kind; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (k.a + 0x7F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.b * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2));
    return h ^ (static_cast<size_t>(k.box) << 7) ^ k.kind;
  }
};
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>& tmap_cache() {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> c;
  return c;
}
std::mutex& tmap_mutex() { static std::mutex m; return m; }
bool tmap_lookup(const TmapKey& k, CUtensorMap* m) {
  std::lock_guard<std::mutex> g(tmap_mutex());
  auto it = tmap_cache().find(k);
  if (it == tmap_cache().end()) return false;
  memcpy(m, &it->second, sizeof(CUtensorMap));
  return true;
}
void tmap_store(const TmapKey& k, const CUtensorMap* m) {
  std::lock_guard<std::mutex> g(tmap_mutex());
  if (tmap_cache().size() > 16384) tmap_cache().clear();  // bounded: a serving loop cycles through a fixed set of buffers
  memcpy(&tmap_cache()[k], m, sizeof(CUtensorMap));
}

// 2-D uint8 tensor [rows, cols] row-major, box {128 bytes, box_rows}, 128-byte swizzle, zero fill out of bounds
int make_tmap_u8(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const TmapKey key{ptr, rows, cols, box_rows, 0u};
  if (tmap_lookup(key, m)) return QS_OK;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%llu cols=%llu box_rows=%u", (int)r, ptr,
                                          (unsigned long long)rows, (unsigned long long)cols, box_rows);
  tmap_store(key, m);
  return QS_OK;
}

// packed INT4 weights [N, K/2] seen as [N/32 bands][K*16 bytes] of uint64 elements; box = 4 bands x 2 KB (one 128-K block)
int make_tmap_w4(CUtensorMap* m, const void* ptr, uint64_t N, uint64_t K, uint32_t box_bands = 4) {
  const TmapKey key{ptr, N, K, box_bands, 1u};
  if (tmap_lookup(key, m)) return QS_OK;
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {K * 16 / 8, N / 32};
  cuuint64_t strides[1] = {K * 16};
  cuuint32_t box[2] = {256, box_bands};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_INT64, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled(w4) failed (%d): ptr=%p N=%llu K=%llu", (int)r, ptr, (unsigned long long)N,
                                          (unsigned long long)K);
  tmap_store(key, m);
  return QS_OK;
}

int choose_split(int tiles, int kb_per_tile, int forced) {
  int s = 1;
  if (forced > 0) {
    s = forced;
  } else {
    const int sms = num_sms();
    // smallest power of two that brings the CTA count to >= ~2/3 of the SMs; every CTA keeps >= 2 k-blocks
    while (s < 8 && tiles * s < (2 * sms) / 3 && kb_per_tile / (2 * s) >= 1) s *= 2;
  }
  if (s > 8) s = 8;
  while (s > 1 && kb_per_tile < s) s /= 2;
  // split launches carry a receive buffer and run one CTA per SM: never more CTAs than SMs (a second wave costs more than it saves)
  if (forced <= 0) while (s > 1 && tiles * s > num_sms()) s /= 2;
  return s;
}

template <int MODE, int NT, int WS, int AS>
int launch_gemm(const GemmArgs& a) {
  using C = Cfg<MODE, NT, WS, AS>;
  GemmParams p{};
  p.s2_scales = static_cast<const uint8_t*>(a.s2_scales);
  p.s2_zeros = static_cast<const uint8_t*>(a.s2_zeros);
  p.wscales = static_cast<const __half*>(a.wscales);
  p.w_szs = static_cast<const __half*>(a.w_szs);
  p.ascales = static_cast<const __half*>(a.ascales);
  p.a_ssums = static_cast<const __half*>(a.a_ssums);
  p.out = static_cast<__half*>(a.out);
  p.acc_out = static_cast<int32_t*>(a.acc_out);
  p.prof = static_cast<unsigned long long*>(a.prof);
  p.M = a.M; p.N = a.N; p.K = a.K;
  const int n_tiles = a.N / kBM;
  p.m_tiles = (a.M + NT - 1) / NT;
  p.kb_per_tile = (a.K + kSub * kBK - 1) / (kSub * kBK);  // pipeline stages of 256 K (the last one may be half zero-filled)
  const int tiles = n_tiles * p.m_tiles;
  p.split = C::kSplitFits ? choose_split(tiles, p.kb_per_tile, a.force_split) : 1;  // NT = 256: no room for the receive buffer

  CUtensorMap tm_act, tm_w;
  int rc = make_tmap_u8(&tm_act, a.act, a.M, a.K, NT);
  if (rc) return rc;
  rc = (MODE == kModeW8) ? make_tmap_u8(&tm_w, a.weight, a.N, a.K, kBM) : make_tmap_w4(&tm_w, a.weight, a.N, a.K);
  if (rc) return rc;

  auto kern = a.acc_out ? gemm_kernel<MODE, NT, WS, AS, true> : gemm_kernel<MODE, NT, WS, AS, false>;
  static bool attr_set[2][kMaxDevices] = {};  // per (instantiation, device): the attribute is a per-device property of the function
  bool& done = attr_set[a.acc_out ? 1 : 0][device_ordinal()];
  if (!done) {
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSplitFits ? C::kSmemBytesSplit : C::kSmemBytes),
                    "cudaFuncSetAttribute(gemm smem)");
    if (rc) return rc;
    done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles * p.split);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = p.split > 1 ? C::kSmemBytesSplit : C::kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(a.stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = p.split;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tm_act, tm_w, p), "gemm launch");
}

// prefill: CTA pairs (cta_group::2), 256 channels x 256 tokens per pair
template <int MODE>
int launch_gemm_pair(const GemmArgs& a) {
  using C = PairCfg<MODE>;
  GemmParams p{};
  p.s2_scales = static_cast<const uint8_t*>(a.s2_scales);
  p.s2_zeros = static_cast<const uint8_t*>(a.s2_zeros);
  p.wscales = static_cast<const __half*>(a.wscales);
  p.w_szs = static_cast<const __half*>(a.w_szs);
  p.ascales = static_cast<const __half*>(a.ascales);
  p.a_ssums = static_cast<const __half*>(a.a_ssums);
  p.out = static_cast<__half*>(a.out);
  p.acc_out = static_cast<int32_t*>(a.acc_out);
  p.M = a.M; p.N = a.N; p.K = a.K;
  const int n_tiles = a.N / kBM;
  p.m_tiles = (a.M + C::NT - 1) / C::NT;
  p.kb_per_tile = (a.K + kSub * kBK - 1) / (kSub * kBK);
  p.split = 1;
  p.prof = static_cast<unsigned long long*>(a.prof);
  CUtensorMap tm_act, tm_w;
  int rc = make_tmap_u8(&tm_act, a.act, a.M, a.K, C::NH);
  if (rc) return rc;
  rc = make_tmap_w4(&tm_w, a.weight, a.N, a.K);
  if (rc) return rc;
  auto kern = a.acc_out ? gemm_pair_kernel<MODE, true> : gemm_pair_kernel<MODE, false>;
  static bool attr_set[2][kMaxDevices] = {};
  bool& done = attr_set[a.acc_out ? 1 : 0][device_ordinal()];
  if (!done) {
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes), "cudaFuncSetAttribute(gemm pair smem)");
    if (rc) return rc;
    done = true;
  }
  cudaLaunchConfig_t cfg{};
  // persistent: one CTA pair per TPC walks the (m fastest) tile list with stride = number of pairs
  const int total_pairs = (n_tiles / 2) * p.m_tiles;
  cfg.gridDim = dim3(2 * std::min(total_pairs, num_sms() / 2));
  cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(a.stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tm_act, tm_w, p), "gemm (pair) launch");
}

template <int MODE>
int dispatch_gemm(const GemmArgs& a) {
  QS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  QS_REQUIRE(a.N % kBM == 0, "gemm: N=%d must be a multiple of %d", a.N, kBM);
  QS_REQUIRE(a.K % kBK == 0, "gemm: K=%d must be a multiple of %d", a.K, kBK);
  QS_REQUIRE((reinterpret_cast<uintptr_t>(a.act) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.weight) & 15) == 0, "gemm: operands must be 16-byte aligned");
  QS_REQUIRE(a.force_split == 0 || a.force_split == 1 || a.force_split == 2 || a.force_split == 4 || a.force_split == 8, "gemm: split must be 1, 2, 4 or 8");
  // weight-ring depths (256-K stages) chosen so that NT <= 128 fits two CTAs per SM (<= 113 KB smem, <= 256 TMEM columns)
  constexpr bool w8 = (MODE == kModeW8), grp = (MODE == kModeW4Grp);
  QS_REQUIRE(a.force_nt == 0 || a.force_nt == 32 || a.force_nt == 64 || a.force_nt == 128, "gemm: tile tokens must be 32, 64 or 128");
  if constexpr (MODE != kModeW8) {
    static const bool no_pair = getenv("QS_GEMM_NO_PAIR") != nullptr;  // A/B hooks
    static const int pair_min_m = getenv("QS_GEMM_PAIR_MIN_M") ? atoi(getenv("QS_GEMM_PAIR_MIN_M")) : 512;
    // Below 512 tokens the pair kernel pays only if its 256-token tiles are as well filled as the 128-token ones (an even number of 128-token
    // tiles: M in (128, 256] or (384, 512)) and there is at least one pair tile per TPC (measured, profiles/r02_notes.md 4b: gate_up at M = 256
    // 30.4 vs 35.5 us, at M = 320 58.1 vs 50.5; o_proj (16 pair tiles) at M = 256 16.6 vs 12.0)
    const int t128 = (a.M + 127) / 128;
    const bool small_ok = a.M > 128 && t128 % 2 == 0 && (a.N / kBM / 2) * ((a.M + 255) / 256) >= num_sms() / 2;
    if (!no_pair && a.force_nt == 0 && a.force_split == 0 && (a.M >= pair_min_m || small_ok) && (a.N / kBM) % 2 == 0 && a.K >= 512) return launch_gemm_pair<MODE>(a);
  }
  int nt = a.force_nt > 0 ? a.force_nt : 0;
  if (nt == 0) {
    // prefill-sized M: 128-token tiles with TWO co-resident CTAs per SM beat the 256-token tile (one CTA per SM) by 14-15 % (M = 1024: 2071 vs
    // 1796 TOP/s, M = 4096: 2388 vs 2100, profiles/r02_notes.md): a tile's setup, pipeline fill and epilogue hide behind the sibling's main loop
    nt = a.M <= 32 ? 32 : a.M <= 64 ? 64 : 128;
  }
  if (nt == 32) return launch_gemm<MODE, 32, (w8 ? 2 : grp ? 4 : 5), 4>(a);
  if (nt == 64) return launch_gemm<MODE, 64, (w8 ? 2 : grp ? 3 : 4), (w8 ? 2 : 3)>(a);
  if constexpr (w8) {
    // W8A8 at 65..128 tokens: the 4-deep weight ring (4 x 32 KB) leaves no room for the split-K receive buffer, so the narrow layers of a batch-128
    // step ran on 32-48 CTAs (down_proj 35 us = 0.25 of the HBM peak).  Layers that want a split take a 2-deep ring (even depth: see the kernel).
    const int tiles = (a.N / kBM) * ((a.M + 127) / 128);
    if (a.force_split != 1 && tiles * 2 <= num_sms()) return launch_gemm<MODE, 128, 2, 2>(a);
  }
  return launch_gemm<MODE, 128, (w8 ? 4 : 2), 2>(a);
}

}  // namespace

int gemm_w4a8_per_chn(const GemmArgs& a) { return dispatch_gemm<kModeW4Chn>(a); }
int gemm_w4a8_per_group(const GemmArgs& a) { return dispatch_gemm<kModeW4Grp>(a); }
int gemm_w8a8(const GemmArgs& a) { return dispatch_gemm<kModeW8>(a); }
size_t gemm_workspace_bytes() { return 4096; }
int gemm_trace_install(void* buf, unsigned cap) { return qs_trace_install(buf, cap); }  // the cluster/DSMEM split-K needs no global workspace; kept for ABI stability

}  // namespace qs
