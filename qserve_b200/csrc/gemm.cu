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
//     The TMA engine (cp.async.bulk) stages 4 bands x 2 KB per 128-K block into shared memory; four unpack warps
//     read one 16-byte lane chunk each (LDS.128), split nibbles in registers (per-group: level-2 dequant
//     q*s2+z2 with the reference's exact 32-bit multiply + vadd4), and write INT8 rows straight into TENSOR MEMORY
//     with tcgen05.st.16x128b -- the mma.sync B-fragment order of the checkpoint is exactly the 16x128b TMEM store
//     pattern, so no shuffle and no shared-memory round trip is needed.
//   * one elected thread issues tcgen05.mma.kind::i8 with A (weights) from TMEM and B (INT8 activations, TMA-loaded
//     with the 128-byte swizzle) from shared memory; INT32 accumulators live in TMEM.
//   * stream-K style decomposition over (tile, k-block) units so that every SM streams an equal share of the weight
//     bytes; INT32 partial tiles are exchanged through an L2-resident workspace and summed by the last-arriving CTA
//     (integer adds: the result is bit-identical for every decomposition).
//   * epilogue fused: acc*s1[n]*sa[m] - s1z[n]*asum[m] (per-channel) or acc*(s1[n]*sa[m]) (per-group, W8A8) -> fp16.
//   * programmatic dependent launch: the weight prefetch is issued before griddepcontrol.wait, so HBM keeps streaming
//     while the preceding activation-quant kernel drains.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "launch.h"

namespace qs {

namespace {

constexpr int kBM = 128;  // output channels per tile (UMMA M)
constexpr int kBK = 128;  // K per pipeline stage (= one g128 group)
constexpr int kNumThreads = 192;
constexpr int kEpiThreads = 128;

enum { kModeW4Chn = 0, kModeW4Grp = 1, kModeW8 = 2 };

struct GemmParams {
  const uint8_t* qweight;    // W4: packed [N, K/2]; W8: unused (tensor map)
  const uint8_t* s2_scales;  // per-group: [K/128, N] (shuffled per 32 columns, as stored in the checkpoint)
  const uint8_t* s2_zeros;   // per-group: [K/128, N]
  const __half* wscales;     // [N]
  const __half* w_szs;       // [N]  (per-channel only)
  const __half* ascales;     // [M]
  const __half* a_ssums;     // [M]  (per-channel only)
  __half* out;               // [M, N]
  int32_t* acc_out;          // optional: raw INT32 accumulators [M, N] (parity tests)
  int32_t* ws_partials;
  uint32_t* ws_counters;
  int M, N, K;
  int m_tiles, kb_per_tile, total_units, units_per_cta, max_contrib;
};

template <int MODE, int NT, int STAGES>
struct Cfg {
  static constexpr int kActBytes = NT * kBK;                                    // int8 activations, 128 B rows, swizzled
  static constexpr int kWBytes = (MODE == kModeW8) ? kBM * kBK : kBM * kBK / 2; // int8 rows or packed int4 tiles
  static constexpr int kS2Bytes = (MODE == kModeW4Grp) ? 2 * kBM : 0;           // scales | zeros for one group
  static constexpr int kStageTx = kActBytes + kWBytes + kS2Bytes;
  static constexpr int kACols = (MODE == kModeW8) ? 0 : STAGES * (kBK / 4);     // TMEM columns of the unpacked-A ring
  static constexpr int kTmemNeed = NT + kACols;
  static constexpr int kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "TMEM overflow");
  // shared memory carve-up (offsets from a 1024-aligned base)
  static constexpr int kOffAct = 0;
  static constexpr int kOffW = kOffAct + STAGES * kActBytes;
  static constexpr int kOffS2 = kOffW + STAGES * kWBytes;
  static constexpr int kOffRow = kOffS2 + STAGES * kS2Bytes;  // float ascales[NT], asums[NT]
  static constexpr int kOffBar = kOffRow + 2 * NT * 4;
  static constexpr int kNumBars = 3 * STAGES + 2;
  static constexpr int kOffMisc = kOffBar + kNumBars * 8;  // tmem ptr, flag
  static constexpr int kSmemBytes = kOffMisc + 16 + 1024;  // + alignment slack
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE>
__device__ __forceinline__ __half epilogue_one(int32_t acc, float ws, float wsz, float as, float asum) {
  // IEEE fp32, reference source order, no FMA contraction (bit-exact against oracle/w4a8.py)
  float ps = __int2float_rn(acc);
  if constexpr (MODE == kModeW4Chn) {
    // w4a8_per_chn/gemm_cuda.cu:586  psum * wscale * ascale - w_sz * a_ssum
    float t = __fmul_rn(__fmul_rn(ps, ws), as);
    float u = __fmul_rn(wsz, asum);
    return __float2half_rn(__fsub_rn(t, u));
  } else {
    // w4a8_per_group/gemm_cuda.cu:619, w8a8_gemm_cuda.cu:522   psum *= wscale * ascale
    return __float2half_rn(__fmul_rn(ps, __fmul_rn(ws, as)));
  }
}

template <int MODE, int NT, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  using C = Cfg<MODE, NT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_act = smem + C::kOffAct;
  uint8_t* s_w = smem + C::kOffW;
  uint8_t* s_s2 = smem + C::kOffS2;
  float* s_asc = reinterpret_cast<float*>(smem + C::kOffRow);
  float* s_asum = s_asc + NT;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* bar_afull = bar_full + STAGES;
  uint64_t* bar_empty = bar_afull + STAGES;
  uint64_t* bar_dfull = bar_empty + STAGES;
  uint64_t* bar_dempty = bar_dfull + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kOffMisc);
  volatile uint32_t* s_flag = s_tmem + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.kb_per_tile;

  const int unit_begin = blockIdx.x * p.units_per_cta;
  const int unit_end = min(unit_begin + p.units_per_cta, p.total_units);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_act);
    if (MODE == kModeW8) tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_afull[i], 4);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_dfull, 1);
    mbar_init(bar_dempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();  // weights are streamed once per step
      auto issue_weights = [&](int u, int it) {
        const int s = it % STAGES;
        const int tile = u / KB, kb = u - tile * KB;
        const int n_tile = tile / p.m_tiles;
        if (it >= STAGES) mbar_wait(&bar_empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_expect_tx(&bar_full[s], C::kStageTx);
        if constexpr (MODE == kModeW8) {
          tma_load_2d(s_w + s * C::kWBytes, &tmap_w, kb * kBK, n_tile * kBM, &bar_full[s]);
        } else {
          // band b of the tile: 4 consecutive 32x32 tiles (2 KB) at ((n32 * K/32) + k32) * 512
          const size_t k32 = static_cast<size_t>(kb) * 4;
          const size_t tiles_per_band = static_cast<size_t>(p.K) / 32;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const size_t n32 = static_cast<size_t>(n_tile) * 4 + b;
            bulk_copy_g2s_hint(s_w + s * C::kWBytes + b * 2048, p.qweight + (n32 * tiles_per_band + k32) * 512, 2048,
                               &bar_full[s], pol_w);
          }
          if constexpr (MODE == kModeW4Grp) {
            const size_t off = static_cast<size_t>(kb) * p.N + static_cast<size_t>(n_tile) * kBM;
            bulk_copy_g2s(s_s2 + s * C::kS2Bytes, p.s2_scales + off, kBM, &bar_full[s]);
            bulk_copy_g2s(s_s2 + s * C::kS2Bytes + kBM, p.s2_zeros + off, kBM, &bar_full[s]);
          }
        }
      };
      auto issue_act = [&](int u, int it) {
        const int s = it % STAGES;
        const int tile = u / KB, kb = u - tile * KB;
        const int m_tile = tile % p.m_tiles;
        tma_load_2d(s_act + s * C::kActBytes, &tmap_act, kb * kBK, m_tile * NT, &bar_full[s]);
      };
      // static weights do not depend on the previous kernel: prefetch a full ring before the PDL wait
      const int n_units = unit_end - unit_begin;
      const int pre = n_units < STAGES ? n_units : STAGES;
      for (int it = 0; it < pre; ++it) issue_weights(unit_begin + it, it);
      pdl_wait();
      for (int it = 0; it < pre; ++it) issue_act(unit_begin + it, it);
      for (int it = pre; it < n_units; ++it) {
        issue_weights(unit_begin + it, it);
        issue_act(unit_begin + it, it);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_i8(kBM, NT, 1u, 1u);
      int it = 0, seg = 0;
      for (int u = unit_begin; u < unit_end; ++seg) {
        const int tile = u / KB, kb0 = u - tile * KB;
        const int kb1 = min(KB, kb0 + (unit_end - u));
        if (seg > 0) {
          mbar_wait(bar_dempty, (seg - 1) & 1);  // epilogue has drained the accumulator of the previous segment
          tc_fence_after();
        }
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&bar_full[s], ph);
          if constexpr (MODE != kModeW8) mbar_wait(&bar_afull[s], ph);
          tc_fence_after();
          const uint64_t bdesc = umma_desc_sw128(smem_u32(s_act + s * C::kActBytes));
#pragma unroll
          for (int t = 0; t < kBK / 32; ++t) {
            const uint32_t acc = (kb > kb0 || t > 0) ? 1u : 0u;
            if constexpr (MODE == kModeW8) {
              const uint64_t adesc = umma_desc_sw128(smem_u32(s_w + s * C::kWBytes));
              umma_i8_ss(tmem_base, adesc + t * 2, bdesc + t * 2, idesc, acc);
            } else {
              umma_i8_ts(tmem_base, tmem_base + NT + s * (kBK / 4) + t * 8, bdesc + t * 2, idesc, acc);
            }
          }
          umma_commit(&bar_empty[s]);
        }
        umma_commit(bar_dfull);
        u += kb1 - kb0;
      }
    }
  } else {
    // ===================================== unpack + epilogue warps =====================================
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int epi_tid = quad * 32 + lane; // 0..127 == channel row inside the tile
    int it = 0, seg = 0;
    bool waited = false;
    for (int u = unit_begin; u < unit_end; ++seg) {
      const int tile = u / KB, kb0 = u - tile * KB;
      const int kb1 = min(KB, kb0 + (unit_end - u));
      const int m_tile = tile % p.m_tiles, n_tile = tile / p.m_tiles;

      if constexpr (MODE != kModeW8) {
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&bar_full[s], (it / STAGES) & 1);
          const uint8_t* wsrc = s_w + s * C::kWBytes + quad * 2048 + lane * 16;
          const uint32_t tdst = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + NT + s * (kBK / 4);
          uint32_t sc4 = 0, zp4 = 0;
          if constexpr (MODE == kModeW4Grp) {
            const uint8_t* s2 = s_s2 + s * C::kS2Bytes + quad * 32 + (lane >> 2) * 4;
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
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_afull[s]);
        }
      }

      // ------------------------------- epilogue of this segment -------------------------------
      if (!waited) {
        pdl_wait();  // ascales / a_ssums / out / workspace belong to the dependency chain
        waited = true;
      }
      const int m0 = m_tile * NT;
      const int n = n_tile * kBM + epi_tid;
      // stage the per-token scales of this token tile
      for (int j = epi_tid; j < NT; j += kEpiThreads) {
        const bool ok = (m0 + j) < p.M;
        s_asc[j] = ok ? __half2float(p.ascales[m0 + j]) : 0.f;
        if constexpr (MODE == kModeW4Chn) s_asum[j] = ok ? __half2float(p.a_ssums[m0 + j]) : 0.f;
      }
      const float ws = __half2float(p.wscales[n]);
      float wsz = 0.f;
      if constexpr (MODE == kModeW4Chn) wsz = __half2float(p.w_szs[n]);

      mbar_wait(bar_dfull, seg & 1);
      tc_fence_after();
      epi_bar_sync();  // s_asc / s_asum visible

      const bool full_k = (kb0 == 0 && kb1 == KB);
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      int ncontrib = 1, first_cta = 0;
      if (!full_k) {
        first_cta = (tile * KB) / p.units_per_cta;
        const int last_cta = ((tile + 1) * KB - 1) / p.units_per_cta;
        ncontrib = last_cta - first_cta + 1;
        int32_t* slot = p.ws_partials + (static_cast<size_t>(tile) * p.max_contrib + (blockIdx.x - first_cta)) * (NT * kBM);
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(trow + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) __stcg(slot + (c * 32 + i) * kBM + epi_tid, static_cast<int32_t>(r[i]));
        }
      }
      bool do_final = full_k;
      if (!full_k) {
        __threadfence();
        epi_bar_sync();
        if (epi_tid == 0) {
          const uint32_t old = atomicAdd(&p.ws_counters[tile], 1u);
          const bool last = (old == static_cast<uint32_t>(ncontrib - 1));
          if (last) p.ws_counters[tile] = 0;  // self-cleaning for the next launch
          *s_flag = last ? 1u : 0u;
        }
        epi_bar_sync();
        do_final = (*s_flag != 0);
        if (do_final) __threadfence();
      }
      if (do_final) {
        const int32_t* slot0 = p.ws_partials + static_cast<size_t>(tile) * p.max_contrib * (NT * kBM);
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t r[32];
          if (full_k) {
            tmem_ld_32x32b_x32(trow + c * 32, r);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0;
            for (int j = 0; j < ncontrib; ++j) {
              const int32_t* sl = slot0 + static_cast<size_t>(j) * (NT * kBM);
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] += static_cast<uint32_t>(__ldcg(sl + (c * 32 + i) * kBM + epi_tid));
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = m0 + c * 32 + i;
            if (m < p.M) {
              const int32_t acc = static_cast<int32_t>(r[i]);
              p.out[static_cast<size_t>(m) * p.N + n] = epilogue_one<MODE>(acc, ws, wsz, s_asc[c * 32 + i], s_asum[c * 32 + i]);
              if (p.acc_out) p.acc_out[static_cast<size_t>(m) * p.N + n] = acc;
            }
          }
        }
      }
      // accumulator drained: release it to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dempty);
      epi_bar_sync();  // s_asc/s_asum/s_flag reuse across segments
      u += kb1 - kb0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
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

// 2-D uint8 tensor [rows, cols] row-major, box {128 bytes, box_rows}, 128-byte swizzle, zero fill out of bounds
int make_tmap_u8(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
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
  return QS_OK;
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

constexpr size_t kCounterBytes = 64 * 1024;  // 16384 tile counters

template <int MODE, int NT, int STAGES>
int launch_gemm(const GemmArgs& a) {
  using C = Cfg<MODE, NT, STAGES>;
  GemmParams p{};
  p.qweight = static_cast<const uint8_t*>(a.weight);
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
  p.m_tiles = (a.M + NT - 1) / NT;
  p.kb_per_tile = a.K / kBK;
  const int tiles = n_tiles * p.m_tiles;
  p.total_units = tiles * p.kb_per_tile;

  // ---- decomposition: equal shares of k-blocks per CTA (stream-K) when the tile count does not fill the machine ----
  const int sms = num_sms();
  int upc = p.kb_per_tile;  // default: one whole tile per CTA
  int max_contrib = 1;
  const size_t slot_bytes = static_cast<size_t>(NT) * kBM * 4;
  if (a.force_units_per_cta > 0) {
    upc = a.force_units_per_cta;
  } else if (tiles < 4 * sms) {
    const int target = (tiles <= sms) ? sms : ((tiles + sms - 1) / sms) * sms;  // CTAs
    upc = (p.total_units + target - 1) / target;
    if (upc < 2) upc = 2;  // keep at least two k-blocks per CTA
    if (upc > p.kb_per_tile) upc = ((upc + p.kb_per_tile - 1) / p.kb_per_tile) * p.kb_per_tile;
  } else {
    // plenty of tiles: whole tiles, several per CTA, persistent-style
    const int per = (tiles + 2 * sms - 1) / (2 * sms);
    upc = per * p.kb_per_tile;
  }
  if (upc % p.kb_per_tile != 0) {
    max_contrib = (p.kb_per_tile + upc - 1) / upc + 1;
    const size_t need = kCounterBytes + static_cast<size_t>(tiles) * max_contrib * slot_bytes;
    if (tiles > static_cast<int>(kCounterBytes / 4) || a.workspace == nullptr || need > a.workspace_bytes) {
      if (a.force_units_per_cta > 0) return set_error(QS_ERR_WORKSPACE, "gemm workspace too small: need %zu have %zu", need, a.workspace_bytes);
      upc = p.kb_per_tile;  // fall back to whole tiles (no exchange)
      max_contrib = 1;
    }
  }
  p.units_per_cta = upc;
  p.max_contrib = max_contrib;
  p.ws_counters = static_cast<uint32_t*>(a.workspace);
  p.ws_partials = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(a.workspace) + kCounterBytes);
  const int grid = (p.total_units + upc - 1) / upc;

  CUtensorMap tm_act, tm_w;
  int rc = make_tmap_u8(&tm_act, a.act, a.M, a.K, NT);
  if (rc) return rc;
  if (MODE == kModeW8) {
    rc = make_tmap_u8(&tm_w, a.weight, a.N, a.K, kBM);
    if (rc) return rc;
  } else {
    tm_w = tm_act;
  }

  auto kern = gemm_kernel<MODE, NT, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes), "cudaFuncSetAttribute(gemm smem)");
    if (rc) return rc;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(a.stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, kern, tm_act, tm_w, p), "gemm launch");
}

template <int MODE>
int dispatch_gemm(const GemmArgs& a) {
  QS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  QS_REQUIRE(a.N % kBM == 0, "gemm: N=%d must be a multiple of %d", a.N, kBM);
  QS_REQUIRE(a.K % kBK == 0, "gemm: K=%d must be a multiple of %d", a.K, kBK);
  QS_REQUIRE((reinterpret_cast<uintptr_t>(a.act) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.weight) & 15) == 0, "gemm: operands must be 16-byte aligned");
  if (a.M <= 32) return launch_gemm<MODE, 32, 8>(a);
  if (a.M <= 64) return launch_gemm<MODE, 64, 8>(a);
  if (a.M <= 128) return launch_gemm<MODE, 128, (MODE == kModeW8 ? 6 : 8)>(a);
  return launch_gemm<MODE, 256, (MODE == kModeW8 ? 4 : 5)>(a);
}

}  // namespace

int gemm_w4a8_per_chn(const GemmArgs& a) { return dispatch_gemm<kModeW4Chn>(a); }
int gemm_w4a8_per_group(const GemmArgs& a) { return dispatch_gemm<kModeW4Grp>(a); }
int gemm_w8a8(const GemmArgs& a) { return dispatch_gemm<kModeW8>(a); }
size_t gemm_workspace_bytes() { return kCounterBytes + static_cast<size_t>(96) * 1024 * 1024; }

}  // namespace qs
