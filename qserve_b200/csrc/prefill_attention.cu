// qserve_b200 -- causal variable-length prefill attention on the sm_100a tensor cores (tcgen05 kind::f16, accumulators and P in tensor memory).
//
// Replaces the third-party call on the reference's prompt path: flash_attn_varlen_func(q, k, v, cu_seqlens, cu_seqlens, max_seqlen, max_seqlen,
// dropout_p=0, causal=True) in qserve/modeling/models/llama_w4a8_unpad.py:232-242, where q / k / v are strided views of the post-RoPE fp16 qkv
// buffer that fused_attention.apply_bias_rope_update_kv_cache has just rotated in place.
//
// One CTA = one (sequence, query head, block of 128 query rows).  Per block of 128 keys:
//     S = Q K^T      8 x tcgen05.mma 128x128x16, both operands K-major in shared memory (TMA, 128-byte swizzle), fp32 accumulators in TMEM
//     softmax        4 warps, thread = query row (TMEM lane): running max / sum in the log2 domain, P = exp2(..) rounded to fp16 and written
//                    back into tensor memory OVER the S columns it came from (P is the A operand of the next MMA)
//     O += P V       8 x tcgen05.mma 128x128x16, A from tensor memory, B = the V tile exactly as TMA delivers it ([key][dim] rows: an MN-major
//                    operand), accumulating in a second TMEM region
// The O accumulator is rescaled only when a row's maximum has grown by more than 2^8 since the last rescale (the sum uses the same stale
// maximum, so the quotient is exact up to rounding); the decision is taken per warp because tcgen05.ld/st are warp-collective.
// 256 TMEM columns and 96 KB of shared memory per CTA: two CTAs per SM, one's softmax overlaps the other's MMAs.
// K and V are single-buffered, each re-filled as soon as the MMA that read it has retired (tcgen05.commit -> mbarrier -> TMA warp).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <mutex>

#include "common.cuh"
#include "launch.h"

namespace qs {
namespace {

constexpr int kD = 128;         // head dim
constexpr int kBQ = 128;        // query rows per CTA (UMMA M)
constexpr int kBKV = 128;       // keys per block (UMMA N of S, K extent of PV)
constexpr int kThreads = 192;   // warps 0..3 softmax / epilogue, warp 4 TMA producer, warp 5 TMEM allocator + MMA issuer
constexpr int kTileBytes = kBQ * kD * 2;      // 32 KB: two swizzled [128 rows x 64 halfs] sub-tiles
constexpr int kSubBytes = kTileBytes / 2;     // 16 KB
constexpr int kOffQ = 0, kOffK = kTileBytes, kOffV = 2 * kTileBytes, kOffBar = 3 * kTileBytes;
constexpr int kSmemBytes = kOffBar + 128;
constexpr uint32_t kTmemCols = 256;           // S / P: columns [0, 128), O: columns [128, 256)
constexpr float kRescaleThreshold = 8.f;      // log2 units

// kind::f16 instruction descriptor: D = f32, A = B = f16, A K-major; B K-major (S = Q K^T) or MN-major (O = P V)
//   bits [4,6) c_format (1 = F32)   [7,10) a_format (0 = F16)   [10,13) b_format   [15] a_major   [16] b_major (1 = MN)
//   bits [17,23) N >> 3             [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t m, uint32_t n, uint32_t b_mn_major) {
  return (1u << 4) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// 2^x on the SFU (MUFU.EX2), flush-to-zero: the argument is <= 8 here and -inf must give 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major shared-memory operand under the 128-byte swizzle: rows of 128 B are 64 contiguous elements of the M/N dimension, one row per K
// index.  In 16-byte units the canonical layout is ((8, n), (8, k)) : ((1, LBO), (8, SBO)): LBO = distance between 64-element column groups
// (here the second 16 KB sub-tile), SBO = distance between groups of 8 K-rows (1024 B).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// thread i of the warp writes 32 consecutive 32-bit columns of ITS lane (lane = 32 * (warp % 4) + i)
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
struct PrefillAttnParams {
  const int* cu_seqlens;  // [B + 1]
  __half* out;            // [T, Hq * 128] rows of out_stride halfs
  long long out_stride;
  int num_heads, num_kv_heads;
  float scale_log2;       // softmax scale * log2(e)
};

__global__ void __launch_bounds__(kThreads, 2)
prefill_attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const PrefillAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq_start = __ldg(p.cu_seqlens + b);
  const int seq_len = __ldg(p.cu_seqlens + b + 1) - seq_start;
  const int n_qb = (seq_len + kBQ - 1) / kBQ;
  const int qb = n_qb - 1 - static_cast<int>(blockIdx.x);  // the longest (last) query blocks of a sequence are scheduled first
  if (qb < 0) return;                                      // uniform per CTA: nothing has been allocated yet
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
  const int hkv = h / (p.num_heads / p.num_kv_heads);
  const int n_kb = qb + 1;  // causal, q and k positions coincide: key blocks 0 .. qb

  uint8_t* s_q = smem + kOffQ;
  uint8_t* s_k = smem + kOffK;
  uint8_t* s_v = smem + kOffV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_q = bar + 0;       // Q tile landed
  uint64_t* bar_kfull = bar + 1;   // K block landed
  uint64_t* bar_kfree = bar + 2;   // S = Q K^T retired: the K buffer may be refilled
  uint64_t* bar_vfull = bar + 3;
  uint64_t* bar_vfree = bar + 4;   // O += P V retired: the V buffer may be refilled
  uint64_t* bar_s = bar + 5;       // S complete in tensor memory
  uint64_t* bar_p = bar + 6;       // P written by the four softmax warps (4 arrivals)
  uint64_t* bar_o = bar + 7;       // O += P V retired: O may be rescaled / read
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], i == 6 ? 4 : 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *s_tmem;
  const uint32_t tmem_o = tmem_s + kBKV;
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp == 4) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      pdl_wait();  // q / k / v are written by the preceding kernel (RoPE + KV append)
      const int q_row = seq_start + qb * kBQ;
      mbar_expect_tx(bar_q, kTileBytes);
      tma_load_2d(s_q, &tmap_q, h * kD, q_row, bar_q);
      tma_load_2d(s_q + kSubBytes, &tmap_q, h * kD + 64, q_row, bar_q);
      for (int j = 0; j < n_kb; ++j) {
        const int k_row = seq_start + j * kBKV;
        const uint32_t ph = static_cast<uint32_t>(j) & 1u;
        mbar_wait(bar_kfree, ph ^ 1u);  // a fresh barrier passes the wait on the "previous" phase
        mbar_expect_tx(bar_kfull, kTileBytes);
        tma_load_2d(s_k, &tmap_k, hkv * kD, k_row, bar_kfull);
        tma_load_2d(s_k + kSubBytes, &tmap_k, hkv * kD + 64, k_row, bar_kfull);
        mbar_wait(bar_vfree, ph ^ 1u);
        mbar_expect_tx(bar_vfull, kTileBytes);
        tma_load_2d(s_v, &tmap_v, hkv * kD, k_row, bar_vfull);
        tma_load_2d(s_v + kSubBytes, &tmap_v, hkv * kD + 64, k_row, bar_vfull);
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(kBQ, kBKV, 0u);
      constexpr uint32_t idesc_pv = umma_idesc_f16(kBQ, kD, 1u);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kb; ++j) {
        const uint32_t ph = static_cast<uint32_t>(j) & 1u;
        mbar_wait(bar_kfull, ph);
        tc_fence_after();
        // S = Q K^T.  The S columns also hold the previous block's P: the tensor pipe executes MMAs in issue order, so this overwrite follows
        // the previous O += P V.
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * kSubBytes;
          const uint64_t adesc = umma_desc_sw128(smem_u32(s_q + off)) + (ks & 3) * 2;
          const uint64_t bdesc = umma_desc_sw128(smem_u32(s_k + off)) + (ks & 3) * 2;
          umma_f16_ss(tmem_s, adesc, bdesc, idesc_qk, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_s);
        umma_commit(bar_kfree);
        mbar_wait(bar_p, ph);
        mbar_wait(bar_vfull, ph);
        tc_fence_after();
        // O += P V: A = P (fp16 pairs, 8 TMEM columns per 16 keys), B = V rows [key][dim] = MN-major, 16 keys (2 KB) per instruction
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          const uint64_t bdesc = umma_desc_sw128_mn(smem_u32(s_v + kk * 2048), kSubBytes);
          umma_f16_ts(tmem_o, tmem_s + kk * 8, bdesc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(bar_vfree);
        umma_commit(bar_o);
      }
    }
  } else {
    // ===================================== softmax + epilogue: thread = query row =====================================
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t t_s = tmem_s + lane_off, t_o = tmem_o + lane_off;
    const int q_pos = qb * kBQ + row;  // position inside the sequence
    float m_used = -CUDART_INF_F;      // the maximum the accumulated O and l are expressed against (log2 domain, scaled)
    float l = 0.f;
    for (int j = 0; j < n_kb; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      const bool diag = (j == qb);
      mbar_wait(bar_s, ph);
      tc_fence_after();
      // ---- the whole S row (128 fp32) comes into registers with ONE TMEM round trip ----
      uint32_t r[kBKV / 32][32];
#pragma unroll
      for (int c = 0; c < kBKV / 32; ++c) tmem_ld_32x32b_x32(t_s + c * 32, r[c]);
      tmem_wait_ld();
      float m_blk = -CUDART_INF_F;
      if (diag) {
#pragma unroll
        for (int c = 0; c < kBKV / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (j * kBKV + c * 32 + i > q_pos) r[c][i] = 0xff800000u;  // -inf: exp2 turns it into an exact 0
            m_blk = fmaxf(m_blk, __uint_as_float(r[c][i]));
          }
      } else {
#pragma unroll
        for (int c = 0; c < kBKV / 32; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) m_blk = fmaxf(m_blk, __uint_as_float(r[c][i]));
      }
      m_blk *= p.scale_log2;
      // ---- lazy rescale of O and l (warp-uniform decision) ----
      const bool grow = (m_blk - m_used) > kRescaleThreshold;  // true on the first block (m_used = -inf)
      if (j > 0) {
        mbar_wait(bar_o, ph ^ 1u);  // the previous O += P V has retired (also: P of the previous block is no longer needed)
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_blk : m_used;
        const float alpha = (m_used == -CUDART_INF_F) ? 0.f : ex2_approx(m_used - m_new);  // 1 for rows that keep their maximum
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < kD / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_o + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(t_o + c * 32, o);
          }
        }
        l *= alpha;
        m_used = m_new;
      }
      // ---- P = exp2(s * scale - m_used), rounded to fp16, written over the S columns (all of S is in registers by now) ----
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(r[hc * 2 + (i >> 5)][i & 31]), p.scale_log2, -m_used));
          const float p1 = ex2_approx(fmaf(__uint_as_float(r[hc * 2 + (i >> 5)][(i & 31) + 1]), p.scale_log2, -m_used));
          const __half2 h2 = __floats2half2_rn(p0, p1);
          // the row sum is taken over the ROUNDED probabilities: numerator (the MMA sees fp16 P) and denominator then agree
          const float2 f = __half22float2(h2);
          l += f.x + f.y;
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        tmem_st_32x32b_x32(t_s + hc * 32, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> fp16 ----
    mbar_wait(bar_o, static_cast<uint32_t>(n_kb - 1) & 1u);
    tc_fence_after();
    const float inv = 1.f / l;
    const bool valid = q_pos < seq_len;
    __half* dst = p.out + static_cast<long long>(seq_start + q_pos) * p.out_stride + h * kD;
#pragma unroll 1
    for (int c = 0; c < kD / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_o + c * 32, r);
      tmem_wait_ld();
      if (valid) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 h2 = __floats2half2_rn(__uint_as_float(r[v * 8 + 2 * e]) * inv, __uint_as_float(r[v * 8 + 2 * e + 1]) * inv);
            ow[e] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(dst + c * 32 + v * 8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kTmemCols>(tmem_s);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// fp16 [rows, cols] with a row pitch of `stride` elements; box = 64 columns (128 B, swizzled) x 128 rows
int make_tmap_f16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t stride) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {stride * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(QS_ERR_CUDA, "cuTensorMapEncodeTiled(f16) failed (%d): ptr=%p rows=%llu cols=%llu stride=%llu", (int)r, ptr, (unsigned long long)rows,
                     (unsigned long long)cols, (unsigned long long)stride);
  return QS_OK;
}

}  // namespace

int prefill_attention(const PrefillAttnArgs& a) {
  QS_REQUIRE(a.head_dim == kD, "prefill_attention: head_dim=%d (only 128 is built, as in the reference)", a.head_dim);
  QS_REQUIRE(a.num_heads > 0 && a.num_kv_heads > 0 && a.num_heads % a.num_kv_heads == 0, "prefill_attention: heads=%d kv_heads=%d", a.num_heads, a.num_kv_heads);
  QS_REQUIRE(a.batch >= 0 && a.num_tokens >= 0 && a.max_seqlen >= 0, "prefill_attention: negative size");
  if (a.batch == 0 || a.num_tokens == 0 || a.max_seqlen == 0) return QS_OK;
  QS_REQUIRE(a.batch <= 65535 && a.num_heads <= 65535, "prefill_attention: batch=%d / heads=%d exceed the grid limits", a.batch, a.num_heads);
  QS_REQUIRE(a.q && a.k && a.v && a.out && a.cu_seqlens, "prefill_attention: null pointer");
  QS_REQUIRE(a.q_stride % 8 == 0 && a.k_stride % 8 == 0 && a.v_stride % 8 == 0 && a.out_stride % 8 == 0, "prefill_attention: row strides must be multiples of 8 halfs");
  QS_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.out)) & 15) == 0,
             "prefill_attention: q, k, v, out must be 16-byte aligned");
  QS_REQUIRE(a.q_stride >= a.num_heads * kD && a.k_stride >= a.num_kv_heads * kD && a.v_stride >= a.num_kv_heads * kD && a.out_stride >= a.num_heads * kD,
             "prefill_attention: row stride smaller than the row");
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_f16(&tq, a.q, a.num_tokens, static_cast<uint64_t>(a.num_heads) * kD, a.q_stride);
  if (rc) return rc;
  rc = make_tmap_f16(&tk, a.k, a.num_tokens, static_cast<uint64_t>(a.num_kv_heads) * kD, a.k_stride);
  if (rc) return rc;
  rc = make_tmap_f16(&tv, a.v, a.num_tokens, static_cast<uint64_t>(a.num_kv_heads) * kD, a.v_stride);
  if (rc) return rc;
  static bool attr_set[kMaxDevices] = {};
  bool& done = attr_set[device_ordinal()];
  if (!done) {
    rc = check_cuda(cudaFuncSetAttribute(prefill_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), "cudaFuncSetAttribute(prefill attention)");
    if (rc) return rc;
    done = true;
  }
  PrefillAttnParams p{};
  p.cu_seqlens = a.cu_seqlens;
  p.out = static_cast<__half*>(a.out);
  p.out_stride = a.out_stride;
  p.num_heads = a.num_heads;
  p.num_kv_heads = a.num_kv_heads;
  p.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((a.max_seqlen + kBQ - 1) / kBQ, a.num_heads, a.batch);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = static_cast<cudaStream_t>(a.stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, prefill_attention_kernel, tq, tk, tv, p), "prefill attention launch");
}

}  // namespace qs
