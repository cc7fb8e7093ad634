// qserve_b200 -- sm_100a device helpers: mbarrier, bulk/TMA copies, tcgen05 (TMEM + UMMA) PTX wrappers.
// Hand-written PTX; nothing here includes CUTLASS/CuTe.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace qs {

// ---------------------------------------------------------------------------------------------
// error plumbing shared by all translation units (capi.cu owns the storage)
// ---------------------------------------------------------------------------------------------
enum : int {
  QS_OK = 0,
  QS_ERR_INVALID = -1,     // bad argument / unsupported shape
  QS_ERR_CUDA = -2,        // CUDA runtime error
  QS_ERR_WORKSPACE = -3,   // workspace too small
  QS_ERR_UNSUPPORTED = -4  // feature present in the reference API but not implemented here
};
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define QS_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::qs::set_error(::qs::QS_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch (PDL)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// step tracing (tools/step_timeline.py): block (0,0,0) of every kernel logs %globaltimer at entry, after the PDL wait and
// at exit into a device buffer [0] = record count, then records of 2 x u64: (kernel_id << 8 | phase), time.  Off by default.
// ---------------------------------------------------------------------------------------------
static __device__ unsigned long long* g_qs_trace = nullptr;
static __device__ unsigned int g_qs_trace_cap = 0;
__device__ __forceinline__ void qs_trace(unsigned kernel_id, unsigned phase, unsigned tid = 0) {
  if (g_qs_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == tid) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    const unsigned long long i = atomicAdd(g_qs_trace, 1ull);
    if (i < g_qs_trace_cap) {
      g_qs_trace[1 + 2 * i] = (static_cast<unsigned long long>(kernel_id) << 8) | phase;
      g_qs_trace[2 + 2 * i] = t;
    }
  }
}
static inline int qs_trace_install(void* buf, unsigned cap) {
  unsigned long long* p = static_cast<unsigned long long*>(buf);
  if (cudaMemcpyToSymbol(g_qs_trace, &p, sizeof(p)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_qs_trace_cap, &cap, sizeof(cap)) != cudaSuccess) return -1;
  return 0;
}
enum : unsigned { QS_K_GEMM = 1, QS_K_ATTN = 2, QS_K_NORM = 3, QS_K_QUANT = 4, QS_K_SILU = 5, QS_K_RMS = 6, QS_K_ADDNORM = 7, QS_K_SILUQ = 8, QS_K_PREFILL = 9 };

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // fast path: the non-blocking probe (the potentially-blocking try_wait costs ~200 cycles even on a completed phase)
  if (mbar_test_wait(bar, parity)) return;
  // bounded wait: a protocol bug must surface as a trapped launch (reported by the host), never as a hung GPU.  The bound is wall-clock
  // (10 s of %globaltimer), not a spin count: under tensor parallelism a kernel legitimately waits -- through its dependency chain -- for a
  // peer rank that may be late by a host scheduling quantum or a whole warm-up phase.
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
      if (t0 == 0) t0 = t;
      if (t - t0 > 10000000000ull) {
        printf("qserve_b200: mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bulk async copies (TMA engine): 1-D linear and 2-D tiled (tensor map)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int32_t x, int32_t y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// fire-and-forget HBM -> L2 prefetch of one tensor-map box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tmap, int32_t x, int32_t y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------------------------
// tcgen05: tensor memory management
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cta_group::2 (a pair of CTAs on the two SMs of a TPC drives ONE 256-row UMMA; only the even CTA issues) ----
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {  // executed by one warp of EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem, 256 rows over both CTAs] (+)= A[tmem of each CTA, its 128 rows] * B[shared memory, each CTA holds half of the N columns]
__device__ __forceinline__ void umma_i8_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread have completed) on the mbarrier at this CTA-relative offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// 2-D tiled load into THIS CTA's shared memory whose completion is signalled on the mbarrier at the same CTA-relative offset in CTA 0 of the pair
// (cta_group::2 form: the barrier may live in either CTA of the pair, so one barrier of the MMA-issuing CTA collects both halves of a stage)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, int32_t x, int32_t y, uint64_t* bar) {
  uint32_t leader_bar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_bar) : "r"(smem_u32(bar)), "r"(0u));
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
// named barrier over `threads` threads (a multiple of 32) of the CTA; id 0 is __syncthreads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// arrive on the mbarrier at the same CTA-relative offset in CTA `rank` of the cluster.  Default semantics (release at CTA scope) on purpose:
// `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive (and `try_wait.acquire.cluster` to a CCTL.IVALL behind the wait),
// ~1 us per pipeline stage.  The data these barriers guard is tensor memory / TMA-written shared memory, ordered by tcgen05.wait::st +
// tcgen05.fence::before_thread_sync on the producer side and tcgen05.fence::after_thread_sync on the consumer side, not by generic-proxy fences;
// the consumer waits with the ordinary mbar_wait.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// 16 lanes x 128 bit (4 columns) pattern, repeated twice along columns: thread t writes
//   r0 -> (lane t/4,   col t%4)   r1 -> (lane t/4+8, col t%4)   r2 -> (lane t/4, col 4+t%4)   r3 -> (lane t/4+8, col 4+t%4)
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
// 32 lanes x 32 bit: thread t reads its own lane, 32 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05.mma kind::i8 (INT8 x INT8 -> INT32), single CTA
//   SS: A and B from shared memory (matrix descriptors);  TS: A from tensor memory, B from shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand stored as rows of 128 B with the 128-byte swizzle
// (the layout a TMA box {128 B, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces; tile base 1024-B aligned).
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   bits [32,46) stride byte offset >> 4 (8 rows * 128 B = 1024)      bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::i8: D = s32, A/B formats (0 = u8, 1 = s8), both K-major, M x N tile.
//   bits [4,6) c_format (2 = S32)   [7,10) a_format   [10,13) b_format   [15] a_major   [16] b_major
//   bits [17,23) N >> 3             [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_i8(uint32_t m, uint32_t n, uint32_t a_signed, uint32_t b_signed) {
  return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

}  // namespace qs
