// Micro-benchmark: tensor-pipe occupancy of one tcgen05.mma.cta_group::1.kind::i8 (SS form) as a function of its shape.
// One CTA per SM issues ITER dependent instructions into one accumulator (and, second variant, alternates two accumulators).
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../qserve_b200/csrc/common.cuh"
using namespace qs;

template <int M, int N, int NACC>
__global__ void __launch_bounds__(128) k(long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* sa = smem;              // A: M rows x 128 B (swizzle irrelevant for timing)
  uint8_t* sb = smem + 32768;      // B: N rows x 128 B
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&s_tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s_tmem;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_i8(M, N, 1u, 1u);
    const uint64_t ad = umma_desc_sw128(smem_u32(sa)), bd = umma_desc_sw128(smem_u32(sb));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int t = 0; t < 4; ++t) umma_i8_ss(tm + ((it * 4 + t) % NACC) * 256, ad + t * 2, bd + t * 2, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int M, int N, int NACC>
void run() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  auto kern = k<M, N, NACC>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 512;
  kern<<<148, 128, 65536>>>(cyc, iters);
  kern<<<148, 128, 65536>>>(cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double per = avg / (iters * 4.0);
  printf("kind::i8 M=%3d N=%3d K=32, %d accumulator(s): %7.1f cycles per instruction  (%6.0f MAC/clk/SM, ideal %d cycles)%s\n", M, N, NACC, per,
         double(M) * N * 32 / per, M * N * 32 / 8192, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  run<128, 32, 1>();
  run<128, 64, 1>();
  run<128, 64, 2>();
  run<128, 128, 1>();
  run<128, 256, 1>();
  run<128, 256, 2>();
  run<64, 64, 1>();
  run<64, 128, 1>();
  run<64, 256, 1>();
  run<64, 256, 2>();
  return 0;
}
