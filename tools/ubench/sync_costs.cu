// Micro-benchmark (prepared in round 1, not yet run): per-operation cost, seen by ONE thread, of the synchronisation
// primitives that pace the GEMM main loop -- mbarrier.try_wait / test_wait on a completed phase, mbarrier.arrive,
// tcgen05.commit with no MMA outstanding, tcgen05.fence -- and the latency of a programmatic-dependent-launch boundary
// (chain of trivial kernels: trigger at entry, griddepcontrol.wait, exit).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../qserve_b200/csrc/common.cuh"
using namespace qs;

__global__ void sync_costs(long long* out) {
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<32>(&s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr int N = 256;
    mbar_arrive(&bar[0]);  // phase 0 of bar[0] is now complete
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) (void)mbar_try_wait(&bar[0], 0);
    long long t1 = clock64();
    for (int i = 0; i < N; ++i) (void)mbar_test_wait(&bar[0], 0);
    long long t2 = clock64();
    for (int i = 0; i < N; ++i) { mbar_arrive(&bar[1]); }
    long long t3 = clock64();
    for (int i = 0; i < N; ++i) { umma_commit(&bar[1]); }
    long long t4 = clock64();
    for (int i = 0; i < N; ++i) { tc_fence_after(); }
    long long t5 = clock64();
    out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = (t4 - t3) / N; out[4] = (t5 - t4) / N;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<32>(s_tmem);
}

__global__ void pdl_link(unsigned long long* stamps, int idx) {
  if (threadIdx.x == 0 && blockIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[idx] = t;
  }
}

// Flag chain: the same trivial kernels, launched with the programmatic attribute (so link i+1 is resident while link i runs) but ordered by a
// global counter instead of griddepcontrol.wait: every CTA of link i bumps ctr[i] when it is done, every CTA of link i+1 spins on ctr[i].
__global__ void flag_link(unsigned long long* stamps, unsigned* ctr, int idx, unsigned target) {
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (idx > 0 && threadIdx.x == 0) {
    unsigned v, spins = 0;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr + idx - 1) : "memory"); } while (v < target && ++spins < 20000000u);  // bounded: never hang the GPU
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[idx] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(ctr + idx, 1u); }
}

template <typename F>
static double graph_chain(F launch_link, unsigned long long* stamps, int links) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < links; ++i) launch_link(st, i);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int rep = 0; rep < 3; ++rep) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  unsigned long long hs[64];
  cudaMemcpy(hs, stamps, sizeof(hs), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 9; i < links; ++i) sum += double(hs[i] - hs[i - 1]);
  return sum / (links - 9) / 1e3;
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  sync_costs<<<1, 128>>>(out);
  cudaDeviceSynchronize();
  long long h[5];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("cycles per op (one thread): try_wait(done) %lld  test_wait(done) %lld  mbarrier.arrive %lld  tcgen05.commit %lld  tcgen05.fence %lld\n", h[0], h[1], h[2],
         h[3], h[4]);

  const int links = 64;
  unsigned long long* stamps;
  cudaMalloc(&stamps, links * 8);
  for (int rep = 0; rep < 2; ++rep) {
    for (int i = 0; i < links; ++i) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(128);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, pdl_link, stamps, i);
    }
    cudaDeviceSynchronize();
  }
  unsigned long long hs[64];
  cudaMemcpy(hs, stamps, sizeof(hs), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int i = 9; i < links; ++i) sum += double(hs[i] - hs[i - 1]);
  printf("PDL chain of trivial 148-CTA kernels, eager launches: %.2f us per boundary (dependency resolved -> next dependency resolved; may be host-launch-bound)\n", sum / (links - 9) / 1e3);
  for (int grid : {64, 148, 296}) {
    auto pdl = [&](cudaStream_t st, int i) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, pdl_link, stamps, i);
    };
    printf("CUDA graph, PDL chain, %3d CTAs per link: %.2f us per boundary\n", grid, graph_chain(pdl, stamps, links));
    auto plain = [&](cudaStream_t st, int i) { pdl_link<<<grid, 128, 0, st>>>(stamps, i); };
    printf("CUDA graph, plain (non-programmatic) chain, %3d CTAs per link: %.2f us per boundary\n", grid, graph_chain(plain, stamps, links));
    if (grid <= 148) {  // all links must be able to be co-resident with their predecessor
      unsigned* ctr;
      cudaMalloc(&ctr, links * 4);
      // counters are monotonic: the graph is instantiated for ONE replay (targets are baked in), so measure a single launch after a reset
      cudaMemset(ctr, 0, links * 4);
      auto flag = [&](cudaStream_t st, int i) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, flag_link, stamps, ctr, i, static_cast<unsigned>(grid));
      };
      cudaStream_t st;
      cudaStreamCreate(&st);
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
      for (int i = 0; i < links; ++i) flag(st, i);
      cudaStreamEndCapture(st, &g);
      cudaGraphInstantiate(&ge, g, 0);
      cudaGraphLaunch(ge, st);
      cudaStreamSynchronize(st);
      unsigned long long hs2[64];
      cudaMemcpy(hs2, stamps, sizeof(hs2), cudaMemcpyDeviceToHost);
      double sum2 = 0;
      for (int i = 9; i < links; ++i) sum2 += double(hs2[i] - hs2[i - 1]);
      printf("CUDA graph, FLAG chain (acquire-spin on a global counter, no griddepcontrol.wait), %3d CTAs per link: %.2f us per boundary\n", grid, sum2 / (links - 9) / 1e3);
    }
  }
  return 0;
}
