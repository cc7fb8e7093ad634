// Micro-benchmark: issue cost (SM cycles per warp instruction per SM sub-partition) of the instruction classes the
// decode kernels lean on.  One CTA of 128 threads per SM (one warp per sub-partition), 8 independent chains per thread.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITER 2048
template <int OP>
__global__ void k(float* out, long long* cyc, int seed) {
  float f[8];
  unsigned u[8];
  for (int i = 0; i < 8; ++i) { f[i] = seed * 0.001f + i + threadIdx.x * 1e-3f; u[i] = seed + i * 77 + threadIdx.x; }
  float c[4][4] = {};
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) f[i] = fmaf(f[i], 1.0001f, 0.5f);
      if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(0x0f0f0f0fu), "r"(0x64006400u));
      if (OP == 2) { __half2 h = *reinterpret_cast<__half2*>(&u[i]); float2 g = __half22float2(h); u[i] = __float_as_uint(g.x + g.y); }  // 2 x HADD2.F32 + FADD
      if (OP == 3) { float g; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(g) : "r"(u[i])); u[i] = __float_as_uint(g); }
      if (OP == 4) { int g; asm volatile("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(g) : "f"(f[i])); f[i] = __int_as_float(g | 0x3f800000); }
      if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 6) { unsigned g; asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(g) : "f"(f[i])); f[i] = __uint_as_float(g | 0x3f800000u); }
      if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[i]) : "r"(0x64006400u));
      if (OP == 8) asm volatile("sub.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(0x3c003c00u));
      if (OP == 9) u[i] = __shfl_xor_sync(0xffffffffu, u[i], 4);
      if (OP == 10) { asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %0;" : "+r"(u[i])); }
      if (OP == 11 && i < 4) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
      }
      if (OP == 12) { float g; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(g) : "h"((unsigned short)u[i])); u[i] = __float_as_uint(g); }
      if (OP == 13) { unsigned short g; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(g) : "f"(f[i])); f[i] = __uint_as_float(g | 0x3f800000u); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(u[i]);
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, int warps_per_smsp) {
  float* out; long long* cyc;
  int threads = 128 * warps_per_smsp;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, threads>>>(out, cyc, 1);
  k<OP><<<148, threads>>>(out, cyc, 2);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-28s warps/smsp=%d  %7.2f cycles per warp-instruction per SMSP\n", name, warps_per_smsp, avg / (double(ITER) * per_iter * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w = 1; w <= 4; w *= 4) {
    run<0>("FFMA", 8, w);
    run<1>("LOP3", 8, w);
    run<2>("2xHADD2.F32+FADD (h2->f2)", 8, w);
    run<3>("I2F (cvt.rn.f32.s32)", 8, w);
    run<4>("F2I (cvt.rni.sat.s8.f32)+LOP", 8, w);
    run<5>("MUFU.EX2", 8, w);
    run<6>("F2FP (cvt.rn.f16x2.f32)+LOP", 8, w);
    run<7>("PRMT", 8, w);
    run<8>("HADD2 (sub.f16x2)", 8, w);
    run<9>("SHFL", 8, w);
    run<10>("MOVMATRIX", 8, w);
    run<11>("HMMA m16n8k16 f32", 4, w);
    run<12>("cvt.f32.f16 (scalar)", 8, w);
    run<13>("cvt.rn.f16.f32 (scalar)+LOP", 8, w);
  }
  return 0;
}
