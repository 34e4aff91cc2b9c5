// Micro-benchmark: issue rate of packed fp32 FMA (fma.rn.f32x2 = FFMA2) and of scalar FFMA on one SM sub-partition, with
// 1 / 2 / 4 warps per scheduler and 16 independent accumulator chains per thread.  Answers: what is the FP32-pipe ceiling
// the depthwise roles of the fused kernels run against?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_ab/probe_ffma2 scripts/probe_ffma2.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  const int t = threadIdx.x;
  f2_t acc[16], a[4], b[4];
  for (int i = 0; i < 16; ++i) acc[i] = (f2_t)(t + i) * 0x0000000100000001ull;
  for (int i = 0; i < 4; ++i) { a[i] = 0x3f8000013f800001ull + i; b[i] = 0x3f7fffff3f7fffffull - i; }
  float s[16];
  for (int i = 0; i < 16; ++i) s[i] = t * 0.001f + i;
  const float fa = 1.0001f, fb = 0.9999f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma2(a[i & 3], b[(i >> 2) & 3], acc[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(s[i]) : "f"(fa), "f"(fb));
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(s[i]) : "f"(fb), "f"(fa));
    }
  }
  const long long t1 = clock64();
  float r = 0;
  for (int i = 0; i < 16; ++i) r += (float)(acc[i] & 0xffff) + s[i];
  out[blockIdx.x * blockDim.x + t] = r;
  if ((t & 31) == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(t1 - t0));   // slowest warp
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      cudaMemset(cyc, 0, 148 * 8);
      if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters); else k<1><<<148, warps * 32>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double instr_per_smsp = (double)iters * (mode == 0 ? 16 : 32) * (warps / 4);
      printf("{\"op\": \"%s\", \"warps_per_scheduler\": %d, \"cycles_per_instruction_per_scheduler\": %.3f, \"fma_per_clk_per_sm\": %.1f}\n",
             mode == 0 ? "FFMA2" : "FFMA", warps / 4, avg / instr_per_smsp, instr_per_smsp / avg * 4 * 32 * (mode == 0 ? 2 : 1));
    }
  return 0;
}
