// Hardware probe + micro-benchmark of csrc/dw_tmem.cuh (depthwise 3x3 taps read straight from tensor memory):
//   1. correctness of tcgen05.ld.32x32b at ODD column offsets (the O row loads start at column 18 r + 8 h + 1);
//   2. cycles per unit (32 channels x 8 x 8 outputs per warp) with 8 depthwise warps per SM, every SM busy.
// Standalone (not part of libirb200.so):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_ab/probe_dw_tmem scripts/probe_dw_tmem.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../image_restoration_models_b200/csrc/dw_tmem.cuh"

using namespace irb::dwt;

__host__ __device__ inline float pat(int lane, int col) { return (float)((lane * 131 + col * 71 + 7) % 257) / 128.0f - 1.0f; }
__host__ __device__ inline float tap(int ch, int t) { return (float)((ch * 37 + t * 101 + 3) % 61) / 61.0f - 0.5f; }

__device__ __forceinline__ uint32_t f2h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cycles, float* sink, int iters, int mode) {
  __shared__ uint32_t tmem_base_s;
  extern __shared__ __align__(1024) uint8_t xs[];       // X tile: [128 channels][128 px] fp16, two 64-px boxes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const int q = warp & 3, h = warp >> 2, ch = q * 32 + lane;
  if (warp < 4) {
    for (int col = 0; col < 192; ++col) {
      const uint32_t v = __float_as_uint(pat(ch, col));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tb + ((uint32_t)(q * 32) << 16) + (uint32_t)col), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  f2_t w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) { const float f = tap(ch, t); w[t] = pack2(f, f); }
  const uint32_t taddr = tb + ((uint32_t)(q * 32) << 16) + (uint32_t)(8 * h);
  const uint32_t xbase = (uint32_t)__cvta_generic_to_shared(xs);
  f2_t nrm = pack2(0.f, 0.f);
  float* o = out + ((size_t)blockIdx.x * 128 + ch) * 128;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const bool last = it == iters - 1;
    unit(taddr, w, [&](int oy, const f2_t (&acc)[4]) {
      float a[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack2(acc[j], a[2 * j], a[2 * j + 1]);
      if (mode >= 1) {
        // q / k path: fp16 row of 8 pixels into the X tile (SWIZZLE_128B, row = channel), norms on the rounded values
        uint4 u;
        u.x = f2h2(a[0], a[1]); u.y = f2h2(a[2], a[3]); u.z = f2h2(a[4], a[5]); u.w = f2h2(a[6], a[7]);
        const uint32_t chunk = (uint32_t)((oy & 3) * 2 + h);
        const uint32_t addr = xbase + (uint32_t)(oy >> 2) * (128u * 128u) + (uint32_t)ch * 128u + ((chunk ^ ((uint32_t)ch & 7u)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
        if (mode >= 2) {
          const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<__half2*>(&u.y));
          const float2 f2 = __half22float2(*reinterpret_cast<__half2*>(&u.z)), f3 = __half22float2(*reinterpret_cast<__half2*>(&u.w));
          nrm = fma2(pack2(f0.x, f0.y), pack2(f0.x, f0.y), nrm); nrm = fma2(pack2(f1.x, f1.y), pack2(f1.x, f1.y), nrm);
          nrm = fma2(pack2(f2.x, f2.y), pack2(f2.x, f2.y), nrm); nrm = fma2(pack2(f3.x, f3.y), pack2(f3.x, f3.y), nrm);
        }
      } else {
        nrm = add2(nrm, add2(add2(acc[0], acc[1]), add2(acc[2], acc[3])));
      }
      if (last) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[oy * 16 + 8 * h + e] = a[e];
      }
    });
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
  float n0, n1;
  unpack2(nrm, n0, n1);
  if (n0 + n1 == 123.456f) sink[0] = n0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const int nblk = 148;
  float *out, *sink; long long* cyc;
  cudaMalloc(&out, (size_t)nblk * 128 * 128 * 4); cudaMalloc(&sink, 16); cudaMalloc(&cyc, nblk * 8 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<float> ho((size_t)nblk * 128 * 128);
  std::vector<long long> hc(nblk * 8);
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(out, 0, ho.size() * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<<<nblk, 256, 200 * 1024>>>(out, cyc, sink, 10, mode);      // warm-up
    cudaEventRecord(e0);
    probe<<<nblk, 256, 200 * 1024>>>(out, cyc, sink, iters, mode);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), cyc, hc.size() * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int b = 0; b < nblk; b += 49)
      for (int ch = 0; ch < 128; ++ch)
        for (int oy = 0; oy < 8; ++oy)
          for (int ox = 0; ox < 16; ++ox) {
            double s = 0;
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) s += (double)tap(ch, ky * 3 + kx) * pat(ch, (oy + ky) * 18 + ox + kx);
            maxerr = fmax(maxerr, fabs(s - ho[((size_t)b * 128 + ch) * 128 + oy * 16 + ox]));
          }
    long long mx = 0; double avg = 0;
    for (auto c : hc) { mx = c > mx ? c : mx; avg += c; }
    avg /= hc.size();
    printf("{\"mode\": %d, \"iters\": %d, \"max_err\": %.3e, \"cycles_per_unit_avg\": %.1f, \"cycles_per_unit_max\": %.1f, \"ms\": %.4f, "
           "\"ffma2_floor_cycles_per_unit_2warps_per_smsp\": 1152}\n", mode, iters, maxerr, avg / iters, (double)mx / iters, ms);
  }
  return 0;
}
