// One-time weight packing: PyTorch-layout fp32 parameters -> the layouts the kernels read.
// (state_dict contract: SURVEY.md Appendix A; BatchNorm eval folding: basicblock.py:69.)
#include "common.cuh"
#include "ffn_tail.cuh"

namespace irb {

// split-pad index map: destination index -> source index or -1 (zero fill).
// GDFN's 2h channels are stored as two halves of hp >= h channels each (hp multiple of 8).
__device__ __forceinline__ int map_src(int i_dst, int src_half, int dst_half) {
  const int hs = i_dst / dst_half, r = i_dst - hs * dst_half;
  return r < src_half ? hs * src_half + r : -1;
}

__global__ void pack_mat_kernel(const PackMat p) {
  const int n_dst = p.n_dst_half * p.n_halves;
  const long long total = (long long)n_dst * p.k_dst;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / p.k_dst), k = (int)(idx - (long long)n * p.k_dst);
    const int ns = map_src(n, p.n_src_half, p.n_dst_half);
    float v = 0.f;
    if (ns >= 0 && (k < p.k_src || p.kind == 2)) {
      if (p.kind == 0) {
        v = p.src[(long long)ns * p.k_src + k];
      } else if (p.kind == 2) {
        // 3x3 weights for tma_conv3.cu: k = tap * kpt + c, every tap padded to kpt = k_dst / 9 channels
        const int kpt = p.k_dst / 9, tap = k / kpt, c = k - tap * kpt;
        v = c < p.cin ? p.src[((long long)ns * p.cin + c) * 9 + tap] : 0.f;
      } else {
        const int tap = k / p.cin, c = k - tap * p.cin;
        v = p.src[((long long)ns * p.cin + c) * 9 + tap];
      }
      if (p.row_scale) v *= p.row_scale[ns];
    }
    if (p.fmt == 1) {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      p.dst[((long long)(k >> 2) * n_dst + n) * 4 + (k & 3)] = __uint_as_float(u);
    } else if (p.fmt == 2) {
      reinterpret_cast<__half*>(p.dst)[((long long)(k >> 3) * n_dst + n) * 8 + (k & 7)] = __float2half_rn(v);
    } else if (p.fmt == 3) {          // [k/32][n][128 B], 16-byte chunk c of row n stored at chunk c ^ (n & 7)
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
      const int kb = k >> 5, c = (k & 31) >> 2;
      p.dst[((long long)kb * n_dst + n) * 32 + ((c ^ (n & 7)) << 2) + (k & 3)] = __uint_as_float(u);
    } else if (p.fmt == 4) {          // [k/64][n][128 B] fp16, same swizzle
      const int kb = k >> 6, c = (k & 63) >> 3;
      reinterpret_cast<__half*>(p.dst)[((long long)kb * n_dst + n) * 64 + ((c ^ (n & 7)) << 3) + (k & 7)] = __float2half_rn(v);
    } else {
      p.dst[idx] = v;
    }
  }
}

int launch_pack_mat(const PackMat& p, cudaStream_t s) {
  const long long total = (long long)p.n_dst_half * p.n_halves * p.k_dst;
  const int blocks = (int)(cdivll(total, 256) < 1024 ? cdivll(total, 256) : 1024);
  pack_mat_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

__global__ void pack_dw_kernel(const float* __restrict__ src, float* __restrict__ dst, int src_half, int dst_half,
                               int n_halves) {
  const int cd = dst_half * n_halves;
  const int total = 9 * cd;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int t = idx / cd, c = idx - t * cd;
    const int cs = map_src(c, src_half, dst_half);
    dst[idx] = cs >= 0 ? src[cs * 9 + t] : 0.f;
  }
}

int launch_pack_dw(const float* src, float* dst, int c_src_half, int c_dst_half, int n_halves, cudaStream_t s) {
  const int total = 9 * c_dst_half * n_halves;
  pack_dw_kernel<<<cdiv(total, 256), 256, 0, s>>>(src, dst, c_src_half, c_dst_half, n_halves);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

__global__ void pack_dw_chunked_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int hp, int kc,
                                       int nsets) {
  const int total = nsets * 9 * hp;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % kc, tap = (idx / kc) % 9, half = (idx / (9 * kc)) % nsets, chunk = idx / (nsets * 9 * kc);
    const int ch = chunk * kc + c;
    dst[idx] = ch < h ? src[(half * h + ch) * 9 + tap] : 0.f;
  }
}

int launch_pack_dw_chunked(const float* src, float* dst, int h, int hp, int kc, int nsets, cudaStream_t s) {
  pack_dw_chunked_kernel<<<cdiv(nsets * 9 * hp, 256), 256, 0, s>>>(src, dst, h, hp, kc, nsets);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

__global__ void pack_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int src_half, int dst_half,
                                int n_halves, const float* __restrict__ scale, const float* __restrict__ shift) {
  const int total = dst_half * n_halves;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int is = map_src(idx, src_half, dst_half);
    float v = 0.f;
    if (is >= 0) {
      v = src ? src[is] : 0.f;
      if (scale) v *= scale[is];
      if (shift) v += shift[is];
    }
    dst[idx] = v;
  }
}

int launch_pack_vec(const float* src, float* dst, int src_half, int dst_half, int n_halves, const float* scale,
                    const float* shift, cudaStream_t s) {
  const int total = dst_half * n_halves;
  pack_vec_kernel<<<cdiv(total, 256), 256, 0, s>>>(src, dst, src_half, dst_half, n_halves, scale, shift);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float sc = g[i] / sqrtf(var[i] + eps);
    scale[i] = sc;
    shift[i] = b[i] - mean[i] * sc;
  }
}

int launch_bn_fold(const float* g, const float* b, const float* mean, const float* var, float eps, float* scale,
                   float* shift, int n, cudaStream_t s) {
  bn_fold_kernel<<<cdiv(n, 256), 256, 0, s>>>(g, b, mean, var, eps, scale, shift, n);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace irb
