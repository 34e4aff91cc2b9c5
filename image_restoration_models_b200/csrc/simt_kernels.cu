// CUDA-core (fp32 FMA) kernels: the memory-bound stages of the forward, plus a plain fp32
// contraction used as the on-device second oracle of the tcgen05 kernels and for shapes the
// tensor-core kernels do not cover.  Channels-last activations: a[pixel*ld + c].
//
// Reference semantics restated here (paths relative to /root/reference/src):
//   LayerNorm            restormer/restormer.py:37-39 (BiasFree), :54-57 (WithBias)
//   depthwise 3x3        restormer/restormer.py:83,106 (groups == channels, zero pad 1)
//   GELU gate            restormer/restormer.py:90-91 (exact erf GELU)
//   q/k normalise, temperature, softmax   restormer/restormer.py:121-125
//   PixelUnshuffle / PixelShuffle          restormer/restormer.py:176,186
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace irb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// ---------------------------------------------------------------------------------------------
// generic fp32 contraction: y[pixel, n] = sum_k A(pixel, k) * w[n, k]
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  static_assert(BM == 64 && BN == 64 && BK == 16, "thread mapping below assumes 64x64x16");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_mu[BM];
  __shared__ float s_rstd[BM];

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int HW = p.H * p.W;
  const int p0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long rowbase = (long long)b * HW;
  const float* __restrict__ w = p.w + (long long)b * p.w_bstride;

  if (p.ln_mode != LN_NONE) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < BM; r += 8) {
      const int pix = p0 + r;
      float mu = 0.f, rstd = 0.f;
      if (pix < HW) {
        const float* src = p.a1 + (rowbase + pix) * (long long)p.lda1;
        float s = 0.f;
        for (int k = lane; k < p.k1; k += 32) s += src[k];
        s = warp_sum(s);
        mu = s / (float)p.k1;
        float v = 0.f;
        for (int k = lane; k < p.k1; k += 32) { const float d = src[k] - mu; v += d * d; }
        v = warp_sum(v);
        rstd = 1.0f / sqrtf(v / (float)p.k1 + 1e-5f);
      }
      if (lane == 0) { s_mu[r] = mu; s_rstd[r] = rstd; }
    }
    __syncthreads();
  }

  const int lrow = tid >> 2;          // 0..63: tile row (A) / tile column (B) this thread stages
  const int lkq = (tid & 3) * 4;      // 0,4,8,12
  const int pixA = p0 + lrow;
  const int py = pixA / p.W, px = pixA - py * p.W;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const int k = k0 + lkq;
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (pixA < HW && k < p.K) {
      if (p.a_mode == A_PLAIN) {
        const float* src = (k < p.k1) ? p.a1 + (rowbase + pixA) * (long long)p.lda1 + k
                                      : p.a2 + (rowbase + pixA) * (long long)p.lda2 + (k - p.k1);
        const float4 t = *reinterpret_cast<const float4*>(src);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        if (p.ln_mode != LN_NONE) {
          const float mu = (p.ln_mode == LN_WITHBIAS) ? s_mu[lrow] : 0.f;
          const float rs = s_rstd[lrow];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v = (av[i] - mu) * rs * p.ln_w[k + i];
            if (p.ln_mode == LN_WITHBIAS) v += p.ln_b[k + i];
            av[i] = v;
          }
        }
      } else if (p.a_mode == A_IM2COL_NHWC) {
        const int tap = k / p.k1, c = k - tap * p.k1;
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          const float4 t = *reinterpret_cast<const float4*>(
              p.a1 + (rowbase + (long long)yy * p.W + xx) * (long long)p.lda1 + c);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        }
      } else {  // A_IM2COL_NCHW, scalar gathers (tiny channel counts at the image boundary of the net)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kk = k + i;
          if (kk < p.K) {
            const int tap = kk / p.k1, c = kk - tap * p.k1;
            const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
            if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
              av[i] = p.a1[(((long long)b * p.k1 + c) * p.H + yy) * p.W + xx];
          }
        }
      }
    }
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 + lrow < p.N && k < p.Kp) {
      const float4 t = *reinterpret_cast<const float4*>(w + (long long)(n0 + lrow) * p.Kp + k);
      wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lkq + i][lrow] = av[i]; Bs[lkq + i][lrow] = wv[i]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float br[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pix = p0 + ty * 4 + i;
    if (pix >= HW) continue;
    const int y = pix / p.W, x = pix - y * p.W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.relu) v = fmaxf(v, 0.f);
      v *= p.acc_sign;
      long long oidx;
      if (p.o_mode == O_NHWC) {
        oidx = (rowbase + pix) * (long long)p.ldy + n;
        if (p.r) v += p.r[(rowbase + pix) * (long long)p.ldr + n];
      } else if (p.o_mode == O_UNSHUFFLE) {
        // out[b, y/2, x/2, n*4 + (y%2)*2 + (x%2)], out extent (H/2, W/2)
        const long long opix = ((long long)b * (p.H / 2) + (y >> 1)) * (p.W / 2) + (x >> 1);
        oidx = opix * p.ldy + n * 4 + (y & 1) * 2 + (x & 1);
      } else if (p.o_mode == O_SHUFFLE) {
        // out[b, 2y+i, 2x+j, n/4] with i = (n%4)/2, j = n%2, out extent (2H, 2W)
        const int q = n & 3;
        const long long opix = ((long long)b * (2 * p.H) + (2 * y + (q >> 1))) * (2 * p.W) + (2 * x + (q & 1));
        oidx = opix * p.ldy + (n >> 2);
      } else {  // O_NCHW
        oidx = (((long long)b * p.N + n) * p.H + y) * p.W + x;
        if (p.r) v += p.r[oidx];
      }
      p.y[oidx] = v;
    }
  }
}

int launch_gemm_simt(const GemmParams& p, cudaStream_t s) {
  IRB_REQUIRE(p.Kp % 4 == 0, "gemm: padded K must be a multiple of 4");
  if (p.a_mode == A_PLAIN) {
    IRB_REQUIRE(p.k1 % 4 == 0 && p.lda1 % 4 == 0, "gemm: plain source 1 must be float4-addressable");
    IRB_REQUIRE(p.k2 == 0 || (p.k2 % 4 == 0 && p.lda2 % 4 == 0), "gemm: plain source 2 must be float4-addressable");
    IRB_REQUIRE(p.K == p.k1 + p.k2, "gemm: K != k1 + k2");
    IRB_REQUIRE(p.ln_mode == LN_NONE || p.k2 == 0, "gemm: LayerNorm prologue needs a single source");
  } else if (p.a_mode == A_IM2COL_NHWC) {
    IRB_REQUIRE(p.k1 % 4 == 0 && p.lda1 % 4 == 0 && p.K == 9 * p.k1, "gemm: im2col NHWC needs Cin % 4 == 0");
  } else {
    IRB_REQUIRE(p.K == 9 * p.k1, "gemm: im2col NCHW K != 9*Cin");
  }
  if (p.o_mode == O_UNSHUFFLE) IRB_REQUIRE(p.H % 2 == 0 && p.W % 2 == 0, "gemm: unshuffle needs even H, W");
  if (p.o_mode == O_SHUFFLE) IRB_REQUIRE(p.N % 4 == 0, "gemm: shuffle needs N % 4 == 0");
  const int HW = p.H * p.W;
  dim3 grid(cdiv(HW, 64), cdiv(p.N, 64), p.B);
  const double rows = (double)p.B * HW;
  const double a_elems = p.a_mode == A_PLAIN ? rows * p.K : rows * p.k1;   // im2col re-reads hit cache, not HBM
  ProfScope prof(p.tag, 4.0 * (a_elems + rows * p.N * (p.r ? 2.0 : 1.0)), 2.0 * rows * p.N * p.K, s);
  gemm_simt_kernel<64, 64, 16><<<grid, 256, 0, s>>>(p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 (+ optional GELU gate), one thread per (pixel, 4 channels)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 dw9(const float* __restrict__ in, int ldi, const float* __restrict__ w, int Cw,
                                      long long imgbase, int y, int x, int H, int W, int c) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(in + (imgbase + (long long)yy * W + xx) * ldi + c);
      const float4 k = *reinterpret_cast<const float4*>(w + ((dy + 1) * 3 + (dx + 1)) * Cw + c);
      acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
      acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
    }
  }
  return acc;
}

__global__ void __launch_bounds__(256) dwconv_kernel(const DwParams p) {
  const int cq = p.C >> 2;
  const long long total = (long long)p.B * p.H * p.W * cq;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cq) * 4;
    const long long pix = idx / cq;
    const int HW = p.H * p.W;
    const int b = (int)(pix / HW);
    const int pp = (int)(pix - (long long)b * HW);
    const int y = pp / p.W, x = pp - y * p.W;
    const long long imgbase = (long long)b * HW;
    float4 a = dw9(p.in, p.ldi, p.w, p.Cw, imgbase, y, x, p.H, p.W, c);
    if (p.bias) {
      const float4 bb = *reinterpret_cast<const float4*>(p.bias + c);
      a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
    }
    if (p.gate) {
      float4 g = dw9(p.in, p.ldi, p.w, p.Cw, imgbase, y, x, p.H, p.W, c + p.gate_off);
      if (p.bias) {
        const float4 bb = *reinterpret_cast<const float4*>(p.bias + c + p.gate_off);
        g.x += bb.x; g.y += bb.y; g.z += bb.z; g.w += bb.w;
      }
      a.x = gelu_erf(a.x) * g.x; a.y = gelu_erf(a.y) * g.y;
      a.z = gelu_erf(a.z) * g.z; a.w = gelu_erf(a.w) * g.w;
    }
    *reinterpret_cast<float4*>(p.out + pix * p.ldo + c) = a;
  }
}

int launch_dwconv_ref(const DwParams& p, cudaStream_t s) {
  IRB_REQUIRE(p.C % 4 == 0 && p.ldi % 4 == 0 && p.ldo % 4 == 0 && p.Cw % 4 == 0 && p.gate_off % 4 == 0,
              "dwconv: channel counts must be multiples of 4");
  const long long total = (long long)p.B * p.H * p.W * (p.C / 4);
  const int blocks = (int)(cdivll(total, 256) < 148LL * 16 ? cdivll(total, 256) : 148LL * 16);
  const double pix = (double)p.B * p.H * p.W;
  ProfScope prof(p.tag, 4.0 * pix * p.C * (p.gate ? 3.0 : 2.0), 2.0 * 9.0 * pix * p.C * (p.gate ? 2.0 : 1.0), s);
  dwconv_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// MDTA Gram partials: S[i][j] = sum_p q[p][i] k[p][j], sum q^2, sum k^2 over a slice of pixels
// grid (nparts, heads, B); block 256; TI = ch/16 outputs per thread per dimension
// ---------------------------------------------------------------------------------------------
template <int TI>
__global__ void __launch_bounds__(256) gram_kernel(const GramParams p) {
  constexpr int CH = TI * 16;
  constexpr int PT = 32;  // pixels staged per step
  __shared__ float qs[PT][CH + 4];
  __shared__ float ks[PT][CH + 4];
  const int part = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int per = cdiv(p.HW, p.nparts);
  const int pbeg = part * per;
  const int pend = min(p.HW, pbeg + per);
  const float* base = p.qkv + (long long)b * p.HW * p.ld;
  const int qoff = head * CH, koff = p.C + head * CH;

  float acc[TI][TI];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TI; ++j) acc[i][j] = 0.f;
  float nrm = 0.f;  // thread tid < CH: sum q[:,tid]^2 ; CH <= tid < 2CH: sum k^2

  for (int ps = pbeg; ps < pend; ps += PT) {
    // stage PT pixels x CH channels of q and k (float4 granularity)
    constexpr int V = CH / 4;
    for (int e = tid; e < PT * V * 2; e += 256) {
      const int which = e / (PT * V);
      const int r = (e % (PT * V)) / V;
      const int c4 = (e % V) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ps + r < pend)
        v = *reinterpret_cast<const float4*>(base + (long long)(ps + r) * p.ld + (which ? koff : qoff) + c4);
      float* dst = which ? &ks[r][c4] : &qs[r][c4];
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < PT; ++r) {
      float qa[TI], ka[TI];
#pragma unroll
      for (int i = 0; i < TI; ++i) { qa[i] = qs[r][ty * TI + i]; ka[i] = ks[r][tx * TI + i]; }
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TI; ++j) acc[i][j] = fmaf(qa[i], ka[j], acc[i][j]);
    }
    if (tid < 2 * CH) {
      const int c = tid < CH ? tid : tid - CH;
      for (int r = 0; r < PT; ++r) {
        const float v = tid < CH ? qs[r][c] : ks[r][c];
        nrm = fmaf(v, v, nrm);
      }
    }
    __syncthreads();
  }
  float* sp = p.s_part + (((long long)b * p.heads + head) * p.nparts + part) * CH * CH;
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TI; ++j) sp[(ty * TI + i) * CH + tx * TI + j] = acc[i][j];
  if (tid < 2 * CH) {
    float* np_ = p.n_part + (((long long)b * p.heads + head) * p.nparts + part) * 2 * CH;
    np_[tid] = nrm;
  }
}

int launch_gram_ref(const GramParams& p, cudaStream_t s) {
  const int ch = p.C / p.heads;
  IRB_REQUIRE(p.C % p.heads == 0 && ch % 16 == 0 && ch <= 128 && p.ld % 4 == 0, "gram: head dim must be a multiple of 16, <= 128");
  dim3 grid(p.nparts, p.heads, p.B);
  switch (ch / 16) {
    case 1: gram_kernel<1><<<grid, 256, 0, s>>>(p); break;
    case 2: gram_kernel<2><<<grid, 256, 0, s>>>(p); break;
    case 3: gram_kernel<3><<<grid, 256, 0, s>>>(p); break;
    case 4: gram_kernel<4><<<grid, 256, 0, s>>>(p); break;
    case 5: gram_kernel<5><<<grid, 256, 0, s>>>(p); break;
    case 6: gram_kernel<6><<<grid, 256, 0, s>>>(p); break;
    case 7: gram_kernel<7><<<grid, 256, 0, s>>>(p); break;
    case 8: gram_kernel<8><<<grid, 256, 0, s>>>(p); break;
    default: IRB_REQUIRE(false, "gram: unsupported head dim");
  }
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// reduce Gram partials, L2-normalise, temperature, softmax, fold project_out:
//   W_eff[b][n][h*ch + j] = sum_i W_proj[n][h*ch + i] * softmax_j(S[i][j] / (|q_i| |k_j|) * T_h)
// grid (heads, B); dynamic smem: ch*ch + 2*ch floats
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fold_kernel(const FoldParams p, const int PG) {
  extern __shared__ float sm[];
  const int head = blockIdx.x, b = blockIdx.y;
  const int ch = p.C / p.heads;
  const int NE = (ch * ch) >> 2;                 // float4 elements of the Gram
  float* A = sm;                                 // [PG][ch][ch]: slab 0 becomes the reduced Gram / the softmax
  float* nq = sm + PG * ch * ch;                 // [ch]
  float* nk = nq + ch;                           // [ch]
  float* wsm = nk + ch;                          // [rows_per][ch]: this CTA's slice of W_proj
  const int tid = threadIdx.x;
  const long long pb = ((long long)b * p.heads + head) * p.nparts;
  const int rows_per = (p.C + gridDim.z - 1) / gridDim.z;
  const int n_lo = blockIdx.z * rows_per, n_hi = min(p.C, n_lo + rows_per);
  const int nw = max(0, n_hi - n_lo) * ch;
  // W_proj is a constant: this CTA's slice goes to shared memory before the wait on the producing kernel (the kernel is a
  // chain of L2 round trips; with the slice in shared memory the product loop below has none)
  for (int e = tid; e < nw; e += 1024) wsm[e] = __ldg(p.w_proj + (long long)(n_lo + e / ch) * p.C + head * ch + e % ch);
  pdl_sync();
  // deterministic reduction of the pixel-slice partials: the parts are dealt round-robin to PG groups of threads (fixed
  // order inside a group, groups summed in order below), so that a thread's dependent L2 round trips are nparts / PG / 4
  const float4* sp4 = reinterpret_cast<const float4*>(p.s_part + pb * ch * ch);
  for (int idx = tid; idx < NE * PG; idx += 1024) {
    const int pg = idx / NE, e4 = idx - pg * NE;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* src = sp4 + e4;
    int part = pg;
    for (; part + 3 * PG < p.nparts; part += 4 * PG) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldg(src + (long long)(part + u * PG) * NE);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
    }
    {
      float4 t[3];
#pragma unroll
      for (int u = 0; u < 3; ++u)
        t[u] = part + u * PG < p.nparts ? __ldg(src + (long long)(part + u * PG) * NE) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 3; ++u) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
    }
    reinterpret_cast<float4*>(A)[idx] = s;
  }
  // squared norms: eight lanes per entry, each a fixed subset of the parts, then a fixed shuffle tree
  // (16 * ch entries-lanes: whole warps enter every iteration)
  for (int idx = tid; idx < 16 * ch; idx += 1024) {
    const int e = idx >> 3, l = idx & 7;
    float s = 0.f;
    for (int part = l; part < p.nparts; part += 8) s += __ldg(p.n_part + (pb + part) * 2 * ch + e);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    // F.normalize: x / max(||x||_2, 1e-12)
    if (l == 0) nq[e] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  if (PG > 1) {
    for (int e4 = tid; e4 < NE; e4 += 1024) {
      float4 s = reinterpret_cast<const float4*>(A)[e4];
      for (int g = 1; g < PG; ++g) {
        const float4 t = reinterpret_cast<const float4*>(A)[g * NE + e4];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      reinterpret_cast<float4*>(A)[e4] = s;
    }
    __syncthreads();
  }
  const float temp = p.temperature[head];
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < ch; i += nwarp) {
    float m = -INFINITY;
    for (int j = lane; j < ch; j += 32) {
      const float v = A[i * ch + j] * nq[i] * nk[j] * temp;
      A[i * ch + j] = v;
      m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < ch; j += 32) {
      const float e = expf(A[i * ch + j] - m);
      A[i * ch + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < ch; j += 32) A[i * ch + j] *= inv;
  }
  __syncthreads();
  // this CTA's slice of the output rows (grid.z row blocks share the softmax work, split the C x ch products)
  float* we = p.w_eff + (long long)b * p.w_eff_bstride;
  for (int e = n_lo * ch + tid; e < n_hi * ch; e += blockDim.x) {
    const int n = e / ch, j = e - n * ch;
    float s = 0.f;
    const float* wrow = wsm + (n - n_lo) * ch;
#pragma unroll 8
    for (int i = 0; i < ch; ++i) s = fmaf(wrow[i], A[i * ch + j], s);
    const int k = head * ch + j;
    if (p.fmt == 1) {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(s));
      we[((long long)(k >> 2) * p.C + n) * 4 + (k & 3)] = __uint_as_float(u);
    } else if (p.fmt == 2) {
      __half* wh = reinterpret_cast<__half*>(p.w_eff) + (long long)b * p.w_eff_bstride;
      wh[((long long)(k >> 3) * p.C + n) * 8 + (k & 7)] = __float2half_rn(s);
    } else if (p.fmt == 3) {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(s));
      we[((long long)(k >> 5) * p.C + n) * 32 + ((((k & 31) >> 2) ^ (n & 7)) << 2) + (k & 3)] = __uint_as_float(u);
    } else if (p.fmt == 4) {
      __half* wh = reinterpret_cast<__half*>(p.w_eff) + (long long)b * p.w_eff_bstride;
      wh[((long long)(k >> 6) * p.C + n) * 64 + ((((k & 63) >> 3) ^ (n & 7)) << 3) + (k & 7)] = __float2half_rn(s);
    } else {
      we[(long long)n * p.C + k] = s;
    }
  }
}

int launch_fold(const FoldParams& p, cudaStream_t s) {
  const int ch = p.C / p.heads;
  IRB_REQUIRE(p.C % p.heads == 0 && ch % 4 == 0 && ch <= 128, "fold: head dim must be a multiple of 4, <= 128");
  // the C x ch products are split over grid.z row blocks so that ONE wave of CTAs shares them (every block repeats the
  // partial reduction and softmax of its head: with 32 .. 64 (image, head) pairs at the low-resolution levels a second
  // wave only re-reads the partials -- measured 18.7 -> 16.8 us at C = 192, 37.9 -> 29.7 us at C = 384)
  static const int zb_env = getenv("IRB_FOLD_ZB") ? atoi(getenv("IRB_FOLD_ZB")) : 0;      // A/B switches for benchmarks
  static const int pg_env = getenv("IRB_FOLD_PG") ? atoi(getenv("IRB_FOLD_PG")) : 0;
  const int zb = zb_env > 0 ? std::min(zb_env, p.C) : std::max(1, std::min(8, 148 / std::max(1, p.heads * p.B)));
  const int rows_per = cdiv(p.C, zb);
  // part groups (PG > 1: the parts dealt to PG groups of threads, more independent load chains per thread; measured no
  // faster at C = 48 / 96 and slower at the low-resolution levels: one group)
  const int PG = pg_env > 0 ? std::max(1, std::min(pg_env, p.nparts)) : 1;
  const size_t smem = (size_t)(PG * ch * ch + 2 * ch + rows_per * ch) * sizeof(float);
  IRB_REQUIRE(smem <= 200 * 1024, "fold: head dim too large");
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(fold_kernel, optin));
  dim3 grid(p.heads, p.B, zb);
  ProfScope prof(TAG_FOLD, 4.0 * (double)p.B * p.C * p.C, 2.0 * (double)p.B * p.C * p.C * ch, s);
  // 1024 threads: the kernel is a chain of L2 round trips (partials -> softmax -> products) over a 48x48 .. 96x96 matrix;
  // four times the threads is a quarter of the trips per thread
  IRB_CUDA(launch_pdl(fold_kernel, dim3(grid), dim3(1024), smem, s, p, PG));
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// 3x3 convolution with 1..4 output channels written as NCHW (+ NCHW residual): the network's `output` conv
// (restormer.py:243,278-281).  Memory-bound: a warp walks 32 consecutive pixels of one image row; lanes split the
// input channels (float4 each), a 3x3 register window slides along x (3 new loads per pixel), the per-pixel result
// is warp-reduced and parked in lane (x mod 32) so the final store / residual load is one coalesced row segment.
// ---------------------------------------------------------------------------------------------
// R output rows per warp: the (R + 2) x 3 register window slides along x, so every input row crosses L2 -> SM (R + 2) / R
// times instead of 3.  The loop is a load -> FMA -> reduce chain per pixel with few warps per SM (the window is 70+
// registers), so the bytes in flight decide its speed: every lane runs its own cp.async queue -- column x + 1 + PD is
// requested (16 bytes per row, zero-filled outside the image: the conv's padding) into a 4-slot ring in shared memory
// while column x + 1 is read back into the window.  A lane only ever reads what it copied itself: no barrier.
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
template <int COUT, int R>
__global__ void __launch_bounds__(256, 2) conv3x3_small_kernel(const float* __restrict__ in, int ld, int cin,
                                                            const float* __restrict__ w, int kp,
                                                            const float* __restrict__ bias, int B, int H, int W,
                                                            const float* __restrict__ r, float sign,
                                                            float* __restrict__ y) {
  constexpr int PD = 3, NS = 4;                   // prefetch distance (columns), ring slots
  extern __shared__ __align__(16) float wsm[];    // [COUT][9][cin] weights, then the per-warp column rings
  for (int i = threadIdx.x; i < COUT * 9 * cin; i += 256) {
    const int co = i / (9 * cin), k = i - co * 9 * cin;
    wsm[i] = w[(long long)co * kp + k];
  }
  __syncthreads();
  pdl_sync();   // the weights (constants) are staged under the previous kernel's tail (common.cuh)
  const int lane = threadIdx.x & 31;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(wsm) + (uint32_t)((COUT * 9 * cin * 4 + 15) & ~15) +
                        (uint32_t)(threadIdx.x >> 5) * (uint32_t)(NS * (R + 2) * 512) + (uint32_t)lane * 16u;
  const int segs = (W + 31) >> 5;
  const int hb = (H + R - 1) / R;                 // row blocks per image
  const long long nwarps = (long long)B * hb * segs;
  const int c4n = cin >> 2;
  for (long long wi = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5; wi < nwarps; wi += ((long long)gridDim.x * 256) >> 5) {
    const int seg = (int)(wi % segs);
    const long long by = wi / segs;
    const int yy = (int)(by % hb) * R, b = (int)(by / hb);
    const int x0 = seg * 32;
    float res[COUT][R];
#pragma unroll
    for (int co = 0; co < COUT; ++co)
#pragma unroll
      for (int rr = 0; rr < R; ++rr) res[co][rr] = 0.f;
    for (int cb = 0; cb < c4n; cb += 32) {        // channel blocks of 32 float4 (one per lane)
      const int c4 = cb + lane;
      const bool cok = c4 < c4n;
      const float* base = in + ((long long)b * H * W) * ld + 4 * c4;
      // Column k of the segment is image column x0 - 1 + k; the columns are issued and fetched strictly in order, so the
      // row pointers, the ring slots and the column's x-validity advance incrementally (the first version recomputed four
      // 64-bit addresses and eight range checks per column: ncu counted 292 instructions per pixel column, 72 of them FMAs).
      const float* rowp[R + 2];
      bool rok[R + 2];
#pragma unroll
      for (int ry = 0; ry < R + 2; ++ry) {
        const int y2 = yy + ry - 1;
        rok[ry] = cok && y2 >= 0 && y2 < H;
        rowp[ry] = base + ((long long)min(max(y2, 0), H - 1) * W + (x0 - 1)) * ld;
      }
      int xk = x0 - 1;
      uint32_t islot = ring, fslot = ring;
      const uint32_t ring_end = ring + (uint32_t)(NS * (R + 2) * 512);
      auto issue = [&]() {
        const bool xok = (unsigned)xk < (unsigned)W;
#pragma unroll
        for (int ry = 0; ry < R + 2; ++ry) {
          const bool ok = rok[ry] && xok;
          cp_async16_zfill(islot + (uint32_t)(ry * 512), xok ? rowp[ry] : in, ok);
          rowp[ry] += ld;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++xk;
        islot += (uint32_t)((R + 2) * 512);
        if (islot == ring_end) islot = ring;
      };
      // the 3-column window lives in three register sets whose roles rotate (no moves): at step j the left / centre /
      // right columns are sets (j+1) % 3, (j+2) % 3, j % 3
      float4 win[3][R + 2];
      auto fetch = [&](auto set_tag) {
        constexpr int SET = decltype(set_tag)::value;
        asm volatile("cp.async.wait_group %0;" ::"n"(PD - 1) : "memory");
#pragma unroll
        for (int ry = 0; ry < R + 2; ++ry)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(win[SET][ry].x), "=f"(win[SET][ry].y), "=f"(win[SET][ry].z), "=f"(win[SET][ry].w)
                       : "r"(fslot + (uint32_t)(ry * 512)));
        fslot += (uint32_t)((R + 2) * 512);
        if (fslot == ring_end) fslot = ring;
      };
#pragma unroll
      for (int k = 0; k < PD; ++k) issue();
      fetch(std::integral_constant<int, 1>{}); issue();
      fetch(std::integral_constant<int, 2>{}); issue();
      // single-output case (gray images): the lane's 9 taps live in registers for the whole segment
      float4 kreg[COUT == 1 ? 9 : 1];
      if constexpr (COUT == 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
          kreg[t] = cok ? *reinterpret_cast<const float4*>(wsm + t * cin + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int jn = min(32, W - x0);
      auto step = [&](auto rot_tag, int j) {
        constexpr int ROT = decltype(rot_tag)::value;
        fetch(std::integral_constant<int, ROT>{});             // the new right column replaces the old left one
        issue();
        float acc[COUT][R];
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) acc[co][rr] = 0.f;
        if (cok) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int co = 0; co < COUT; ++co) {
                float4 k;
                if constexpr (COUT == 1) k = kreg[ky * 3 + dx];
                else k = *reinterpret_cast<const float4*>(wsm + (co * 9 + ky * 3 + dx) * cin + 4 * c4);
#pragma unroll
                for (int rr = 0; rr < R; ++rr) {
                  const float4 v = win[(ROT + 1 + dx) % 3][rr + ky];
                  acc[co][rr] = fmaf(v.x, k.x, fmaf(v.y, k.y, fmaf(v.z, k.z, fmaf(v.w, k.w, acc[co][rr]))));
                }
              }
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) {
            const float t = warp_sum(acc[co][rr]);
            if (lane == j) res[co][rr] += t;
          }
      };
      for (int j = 0; j < jn; j += 3) {
        step(std::integral_constant<int, 0>{}, j);
        if (j + 1 < jn) step(std::integral_constant<int, 1>{}, j + 1);
        if (j + 2 < jn) step(std::integral_constant<int, 2>{}, j + 2);
      }
      // the ring is reused by the next channel block / segment: nothing of this one may still be in flight
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (x0 + lane < W) {
#pragma unroll
      for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          if (yy + rr >= H) continue;
          const long long o = (((long long)b * COUT + co) * H + yy + rr) * W + x0 + lane;
          float v = res[co][rr] + (bias ? bias[co] : 0.f);
          v *= sign;
          if (r) v += r[o];
          y[o] = v;
        }
    }
  }
}

template <int COUT, int R>
static int launch_conv3x3_small_inst(int blocks, size_t smem, cudaStream_t s, const float* in, int ld, int cin, const float* w,
                                     int kp, const float* bias, int B, int H, int W, const float* r, float sign, float* y) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(conv3x3_small_kernel<COUT, R>, optin, 160 * 1024));
  IRB_CUDA(launch_pdl(conv3x3_small_kernel<COUT, R>, dim3(blocks), dim3(256), smem, s, in, ld, cin, w, kp, bias, B, H, W, r, sign, y));
  return IR_OK;
}

int launch_conv3x3_small(const float* in, int ld, int cin, const float* w, int kp, const float* bias, int cout, int B,
                         int H, int W, const float* r, float sign, float* y, cudaStream_t s) {
  IRB_REQUIRE(cout >= 1 && cout <= 4 && cin % 4 == 0 && ld % 4 == 0, "conv3x3_small: cout in 1..4, cin % 4 == 0");
  const size_t wbytes = ((size_t)cout * 9 * cin * sizeof(float) + 15) & ~(size_t)15;
  IRB_REQUIRE(wbytes <= 48 * 1024, "conv3x3_small: weights do not fit shared memory");
  // two rows per warp: four (255 registers, one block per SM) measured 420-430 us against 377-386 us on the bench's output conv
  const int R = 2;
  const size_t smem = wbytes + (size_t)8 * 4 * (R + 2) * 512;      // + 8 warps x 4 ring slots x (R + 2) rows x 32 lanes x 16 B
  const long long nwarps = (long long)B * cdiv(H, R) * ((W + 31) / 32);
  const int blocks = (int)std::min<long long>(cdivll(nwarps, 8), 148LL * 32);
  const double pix = (double)B * H * W;
  ProfScope prof(TAG_CONV3, 4.0 * pix * (cin + cout * (r ? 2.0 : 1.0)), 2.0 * 9.0 * pix * cin * cout, s);
  switch (cout) {
    case 1: IRB_TRY((launch_conv3x3_small_inst<1, 2>(blocks, smem, s, in, ld, cin, w, kp, bias, B, H, W, r, sign, y))); break;
    case 2: IRB_TRY((launch_conv3x3_small_inst<2, 2>(blocks, smem, s, in, ld, cin, w, kp, bias, B, H, W, r, sign, y))); break;
    case 3: IRB_TRY((launch_conv3x3_small_inst<3, 2>(blocks, smem, s, in, ld, cin, w, kp, bias, B, H, W, r, sign, y))); break;
    default: IRB_TRY((launch_conv3x3_small_inst<4, 2>(blocks, smem, s, in, ld, cin, w, kp, bias, B, H, W, r, sign, y))); break;
  }
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// First 3x3 convolution of a network (patch_embed restormer.py:160-165, DnCNN head network_dncnn.py:63): an NCHW image
// with 1-8 channels -> channels-last [pixel][cout] (+ bias, ReLU).  Pure store bandwidth (48-64 floats out per 1-6 in):
// cout/4 consecutive threads own one pixel and write one float4 each, so a warp's store is one contiguous segment.
// A thread walks DOWN a band of rows with its 3x3 input window in registers: per output float4 it issues three loads (the
// window's new row; neighbouring threads share them through L1), 36 FMAs per input channel and one store.  (The previous
// version re-read all nine taps with nine predicates and nine 64-bit addresses per output -- ncu: issue-bound, ~150
// instructions per float4 stored, 1.6 TB/s.)
// ---------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w, int kp,
                                                            const float* __restrict__ bias, int relu, int cout, int B, int H,
                                                            int W, float* __restrict__ y, int ldy, int band) {
  extern __shared__ float wsm[];                  // [9*CIN][cout] (transposed: a thread's 4 outputs are one float4)
  for (int i = threadIdx.x; i < cout * 9 * CIN; i += 256) {
    const int co = i / (9 * CIN), k = i - co * 9 * CIN;      // k = tap * CIN + c (PackMat kind 1 order)
    wsm[k * cout + co] = w[(long long)co * kp + k];
  }
  __syncthreads();
  pdl_sync();   // the weights (constants) are staged under the previous kernel's tail (common.cuh)
  // grid.x walks the W * cout/4 (pixel, quad) pairs of a row, grid.y the (image, row band) pairs
  const int q4 = cout >> 2;
  const int in_row = blockIdx.x * 256 + threadIdx.x;
  if (in_row >= W * q4) return;
  const int xx = in_row / q4, q = in_row - xx * q4;
  const int nbands = (H + band - 1) / band;
  const int b = blockIdx.y / nbands, y0 = (blockIdx.y - b * nbands) * band, y1 = min(H, y0 + band);
  // a thread's column never changes: the x-validity of the taps and the (clamped) column offsets are loop invariants
  const bool xl = xx > 0, xr = xx + 1 < W;
  const int cl = xl ? -1 : 0, cr = xr ? 1 : 0;
  const long long plane = (long long)H * W;
  const float* xc = x + (long long)b * CIN * plane + xx;              // channel 0, row 0, this column
  const float4 bq = bias ? *reinterpret_cast<const float4*>(bias + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  // gray images: the thread's nine weight quads live in registers for its whole band
  float4 wreg[CIN == 1 ? 9 : 1];
  if constexpr (CIN == 1) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wreg[tap] = *reinterpret_cast<const float4*>(wsm + tap * cout + 4 * q);
  }
  // one row of the window: columns xx-1, xx, xx+1 of every input channel; zeros outside the image (the loads themselves
  // are unconditional, from clamped addresses)
  auto load_row = [&](int yy, float (&v)[CIN][3]) {
    const bool ok = yy >= 0 && yy < H;
    const float* rp = xc + (long long)min(max(yy, 0), H - 1) * W;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float l = __ldg(rp + c * plane + cl), m = __ldg(rp + c * plane), r = __ldg(rp + c * plane + cr);
      v[c][0] = (ok && xl) ? l : 0.f; v[c][1] = ok ? m : 0.f; v[c][2] = (ok && xr) ? r : 0.f;
    }
  };
  float win[3][CIN][3];
  load_row(y0 - 1, win[0]);
  load_row(y0, win[1]);
  float* yp = y + (((long long)b * H + y0) * W + xx) * ldy + 4 * q;
  const long long ystep = (long long)W * ldy;
  for (int yy = y0; yy < y1; ++yy) {
    load_row(yy + 1, win[2]);
    float4 acc = bq;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float v = win[tap / 3][c][tap % 3];
        float4 ww;
        if constexpr (CIN == 1) ww = wreg[tap];
        else ww = *reinterpret_cast<const float4*>(wsm + (tap * CIN + c) * cout + 4 * q);
        acc.x = fmaf(v, ww.x, acc.x); acc.y = fmaf(v, ww.y, acc.y); acc.z = fmaf(v, ww.z, acc.z); acc.w = fmaf(v, ww.w, acc.w);
      }
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    *reinterpret_cast<float4*>(yp) = acc;
    yp += ystep;
#pragma unroll
    for (int c = 0; c < CIN; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) { win[0][c][k] = win[1][c][k]; win[1][c][k] = win[2][c][k]; }
  }
}

bool conv3x3_first_supported(int cin, int cout) {
  return (cin == 1 || cin == 3 || cin == 6) && cout % 4 == 0 && (size_t)cout * 9 * cin * sizeof(float) <= 48 * 1024;
}

int launch_conv3x3_first(const float* x_nchw, int cin, const float* w, int kp, const float* bias, int relu, int cout, int B,
                         int H, int W, float* y, int ldy, cudaStream_t s) {
  IRB_REQUIRE(conv3x3_first_supported(cin, cout) && ldy % 4 == 0, "conv3x3_first: unsupported shape");
  IRB_REQUIRE((long long)B * H < (1LL << 31), "conv3x3_first: too many image rows");
  // row bands of 16 (the window's two halo rows cost 12 % more loads); shorter bands when the image alone would not give
  // every SM about four blocks (the batch-1 latency path)
  const int gx = cdiv(W * (cout / 4), 256);
  int band = 16;
  while (band > 4 && (long long)gx * B * cdiv(H, band) < 4 * 148) band >>= 1;
  IRB_REQUIRE((long long)B * cdiv(H, band) <= 65535, "conv3x3_first: too many row bands for one launch");
  const dim3 blocks(gx, (unsigned)(B * cdiv(H, band)));
  const size_t smem = (size_t)cout * 9 * cin * sizeof(float);
  const double pix = (double)B * H * W;
  ProfScope prof(TAG_CONV3, 4.0 * pix * (cin + cout), 2.0 * 9.0 * pix * cin * cout, s);
  switch (cin) {
    case 1: IRB_CUDA(launch_pdl(conv3x3_first_kernel<1>, dim3(blocks), dim3(256), smem, s, x_nchw, w, kp, bias, relu, cout, B, H, W, y, ldy, band)); break;
    case 3: IRB_CUDA(launch_pdl(conv3x3_first_kernel<3>, dim3(blocks), dim3(256), smem, s, x_nchw, w, kp, bias, relu, cout, B, H, W, y, ldy, band)); break;
    default: IRB_CUDA(launch_pdl(conv3x3_first_kernel<6>, dim3(blocks), dim3(256), smem, s, x_nchw, w, kp, bias, relu, cout, B, H, W, y, ldy, band)); break;
  }
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// standalone channel LayerNorm, one warp per pixel (C <= 1024), two-pass statistics in registers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

template <typename TY>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int ldx, TY* __restrict__ y,
                                                        int ldy, long long rows, int C, int ln_mode,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        int round_tf32) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int f4n = C >> 2;
  pdl_sync();
  for (long long row = warp0; row < rows; row += nwarps) {
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = lane + 32 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < f4n) v[i] = *reinterpret_cast<const float4*>(x + row * ldx + 4 * f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    s = warp_sum(s);
    const float mu = s / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < f4n) {
        const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
        ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    ss = warp_sum(ss);
    const float rstd = 1.0f / sqrtf(ss / (float)C + 1e-5f);
    const float sub = ln_mode == LN_WITHBIAS ? mu : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = lane + 32 * i;
      if (f < f4n) {
        const float4 g = *reinterpret_cast<const float4*>(w + 4 * f);
        float4 o;
        o.x = (v[i].x - sub) * rstd * g.x; o.y = (v[i].y - sub) * rstd * g.y;
        o.z = (v[i].z - sub) * rstd * g.z; o.w = (v[i].w - sub) * rstd * g.w;
        if (ln_mode == LN_WITHBIAS) {
          const float4 bb = *reinterpret_cast<const float4*>(b + 4 * f);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        if constexpr (sizeof(TY) == 4) {
          if (round_tf32) o = make_float4(rna_tf32(o.x), rna_tf32(o.y), rna_tf32(o.z), rna_tf32(o.w));
          *reinterpret_cast<float4*>(y + row * ldy + 4 * f) = o;
        } else {
          uint2 t;
          *reinterpret_cast<__half2*>(&t.x) = f2h2_sat(o.x, o.y);
          *reinterpret_cast<__half2*>(&t.y) = f2h2_sat(o.z, o.w);
          *reinterpret_cast<uint2*>(y + row * ldy + 4 * f) = t;
        }
      }
    }
  }
}

// Narrow rows (C = 4 * LPR * V, e.g. 48 / 96 / 192 / 384 with V = 3): LPR lanes share a pixel, a warp normalises
// 32 / LPR pixels per pass with every load a full 16-byte lane access and 64-128 contiguous bytes per pixel segment.
// Same two-pass statistics as above.
template <typename TY, int LPR, int V>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, TY* __restrict__ y, long long rows,
                                                             int ln_mode, const float* __restrict__ w,
                                                             const float* __restrict__ b, int round_tf32) {
  constexpr int C = 4 * LPR * V, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, l = lane % LPR, r = lane / LPR;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 g[V], bb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(w) + i * LPR + l);
    bb[i] = ln_mode == LN_WITHBIAS ? __ldg(reinterpret_cast<const float4*>(b) + i * LPR + l) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_sync();
  for (long long row0 = warp0 * RPW; row0 < rows; row0 += nwarps * RPW) {
    const long long row = row0 + r;
    const bool ok = row < rows;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = ok ? __ldcs(reinterpret_cast<const float4*>(x + row * C) + i * LPR + l) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
      ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = 1.0f / sqrtf(ss / (float)C + 1e-5f);
    const float sub = ln_mode == LN_WITHBIAS ? mu : 0.f;
    if (!ok) continue;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = (v[i].x - sub) * rstd * g[i].x + bb[i].x; o.y = (v[i].y - sub) * rstd * g[i].y + bb[i].y;
      o.z = (v[i].z - sub) * rstd * g[i].z + bb[i].z; o.w = (v[i].w - sub) * rstd * g[i].w + bb[i].w;
      if constexpr (sizeof(TY) == 4) {
        if (round_tf32) o = make_float4(rna_tf32(o.x), rna_tf32(o.y), rna_tf32(o.z), rna_tf32(o.w));
        reinterpret_cast<float4*>(y + row * C)[i * LPR + l] = o;
      } else {
        uint2 t;
        *reinterpret_cast<__half2*>(&t.x) = f2h2_sat(o.x, o.y);
        *reinterpret_cast<__half2*>(&t.y) = f2h2_sat(o.z, o.w);
        reinterpret_cast<uint2*>(y + row * C)[i * LPR + l] = t;
      }
    }
  }
}

template <int LPR>
static int launch_layernorm_rows(const float* x, void* y, int y_fmt, long long rows, int ln_mode, const float* w,
                                 const float* b, cudaStream_t s) {
  const int y_half = y_fmt == 1, rnd = y_fmt == 2;
  constexpr int RPB = 8 * (32 / LPR);      // pixels per block pass
  const long long need = cdivll(rows, RPB);
  const int blocks = (int)(need < 148LL * 8 ? (need > 0 ? need : 1) : 148LL * 8);
  if (y_half) IRB_CUDA(launch_pdl(layernorm_rows_kernel<__half, LPR, 3>, dim3(blocks), dim3(256), 0, s, x, (__half*)y, rows, ln_mode, w, b, 0));
  else        IRB_CUDA(launch_pdl(layernorm_rows_kernel<float, LPR, 3>, dim3(blocks), dim3(256), 0, s, x, (float*)y, rows, ln_mode, w, b, rnd));
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// y_fmt: 0 fp32, 1 fp16, 2 fp32 rounded to tf32 (round-to-nearest; the consumer is a kind::tf32 MMA reading the tensor
// straight from its TMA box, and the tensor core TRUNCATES the low 13 mantissa bits of an unrounded operand)
int launch_layernorm(const float* x, int ldx, void* y, int ldy, int y_fmt, long long rows, int C, int ln_mode,
                     const float* w, const float* b, cudaStream_t s) {
  const int y_half = y_fmt == 1;
  IRB_REQUIRE(C % 4 == 0 && C <= 1024 && ldx % 4 == 0 && ldy % 4 == 0, "layernorm: C must be a multiple of 4, <= 1024");
  IRB_REQUIRE(ln_mode == LN_BIASFREE || ln_mode == LN_WITHBIAS, "layernorm: bad mode");
  const long long blocks_needed = cdivll(rows, 8);
  const int blocks = (int)(blocks_needed < 148LL * 8 ? blocks_needed : 148LL * 8);
  ProfScope prof(TAG_LAYERNORM, (y_half ? 6.0 : 8.0) * (double)rows * C, 0.0, s);
  if (ldx == C && ldy == C && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0) {
    switch (C) {
      case 48:  return launch_layernorm_rows<4>(x, y, y_fmt, rows, ln_mode, w, b, s);
      case 96:  return launch_layernorm_rows<8>(x, y, y_fmt, rows, ln_mode, w, b, s);
      case 192: return launch_layernorm_rows<16>(x, y, y_fmt, rows, ln_mode, w, b, s);
      case 384: return launch_layernorm_rows<32>(x, y, y_fmt, rows, ln_mode, w, b, s);
      default: break;
    }
  }
  if (y_half) IRB_CUDA(launch_pdl(layernorm_kernel<__half>, dim3(blocks > 0 ? blocks : 1), dim3(256), 0, s, x, ldx, (__half*)y, ldy, rows, C, ln_mode, w, b, 0));
  else        IRB_CUDA(launch_pdl(layernorm_kernel<float>, dim3(blocks > 0 ? blocks : 1), dim3(256), 0, s, x, ldx, (float*)y, ldy, rows, C, ln_mode, w, b, y_fmt == 2));
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

// ---------------------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------------------
__global__ void copy_channels_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                     long long rows, int c4) {
  pdl_sync();
  const long long total = rows * c4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / c4;
    const int c = (int)(idx - r * c4) * 4;
    *reinterpret_cast<float4*>(dst + r * ldd + c) = *reinterpret_cast<const float4*>(src + r * lds + c);
  }
}

int launch_copy_channels(const float* src, int lds, float* dst, int ldd, long long rows, int C, cudaStream_t s) {
  IRB_REQUIRE(C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "copy_channels: multiples of 4 required");
  const long long total = rows * (C / 4);
  const int blocks = (int)(cdivll(total, 256) < 148LL * 16 ? cdivll(total, 256) : 148LL * 16);
  ProfScope prof(TAG_COPY, 8.0 * (double)rows * C, 0.0, s);
  IRB_CUDA(launch_pdl(copy_channels_kernel, dim3(blocks > 0 ? blocks : 1), dim3(256), 0, s, src, lds, dst, ldd, rows, C / 4));
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
  const long long total = (long long)B * C * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long pix = idx / C;           // b*HW + p
    const long long b = pix / HW, pp = pix - b * HW;
    dst[idx] = src[(b * C + c) * HW + pp];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW) {
  const long long total = (long long)B * C * HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pp = idx % HW;
    const long long bc = idx / HW;
    const long long b = bc / C, c = bc - b * C;
    dst[idx] = src[(b * HW + pp) * C + c];
  }
}
int launch_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t s) {
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)(cdivll(total, 256) < 148LL * 16 ? cdivll(total, 256) : 148LL * 16);
  nchw_to_nhwc_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(src, dst, B, C, H * W);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}
int launch_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t s) {
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)(cdivll(total, 256) < 148LL * 16 ? cdivll(total, 256) : 148LL * 16);
  nhwc_to_nchw_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(src, dst, B, C, H * W);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace irb
