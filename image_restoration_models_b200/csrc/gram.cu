// MDTA Gram partials on tensor cores:  S[i][j] = sum_p q[p][i] * k[p][j]  per (image, head, pixel slice),
// plus the squared row norms sum_p q[p][i]^2, sum_p k[p][j]^2 that F.normalize needs (restormer.py:121-124).
// The normalisation itself is applied afterwards as a rank-1 scaling of S (softmax_fold kernel), so q and k
// are read exactly once here.
//
// HBM-bound (2*ch FLOP per byte, ch = 48 or 96): the kernel streams 64-pixel chunks of the q and k channel
// groups of one head through a double-buffered shared-memory tile (global loads of chunk i+1 are in flight
// while chunk i is multiplied), and runs the ch x ch x 64 product with mma.sync m16n8k8 tf32 with fp32
// accumulation.  The 8 warps split the pixels (and, for ch = 96, the output quadrants); accumulators stay in
// registers across all chunks of the slice and are reduced once through shared memory at the end.
// Head dims other than 48 / 96 use the CUDA-core kernel in simt_kernels.cu.
#include "common.cuh"

namespace irb {

namespace {

constexpr int PT = 64;   // pixels per staged chunk

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// raw 4-channel vectors: loaded unconditionally from a clamped row, converted and masked when staged (a load or
// its conversion under a condition serialises the loads of a chunk, see dwconv.cu)
template <typename T> struct Raw4;
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__half> { using type = uint2; };
template <typename T> __device__ __forceinline__ typename Raw4<T>::type ldraw(const T* p) {
  return __ldg(reinterpret_cast<const typename Raw4<T>::type*>(p));
}
__device__ __forceinline__ float4 cvt4(const float4& t) { return t; }
__device__ __forceinline__ float4 cvt4(const uint2& t) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// CH = head dim; QS = quadrant split per dimension (1 for 48, 2 for 96); PG = pixel groups = 8 / (QS*QS)
template <typename T, int CH, int QS>
__global__ void __launch_bounds__(256) gram_mma_kernel(const GramParams p) {
  constexpr int LD = CH + 4;                 // smem row stride: pixels are 2*LD apart, 2*LD % 32 == 8 -> conflict-free fragment loads
  constexpr int PG = 8 / (QS * QS);
  constexpr int MT = CH / 16 / QS;           // 16-row tiles per warp
  constexpr int NT = CH / 8 / QS;            // 8-column tiles per warp
  constexpr int V = 2 * CH / 4;              // float4 per pixel (q and k)
  constexpr int LPT = PT * V / 256;          // float4 loads per thread per chunk
  static_assert(PT * V % 256 == 0, "chunk must divide evenly over the block");
  extern __shared__ float sm[];              // [2 buffers][PT][2][LD]  (q row, k row per pixel)
  const T* __restrict__ qkv = reinterpret_cast<const T*>(p.qkv);

  const int part = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int pg = warp % PG, quad = warp / PG;
  const int m0 = (quad / QS) * (CH / QS), n0 = (quad % QS) * (CH / QS);
  const int per = cdiv(p.HW, p.nparts);
  const int pbeg = part * per, pend = min(p.HW, pbeg + per);
  const T* base = qkv + (long long)b * p.HW * p.ld;
  const int qoff = head * CH, koff = p.C + head * CH;

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  float nrm = 0.f;                            // tid < 2*CH: running sum of squares of column tid (q then k)

  typename Raw4<T>::type stage[LPT];
  auto issue_loads = [&](int ps) {
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
      const int e = tid + l * 256;
      const int r = e / V, c = e % V;         // pixel in chunk, float4 column (q: [0, CH/4), k: [CH/4, CH/2))
      const int row = min(ps + r, pend - 1);  // clamped; rows past the slice are zeroed when staged
      const int off = (c < CH / 4) ? qoff + 4 * c : koff + 4 * (c - CH / 4);
      stage[l] = ldraw<T>(base + (long long)row * p.ld + off);
    }
  };
  auto store_stage = [&](int buf, int ps) {
    float* dst = sm + (size_t)buf * PT * 2 * LD;
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
      const int e = tid + l * 256;
      const int r = e / V, c = e % V;
      float* d = dst + (r * 2 + (c < CH / 4 ? 0 : 1)) * LD + 4 * (c < CH / 4 ? c : c - CH / 4);
      const float4 t = cvt4(stage[l]);
      const bool live = ps + r < pend;
      // operands are rounded to tf32 once here; the row norms below use the same rounded values
      *reinterpret_cast<float4*>(d) =
          make_float4(live ? __uint_as_float(f2tf32(t.x)) : 0.f, live ? __uint_as_float(f2tf32(t.y)) : 0.f,
                      live ? __uint_as_float(f2tf32(t.z)) : 0.f, live ? __uint_as_float(f2tf32(t.w)) : 0.f);
    }
  };

  int buf = 0;
  if (pbeg < pend) { issue_loads(pbeg); store_stage(0, pbeg); }
  __syncthreads();
  for (int ps = pbeg; ps < pend; ps += PT) {
    const bool more = ps + PT < pend;
    if (more) issue_loads(ps + PT);           // in flight during the MMAs of this chunk
    const float* cur = sm + (size_t)buf * PT * 2 * LD;
    // this warp's pixels: PT / PG consecutive pixels, in k-steps of 8
#pragma unroll
    for (int ks = 0; ks < PT / PG / 8; ++ks) {
      const int pr = pg * (PT / PG) + ks * 8;
      uint32_t af[MT][4], bf[NT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float* q0 = cur + ((pr + t) * 2) * LD + m0 + i * 16 + g;
        const float* q1 = cur + ((pr + t + 4) * 2) * LD + m0 + i * 16 + g;
        af[i][0] = __float_as_uint(q0[0]); af[i][1] = __float_as_uint(q0[8]);
        af[i][2] = __float_as_uint(q1[0]); af[i][3] = __float_as_uint(q1[8]);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        bf[j][0] = __float_as_uint(cur[((pr + t) * 2 + 1) * LD + n0 + j * 8 + g]);
        bf[j][1] = __float_as_uint(cur[((pr + t + 4) * 2 + 1) * LD + n0 + j * 8 + g]);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_tf32(acc[i][j], af[i], bf[j]);
    }
    if (tid < 2 * CH) {
      const float* col = cur + (tid < CH ? 0 : LD) + (tid < CH ? tid : tid - CH);
#pragma unroll 8
      for (int r = 0; r < PT; ++r) { const float v = col[r * 2 * LD]; nrm = fmaf(v, v, nrm); }
    }
    if (more) store_stage(buf ^ 1, ps + PT);
    __syncthreads();
    buf ^= 1;
  }

  // cross-warp reduction over the pixel groups through shared memory (disjoint quadrants, PG rounds)
  float* red = sm;                            // [CH][CH]
  for (int e = tid; e < CH * CH; e += 256) red[e] = 0.f;
  __syncthreads();
  for (int round = 0; round < PG; ++round) {
    if (pg == round) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const int r = m0 + i * 16 + g, c = n0 + j * 8 + 2 * t;
          red[r * CH + c] += acc[i][j][0];       red[r * CH + c + 1] += acc[i][j][1];
          red[(r + 8) * CH + c] += acc[i][j][2]; red[(r + 8) * CH + c + 1] += acc[i][j][3];
        }
    }
    __syncthreads();
  }
  float* sp = p.s_part + (((long long)b * p.heads + head) * p.nparts + part) * CH * CH;
  for (int e = tid; e < CH * CH; e += 256) sp[e] = red[e];
  if (tid < 2 * CH) {
    float* np_ = p.n_part + (((long long)b * p.heads + head) * p.nparts + part) * 2 * CH;
    np_[tid] = nrm;
  }
}

template <typename T, int CH, int QS>
int launch_mma(const GramParams& p, cudaStream_t s) {
  constexpr int LD = CH + 4;
  const size_t smem = std::max((size_t)2 * PT * 2 * LD * sizeof(float), (size_t)CH * CH * sizeof(float));
  static SmemOptIn optin;
  if (smem > 48 * 1024) IRB_TRY(opt_in_smem(gram_mma_kernel<T, CH, QS>, optin, (int)smem));
  dim3 grid(p.nparts, p.heads, p.B);
  gram_mma_kernel<T, CH, QS><<<grid, 256, smem, s>>>(p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace

int launch_gram(const GramParams& p, cudaStream_t s) {
  const int ch = p.C / p.heads;
  IRB_REQUIRE(p.C % p.heads == 0 && ch % 16 == 0 && ch <= 128 && p.ld % 4 == 0, "gram: head dim must be a multiple of 16, <= 128");
  const double es = p.in_half ? 2.0 : 4.0;
  ProfScope prof(TAG_GRAM, es * (double)p.B * p.HW * 2.0 * p.C, 2.0 * (double)p.B * p.HW * p.C * ch, s);
  if (ch == 48) return p.in_half ? launch_mma<__half, 48, 1>(p, s) : launch_mma<float, 48, 1>(p, s);
  if (ch == 96) return p.in_half ? launch_mma<__half, 96, 2>(p, s) : launch_mma<float, 96, 2>(p, s);
  IRB_REQUIRE(!p.in_half, "gram: fp16 input is implemented for head dims 48 and 96");
  return launch_gram_ref(p, s);
}

}  // namespace irb
