// Depthwise 3x3 convolution on channels-last activations (zero padding 1, cross-correlation), optionally fused
// with the GDFN gate gelu(x1) * x2.  Replaces Attention.qkv_dwconv (restormer.py:106) and FeedForward.dwconv +
// gating (restormer.py:83,90-91).  Purely HBM-bound (18..36 FLOP per 16..24 bytes), so the kernel is built to
// touch every input vector once per band instead of nine times:
//
//   * a thread owns 4 consecutive channels x WT consecutive pixels of a row and walks down a band of R rows;
//   * for each input row it loads WT+2 16-byte vectors (8-byte for fp16) and scatters them into three rolling
//     accumulator rows (the output rows above / at / below), so each loaded vector is used for 9 FMAs per
//     channel and re-read only by the neighbouring x-tile (L1) and the neighbouring band (L2);
//   * consecutive lanes own consecutive channel vectors of the same pixels -> fully coalesced 512-byte segments;
//   * the 3x3 weights of the block's channel slice live in shared memory.
#include "common.cuh"
#include "gdfn_math.cuh"

#include <type_traits>

namespace irb {

namespace {

// erf-form GELU (F.gelu default, restormer.py:91): 0.5 x (1 + erf(x / sqrt 2)).  The gated kernel is instruction-issue
// bound, and libdevice erff costs ~40 instructions per element; erf is evaluated with the Abramowitz-Stegun 7.1.26
// rational form instead (1 MUFU.RCP + 1 MUFU.EX2 + 7 FMA).  Measured in fp32 against the exact function over
// [-8, 8]: |erf error| <= 6.1e-7, |gelu error| <= 2.6e-7 -- three orders of magnitude inside the 1e-3 parity budget.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));   // MUFU.RCP (the IEEE __frcp_rn is a subroutine call)
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f),
                              0.254829592f);
  const float y = fmaf(-poly, __expf(-ax * ax), 1.0f);          // erf(|x| / sqrt 2)
  return 0.5f * x * (1.0f + copysignf(y, x));
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Raw 4-channel vectors: loads are issued unconditionally (clamped address) and converted / masked afterwards.
// A load placed under a condition, or with its conversion inside the condition, gets funnelled through one
// register set by the compiler and the loads of a row then serialise (measured: 3x slower).
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
template <typename T> __device__ __forceinline__ void st4(T* p, const float4& v);
template <> __device__ __forceinline__ void st4<float>(float* p, const float4& v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void st4<__half>(__half* p, const float4& v) {
  uint2 t;
  *reinterpret_cast<__half2*>(&t.x) = f2h2_sat(v.x, v.y);
  *reinterpret_cast<__half2*>(&t.y) = f2h2_sat(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ void fma4(float4& acc, const float4& w, const float4& x) {
  acc.x = fmaf(w.x, x.x, acc.x); acc.y = fmaf(w.y, x.y, acc.y);
  acc.z = fmaf(w.z, x.z, acc.z); acc.w = fmaf(w.w, x.w, acc.w);
}

struct DwGeom { int cvb, xb, rows; };   // channel vectors per block, x-tiles per block, band height

template <typename TI, typename TO, int WT, bool GATE>
__global__ void __launch_bounds__(256, 2) dw_roll_kernel(const DwParams p, const DwGeom g) {
  constexpr int NS = GATE ? 2 : 1;
  constexpr int WS = NS * 9 + 1;                       // float4 per thread slot (odd -> conflict-free LDS.128)
  extern __shared__ float4 wsm[];                      // [cvb][WS]: a thread's 9 (18) tap vectors are contiguous
  const TI* __restrict__ in = reinterpret_cast<const TI*>(p.in);
  TO* __restrict__ out = reinterpret_cast<TO*>(p.out);
  const int tid = threadIdx.x;
  const int cv_total = p.C >> 2;
  const int cv0 = blockIdx.x * g.cvb;
  const int ncv = min(g.cvb, cv_total - cv0);

  for (int idx = tid; idx < NS * 9 * g.cvb; idx += 256) {
    const int cvl = idx % g.cvb, tap = (idx / g.cvb) % 9, set = idx / (9 * g.cvb);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cvl < ncv) w = *reinterpret_cast<const float4*>(p.w + tap * p.Cw + set * p.gate_off + (cv0 + cvl) * 4);
    wsm[cvl * WS + set * 9 + tap] = w;
  }
  __syncthreads();
  pdl_sync();   // the taps (constants) are staged under the previous kernel's tail (common.cuh)

  const int xl = tid / g.cvb, cvl = tid - xl * g.cvb;
  const int x0 = (blockIdx.y * g.xb + xl) * WT;
  if (xl >= g.xb || cvl >= ncv || x0 >= p.W) return;
  const int nbands = (p.H + g.rows - 1) / g.rows;
  const int b = blockIdx.z / nbands, band = blockIdx.z - b * nbands;
  const int y0 = band * g.rows, y1 = min(p.H, y0 + g.rows);
  const int c = (cv0 + cvl) * 4;
  const float4* __restrict__ wq = wsm + cvl * WS;

  // everything that does not change from row to row is computed once: clamped column offsets (zero padding is
  // applied by masking the two edge columns), the row pitch, the output offsets
  int coff[WT + 2];
  bool cmask[WT + 2];
#pragma unroll
  for (int i = 0; i < WT + 2; ++i) {
    const int x = x0 - 1 + i;
    cmask[i] = x >= 0 && x < p.W;
    coff[i] = min(max(x, 0), p.W - 1) * p.ldi;
  }
  const long long in_pitch = (long long)p.W * p.ldi, out_pitch = (long long)p.W * p.ldo;
  TO* orow = out + ((long long)b * p.H + (y0 - 1)) * out_pitch + (long long)x0 * p.ldo + c;

  float4 bias[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s)
    bias[s] = p.bias ? *reinterpret_cast<const float4*>(p.bias + s * p.gate_off + c) : make_float4(0.f, 0.f, 0.f, 0.f);

  float4 acc[NS][3][WT];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int i = 0; i < WT; ++i) acc[s][r][i] = bias[s];

  // One step consumes input row yy.  The three accumulator rows rotate roles; the rotation is resolved at compile
  // time (K = step index mod 3: row (K+0)%3 completes output yy-1, (K+1)%3 is output yy, (K+2)%3 is output yy+1),
  // so no register moves are needed between rows.
  // The raw vectors of row yy+1 are requested BEFORE row yy's taps run (ncu: 32 % of the warp samples of the gated
  // kernel sat on the first conversion of a freshly loaded vector -- with 16 warps per SM and ~200 instructions per row
  // nothing else covered the load latency).  Rows outside the image are fetched from a clamped address and ignored.
  const TI* imgp = in + (long long)b * p.H * in_pitch + c;
  typename Raw4<TI>::type raw[NS][WT + 2];
  auto fetch = [&](int yy) {
    const TI* rp = imgp + (long long)min(max(yy, 0), p.H - 1) * in_pitch;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < WT + 2; ++i) raw[s][i] = ldraw<TI>(rp + coff[i] + s * p.gate_off);
  };
  // (fp16 rows only: the fp32 vectors of two rows in flight do not fit the 128 registers of two resident blocks)
  constexpr bool PREFETCH = std::is_same<TI, __half>::value;
  if (PREFETCH) fetch(y0 - 1);
  auto step = [&](auto kc, int yy) {
    constexpr int K = decltype(kc)::value;
    if (!PREFETCH) fetch(yy);
    float4 v[NS][WT + 2];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < WT + 2; ++i) {
        const float4 t = cvt4(raw[s][i]);
        v[s][i] = cmask[i] ? t : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    if (PREFETCH && yy < y1) fetch(yy + 1);  // the next step's row (the band ends with row y1)
    if (yy >= 0 && yy < p.H) {
      // input row yy feeds output rows yy-1 (tap row 2), yy (tap row 1), yy+1 (tap row 0)
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 w = wq[s * 9 + (2 - r) * 3 + dx];
#pragma unroll
            for (int i = 0; i < WT; ++i) fma4(acc[s][(K + r) % 3][i], w, v[s][i + dx]);
          }
    }
    if (yy - 1 >= y0) {                     // output row yy-1 is complete (orow points at row yy)
      TO* o_ = orow - out_pitch;
#pragma unroll
      for (int i = 0; i < WT; ++i) {
        if (x0 + i < p.W) {
          float4 o = acc[0][K % 3][i];
          if (GATE) {
            const float4 gt = acc[NS - 1][K % 3][i];
            // gelu(x1) * x2 on packed pairs, one MUFU per element (gdfn_math.cuh; max error 7.1e-7)
            const gdfn::f2_t g0 = gdfn::gelu_gate2e(gdfn::pack2(o.x, o.y), gdfn::pack2(gt.x, gt.y));
            const gdfn::f2_t g1 = gdfn::gelu_gate2e(gdfn::pack2(o.z, o.w), gdfn::pack2(gt.z, gt.w));
            gdfn::unpack2(g0, o.x, o.y);
            gdfn::unpack2(g1, o.z, o.w);
          }
          if (std::is_same<TO, float>::value && p.round_tf32) { o.x = rna_tf32(o.x); o.y = rna_tf32(o.y); o.z = rna_tf32(o.z); o.w = rna_tf32(o.w); }
          st4<TO>(o_ + (long long)i * p.ldo, o);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int i = 0; i < WT; ++i) acc[s][K % 3][i] = bias[s];      // becomes output row yy+2
    orow += out_pitch;
  };
  for (int yy = y0 - 1; yy <= y1; yy += 3) {
    step(std::integral_constant<int, 0>{}, yy);
    if (yy + 1 <= y1) step(std::integral_constant<int, 1>{}, yy + 1);
    if (yy + 2 <= y1) step(std::integral_constant<int, 2>{}, yy + 2);
  }
}

template <typename TI, typename TO>
int launch_typed(const DwParams& p, cudaStream_t s) {
  const int cv = p.C / 4;
  DwGeom g;
  if (cv <= 64) {
    g.cvb = cv;
  } else {
    g.cvb = 32;
    for (int d = 64; d >= 16; --d)
      if (cv % d == 0) { g.cvb = d; break; }
  }
  g.xb = 256 / g.cvb;
  g.rows = 16;
  const int wt = p.gate ? 2 : 4;
  dim3 grid(cdiv(cv, g.cvb), cdiv(p.W, g.xb * wt), p.B * cdiv(p.H, g.rows));
  const size_t smem = (size_t)((p.gate ? 2 : 1) * 9 + 1) * g.cvb * sizeof(float4);
  if (p.gate) IRB_CUDA(launch_pdl(dw_roll_kernel<TI, TO, 2, true>, grid, dim3(256), smem, s, p, g));
  else        IRB_CUDA(launch_pdl(dw_roll_kernel<TI, TO, 4, false>, grid, dim3(256), smem, s, p, g));
  return IR_OK;
}

}  // namespace

int launch_dwconv(const DwParams& p, cudaStream_t s) {
  IRB_REQUIRE(p.C % 4 == 0 && p.ldi % 4 == 0 && p.ldo % 4 == 0 && p.Cw % 4 == 0 && p.gate_off % 4 == 0,
              "dwconv: channel counts must be multiples of 4");
  IRB_REQUIRE((long long)p.B * cdiv(p.H, 16) <= 65535, "dwconv: too many row bands for one launch");
  const double pix = (double)p.B * p.H * p.W;
  const double ies = p.in_half ? 2.0 : 4.0, oes = p.out_half ? 2.0 : 4.0;
  ProfScope prof(p.tag, pix * p.C * ((p.gate ? 2.0 : 1.0) * ies + oes), 2.0 * 9.0 * pix * p.C * (p.gate ? 2.0 : 1.0), s);
  if (!p.in_half && !p.out_half) return launch_typed<float, float>(p, s);
  if (p.in_half && p.out_half) return launch_typed<__half, __half>(p, s);
  IRB_REQUIRE(false, "dwconv: unsupported type combination");
  return IR_OK;
}

}  // namespace irb
