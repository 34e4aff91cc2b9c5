// Packed-pair fp32 arithmetic and the exact-erf GELU gate shared by the GDFN kernels (ffn_tail.cu, ffn_fused.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace irb {
namespace gdfn {

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2): a dw thread owns two adjacent channels, so every multiply-add of the
// depthwise taps and of the GELU polynomial is one instruction for both channels.  The kernel is bound by FP32
// instruction issue, not by bytes: this halves the issue slots of its inner loops.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// gelu(x) * gate for a channel pair.  Exact-erf GELU (F.gelu default, restormer.py:91) through the Abramowitz-Stegun
// 7.1.26 rational form, erf(u) = 1 - (a1 t + .. + a5 t^5) exp(-u^2), t = 1/(1 + p u), u = |x|/sqrt2 (|gelu error| <=
// 3e-7 in fp32, see dwconv.cu).  z = u * sqrt(log2 e), so that exp(-u^2) = 2^(-z^2) is a bare MUFU.EX2.
__device__ __forceinline__ f2_t gelu_gate2(f2_t x, f2_t gate) {
  const f2_t C1 = pack2(0.84932180028801904272f, 0.84932180028801904272f);       // sqrt(log2 e) / sqrt 2
  const float PZ = 0.27273629f;                                                    // 0.3275911 / sqrt(log2 e)
  const f2_t A5 = pack2(-1.061405429f, -1.061405429f), A4 = pack2(1.453152027f, 1.453152027f),
             A3 = pack2(-1.421413741f, -1.421413741f), A2 = pack2(0.284496736f, 0.284496736f),
             A1 = pack2(-0.254829592f, -0.254829592f), ONE = pack2(1.0f, 1.0f), HALF = pack2(0.5f, 0.5f);
  const f2_t z = mul2(x, C1);
  float zx, zy, xx, xy;
  unpack2(z, zx, zy);
  unpack2(x, xx, xy);
  const f2_t t = pack2(rcp_approx(fmaf(PZ, fabsf(zx), 1.0f)), rcp_approx(fmaf(PZ, fabsf(zy), 1.0f)));
  const f2_t sq = mul2(z, z);
  float sx, sy;
  unpack2(sq, sx, sy);
  const f2_t e = pack2(ex2_approx(-sx), ex2_approx(-sy));
  f2_t poly = fma2(t, A5, A4);
  poly = fma2(t, poly, A3);
  poly = fma2(t, poly, A2);
  poly = fma2(t, poly, A1);
  poly = mul2(poly, t);                                  // -(a1 t + ... + a5 t^5)
  const f2_t y = fma2(poly, e, ONE);                     // erf(|x| / sqrt 2)
  float yx, yy;
  unpack2(y, yx, yy);
  const f2_t ys = pack2(copysignf(yx, xx), copysignf(yy, xy));
  const f2_t hx = mul2(x, HALF);
  return mul2(fma2(hx, ys, hx), gate);                   // 0.5 x (1 + erf(x / sqrt 2)) * gate
}

// gelu(x) * gate with ONE special-function instruction per element.  For a = |x| the tail of the Gaussian is
// erfc(a / sqrt 2) = 2^(-Q(a)) with Q smooth and Q(0) = 0, and
//     gelu(x) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2) = max(x, 0) - a * 2^(-(Q(a) + 1))
// (x > 0: x - 0.5 x erfc; x <= 0: 0.5 x erfc).  Q is a polynomial without constant term: one packed FMA per degree (the last
// one adds the 1 that halves the result), one MUFU.EX2, no reciprocal, no select.  The error of the polynomial is weighted by
// erfc itself, i.e. it vanishes where gelu is large.  Fit: scripts/fit_gelu_exp2.py (weighted minimax on [0, 6], |x| clamped
// to 6 where erfc < 2e-9).  max |gelu error| in fp32 arithmetic: degree 6 3.1e-7 (the rounding level of the subtraction
// itself), degree 5 7.1e-7, degree 4 8.7e-6.  The result is rounded to a 16-bit tensor-core operand right away (2^-11
// relative), so degree 5 is three orders of magnitude below what the consumer can see.  Measured on B200: with the first
// TMEM-direct kernels the degree bought nothing (fused GDFN C = 96 0.765 ms against 0.751-0.753 ms: the gate's FMAs were not
// on the critical path of the lock-step phases); after the depthwise warps' instruction diet degree 5 is 1-2 % faster than
// degree 6 (C = 96 0.713 against 0.721 ms, C = 48 0.360 against 0.367 ms, A/B on one box), so the library uses degree 5;
// IRB_GELU_DEG = 4 / 6 select the others.  The A&S form above costs two MUFU per element (rcp + ex2) and four more packed
// operations.
#ifndef IRB_GELU_DEG
#define IRB_GELU_DEG 5
#endif
__device__ __forceinline__ f2_t gelu_gate2e(f2_t x, f2_t gate) {
#if IRB_GELU_DEG == 6
  const f2_t C6 = pack2(-2.992467082811e-05f, -2.992467082811e-05f), C5 = pack2(7.398781788458e-04f, 7.398781788458e-04f),
             C4 = pack2(-7.977474556594e-03f, -7.977474556594e-03f), C3 = pack2(5.323820251441e-02f, 5.323820251441e-02f),
             C2 = pack2(4.589156769318e-01f, 4.589156769318e-01f), C1 = pack2(1.151147084379f, 1.151147084379f);
#elif IRB_GELU_DEG == 5
  const f2_t C5 = pack2(4.88102132e-04f, 4.88102132e-04f), C4 = pack2(-7.19871846e-03f, -7.19871846e-03f),
             C3 = pack2(5.21466320e-02f, 5.21466320e-02f), C2 = pack2(4.59595844e-01f, 4.59595844e-01f),
             C1 = pack2(1.15100054f, 1.15100054f);
#else
  const f2_t C4 = pack2(-4.16167e-03f, -4.16167e-03f), C3 = pack2(4.573546e-02f, 4.573546e-02f),
             C2 = pack2(4.6493045e-01f, 4.6493045e-01f), C1 = pack2(1.14956698f, 1.14956698f);
#endif
  const f2_t ONE = pack2(1.0f, 1.0f), NEG1 = pack2(-1.0f, -1.0f);
  float xx, xy;
  unpack2(x, xx, xy);
  const f2_t a = pack2(fminf(fabsf(xx), 6.0f), fminf(fabsf(xy), 6.0f));
#if IRB_GELU_DEG == 6
  f2_t q = fma2(C6, a, C5);
  q = fma2(q, a, C4);
  q = fma2(q, a, C3);
#elif IRB_GELU_DEG == 5
  f2_t q = fma2(C5, a, C4);
  q = fma2(q, a, C3);
#else
  f2_t q = fma2(C4, a, C3);
#endif
  q = fma2(q, a, C2);
  q = fma2(q, a, C1);
  q = fma2(q, a, ONE);                                          // Q(a) + 1
  float qx, qy;
  unpack2(q, qx, qy);
  const f2_t t = mul2(a, pack2(ex2_approx(-qx), ex2_approx(-qy)));   // |x| * 0.5 erfc(|x| / sqrt 2)
  const f2_t g = fma2(t, NEG1, pack2(fmaxf(xx, 0.f), fmaxf(xy, 0.f)));
  return mul2(g, gate);
}

// Measured on ffn_fused.cu: alternating packed FFMA2 with pairs of scalar FFMA (to use both halves of the FP32 datapath)
// is 6 % slower than packed-only -- FFMA with three register operands issues every second cycle, so two scalar
// instructions cost what one FFMA2 costs and take twice the issue slots.
}  // namespace gdfn
}  // namespace irb
