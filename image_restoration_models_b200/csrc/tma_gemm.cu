// TMA-fed tcgen05 contraction for the 1x1 convolutions of the Restormer block at the high-resolution levels:
//
//     y[pixel, n] = (r[pixel, n]) + sum_k A(pixel, k) * W[n, k]        M = 128 pixels per tile, N <= 256 per CTA
//
// Same reference scope as tc_gemm.cu (restormer.py:105 qkv + LayerNorm :37-39/:54-57, :127-131 attention output,
// :82 project_in, :86 project_out, residual adds :147-148).  The first kernel moved every byte through registers
// (LDG -> LayerNorm -> STS, LDTM -> STS -> LDS -> STG); ncu showed its four epilogue warps stalled on the
// write-after-read scoreboard of their own STG instructions and the producers idle between load bursts
// (profiles/r01p_*).  Here every global access is an asynchronous bulk-tensor copy:
//
//   A-producer (1 thread)  cp.async.bulk.tensor loads of [128 px][128 B] boxes (SWIZZLE_128B) into a ring, many in flight
//   transform (4 warps)    LayerNorm prologue only: one thread per pixel row reads its row from the swizzled boxes
//                          (conflict-free), two-pass statistics, normalises IN PLACE (tf32 operands) or into a
//                          second fp16 operand ring
//   MMA (1 thread)         tcgen05.mma kind::tf32 / kind::f16, SWIZZLE_128B K-major descriptors, weights CTA-resident
//                          (pre-swizzled in HBM, fetched with one bulk copy per K box), double-buffered TMEM accumulator
//   epilogue (4 warps)     tcgen05.ld -> (+bias, +residual) -> swizzled smem box -> cp.async.bulk.tensor STORE;
//                          the residual arrives through its own TMA ring (R-producer thread) and is updated in place
//
// Out-of-range rows / columns need no code: TMA zero-fills loads and clips stores at the tensor extents
// ([B][H*W][C] maps, so a tile never straddles two images).
#include "common.cuh"
#include "tc_gemm.cuh"

#include "sm100.cuh"
#include "tmap.cuh"

#include <type_traits>

#include <cstdlib>

namespace irb {

namespace {

constexpr int TM = 128;                  // pixels per tile == UMMA M
constexpr int BOX = TM * 128;            // bytes of one [128 rows][128 B] box
constexpr int WBOX = 32 * 128;           // one warp's [32 rows][128 B] slice
constexpr int EPI_WARPS = 4, XF_WARPS = 4;
constexpr int WARP_A = 8, WARP_MMA = 9, WARP_R = 10;
constexpr int NTHREADS = 11 * 32;
constexpr int MAX_S = 12, MAX_OP = 4, MAX_RB = 8;
constexpr int HDR = 1024;

struct Bars {
  unsigned long long a_full[MAX_S], a_empty[MAX_S], a_ready[MAX_S];
  unsigned long long op_ready[MAX_OP], op_empty[MAX_OP];
  unsigned long long acc_full[2], acc_empty[2];
  unsigned long long r_full[MAX_RB], r_empty[MAX_RB];
  unsigned long long w_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= HDR, "barrier block too large");

struct TmaGemmParams {
  const uint8_t* w; long long w_bstride_bytes;    // swizzled weights [nob][N][128 B]; per-image stride (0 = shared)
  int B, HW, K, N;
  int nc;                  // output columns per CTA (grid.y chunks)
  int nkb, nob;            // raw A boxes / operand boxes per tile
  int S, SOP, RB;          // ring depths (boxes)
  int ln_mode; const float* ln_w; const float* ln_b;
  const float* bias; int has_r;
  int tiles_per_img, ntiles, per_image;
  int split;               // > 0: 1-D grid, CTAs [0, split) own N-chunk 0 and the rest chunk 1 (uneven chunks get
                           // CTA counts in proportion to their bytes); 0: blockIdx.y is the chunk
  int tmem_cols, acc_stride;
  int xn_mode;             // LnMode of the second output xn = LayerNorm(y) (fp16), 0 = none
  const float* xn_w; const float* xn_b;
  int xn_boxes;            // 64-channel boxes of an xn row (1 or 2)
  int egroups;             // epilogue warpgroups of the xn path (2 at N = 48: see xn_epilogue)
  uint32_t off_xn;         // per-warp xn staging boxes
  int wstream;             // 1: the weight chunk does not fit next to the rings (K * nc too large): every stage carries
                           // the [nc][128 B] weight box of its K box next to the A box (no LayerNorm prologue in this mode)
  uint32_t a_stride;       // bytes from one A stage to the next (BOX, or BOX + the weight box when streaming)
  uint32_t off_w, off_a, off_op, off_ro, off_ln;  // byte offsets from the 1024-aligned smem base
};

using namespace sm100;

// tile sequence of one CTA (identical in every role)
struct TileIter {
  int b, t, step, end, tpi; bool per_image;
  __device__ TileIter(const TmaGemmParams& p) {
    per_image = p.per_image != 0; tpi = p.tiles_per_img; step = gridDim.x;
    if (per_image) { b = blockIdx.z; t = blockIdx.x; end = p.tiles_per_img; }
    else { t = blockIdx.x; end = p.ntiles; b = 0; }
    if (p.split > 0) {
      const bool second = (int)blockIdx.x >= p.split;
      t = second ? blockIdx.x - p.split : blockIdx.x;
      step = second ? gridDim.x - p.split : p.split;
    }
  }
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int img() const { return per_image ? b : t / tpi; }
  __device__ int row0() const { return (per_image ? t : t % tpi) * TM; }
};

// LayerNorm prologue (restormer.py:37-39 BiasFree, :54-57 WithBias): thread r owns pixel row r of the tile.  The row
// (NKB boxes of 32 fp32 channels, K <= 32*NKB) is read once into registers -- all loads of a box are independent and
// conflict-free under the 128-byte swizzle -- the population variance is taken about the mean (two passes over the
// registers), and the normalised row goes back IN PLACE rounded to tf32, or into the fp16 operand ring.
template <typename TOp, int NKB>
__device__ __forceinline__ void ln_transform(const TmaGemmParams& p, Bars* bars, uint32_t sA, uint32_t sOP,
                                             const float* lnv, int r) {
  constexpr bool OPRING = sizeof(TOp) == 2;
  const uint32_t rsw = (uint32_t)(r & 7);
  const float inv_k = 1.0f / (float)p.K;
  const bool withbias = p.ln_mode == LN_WITHBIAS;
  const uint32_t S = (uint32_t)p.S, SOP = (uint32_t)p.SOP;
  uint32_t it = 0, ot = 0;
  for (TileIter ti(p); ti.valid(); ti.next()) {
    float4 x[NKB * 8];
    uint32_t rowa[NKB];
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      const uint32_t s = (it + kb) % S, ph = ((it + kb) / S) & 1u;
      mbar_wait(smem_u32(&bars->a_full[s]), ph);
      rowa[kb] = sA + s * p.a_stride + (uint32_t)r * 128u;
#pragma unroll
      for (int c = 0; c < 8; ++c) x[kb * 8 + c] = lds128(rowa[kb] + (((uint32_t)c ^ rsw) << 4));   // columns >= K are zero (TMA fill)
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < NKB * 8; ++i) { s0 += x[i].x; s1 += x[i].y; s2 += x[i].z; s3 += x[i].w; }
    const float mu = ((s0 + s1) + (s2 + s3)) * inv_k;
    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
    for (int i = 0; i < NKB * 8; ++i) {
      if (i * 4 < p.K) {
        const float d0 = x[i].x - mu, d1 = x[i].y - mu, d2 = x[i].z - mu, d3 = x[i].w - mu;
        s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
      }
    }
    const float rstd = rsqrtf(((s0 + s1) + (s2 + s3)) * inv_k + 1e-5f);
    const float sub = withbias ? mu : 0.f;
#pragma unroll
    for (int i = 0; i < NKB * 8; ++i) {
      if (i * 4 < p.K) {
        const float4 g = *reinterpret_cast<const float4*>(lnv + i * 4);
        const float4 bb = *reinterpret_cast<const float4*>(lnv + p.K + i * 4);
        x[i].x = fmaf((x[i].x - sub) * rstd, g.x, bb.x); x[i].y = fmaf((x[i].y - sub) * rstd, g.y, bb.y);
        x[i].z = fmaf((x[i].z - sub) * rstd, g.z, bb.z); x[i].w = fmaf((x[i].w - sub) * rstd, g.w, bb.w);
      }
    }
    if constexpr (!OPRING) {
#pragma unroll
      for (int i = 0; i < NKB * 8; ++i) {
        if (i * 4 < p.K)
          sts128(rowa[i >> 3] + (((uint32_t)(i & 7) ^ rsw) << 4),
                 make_float4(to_tf32(x[i].x), to_tf32(x[i].y), to_tf32(x[i].z), to_tf32(x[i].w)));
      }
      fence_async_smem();
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) mbar_arrive(smem_u32(&bars->a_ready[(it + kb) % S]));
    } else {
      // the raw boxes are free as soon as the row sits in registers
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) mbar_arrive(smem_u32(&bars->a_empty[(it + kb) % S]));
#pragma unroll
      for (int ob = 0; ob < (NKB + 1) / 2; ++ob, ++ot) {
        const uint32_t o = ot % SOP, ph = (ot / SOP) & 1u;
        mbar_wait(smem_u32(&bars->op_empty[o]), ph ^ 1u);
        const uint32_t orow = sOP + o * BOX + (uint32_t)r * 128u;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const int i = ob * 16 + c8 * 2;                    // float4 index of the first of the two source chunks
          uint4 t = make_uint4(0u, 0u, 0u, 0u);
          if (i + 1 < NKB * 8 && i * 4 < p.K) {
            __half2* h = reinterpret_cast<__half2*>(&t);
            h[0] = f2h2_sat(x[i].x, x[i].y); h[1] = f2h2_sat(x[i].z, x[i].w);
            h[2] = f2h2_sat(x[i + 1].x, x[i + 1].y); h[3] = f2h2_sat(x[i + 1].z, x[i + 1].w);
          }
          sts128u(orow + (((uint32_t)c8 ^ rsw) << 4), t);
        }
        fence_async_smem();
        mbar_arrive(smem_u32(&bars->op_ready[o]));
      }
    }
    it += NKB;
  }
}

// Residual epilogue that also emits xn = LayerNorm(y) (norm2, restormer.py:148) as the fp16 operand of the fused GDFN: the
// thread owns a whole pixel row (<= 3 groups of 32 columns), so the row stays in registers, the statistics are two passes
// over them, and the LayerNorm pass over HBM disappears.  NC = N as a compile-time constant (48 or 96): no predicates in the
// statistics, four independent sums.
// p.egroups == 2 (N = 48): TWO epilogue warpgroups -- warps 0-3 take the even tiles (accumulator slot 0), the otherwise idle
// transform warps 4-7 the odd ones (slot 1; a warp reaches TMEM lane quarter warp % 4 either way).  At N = 48 a tile is 74 KB
// of traffic against the same fixed chain of waits per tile, and one warpgroup could not keep up with the loads.
template <int NC>
__device__ __forceinline__ void xn_epilogue(const TmaGemmParams& p, Bars* bars, uint32_t tmem_base, uint32_t base, uint32_t sRO,
                                            const float* lnv, const CUtensorMap& tmY, const CUtensorMap& tmXn, int n0, int q,
                                            int lane, int eg) {
  constexpr int NG = (NC + 31) / 32, NV = NC / 4, GC = 32;
  const float inv_n = 1.0f / (float)NC;
  const uint32_t lsw = (uint32_t)(lane & 7);
  const uint32_t sXN = base + p.off_xn + (uint32_t)(eg * EPI_WARPS + q) * (uint32_t)p.xn_boxes * WBOX + (uint32_t)lane * 128u;
  if (lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmXn)) : "memory");
  }
  const uint32_t neg = (uint32_t)p.egroups;
  uint32_t j = 0;
  bool have_prev = false;
  uint32_t prev_gc = 0;
  for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
    if (neg == 2u && (j & 1u) != (uint32_t)eg) continue;
    const int b = ti.img(), row0 = ti.row0();
    const uint32_t slot = j & 1u;
    mbar_wait(smem_u32(&bars->acc_full[slot]), (j >> 1) & 1u);
    tc_fence_after();
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.acc_stride;
    float4 xr[NG * 8];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const uint32_t gc = j * (uint32_t)NG + (uint32_t)g;          // the box's position in the residual ring
      const int col = n0 + g * GC;
      uint32_t box;
      if (p.has_r) {
        const uint32_t s = gc % (uint32_t)p.RB, ph = (gc / (uint32_t)p.RB) & 1u;
        mbar_wait(smem_u32(&bars->r_full[s]), ph);
        box = sRO + s * BOX + (uint32_t)q * WBOX;
      } else {
        box = sRO + (uint32_t)((eg * EPI_WARPS + q) * 2 + (int)(gc & 1u)) * WBOX;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
      }
      const uint32_t myrow = box + (uint32_t)lane * 128u;
      float v[32];
      tmem_ld32(tacc + (uint32_t)(g * 32), v);
      tmem_ld_wait();
      if (g == NG - 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
      if (p.bias) {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] += (col + e < p.N) ? __ldg(p.bias + col + e) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t a = myrow + (((uint32_t)c ^ lsw) << 4);
        float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        if (p.has_r) {
          const float4 rr = lds128(a);
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        sts128(a, o);
        xr[g * 8 + c] = o;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmY, box, col, row0 + q * 32, b);
        bulk_commit();
        if (p.has_r && have_prev) {
          // this warpgroup's previous box: its store (older than the one just committed) has finished reading it
          bulk_wait_read<1>();
          mbar_arrive(smem_u32(&bars->r_empty[prev_gc % (uint32_t)p.RB]));
        }
      }
      have_prev = true; prev_gc = gc;
    }
    // statistics over the NC valid columns (TMEM columns past N were never written: not read here)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { s0 += xr[i].x; s1 += xr[i].y; s2 += xr[i].z; s3 += xr[i].w; }
    const float mu = ((s0 + s1) + (s2 + s3)) * inv_n;
    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float d0 = xr[i].x - mu, d1 = xr[i].y - mu, d2 = xr[i].z - mu, d3 = xr[i].w - mu;
      s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
    }
    const float rstd = 1.0f / sqrtf(((s0 + s1) + (s2 + s3)) * inv_n + 1e-5f);
    const float sub = p.xn_mode == LN_WITHBIAS ? mu : 0.f;
    // the previous tile's xn store has read the staging boxes (it is older than this tile's y stores)
    if (lane == 0) bulk_wait_read<NG>();
    __syncwarp();
#pragma unroll
    for (int c8 = 0; c8 < NC / 8; ++c8) {                   // 16-byte chunks of 8 fp16 channels
      uint4 t;
      __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 x = xr[c8 * 2 + e];
        const float4 gw = *reinterpret_cast<const float4*>(lnv + c8 * 8 + e * 4);
        const float4 gb = *reinterpret_cast<const float4*>(lnv + NC + c8 * 8 + e * 4);
        h[2 * e] = f2h2_sat(fmaf((x.x - sub) * rstd, gw.x, gb.x), fmaf((x.y - sub) * rstd, gw.y, gb.y));
        h[2 * e + 1] = f2h2_sat(fmaf((x.z - sub) * rstd, gw.z, gb.z), fmaf((x.w - sub) * rstd, gw.w, gb.w));
      }
      sts128u(sXN + (uint32_t)(c8 >> 3) * WBOX + ((((uint32_t)c8 & 7u) ^ lsw) << 4), t);
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      for (int xb = 0; xb < p.xn_boxes; ++xb)
        tma_store_3d(&tmXn, sXN - (uint32_t)lane * 128u + (uint32_t)xb * WBOX, xb * 64, row0 + q * 32, b);
      bulk_commit();
    }
  }
  if (lane == 0) {
    bulk_wait_read<0>();
    // with two warpgroups the other one may still need ring slots behind this group's last box
    if (p.has_r && have_prev && neg == 2u) mbar_arrive(smem_u32(&bars->r_empty[prev_gc % (uint32_t)p.RB]));
  }
}

// TOp: tensor-core operand type (float = tf32, __half = f16).  LN: LayerNorm prologue over an fp32 source.
// TY: output element type.  Without LN the A boxes are already operands (fp32 pre-rounded to tf32 by the producing
// kernel, or fp16).
template <typename TOp, bool LN, typename TY>
__global__ void __launch_bounds__(NTHREADS, 1)
tma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmXn, const TmaGemmParams p) {
  constexpr bool OPRING = LN && sizeof(TOp) == 2;          // fp32 raw boxes -> separate fp16 operand boxes
  constexpr int OPCOLS = 128 / (int)sizeof(TOp);           // K elements per operand box
  constexpr int GC = 128 / (int)sizeof(TY);                // output columns per store box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (base - smem_u32(smem_raw)));
  const uint32_t sW = base + p.off_w, sA = base + p.off_a, sOP = base + p.off_op, sRO = base + p.off_ro;
  float* lnv = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + p.off_ln);

  // the warp index through a warp reduction lives in a UNIFORM register: the role branches become uniform branches and the
  // code under them uses the uniform datapath (memory descriptors, TMEM / barrier addresses) without one R2UR per use
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));
  const int n0 = (p.split > 0 ? ((int)blockIdx.x >= p.split ? 1 : 0) : (int)blockIdx.y) * p.nc;
  const int ncur = min(p.nc, p.N - n0);

  if (tid == 0) {
    for (int s = 0; s < p.S; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), OPRING ? XF_WARPS * 32 : 1);
      mbar_init(smem_u32(&bars->a_ready[s]), XF_WARPS * 32);
    }
    for (int s = 0; s < MAX_OP; ++s) {
      mbar_init(smem_u32(&bars->op_ready[s]), XF_WARPS * 32);
      mbar_init(smem_u32(&bars->op_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), EPI_WARPS * 32);
    }
    for (int s = 0; s < MAX_RB; ++s) {
      mbar_init(smem_u32(&bars->r_full[s]), 1);
      mbar_init(smem_u32(&bars->r_empty[s]), EPI_WARPS);
    }
    mbar_init(smem_u32(&bars->w_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (!LN && p.xn_mode && warp < EPI_WARPS) {
    for (int i = tid; i < p.N; i += EPI_WARPS * 32) {
      lnv[i] = p.xn_w[i];
      lnv[p.N + i] = p.xn_mode == LN_WITHBIAS ? p.xn_b[i] : 0.f;
    }
  }
  if (LN && warp >= EPI_WARPS && warp < EPI_WARPS + XF_WARPS) {
    for (int i = tid - EPI_WARPS * 32; i < p.K; i += XF_WARPS * 32) {
      lnv[i] = p.ln_w[i];
      lnv[p.K + i] = p.ln_mode == LN_WITHBIAS ? p.ln_b[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_sync();   // set-up (incl. the constant LayerNorm vectors) done under the previous kernel's tail (common.cuh)

  if (warp == WARP_A) {
    // =============================== A-producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      TileIter ti(p);
      const uint8_t* wsrc = p.w + (long long)(p.per_image ? (int)blockIdx.z : 0) * p.w_bstride_bytes;
      if (ti.valid() && !p.wstream) {
        const uint32_t wb = smem_u32(&bars->w_full);
        mbar_expect_tx(wb, (uint32_t)(p.nob * ncur * 128));
        for (int ob = 0; ob < p.nob; ++ob)
          bulk_load(sW + (uint32_t)(ob * ncur * 128), wsrc + ((size_t)ob * p.N + n0) * 128, (uint32_t)(ncur * 128), wb);
      }
      constexpr int RAWCOLS = LN ? 32 : OPCOLS;
      uint32_t it = 0;
      for (; ti.valid(); ti.next()) {
        const int b = ti.img(), row0 = ti.row0();
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          const uint32_t s = it % (uint32_t)p.S, ph = (it / (uint32_t)p.S) & 1u;
          mbar_wait(smem_u32(&bars->a_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->a_full[s]);
          if (p.wstream) {
            mbar_expect_tx(fb, BOX + (uint32_t)(ncur * 128));
            bulk_load(sA + s * p.a_stride + BOX, wsrc + ((size_t)kb * p.N + n0) * 128, (uint32_t)(ncur * 128), fb);
          } else {
            mbar_expect_tx(fb, BOX);
          }
          tma_load_3d(&tmA, fb, sA + s * p.a_stride, kb * RAWCOLS, row0, b);
        }
      }
    }
  } else if (warp == WARP_R) {
    // =============================== residual producer ===============================
    if (lane == 0 && p.has_r) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
      const int ngroups = (ncur + 31) / 32;
      uint32_t gc = 0;
      for (TileIter ti(p); ti.valid(); ti.next()) {
        const int b = ti.img(), row0 = ti.row0();
        for (int g = 0; g < ngroups; ++g, ++gc) {
          const uint32_t s = gc % (uint32_t)p.RB, ph = (gc / (uint32_t)p.RB) & 1u;
          mbar_wait(smem_u32(&bars->r_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->r_full[s]);
          mbar_expect_tx(fb, BOX);
          tma_load_3d(&tmR, fb, sRO + s * BOX, n0 + g * 32, row0, b);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    TileIter ti(p);
    if (ti.valid() && !p.wstream) {
      mbar_wait(smem_u32(&bars->w_full), 0);
      tc_fence_after();
    }
    const uint32_t idesc = make_idesc<TOp>(ncur);
    const uint32_t ring = OPRING ? (uint32_t)p.SOP : (uint32_t)p.S;
    uint32_t it = 0, j = 0;
    for (; ti.valid(); ti.next(), ++j) {
      const uint32_t slot = j & 1u;
      mbar_wait(smem_u32(&bars->acc_empty[slot]), ((j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int ob = 0; ob < p.nob; ++ob, ++it) {
        const uint32_t s = it % ring, ph = (it / ring) & 1u;
        const uint32_t rb = OPRING ? smem_u32(&bars->op_ready[s]) : LN ? smem_u32(&bars->a_ready[s]) : smem_u32(&bars->a_full[s]);
        mbar_wait(rb, ph);
        tc_fence_after();
        {
          // the whole converged warp runs this; one elected lane issues (no per-instruction election loop)
          const uint32_t a_addr = OPRING ? sOP + s * BOX : sA + s * p.a_stride;
          const uint32_t w_addr = p.wstream ? a_addr + BOX : sW + (uint32_t)(ob * ncur * 128);
          const int kbytes = min(128, (p.K - ob * OPCOLS) * (int)sizeof(TOp));
          for (int kk = 0; kk < kbytes / 32; ++kk)
            umma_elect<TOp>(tmem_base + slot * (uint32_t)p.acc_stride, sw128_desc(a_addr + kk * 32),
                            sw128_desc(w_addr + kk * 32), idesc, (ob > 0 || kk > 0) ? 1u : 0u);
          umma_commit_elect(OPRING ? smem_u32(&bars->op_empty[s]) : smem_u32(&bars->a_empty[s]));
          if (ob == p.nob - 1) umma_commit_elect(smem_u32(&bars->acc_full[slot]));
        }
        __syncwarp();
      }
    }
  } else if (warp >= EPI_WARPS) {
    // =============================== LayerNorm transform ===============================
    if constexpr (!LN && sizeof(TY) == 4) {
      // no prologue to run: with two epilogue warpgroups these warps take the odd tiles of the xn path
      if (p.xn_mode && p.egroups == 2) {
        if (p.N == 96) xn_epilogue<96>(p, bars, tmem_base, base, sRO, lnv, tmY, tmXn, n0, warp - EPI_WARPS, lane, 1);
        else xn_epilogue<48>(p, bars, tmem_base, base, sRO, lnv, tmY, tmXn, n0, warp - EPI_WARPS, lane, 1);
      }
    }
    if constexpr (LN) {
      switch (p.nkb) {
        case 1: ln_transform<TOp, 1>(p, bars, sA, sOP, lnv, tid - EPI_WARPS * 32); break;
        case 2: ln_transform<TOp, 2>(p, bars, sA, sOP, lnv, tid - EPI_WARPS * 32); break;
        case 3: ln_transform<TOp, 3>(p, bars, sA, sOP, lnv, tid - EPI_WARPS * 32); break;
        default: ln_transform<TOp, 4>(p, bars, sA, sOP, lnv, tid - EPI_WARPS * 32); break;
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp;                                       // TMEM lane quarter == tile rows [32q, 32q+32)
    const uint32_t lsw = (uint32_t)(lane & 7);
    const int ngroups = (ncur + GC - 1) / GC;
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    uint32_t j = 0, gc = 0;
    bool xn_done = false;
    if constexpr (!LN && sizeof(TY) == 4) {
      if (p.xn_mode) {
        // ---- residual epilogue that also emits xn = LayerNorm(y) (norm2, restormer.py:148) as the fp16 operand of the
        //      fused GDFN: the thread owns a whole pixel row (<= 3 groups of 32 columns), so the row stays in registers,
        //      the statistics are two passes over them, and the LayerNorm pass over HBM disappears ----
        xn_done = true;
        if (p.N == 96) xn_epilogue<96>(p, bars, tmem_base, base, sRO, lnv, tmY, tmXn, n0, q, lane, 0);
        else xn_epilogue<48>(p, bars, tmem_base, base, sRO, lnv, tmY, tmXn, n0, q, lane, 0);
      }
    }
    if (!xn_done)
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int b = ti.img(), row0 = ti.row0();
      const uint32_t slot = j & 1u;
      mbar_wait(smem_u32(&bars->acc_full[slot]), (j >> 1) & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.acc_stride;
      for (int g = 0; g < ngroups; ++g, ++gc) {
        const int col = n0 + g * GC;
        uint32_t box;
        if (p.has_r) {
          const uint32_t s = gc % (uint32_t)p.RB, ph = (gc / (uint32_t)p.RB) & 1u;
          mbar_wait(smem_u32(&bars->r_full[s]), ph);
          box = sRO + s * BOX + (uint32_t)q * WBOX;
        } else {
          box = sRO + (uint32_t)(q * 2 + (int)(gc & 1u)) * WBOX;
          if (lane == 0) bulk_wait_read<1>();                 // the store that last read this buffer has drained it
          __syncwarp();
        }
        const uint32_t myrow = box + (uint32_t)lane * 128u;
        if constexpr (sizeof(TY) == 4) {
          float v[32];
          tmem_ld32(tacc + (uint32_t)(g * 32), v);
          tmem_ld_wait();
          if (g == ngroups - 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += (col + e < p.N) ? __ldg(p.bias + col + e) : 0.f;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = myrow + (((uint32_t)c ^ lsw) << 4);
            float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            if (p.has_r) {
              const float4 rr = lds128(a);
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            sts128(a, o);
          }
        } else {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float v[32];
            tmem_ld32(tacc + (uint32_t)(g * 64 + hf * 32), v);
            tmem_ld_wait();
            if (g == ngroups - 1 && hf == 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
            if (p.bias) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] += (col + hf * 32 + e < p.N) ? __ldg(p.bias + col + hf * 32 + e) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 t;
              __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = f2h2_sat(v[8 * c + 2 * e], v[8 * c + 2 * e + 1]);
              sts128u(myrow + (((uint32_t)(hf * 4 + c) ^ lsw) << 4), t);
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmY, box, col, row0 + q * 32, b);
          bulk_commit();
          if (p.has_r && gc > 0) {                           // the previous group's store has finished reading its box
            bulk_wait_read<1>();
            mbar_arrive(smem_u32(&bars->r_empty[(gc - 1) % (uint32_t)p.RB]));
          }
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
// [B][rows][inner] view with a 128-byte-wide, SWIZZLE_128B box of box_rows rows
int make_map(CUtensorMap* tm, const void* ptr, bool half, int inner, long long row_stride_elems, int rows, int B,
             int box_rows) {
  const int es = half ? 2 : 4;
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)row_stride_elems * es, (cuuint64_t)row_stride_elems * es * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1u};
  return make_tmap(tm, ptr, half, 3, gdim, gstr, box, true);
}

struct TmaCfg { int nc, nchunks, nkb, nob, S, SOP, RB, wstream, egroups; uint32_t a_stride, off_w, off_a, off_op, off_ro, off_ln, off_xn; size_t smem; };

// Shape-only feasibility + shared-memory carve-up.  ln: fused LayerNorm prologue.
bool configure(int K, int N, bool op_half, bool ln, bool has_r, bool y_half, TmaCfg& c, bool xn = false) {
  const int op_es = op_half ? 2 : 4;
  if (K <= 0 || N <= 0 || N % 16 != 0 || (K * op_es) % 32 != 0) return false;
  if (ln && (K % 4 != 0 || K > 128)) return false;          // the transform keeps a row of <= 4 boxes in registers
  if (has_r && y_half) return false;
  const int opcols = 128 / op_es, rawcols = ln ? 32 : opcols;
  c.nkb = (K + rawcols - 1) / rawcols;
  c.nob = (K + opcols - 1) / opcols;
  const bool opring = ln && op_half;
  const int gcs = y_half ? 64 : 32;
  // xn second output: per-warp staging boxes + the LayerNorm parameters; the row must be whole (one N-chunk, <= 3 groups)
  if (xn && (ln || y_half || (N != 48 && N != 96))) return false;
  static const bool one_group = getenv("IRB_ONE_EPI_GROUP") != nullptr;       // A/B switch for benchmarks
  static const bool two96 = getenv("IRB_TWO_EPI_GROUPS_96") != nullptr;     // A/B switch: two epilogue warpgroups at N = 96 too
  c.egroups = xn && has_r && (N == 48 || (N == 96 && two96)) && !one_group ? 2 : 1;
  const size_t xn_bytes = xn ? (size_t)c.egroups * EPI_WARPS * ((N + 63) / 64) * WBOX + (size_t)2 * N * 4 + 16 : 0;
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - xn_bytes;
  for (int chunks = 1; chunks <= N / 16; ++chunks) {
    int nc = (N + chunks - 1) / chunks;
    nc = chunks == 1 ? N : (nc + gcs - 1) / gcs * gcs;
    if (nc > 256) continue;
    size_t off = HDR;
    c.off_w = (uint32_t)off; off += (size_t)c.nob * nc * 128; off = (off + 1023) / 1024 * 1024;
    const size_t w_end = off;
    const size_t ro = has_r ? 0 : (size_t)EPI_WARPS * 2 * WBOX;
    const size_t lnb = ln ? (size_t)2 * K * 4 + 16 : 0;
    if (w_end + ro + lnb > budget) continue;
    size_t rest = budget - w_end - ro - lnb;
    int S, SOP = 0, RB = 0;
    if (opring) {
      // raw fp32 ring (two tiles deep if it fits) + fp16 operand ring
      SOP = std::min(MAX_OP, 2 * c.nob);
      if (rest < (size_t)(c.nkb + c.nob) * BOX) continue;
      S = (int)std::min<size_t>(MAX_S, (rest - (size_t)SOP * BOX) / BOX);
      if (S < 2 * c.nkb) { SOP = c.nob; S = (int)std::min<size_t>(MAX_S, (rest - (size_t)SOP * BOX) / BOX); }
      if (S < c.nkb) continue;
    } else if (has_r) {
      const int boxes = (int)(rest / BOX);
      if (boxes < 4) continue;
      // the two rings split the boxes in proportion to the bytes a tile pulls through them (fp32 residual columns against
      // operand bytes): with a third of the boxes the residual ring held ONE tile at C = 96 -- load, epilogue and store of
      // a tile's residual ran back to back with nothing in flight behind them
      static const bool old_split = getenv("IRB_OLD_RING_SPLIT") != nullptr;    // A/B switch for benchmarks
      const double rb = 4.0 * nc, ab = (double)K * op_es;
      RB = old_split ? boxes / 3 : (int)((double)boxes * rb / (rb + ab));
      RB = std::min(MAX_RB, std::max(2, std::min(RB, boxes - 2)));
      S = std::min(MAX_S, boxes - RB);
      if (S < 2) continue;
    } else {
      S = (int)std::min<size_t>(MAX_S, rest / BOX);
      if (S < (ln ? c.nkb : 2)) continue;
      if (ln && S < 2 * c.nkb && chunks < N / 64) continue;      // prefer two tiles in flight: split N further
    }
    c.nc = nc; c.nchunks = (N + nc - 1) / nc; c.S = S; c.SOP = SOP; c.RB = RB; c.wstream = 0; c.a_stride = BOX;
    c.off_a = (uint32_t)off; off += (size_t)S * BOX;
    c.off_op = (uint32_t)off; off += (size_t)SOP * BOX;
    c.off_ro = (uint32_t)off; off += has_r ? (size_t)RB * BOX : ro;
    c.off_ln = (uint32_t)off; off += lnb;
    if (xn) {
      if (chunks != 1) return false;
      c.off_xn = (uint32_t)off; off += (size_t)c.egroups * EPI_WARPS * ((N + 63) / 64) * WBOX;
      c.off_ln = (uint32_t)off; off += (size_t)2 * N * 4 + 16;
    }
    c.smem = off + 1024;
    return true;
  }
  return false;
}

// Streamed-weights carve-up for the wide low-resolution layers (K * N too large for a CTA-resident weight chunk):
// stages of [A box | weight box of <= 128 columns], at least four deep; the A tensor is re-read once per N-chunk (from L2
// at these levels).
bool configure_stream(int K, int N, bool op_half, bool has_r, bool y_half, TmaCfg& c) {
  const int op_es = op_half ? 2 : 4;
  if (K <= 0 || N <= 0 || N % 16 != 0 || (K * op_es) % 32 != 0 || (has_r && y_half)) return false;
  const int opcols = 128 / op_es, gcs = y_half ? 64 : 32;
  c.nkb = c.nob = (K + opcols - 1) / opcols;
  const size_t budget = 227 * 1024 - 1024;
  for (int nc = 256; nc >= 64; nc -= gcs) {
    if (nc < N && N % nc != 0) continue;                        // equal chunks only
    const int ncc = std::min(nc, N);
    const size_t wbox = ((size_t)ncc * 128 + 1023) / 1024 * 1024;
    const size_t stage = BOX + wbox;
    const size_t ro = has_r ? (size_t)3 * BOX : (size_t)EPI_WARPS * 2 * WBOX;
    if (HDR + 4 * stage + ro > budget) continue;
    const int S = (int)std::min<size_t>(MAX_S, (budget - HDR - ro) / stage);
    size_t off = HDR;
    c.off_w = (uint32_t)off;
    c.nc = ncc; c.nchunks = (N + ncc - 1) / ncc; c.S = S; c.SOP = 0; c.RB = has_r ? 3 : 0; c.wstream = 1;
    c.a_stride = (uint32_t)stage;
    c.off_a = (uint32_t)off; off += (size_t)S * stage;
    c.off_op = (uint32_t)off;
    c.off_ro = (uint32_t)off; off += ro;
    c.off_ln = (uint32_t)off;
    c.smem = off + 1024;
    return true;
  }
  return false;
}

// resident weights where the chunk count stays small, streamed weights otherwise
bool configure_any(int K, int N, bool op_half, bool ln, bool has_r, bool y_half, TmaCfg& c) {
  const int lim = K <= 128 ? 2 : has_r ? 6 : 24;
  if (configure(K, N, op_half, ln, has_r, y_half, c)) {
    if (c.nchunks <= lim) {
      if (ln || K <= 128 || c.nchunks <= 2) return true;
      TmaCfg st;                                                 // many small chunks: prefer wide streamed chunks
      if (c.nchunks > 8 && configure_stream(K, N, op_half, has_r, y_half, st) && st.nchunks * 2 <= c.nchunks) { c = st; return true; }
      return true;
    }
  }
  if (ln || K <= 128) return false;
  return configure_stream(K, N, op_half, has_r, y_half, c);
}

template <typename TOp, bool LN, typename TY>
int launch_inst(const CUtensorMap& tA, const CUtensorMap& tR, const CUtensorMap& tY, const CUtensorMap& tXn,
                const TmaGemmParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(tma_gemm_kernel<TOp, LN, TY>, optin));
  IRB_CUDA(launch_pdl(tma_gemm_kernel<TOp, LN, TY>, grid, dim3(NTHREADS), smem, s, tA, tR, tY, tXn, p));
  return IR_OK;
}

}  // namespace

// The plan uses this kernel when the A tile is re-read by few N-chunks: at most two at the high-resolution levels
// (K <= 128), more at the low-resolution levels, whose whole A tensor (1/16 or 1/64 of the pixels) stays in the 126 MB L2.
bool tma_gemm_shape_supported(int K, int N, bool op_half, bool ln, bool has_r, bool y_half) {
  TmaCfg c;
  const char* e = getenv("IRB_TMA_CHUNK_LIMIT");      // bring-up knob: 0 keeps the low-resolution levels on the first kernel
  if (e && K > 128) return configure(K, N, op_half, ln, has_r, y_half, c) && c.nchunks <= atoi(e);
  return configure_any(K, N, op_half, ln, has_r, y_half, c);
}

bool tma_gemm_xn_supported(int C, bool op_half) {
  TmaCfg c;
  return configure(C, C, op_half, false, true, false, c, true);
}

int tma_gemm_kpad(int K, bool op_half) { const int oc = op_half ? 64 : 32; return (K + oc - 1) / oc * oc; }

// Returns IR_OK, an error, or IR_UNSUPPORTED_SHAPE (> 0) when the caller should use the first-generation kernel.
int launch_gemm_tma(const TcGemmParams& t, cudaStream_t s) {
  const bool ln = t.ln_mode != LN_NONE;
  const bool op_half = t.op_half != 0, a_half = t.a_half != 0, y_half = t.y_half != 0, has_r = t.r != nullptr;
  if (t.k2 != 0 || t.a_mode != 0 || t.o_mode != O_NHWC || t.relu || (t.acc_sign != 0.f && t.acc_sign != 1.f))
    return IR_UNSUPPORTED_SHAPE;
  if (ln ? a_half : (a_half != op_half)) return IR_UNSUPPORTED_SHAPE;
  TmaCfg c;
  const bool xn = t.xn != nullptr;
  if (xn) {
    if (!configure(t.K, t.N, op_half, ln, has_r, y_half, c, true)) return IR_UNSUPPORTED_SHAPE;
    c.wstream = 0; c.a_stride = BOX;
  } else if (!configure_any(t.K, t.N, op_half, ln, has_r, y_half, c)) return IR_UNSUPPORTED_SHAPE;
  const int a_es = a_half ? 2 : 4, y_es = y_half ? 2 : 4;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!al16(t.a1) || !al16(t.y) || !al16(t.w) || (has_r && !al16(t.r)) || (t.lda1 * a_es) % 16 != 0 ||
      (t.ldy * y_es) % 16 != 0 || (has_r && (t.ldr * 4) % 16 != 0) || (t.w_bstride * (op_half ? 2 : 4)) % 16 != 0)
    return IR_UNSUPPORTED_SHAPE;

  CUtensorMap tA, tR, tY;
  IRB_TRY(make_map(&tA, t.a1, a_half, t.K, t.lda1, t.HW, t.B, TM));
  IRB_TRY(make_map(&tY, t.y, y_half, t.N, t.ldy, t.HW, t.B, 32));
  if (has_r) IRB_TRY(make_map(&tR, t.r, false, t.N, t.ldr, t.HW, t.B, TM));
  else tR = tY;
  CUtensorMap tXn = tY;
  if (xn) {
    if (!al16(t.xn) || (t.ldxn * 2) % 16 != 0) return IR_UNSUPPORTED_SHAPE;
    IRB_TRY(make_map(&tXn, t.xn, true, t.N, t.ldxn, t.HW, t.B, 32));
  }

  TmaGemmParams p{};
  p.w = reinterpret_cast<const uint8_t*>(t.w);
  p.w_bstride_bytes = t.w_bstride * (op_half ? 2 : 4);
  p.B = t.B; p.HW = t.HW; p.K = t.K; p.N = t.N;
  p.nc = c.nc; p.nkb = c.nkb; p.nob = c.nob; p.S = c.S; p.SOP = c.SOP; p.RB = c.RB;
  p.ln_mode = t.ln_mode; p.ln_w = t.ln_w; p.ln_b = t.ln_b; p.bias = t.bias; p.has_r = has_r ? 1 : 0;
  p.tiles_per_img = cdiv(t.HW, TM); p.ntiles = p.tiles_per_img * t.B; p.per_image = t.w_bstride != 0 ? 1 : 0;
  p.acc_stride = y_half ? (c.nc + 63) / 64 * 64 : (c.nc + 31) / 32 * 32;   // whole store groups: the tail group reads (and TMA clips) past nc
  int cols = 32; while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.off_w = c.off_w; p.off_a = c.off_a; p.off_op = c.off_op; p.off_ro = c.off_ro; p.off_ln = c.off_ln;
  p.wstream = c.wstream; p.a_stride = c.a_stride;
  p.xn_mode = xn ? t.xn_ln_mode : 0; p.xn_w = t.xn_w; p.xn_b = t.xn_b; p.xn_boxes = (t.N + 63) / 64; p.off_xn = c.off_xn; p.egroups = xn ? c.egroups : 1;

  dim3 grid;
  if (p.per_image) grid = dim3(std::max(1, std::min(p.tiles_per_img, 148 / (c.nchunks * t.B))), c.nchunks, t.B);
  else grid = dim3(std::max(1, std::min(p.ntiles, 148 / c.nchunks)), c.nchunks, 1);
  if (!p.per_image && c.nchunks == 2 && p.ntiles >= 148) {
    // two uneven chunks (e.g. 288 = 160 + 128 columns): split the 148 CTAs in proportion to the bytes each chunk moves
    const double per_col = y_half ? 2.0 : 4.0 + (has_r ? 4.0 : 0.0);
    const double c0 = (double)t.K * a_es + c.nc * per_col, c1 = (double)t.K * a_es + (t.N - c.nc) * per_col;
    p.split = std::min(147, std::max(1, (int)(148.0 * c0 / (c0 + c1) + 0.5)));
    grid = dim3(148, 1, 1);
  }
  // one CTA per SM: the kernel owns up to 512 TMEM columns
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);

  const double rows = (double)t.B * t.HW;
  ProfScope prof(t.tag, rows * ((double)t.K * a_es + (double)t.N * (y_es + (has_r ? 4.0 : 0.0) + (xn ? 2.0 : 0.0))), 2.0 * rows * t.N * t.K, s);
  if (!op_half && !ln && !y_half) return launch_inst<float, false, float>(tA, tR, tY, tXn, p, grid, smem, s);
  if (!op_half && ln && !y_half) return launch_inst<float, true, float>(tA, tR, tY, tXn, p, grid, smem, s);
  if (op_half && ln && y_half) return launch_inst<__half, true, __half>(tA, tR, tY, tXn, p, grid, smem, s);
  if (op_half && ln && !y_half) return launch_inst<__half, true, float>(tA, tR, tY, tXn, p, grid, smem, s);
  if (op_half && !ln && !y_half) return launch_inst<__half, false, float>(tA, tR, tY, tXn, p, grid, smem, s);
  if (op_half && !ln && y_half) return launch_inst<__half, false, __half>(tA, tR, tY, tXn, p, grid, smem, s);
  return IR_UNSUPPORTED_SHAPE;
}

}  // namespace irb
