// The whole front of the transposed attention (MDTA, restormer.py:111-124 and norm1 :147) in ONE kernel:
//
//     xn      = LayerNorm(x)                                  (norm1)
//     qkv     = dw3x3(W_qkv . xn)                             (qkv 1x1 + qkv_dwconv :114)
//     S[b]   += q . k^T over the pixels, |q_i|^2, |k_j|^2     (Gram + the norms F.normalize needs :121-124)
//     v      -> HBM (fp16)                                    (the only tensor written)
//
// The 3C-wide qkv tensor -- written once and read once by the two-kernel version (24C of the attention path's 46C bytes
// per pixel) -- never exists in HBM: the 1x1 contraction is RECOMPUTED per 8 x 16 pixel tile over the (8+2) x (16+2) halo
// the depthwise conv needs (1.4 x the contraction work on a tensor pipe that is mostly idle) and lives in shared memory
// only, exactly as ffn_fused.cu does for the GDFN.  Per pixel the kernel reads 4C bytes of x (1.4 x with the halo, served
// by L2) and writes 2C bytes of v.
//
// One persistent CTA per SM, bound to one image (the Gram accumulates per image), 20 warps.  The pipeline unit of the
// CUDA-core roles is 32 qkv channels; the tensor core works in GROUPS of three units (N = 96): with one MMA1 per unit the
// single issuing warp -- ~33 instructions of descriptor / uniform-register traffic per tcgen05.mma, 116 of them per tile --
// paced the whole kernel (ncu: every other role waiting on it, whatever else was removed; profiles/r02_attn_fused_*).
//
//   LayerNorm (4 warps)   per tile: the fp32 x halo rows straight from global memory (C/12 lanes per pixel, 16-byte
//                         loads), two-pass statistics, fp16 xn rows into the [192 px][128 B] SWIZZLE_128B operand boxes
//                         (double-buffered: tile j+1 is normalised while tile j computes).  Pixels outside the image
//                         are written as ZERO rows: the qkv conv has no bias, so qkv = 0 there, which IS the depthwise
//                         conv's zero padding (also with a WithBias LayerNorm, whose output on a zero row is not zero)
//                         A bulk-tensor L2 PREFETCH of the x halo box runs two tiles ahead, so that these register loads
//                         see L2 latency instead of HBM latency (a lane holds only 9 x 16 bytes in flight)
//   producer (1 thread)   once: all of W_qkv (fp16 operand image, 20 / 74 KB) into shared memory, where it stays
//   MMA (1 warp)          MMA1: D1[192 px][96] = xn_patch . W_group^T as two M = 128 instructions per K step (patch rows
//                         0-127 and 64-191) into a double-buffered TMEM accumulator; everything about the issue loop is
//                         a compile-time constant (descriptors are a base plus a constant);
//                         Gram: S[i][j] += sum_p q_i[p] k_j[p] with A = the q rows and B = the k rows of the X tile
//   convert (4 + 2 warps) tcgen05.ld of D1 -> fp16 -> qkv patch [180 px][80 B] in shared memory
//   dw warps (8)          thread = 2 channels x a 2 x 4 pixel block, packed FFMA2 taps.  q / k units are written
//                         TRANSPOSED ([channel][pixel], pixels = the Gram's K dimension) into the X tile and their
//                         squares accumulate in registers; v units go to a per-warp staging box -> bulk-tensor store
//   epilogue (warps 0-3)  once, at the end: TMEM -> the CTA's partial S; the dw warps reduce the norm partials
//
// fold_kernel then reduces the partials in a fixed order (deterministic, no atomics), as for attn_front.cu.
#include "attn_fused.cuh"
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace irb {

namespace {

using namespace sm100;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int PW = TW + 2;
constexpr int HPIX = (TH + 2) * PW;            // 180 halo pixels
constexpr int AROWS = 192;                     // rows the two M = 128 instructions read
constexpr int ABOX = AROWS * 128;              // one 64-channel box of the xn patch
constexpr int UC = 32;                         // qkv channels per unit
constexpr int HROW = 80;                       // qkv patch row pitch: 64 B of channels + 16 B pad (conflict-free 16-byte stores)
constexpr int HSTAGE = HPIX * HROW;

constexpr int CP = 16, BH = 2, BW = 4;         // dw thread: channel pair x 2x4 pixel block; 16 x 16 = 256 threads
constexpr int CVA_WARPS = 4, DW_WARPS = 8, CVB_WARPS = 2;
constexpr int WARP_DW = CVA_WARPS, WARP_MMA = WARP_DW + DW_WARPS, WARP_PROD = WARP_MMA + 1, WARP_CVB = WARP_PROD + 1,
              WARP_LN = WARP_CVB + CVB_WARPS;
template <int LNW> constexpr int nthreads() { return (WARP_LN + LNW) * 32; }
static_assert((WARP_CVB & 3) == 2, "the two extra convert warps must own TMEM lane quarters 2 and 3");
constexpr int GU = 3;                          // units per MMA1 group
constexpr int GN = GU * UC;                    // 96 accumulator columns per group and row half
constexpr int D1_COLS = 2 * GN;                // per buffer: rows 0-127 in columns [0, 96), rows 64-191 in [96, 192)
constexpr int ND = 2;                          // D1 group accumulators in TMEM
constexpr int S_COL0 = ND * D1_COLS;           // the Gram accumulator (<= 96 columns)
constexpr int TMEM_COLS = 512;
constexpr int NHMAX = 4;                       // qkv patch stages (at most)
constexpr int VBOXB = TM * UC * 2;             // v staging bytes per buffer (all 8 warps)

struct Bars {
  unsigned long long a_full[2], a_empty[2];
  unsigned long long w_full;
  unsigned long long d1_full[ND], d1_empty[ND];
  unsigned long long h_full[NHMAX], h_empty[NHMAX];
  unsigned long long x_ready, x_empty, acc_done;
  uint32_t tmem_base;
};

struct FusedFrontParams {
  const float* x;
  const float* ln_w;
  const float* ln_b;
  const uint8_t* w_qkv;    // fp16 SWIZZLE_128B image [nkb][NP][128 B]
  const float* dw;         // taps [nunits][9][32]
  float* s_part;           // [B][heads][parts][ch][ch]
  float* n_part;           // [B][heads][parts][2][ch]
  int ln_mode;
  int B, H, W, C, heads, parts;
  int nv, NP;              // units of v; NP = padded rows of W_qkv
  int tiles_x, tiles_y, tiles_per_img;
  int xrows;
  uint32_t off_a, off_w, off_h, off_x, off_v, off_bars;
};

// compile-time geometry of one channel width
template <int CW> struct Geo {
  static constexpr int NKB = (CW + 63) / 64;                 // 64-channel K boxes of xn / W_qkv
  static constexpr int KS_LAST = (CW - 64 * (NKB - 1)) / 16; // K steps (16 channels) in the last box
  static constexpr int NUNITS = (3 * CW + UC - 1) / UC;      // 5 / 9
  static constexpr int NQK = 2 * CW / UC;                    // 3 / 6 units of q|k
  static constexpr int NG = (NUNITS + GU - 1) / GU;          // 2 / 3 MMA1 groups
  static constexpr int NP = NUNITS * UC;                     // rows of the W_qkv image
  static constexpr int NA = CW <= 48 ? 2 : 1;                // xn patch buffers (C = 96: the patch retires a third into its
                                                             // tile, once the last group's MMA1 has read it)
  static constexpr int NH = CW <= 48 ? 4 : 2;                // qkv patch stages
  static constexpr int LNW = CW <= 48 ? 4 : 8;               // LayerNorm warps (C = 96: single xn buffer, LayerNorm of the next
                                                             // tile is on the critical path -> twice the warps; 80 registers)
  static constexpr int NTHREADS = nthreads<LNW>();
  // where the MMA warp issues the Gram of a tile: behind the tile's own groups (C = 48: measured 0.47 vs 0.55 ms), or in
  // front of the NEXT tile's last group (C = 96, single xn buffer: 1.00 vs 1.14 ms)
  static constexpr bool GRAM_LATE = CW > 48;
  static constexpr uint32_t A_BYTES = NKB * ABOX;
  static constexpr uint32_t W_BYTES = NKB * NP * 128;
  __host__ __device__ static constexpr int group_units(int grp) { return grp == NG - 1 ? NUNITS - GU * (NG - 1) : GU; }
};

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t ld_h2(uint32_t a) {
  uint32_t t;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(a));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&t));
  return pack2(f.x, f.y);
}
__device__ __forceinline__ uint32_t h2_bits(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }

struct TileIter {
  int t, step, end, tx_n;
  __device__ TileIter(const FusedFrontParams& p) : t(blockIdx.x), step(gridDim.x), end(p.tiles_per_img), tx_n(p.tiles_x) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int y0() const { return (t / tx_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * TW; }
};

// D1 -> fp16 qkv patch for the 32 patch rows this warp owns (TMEM lane quarter `q`, rows row0 .. row0 + 31), one MMA1 group
// (up to three units) at a time; `half_off` selects the row half's accumulator columns
template <int CW>
__device__ __forceinline__ void convert_tiles(Bars* bars, const FusedFrontParams& p, uint32_t tmem_base, int q, int half_off,
                                              int row0, uint32_t sH, int lane) {
  using G = Geo<CW>;
  const int row = row0 + lane;
  uint32_t g = 0, gg = 0;
  for (TileIter ti(p); ti.valid(); ti.next()) {
#pragma unroll
    for (int grp = 0; grp < G::NG; ++grp, ++gg) {
      const uint32_t sd = gg & 1u;
      mbar_wait(smem_u32(&bars->d1_full[sd]), (gg >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int ug = 0; ug < G::group_units(grp); ++ug, ++g) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + sd * D1_COLS + (uint32_t)(half_off + ug * UC), v);
        tmem_ld_wait();
        if (ug == G::group_units(grp) - 1) {                  // the group's accumulator is free again
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->d1_empty[sd]));
        }
        const uint32_t s = g % G::NH;
        mbar_wait(smem_u32(&bars->h_empty[s]), ((g / G::NH) & 1u) ^ 1u);
        if (row < HPIX) {
          const uint32_t dst = sH + s * HSTAGE + (uint32_t)row * HROW;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = h2_bits(f2h2_sat(v[8 * c + 0], v[8 * c + 1]));
            u.y = h2_bits(f2h2_sat(v[8 * c + 2], v[8 * c + 3]));
            u.z = h2_bits(f2h2_sat(v[8 * c + 4], v[8 * c + 5]));
            u.w = h2_bits(f2h2_sat(v[8 * c + 6], v[8 * c + 7]));
            sts128u(dst + c * 16, u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_full[s]));
      }
    }
  }
}

// this thread's taps of one unit, straight from L1 / L2 (10 KB in all); issued BEFORE the wait for the unit's patch
__device__ __forceinline__ void load_taps(const float* __restrict__ taps, f2_t (&w)[9]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(taps + t * UC));
    w[t] = pack2(f.x, f.y);
  }
}

// depthwise 3x3 of one unit for this thread's 2 channels x (2 x 4) pixels
template <int DBG>
__device__ __forceinline__ void dw_unit(uint32_t patch, const f2_t (&w)[9], f2_t (&acc)[BH][BW]) {
  if (DBG & 2) {
#pragma unroll
    for (int oy = 0; oy < BH; ++oy)
#pragma unroll
      for (int ox = 0; ox < BW; ++ox) acc[oy][ox] = w[oy * 4 + ox];
    return;
  }
#pragma unroll
  for (int iy = 0; iy < BH + 2; ++iy) {
    f2_t v[BW + 2];
#pragma unroll
    for (int ix = 0; ix < BW + 2; ++ix) v[ix] = ld_h2(patch + (uint32_t)(iy * PW + ix) * HROW);
#pragma unroll
    for (int oy = 0; oy < BH; ++oy) {
      const int ky = iy - oy;
      if (ky < 0 || ky > 2) continue;
#pragma unroll
      for (int ox = 0; ox < BW; ++ox) {
        if (ky == 0) acc[oy][ox] = mul2(w[0], v[ox]);
        else acc[oy][ox] = fma2(w[ky * 3], v[ox], acc[oy][ox]);
        acc[oy][ox] = fma2(w[ky * 3 + 1], v[ox + 1], acc[oy][ox]);
        acc[oy][ox] = fma2(w[ky * 3 + 2], v[ox + 2], acc[oy][ox]);
      }
    }
  }
}

// CW: channel count (48 or 96).  NQK = 2C/32 q|k units per tile.
// DBG != 0: timing experiments only (results are garbage): 1 LayerNorm without its global loads, 2 no depthwise taps,
// 8 no MMA1 instructions (compiled with -DIRB_FUSED_EXPERIMENTS, selected by IRB_AF_DBG)
template <int CW, int DBG>
__global__ void __launch_bounds__(Geo<CW>::NTHREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmX, const FusedFrontParams p) {
  using G = Geo<CW>;
  constexpr int NQK = G::NQK, NH = G::NH, NA = G::NA;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  Bars* bars = reinterpret_cast<Bars*>(gbase + p.off_bars);
  const uint32_t sA = base + p.off_a, sW = base + p.off_w, sH = base + p.off_h, sX = base + p.off_x, sV = base + p.off_v;
  float* red = reinterpret_cast<float*>(gbase + p.off_v);      // aliases the v staging (free once the stores have drained)
  const uint32_t xbox = (uint32_t)p.xrows * 128u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, part = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), G::LNW);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
    }
    for (int s = 0; s < NHMAX; ++s) {
      mbar_init(smem_u32(&bars->h_full[s]), CVA_WARPS + CVB_WARPS);
      mbar_init(smem_u32(&bars->h_empty[s]), DW_WARPS);
    }
    mbar_init(smem_u32(&bars->w_full), 1);
    for (int s = 0; s < ND; ++s) {
      mbar_init(smem_u32(&bars->d1_full[s]), 1);
      mbar_init(smem_u32(&bars->d1_empty[s]), CVA_WARPS + CVB_WARPS);
    }
    mbar_init(smem_u32(&bars->x_ready), DW_WARPS);
    mbar_init(smem_u32(&bars->x_empty), 1);
    mbar_init(smem_u32(&bars->acc_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= WARP_LN) {
    // =============================== LayerNorm: x halo rows -> fp16 operand boxes ===============================
    constexpr int LPR = CW / 12, RPW = 32 / LPR;           // lanes per pixel row (3 float4 each), rows per warp pass
    constexpr int RPP = RPW * G::LNW;                      // rows per pass of all LayerNorm warps (32)
    constexpr int NPASS = AROWS / RPP;                     // 6 / 12: rows 180 .. 191 are written as zeros
    constexpr int GP = 3;                                  // passes in flight (loads issued before the arithmetic)
    static_assert(AROWS % RPP == 0 && NPASS % GP == 0, "pass geometry");
    const int lw = warp - WARP_LN, l = lane % LPR, r = lane / LPR;
    float4 gw[3], gb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gw[i] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + i * LPR + l);
      gb[i] = p.ln_mode == LN_WITHBIAS ? __ldg(reinterpret_cast<const float4*>(p.ln_b) + i * LPR + l)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // L2 prefetch of the x halo boxes, two tiles ahead of the tile being normalised
    TileIter tp(p);
    const bool pf = lw == 0 && lane == 0 && !(DBG & 1);
    if (pf) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      for (int k = 0; k < 2 && tp.valid(); ++k, tp.next()) tma_prefetch_l2_4d(&tmX, 0, tp.x0() - 1, tp.y0() - 1, b);
    }
    uint32_t j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int y0 = ti.y0(), x0 = ti.x0();
      if (pf && tp.valid()) { tma_prefetch_l2_4d(&tmX, 0, tp.x0() - 1, tp.y0() - 1, b); tp.next(); }
      const uint32_t ab = j % NA;
      const uint32_t abase = sA + ab * G::A_BYTES;
      float4 v[GP][3];
      bool inside[GP];
      auto load_group = [&](int pg) {
#pragma unroll
        for (int q = 0; q < GP; ++q) {
          const int row = (pg + q) * RPP + lw * RPW + r;
          const int py = row / PW, px = row - py * PW;
          const int gy = y0 - 1 + py, gx = x0 - 1 + px;
          inside[q] = row < HPIX && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
          const float4* src = reinterpret_cast<const float4*>(p.x + (((long long)b * p.H + gy) * p.W + gx) * CW);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            v[q][i] = (DBG & 1) ? make_float4(0.5f * l, 1.f, -1.f, 0.25f * i)
                                : inside[q] ? __ldg(src + i * LPR + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto compute_group = [&](int pg) {
#pragma unroll
        for (int q = 0; q < GP; ++q) {
          const int row = (pg + q) * RPP + lw * RPW + r;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) s += (v[q][i].x + v[q][i].y) + (v[q][i].z + v[q][i].w);
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mu = s * (1.0f / CW);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float dx = v[q][i].x - mu, dy = v[q][i].y - mu, dz = v[q][i].z - mu, dw = v[q][i].w - mu;
            ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
          }
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const float rstd = inside[q] ? 1.0f / sqrtf(ss * (1.0f / CW) + 1e-5f) : 0.f;
          const float sub = p.ln_mode == LN_WITHBIAS ? mu : 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float4 o;
            o.x = (v[q][i].x - sub) * rstd * gw[i].x; o.y = (v[q][i].y - sub) * rstd * gw[i].y;
            o.z = (v[q][i].z - sub) * rstd * gw[i].z; o.w = (v[q][i].w - sub) * rstd * gw[i].w;
            if (inside[q]) { o.x += gb[i].x; o.y += gb[i].y; o.z += gb[i].z; o.w += gb[i].w; }
            uint2 t;
            t.x = h2_bits(f2h2_sat(o.x, o.y));
            t.y = h2_bits(f2h2_sat(o.z, o.w));
            const int k0 = (i * LPR + l) * 4;                 // first of this lane's four channels
            const uint32_t kb = (uint32_t)k0 >> 6, kk = (uint32_t)k0 & 63u;
            sts64u(abase + kb * ABOX + (uint32_t)row * 128u + (((kk >> 3) ^ ((uint32_t)row & 7u)) << 4) + (kk & 7u) * 2u, t);
          }
        }
      };
      // the first group's loads go out BEFORE the wait for the buffer: their latency hides behind it
      load_group(0);
      mbar_wait(smem_u32(&bars->a_empty[ab]), ((j / NA) & 1u) ^ 1u);
      compute_group(0);
#pragma unroll 1
      for (int pg = GP; pg < NPASS; pg += GP) {
        load_group(pg);
        compute_group(pg);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[ab]));
    }
  } else if (warp == WARP_PROD) {
    // =============================== producer: W_qkv, once ===============================
    if (lane == 0) {
      const uint32_t fb = smem_u32(&bars->w_full);
      mbar_expect_tx(fb, G::W_BYTES);
      constexpr uint32_t CHUNK = 8192;
      for (uint32_t o = 0; o < G::W_BYTES; o += CHUNK)
        bulk_load(sW + o, p.w_qkv + o, (G::W_BYTES - o) < CHUNK ? (G::W_BYTES - o) : CHUNK, fb);
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    // Issue order: the groups of tile j with the Gram of tile j-1 in front of the LAST group.  (Groups of tile j, then
    // Gram j, left a bubble at every tile boundary: the first group of tile j+1 sat behind the wait for tile j's q | k rows.
    // The last group cannot go earlier anyway: its accumulator is freed by converts whose qkv patch stages wait for dw
    // warps that wait for this very Gram.  A polling event loop over both conditions was measured 20-100 % slower.)
    const uint32_t idescS = make_idesc<__half>(CW);
    const uint64_t wdesc = sw128_desc(sW), xdesc = sw128_desc(sX);
    mbar_wait(smem_u32(&bars->w_full), 0);
    auto issue_group = [&](auto grp_tag, uint32_t gg, uint32_t jg) {
      constexpr int grp = decltype(grp_tag)::value;
      const uint32_t idesc1 = make_idesc<__half>(G::group_units(grp) * UC);
      const uint32_t sd = gg & 1u, ab = jg % NA;
      tc_fence_after();
      const uint32_t d = tmem_base + sd * D1_COLS;
      const uint64_t adesc = sw128_desc(sA + ab * G::A_BYTES);
#pragma unroll
      for (int kb = 0; kb < G::NKB; ++kb) {
#pragma unroll
        for (int kk = 0; kk < (kb == G::NKB - 1 ? G::KS_LAST : 4); ++kk) {
          if (DBG & 8) continue;
          const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
          // descriptors: base + (byte offset >> 4); the whole offset is a compile-time constant
          const uint64_t bd = wdesc + (uint64_t)((kb * G::NP * 128 + grp * GU * UC * 128 + kk * 32) >> 4);
          const uint64_t ad = adesc + (uint64_t)((kb * ABOX + kk * 32) >> 4);
          umma_elect<__half>(d, ad, bd, idesc1, acc);
          umma_elect<__half>(d + GN, ad + (uint64_t)((64 * 128) >> 4), bd, idesc1, acc);
        }
      }
      umma_commit_elect(smem_u32(&bars->d1_full[sd]));
      if (grp == G::NG - 1) umma_commit_elect(smem_u32(&bars->a_empty[ab]));
      __syncwarp();
    };
    auto gram = [&](uint32_t jx) {
      mbar_wait(smem_u32(&bars->x_ready), jx & 1u);          // the dw warps have written tile jx's q | k rows
      tc_fence_after();
#pragma unroll
      for (int a = 0; a < 2; ++a)                            // X boxes: 64 pixels (fp16) each
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t qa = xdesc + (uint64_t)((a * (int)xbox + kk * 32) >> 4);   // q rows start at row 0
          const uint64_t ka = qa + (uint64_t)((CW * 128) >> 4);                     // k rows start at row C
          umma_elect<__half>(tmem_base + S_COL0, qa, ka, idescS, (jx > 0 || a > 0 || kk > 0) ? 1u : 0u);
        }
      umma_commit_elect(smem_u32(&bars->x_empty));
      __syncwarp();
    };
    uint32_t gg = 0, j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      auto one = [&](auto grp_tag) {
        constexpr int grp = decltype(grp_tag)::value;
        if (G::GRAM_LATE && grp == G::NG - 1 && j > 0) gram(j - 1);
        if (grp == 0) mbar_wait(smem_u32(&bars->a_full[j % NA]), (j / NA) & 1u);
        mbar_wait(smem_u32(&bars->d1_empty[gg & 1u]), ((gg >> 1) & 1u) ^ 1u);
        issue_group(grp_tag, gg, j);
        ++gg;
      };
      one(std::integral_constant<int, 0>{});
      if (G::NG > 2) one(std::integral_constant<int, 1>{});
      one(std::integral_constant<int, G::NG - 1>{});
      if (!G::GRAM_LATE) gram(j);
    }
    if (G::GRAM_LATE && j > 0) gram(j - 1);
    umma_commit_elect(smem_u32(&bars->acc_done));
  } else if (warp >= WARP_CVB) {
    // =============================== convert: patch rows 128 .. 179 (second MMA, TMEM lanes 64 .. 127) ===============================
    const int q = warp & 3;
    convert_tiles<CW>(bars, p, tmem_base, q, GN, 128 + (q - 2) * 32, sH, lane);
  } else if (warp >= WARP_DW) {
    // =============================== depthwise 3x3 -> X tile (q, k) / v staging ===============================
    const int ctid = tid - WARP_DW * 32;
    const int cp = ctid % CP, blk = ctid / CP;
    const int by = blk / (TW / BW), bx = blk % (TW / BW);
    const uint32_t win0 = (uint32_t)((BH * by) * PW + BW * bx) * HROW + (uint32_t)cp * 4u;
    f2_t nrm[NQK];
#pragma unroll
    for (int i = 0; i < NQK; ++i) nrm[i] = pack2(0.f, 0.f);
    uint32_t g = 0, j = 0, vc = 0;
    const int dwi = warp - WARP_DW;                         // this warp's pixels: rows 2(dwi/2)..+1, columns 8(dwi%2)..+7
    const uint32_t vwarp = sV + (uint32_t)dwi * (2u * 16u * 64u);
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int y0 = ti.y0(), x0 = ti.x0();
      const bool partial = (y0 + TH > p.H) || (x0 + TW > p.W);
      // Output pixels past the right / bottom image edge still see taps from inside the image; they must not reach
      // the Gram or the norms (the v store is clipped by TMA).
      bool okp[BH][BW];
#pragma unroll
      for (int oy = 0; oy < BH; ++oy)
#pragma unroll
        for (int ox = 0; ox < BW; ++ox) okp[oy][ox] = (y0 + BH * by + oy < p.H) && (x0 + BW * bx + ox < p.W);
      // the Gram MMA of the previous tile has read the X tile
      mbar_wait(smem_u32(&bars->x_empty), (j & 1u) ^ 1u);
#pragma unroll
      for (int ch = 0; ch < NQK; ++ch, ++g) {
        const uint32_t s = g % NH;
        f2_t w[9];
        load_taps(p.dw + (size_t)(ch * 9) * UC + cp * 2, w);
        mbar_wait(smem_u32(&bars->h_full[s]), (g / NH) & 1u);
        f2_t acc[BH][BW];
        dw_unit<DBG>(sH + s * HSTAGE + win0, w, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        // transposed store: rows = channels ch*32 + 2cp (+1), columns = the block's pixels (4 consecutive per row)
        const uint32_t r0 = (uint32_t)(ch * UC + 2 * cp);
#pragma unroll
        for (int oy = 0; oy < BH; ++oy) {
          float a0, a1, b0, b1, c0, c1, d0, d1;
          unpack2(acc[oy][0], a0, a1); unpack2(acc[oy][1], b0, b1);
          unpack2(acc[oy][2], c0, c1); unpack2(acc[oy][3], d0, d1);
          if (partial) {
            if (!okp[oy][0]) { a0 = 0.f; a1 = 0.f; }
            if (!okp[oy][1]) { b0 = 0.f; b1 = 0.f; }
            if (!okp[oy][2]) { c0 = 0.f; c1 = 0.f; }
            if (!okp[oy][3]) { d0 = 0.f; d1 = 0.f; }
          }
          // pixel p = (2by+oy)*16 + 4bx: X box p/64 = by/2, byte offset in the 128-byte row = (p%64)*2
          const __half2 h01 = f2h2_sat(a0, b0), h23 = f2h2_sat(c0, d0);
          const __half2 g01 = f2h2_sat(a1, b1), g23 = f2h2_sat(c1, d1);
          // the norms use the rounded operands, like the MMA
          a0 = __low2float(h01); b0 = __high2float(h01); c0 = __low2float(h23); d0 = __high2float(h23);
          a1 = __low2float(g01); b1 = __high2float(g01); c1 = __low2float(g23); d1 = __high2float(g23);
          const uint32_t boxa = sX + (uint32_t)(by >> 1) * xbox;
          const uint32_t off = (uint32_t)(((by & 1) * 32 + oy * 16 + 4 * bx) * 2);
          const uint32_t c16 = off >> 4, sub = off & 15u;
          sts64u(boxa + r0 * 128u + ((c16 ^ (r0 & 7u)) << 4) + sub, make_uint2(h2_bits(h01), h2_bits(h23)));
          sts64u(boxa + (r0 + 1) * 128u + ((c16 ^ ((r0 + 1) & 7u)) << 4) + sub, make_uint2(h2_bits(g01), h2_bits(g23)));
          f2_t n = nrm[ch];
          n = fma2(pack2(a0, a1), pack2(a0, a1), n); n = fma2(pack2(b0, b1), pack2(b0, b1), n);
          n = fma2(pack2(c0, c1), pack2(c0, c1), n); n = fma2(pack2(d0, d1), pack2(d0, d1), n);
          nrm[ch] = n;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->x_ready));
      // ---- v units: each warp stages its own 2 x 8 pixel region ([pixel][channel], 64-byte rows) and stores it with its
      //      own bulk-tensor copy: no block-wide barrier on the path ----
      for (int ch = 0; ch < p.nv; ++ch, ++vc, ++g) {
        const uint32_t s = g % NH;
        f2_t w[9];
        load_taps(p.dw + (size_t)((NQK + ch) * 9) * UC + cp * 2, w);
        mbar_wait(smem_u32(&bars->h_full[s]), (g / NH) & 1u);
        f2_t acc[BH][BW];
        dw_unit<DBG>(sH + s * HSTAGE + win0, w, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        const uint32_t vb = vwarp + (vc & 1u) * (16u * 64u);
        if (lane == 0) bulk_wait_read<1>();                  // the store that last read this buffer has drained it
        __syncwarp();
#pragma unroll
        for (int oy = 0; oy < BH; ++oy)
#pragma unroll
          for (int ox = 0; ox < BW; ++ox) {
            float gx, gy;
            unpack2(acc[oy][ox], gx, gy);
            const uint32_t lp = (uint32_t)(oy * 8 + (lane >> 4) * 4 + ox);     // pixel inside the warp's 2 x 8 region
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(vb + lp * 64u + (uint32_t)cp * 4u), "r"(h2_bits(f2h2_sat(gx, gy)))
                         : "memory");
          }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmV, vb, ch * UC, x0 + 8 * (dwi & 1), y0 + 2 * (dwi >> 1), b);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    // ---- norm partials: reduce the 16 pixel blocks of every channel in a fixed order (red aliases the v staging:
    //      every warp's stores have drained before anyone writes it) ----
    asm volatile("bar.sync 1, %0;" ::"n"(DW_WARPS * 32) : "memory");
#pragma unroll
    for (int ch = 0; ch < NQK; ++ch) {
      float n0, n1;
      unpack2(nrm[ch], n0, n1);
      red[blk * (NQK * UC) + ch * UC + 2 * cp] = n0;
      red[blk * (NQK * UC) + ch * UC + 2 * cp + 1] = n1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DW_WARPS * 32) : "memory");
    if (ctid < 2 * CW) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sum += red[k * (NQK * UC) + ctid];
      // channel ctid of q|k -> n_part[b][head][part][which][c]
      const int which = ctid >= CW ? 1 : 0, c = ctid - which * CW;
      const int chd = CW / p.heads, head = c / chd, cc = c - head * chd;
      p.n_part[((((long long)b * p.heads + head) * p.parts + part) * 2 + which) * chd + cc] = sum;
    }
  } else {
    // =============================== convert (patch rows 0 .. 127), then the CTA's partial Gram ===============================
    const int q = warp;
    const bool any = TileIter(p).valid();
    convert_tiles<CW>(bars, p, tmem_base, q, 0, q * 32, sH, lane);
    const int chd = CW / p.heads;
    const int i = warp * 32 + lane;                         // TMEM lane == q channel
    if (any) {
      mbar_wait(smem_u32(&bars->acc_done), 0);
      tc_fence_after();
    }
    const int head = i / chd, ii = i - head * chd;
    for (int c0 = 0; c0 < CW; c0 += 32) {
      float v[32];
      if (any) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + S_COL0 + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
      if (i < CW) {
        float* dst = p.s_part + ((((long long)b * p.heads + head) * p.parts + part) * chd + ii) * chd;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int jn = c0 + e - head * chd;                // column inside this head's diagonal block
          if (c0 + e < CW && jn >= 0 && jn < chd) dst[jn] = v[e];
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

struct FusedFrontCfg {
  int xrows;
  uint32_t off_a, off_w, off_h, off_x, off_v, off_bars;
  size_t smem;
};

template <int CW>
bool configure_w(int heads, FusedFrontCfg& c) {
  using G = Geo<CW>;
  if (heads <= 0 || CW % heads != 0) return false;
  c.xrows = std::max(2 * CW, 128);                      // A reads 128 rows from row 0 (M = 128), B reads C rows from row C
  size_t off = 0;
  c.off_a = (uint32_t)off; off += (size_t)G::NA * G::A_BYTES;
  c.off_w = (uint32_t)off; off += align_up((size_t)G::W_BYTES, 1024);
  c.off_x = (uint32_t)off; off += (size_t)2 * c.xrows * 128;
  c.off_v = (uint32_t)off; off += (size_t)2 * VBOXB;
  c.off_h = (uint32_t)off; off += (size_t)G::NH * HSTAGE;
  off = align_up(off, 16);
  c.off_bars = (uint32_t)off; off += sizeof(Bars);
  c.smem = off + 1024;          // alignment slack
  return c.smem <= 227 * 1024 && (size_t)16 * 2 * CW * 4 <= (size_t)2 * VBOXB;
}

bool configure(int C, int heads, FusedFrontCfg& c) {
  return C == 48 ? configure_w<48>(heads, c) : C == 96 ? configure_w<96>(heads, c) : false;
}

template <int CW, int DBG = 0>
int launch_inst(const CUtensorMap& tV, const CUtensorMap& tX, const FusedFrontParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(attn_fused_kernel<CW, DBG>, optin));
  attn_fused_kernel<CW, DBG><<<grid, Geo<CW>::NTHREADS, smem, s>>>(tV, tX, p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace

bool attn_fused_v1_supported(int C, int heads) {
  FusedFrontCfg c;
  return configure(C, heads, c);
}

static int attn_fused_wrows_v1(int C) { return (3 * C + UC - 1) / UC * UC; }

static int attn_fused_parts_v1(int B, int H, int W) {
  const int tiles = cdiv(W, TW) * cdiv(H, TH);
  return std::max(1, std::min(tiles, 148 / std::max(1, B)));
}

int launch_attn_fused_v1(const AttnFusedArgs& a, cudaStream_t s) {
  FusedFrontCfg c;
  IRB_REQUIRE(configure(a.C, a.heads, c), "attn_fused: unsupported shape");
  IRB_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0 && a.B <= 65535, "attn_fused: bad extent");
  IRB_REQUIRE(a.ln_mode == LN_BIASFREE || a.ln_mode == LN_WITHBIAS, "attn_fused: bad LayerNorm mode");
  IRB_REQUIRE((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0, "attn_fused: x must be 16-byte aligned");
  CUtensorMap tV, tX;
  {
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.C * 2 * a.W, (cuuint64_t)a.C * 2 * a.W * a.H};
    cuuint32_t box[4] = {UC, 8, 2, 1};                 // one dw warp's region
    IRB_TRY(make_tmap(&tV, a.v, true, 4, d, st, box, false));
  }
  {
    // the x halo box of a tile, for the L2 prefetch only (never a shared-memory destination)
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)a.C * 4 * a.W, (cuuint64_t)a.C * 4 * a.W * a.H};
    cuuint32_t box[4] = {(cuuint32_t)a.C, PW, TH + 2, 1};
    IRB_TRY(make_tmap(&tX, a.x, false, 4, d, st, box, false));
  }
  FusedFrontParams p{};
  p.x = a.x; p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_mode = a.ln_mode;
  p.w_qkv = reinterpret_cast<const uint8_t*>(a.w_qkv); p.dw = a.dw_chunked;
  p.s_part = a.s_part; p.n_part = a.n_part;
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.heads = a.heads;
  p.parts = attn_fused_parts_v1(a.B, a.H, a.W);
  IRB_REQUIRE(p.parts == a.parts, "attn_fused: partial count mismatch");
  p.nv = cdiv(a.C, UC); p.NP = attn_fused_wrows_v1(a.C);
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.tiles_per_img = p.tiles_x * p.tiles_y;
  p.xrows = c.xrows;
  p.off_a = c.off_a; p.off_w = c.off_w; p.off_h = c.off_h; p.off_x = c.off_x; p.off_v = c.off_v; p.off_bars = c.off_bars;
  dim3 grid(p.parts, a.B, 1);
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);     // one CTA per SM
  const double pix = (double)a.B * a.H * a.W;
  // algorithmic bytes: x read (fp32) + v write (fp16); flops: qkv 1x1 + depthwise + Gram
  ProfScope prof(TAG_ATTN_FUSED, pix * (4.0 * a.C + 2.0 * a.C),
                 pix * (2.0 * 3 * a.C * a.C + 2.0 * 9 * 3 * a.C + 2.0 * a.C * (a.C / a.heads)), s);
#ifdef IRB_FUSED_EXPERIMENTS
  static const int dbg = getenv("IRB_AF_DBG") ? atoi(getenv("IRB_AF_DBG")) : 0;
  if (a.C == 96) {
    switch (dbg) {
      case 1: return launch_inst<96, 1>(tV, tX, p, grid, smem, s);
      case 2: return launch_inst<96, 2>(tV, tX, p, grid, smem, s);
      case 3: return launch_inst<96, 3>(tV, tX, p, grid, smem, s);
      case 8: return launch_inst<96, 8>(tV, tX, p, grid, smem, s);
      case 11: return launch_inst<96, 11>(tV, tX, p, grid, smem, s);
      default: break;
    }
  }
#endif
  return a.C == 48 ? launch_inst<48>(tV, tX, p, grid, smem, s) : launch_inst<96>(tV, tX, p, grid, smem, s);
}

}  // namespace irb
