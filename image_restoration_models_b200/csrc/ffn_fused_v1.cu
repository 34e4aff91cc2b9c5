// The whole gated-dconv feed-forward network (GDFN, restormer.py:88-93 and the residual add :148) in ONE kernel:
//
//     x[pixel, :] += W_out . ( gelu(dw3x3(W_in . xn)[pixel, 0:hp]) * dw3x3(W_in . xn)[pixel, hp:2hp] )
//
// xn = LayerNorm(x) (norm2) arrives as an fp16 tensor.  The 2*hp-wide hidden tensor -- 5.3 x the residual stream, written
// and re-read once per block by the two-kernel version (16*hp of the block's ~55*C bytes per pixel) -- never exists in
// HBM: project_in is recomputed per 8 x 16 pixel tile over the (8+2) x (16+2) halo the depthwise conv needs
// (1.4 x the contraction work; the tensor pipe is nearly idle on this path) and lives in shared memory only.
//
// One persistent CTA per SM, 16 warps.  The pipeline unit is 64 hidden channels of ONE GDFN half (x1 or x2):
//
//   producer (1 thread)   per tile: the xn halo patch as [180 px][128 B] SWIZZLE_128B boxes (4-D bulk-tensor load; the
//                         zero fill outside the image IS the conv's zero padding: project_in has no bias, so xn = 0
//                         gives hidden = 0); per unit: the unit's 64 project_in rows; per chunk: the project_out rows
//   MMA (1 warp)          MMA1: D1[192 px][64] = xn_patch . W_in_unit^T as two M = 128 instructions (rows 0-127 and
//                         64-191) into a double-buffered TMEM accumulator, issued two units ahead of the consumers;
//                         MMA2: D2[128 px][C] += gated_chunk . W_out_chunk^T, accumulated over the hp/64 chunks
//   convert (4 + 2 warps) tcgen05.ld of D1 -> fp16 -> hidden patch [180 px][144 B] in shared memory (row pitch padded
//                         to 144 B: conflict-free 16-byte stores, one pixel per lane)
//   dw warps (8)          thread = 2 channels x a 4x4 pixel block: 3x3 taps (packed FFMA2) over the smem patch; after the
//                         x2 unit the exact-erf GELU gate; result stored as MMA2's [128 px][128 B] operand box
//   epilogue (warps 0-3)  tcgen05.ld of D2 -> swizzled staging boxes -> bulk-tensor REDUCTION (x += ...): the residual is
//                         never loaded.  At C = 96 the staging boxes are the tile's retired xn buffer.
//
// Channels-last: x fp32 [B][H][W][C], xn fp16 [B][H][W][C]; every access to either is a bulk-tensor copy.
#include "common.cuh"
#include "ffn_fused.cuh"
#include "gdfn_math.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace irb {

namespace {

using namespace sm100;
using namespace gdfn;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int PW = TW + 2;                     // halo patch width
constexpr int HPIX = (TH + 2) * PW;            // 180 halo pixels
constexpr int AROWS = 192;                     // rows the two M = 128 instructions read (64 .. 191 for the second)
constexpr int ABOX = AROWS * 128;              // one 64-channel box of the xn patch
constexpr int HROW = 144;                      // hidden patch row pitch (128 B of channels + 16 B pad)
constexpr int HSTAGE = HPIX * HROW;
constexpr int UC = 64;                         // hidden channels per unit
constexpr int WINBOX = UC * 128;               // one 64-channel K box of a unit's project_in rows
constexpr int OPBOX = TM * 128;                // MMA2 operand box, and one 32-channel staging group of the epilogue
constexpr int EPI_WARPS = 4, CVB_WARPS = 2;
// warp roles; BH = rows of a depthwise thread's pixel block: 4 -> 8 dw warps (4x4 blocks), 2 -> 16 dw warps (2x4 blocks).
// Measured: 16 warps of 2x4 blocks are 9 % SLOWER (more halo loads and conversions per output; the kernel is bound by
// the FP32 pipe, not by latency), so only BH = 4 is instantiated.
template <int BH> struct Roles {
  static constexpr int DW_WARPS = 32 / BH;
  static constexpr int WARP_DW = EPI_WARPS, WARP_MMA = WARP_DW + DW_WARPS, WARP_PROD = WARP_MMA + 1, WARP_CVB = WARP_PROD + 1;
  static constexpr int NWARPS = WARP_CVB + CVB_WARPS;
  static_assert((WARP_CVB & 3) == 2, "the two extra convert warps must own TMEM lane quarters 2 and 3");
};
constexpr int D1_COLS = 128;                   // per buffer: rows 0-127 in columns [0, 64), rows 64-191 in [64, 128)
constexpr int D2_COL0 = 2 * D1_COLS;

struct Bars {
  unsigned long long a_full[2], a_empty[2];
  unsigned long long w1_full[2], w1_empty[2];
  unsigned long long w2_full[2], w2_empty[2];
  unsigned long long d1_full[2], d1_empty[2];
  unsigned long long h_full[2], h_empty[2];
  unsigned long long op_ready, op_empty;
  unsigned long long acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct FusedParams {
  const uint8_t* w_in;
  const uint8_t* w_out;
  const float* dw;
  int B, H, W, C, hp, nchunk, nunits;
  int nkb;                 // 64-channel K boxes of xn / W_in
  int ks_last;             // K steps (16 channels) in the last box; 4 in the others
  int tiles_x, tiles_y, ntiles;
  int ngroups;             // 32-channel groups of the output
  int acc_stride;
  int stg_in_a;
  uint32_t a_bytes, win_bytes, wout_bytes;
  uint32_t off_a, off_win, off_op, off_wout, off_stg, off_h, off_bars;
};

struct TileIter {
  int t, step, end, tx_n, ty_n;
  __device__ TileIter(const FusedParams& p) : t(blockIdx.x), step(gridDim.x), end(p.ntiles), tx_n(p.tiles_x), ty_n(p.tiles_y) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int img() const { return t / (tx_n * ty_n); }
  __device__ int y0() const { return ((t / tx_n) % ty_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * TW; }
};

__device__ __forceinline__ f2_t ld_h2(uint32_t a) {
  uint32_t t;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(a));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&t));
  return pack2(f.x, f.y);
}

// Measured alternatives for the tap arithmetic, all slower than packed FFMA2 on converted operands: mixed-precision
// FHFMA (fma.rn.f32.f16, no conversions, two scalar instructions per channel pair): +18 %; FFMA2 alternated with
// scalar FFMA pairs: +6 %; 16 dw warps of 2x4 blocks: +9 %.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = f2h2_sat(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// D1 -> fp16 hidden patch for the 32 patch rows this warp owns (TMEM lane quarter `q`, rows row0 .. row0 + 31)
__device__ __forceinline__ void convert_unit(Bars* bars, uint32_t g, uint32_t tmem_base, int q, int col_off, int row0,
                                             uint32_t sH, int lane) {
  const uint32_t s = g & 1u, ph = (g >> 1) & 1u;
  mbar_wait(smem_u32(&bars->d1_full[s]), ph);
  tc_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * D1_COLS + (uint32_t)col_off;
  mbar_wait(smem_u32(&bars->h_empty[s]), ph ^ 1u);
  const int row = row0 + lane;
  const uint32_t dst = sH + s * HSTAGE + (uint32_t)row * HROW;
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    float v[32];
    tmem_ld32(taddr + hlf * 32, v);
    tmem_ld_wait();
    if (hlf == 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->d1_empty[s]));
    }
    if (row < HPIX) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_h2(v[8 * c + 0], v[8 * c + 1]);
        u.y = pack_h2(v[8 * c + 2], v[8 * c + 3]);
        u.z = pack_h2(v[8 * c + 4], v[8 * c + 5]);
        u.w = pack_h2(v[8 * c + 6], v[8 * c + 7]);
        sts128u(dst + (hlf * 4 + c) * 16, u);
      }
    }
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(smem_u32(&bars->h_full[s]));
}

// DBG != 0: timing experiments only (results are garbage): 1 no W_in reloads, 2 no xn patch reloads, 4 no depthwise taps,
// 8 no W_out reloads
// CW: channel count when it is one of the shipped widths (48 / 96), else 0 (generic: runtime K loops).  With CW known the
// MMA warp's issue loop is straight-line code: descriptors are a base plus a compile-time constant and the unit / chunk /
// tile counters advance by compare-and-wrap.  (ncu on attn_fused.cu: ~33 instructions of descriptor and uniform-register
// traffic per tcgen05.mma when the loop bounds are runtime values, and the issuing warp paced the kernel.)
template <int BH, int DBG, int CW>
__global__ void __launch_bounds__(Roles<BH>::NWARPS * 32, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  Bars* bars = reinterpret_cast<Bars*>(gbase + p.off_bars);
  const uint32_t sA = base + p.off_a, sWin = base + p.off_win, sOP = base + p.off_op, sWout = base + p.off_wout,
                 sStg = base + p.off_stg, sH = base + p.off_h;

  using R = Roles<BH>;
  constexpr int DW_WARPS = R::DW_WARPS, WARP_DW = R::WARP_DW, WARP_MMA = R::WARP_MMA, WARP_PROD = R::WARP_PROD,
                WARP_CVB = R::WARP_CVB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), 1 + (p.stg_in_a ? EPI_WARPS : 0));   // last MMA1 of the tile (+ the epilogue's staging)
      mbar_init(smem_u32(&bars->w1_full[s]), 1);
      mbar_init(smem_u32(&bars->w1_empty[s]), 1);
      mbar_init(smem_u32(&bars->w2_full[s]), 1);
      mbar_init(smem_u32(&bars->w2_empty[s]), 1);
      mbar_init(smem_u32(&bars->d1_full[s]), 1);
      mbar_init(smem_u32(&bars->d1_empty[s]), EPI_WARPS + CVB_WARPS);
      mbar_init(smem_u32(&bars->h_full[s]), EPI_WARPS + CVB_WARPS);
      mbar_init(smem_u32(&bars->h_empty[s]), DW_WARPS);
      mbar_init(smem_u32(&bars->acc_full[s]), 1);
      mbar_init(smem_u32(&bars->acc_empty[s]), EPI_WARPS);
    }
    mbar_init(smem_u32(&bars->op_ready), DW_WARPS);
    mbar_init(smem_u32(&bars->op_empty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t nunits = (uint32_t)p.nunits;

  if (warp == WARP_PROD) {
    // =============================== producer: xn patches, project_in rows, project_out rows ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      const uint32_t patch_bytes = (uint32_t)p.nkb * (uint32_t)(HPIX * 128);
      TileIter ta(p);          // the next tile whose patch is to be loaded
      uint32_t ja = 0;
      auto load_a = [&]() {
        const uint32_t buf = ja & 1u, fb = smem_u32(&bars->a_full[buf]);
        if ((DBG & 2) && ja >= 2) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, patch_bytes);
          for (int kb = 0; kb < p.nkb; ++kb)
            tma_load_4d(&tmA, fb, sA + buf * p.a_bytes + (uint32_t)kb * ABOX, kb * 64, ta.x0() - 1, ta.y0() - 1, ta.img());
        }
        ta.next();
        ++ja;
      };
      if (ta.valid()) load_a();     // first tile: the buffer is free
      uint32_t g = 0, cc = 0, j = 0;
      auto load_wout = [&]() {
        const uint32_t s = cc & 1u, fb = smem_u32(&bars->w2_full[s]), ch = cc % (uint32_t)p.nchunk;
        mbar_wait(smem_u32(&bars->w2_empty[s]), ((cc >> 1) & 1u) ^ 1u);
        if ((DBG & 8) && cc >= 2) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, (uint32_t)p.C * 128u);
          bulk_load(sWout + s * p.wout_bytes, p.w_out + (size_t)ch * p.C * 128, (uint32_t)p.C * 128u, fb);
        }
        ++cc;
      };
      for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
        for (uint32_t u = 0; u < nunits; ++u, ++g) {
          const uint32_t set = u & 1u, ch = u >> 1;
          {
            const uint32_t s = g & 1u, fb = smem_u32(&bars->w1_full[s]);
            mbar_wait(smem_u32(&bars->w1_empty[s]), ((g >> 1) & 1u) ^ 1u);
            if ((DBG & 1) && g >= 2) {
              mbar_arrive(fb);
            } else {
              mbar_expect_tx(fb, (uint32_t)p.nkb * WINBOX);
              const size_t n0 = (size_t)set * p.hp + (size_t)ch * UC;
              for (int kb = 0; kb < p.nkb; ++kb)
                bulk_load(sWin + s * p.win_bytes + (uint32_t)kb * WINBOX, p.w_in + ((size_t)kb * 2 * p.hp + n0) * 128, WINBOX, fb);
            }
          }
          // project_out rows of the PREVIOUS chunk: its slot was released by the MMA2 three chunks back, so this wait
          // never holds up the project_in loads behind it (the chunk's own MMA2 is still several units away)
          if (set == 0 && g > 0) load_wout();
          // the next tile's patch: as soon as its buffer retires (polled), at the latest with the tile's last unit
          if (ja == j + 1 && ta.valid() && u >= 1) {
            const uint32_t eb = smem_u32(&bars->a_empty[ja & 1u]), par = ((ja >> 1) & 1u) ^ 1u;
            if (u == nunits - 1) { mbar_wait(eb, par); load_a(); }
            else if (mbar_test(eb, par)) load_a();
          }
        }
      }
      if (g > 0) load_wout();       // the last chunk's project_out rows
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc1 = make_idesc<__half>(UC), idesc2 = make_idesc<__half>(p.C);
    uint32_t ntl = 0;
    for (TileIter ti(p); ti.valid(); ti.next()) ++ntl;
    const uint32_t G = ntl * nunits;
    const uint64_t adesc0 = sw128_desc(sA), wdesc0 = sw128_desc(sWin), opdesc = sw128_desc(sOP), wodesc0 = sw128_desc(sWout);
    // MMA1 of unit g: (tile j1, unit u1) advance with it
    uint32_t j1 = 0, u1 = 0;
    auto issue1 = [&](uint32_t g) {
      const uint32_t s = g & 1u, ph = (g >> 1) & 1u, ab = j1 & 1u;
      if (u1 == 0) mbar_wait(smem_u32(&bars->a_full[ab]), (j1 >> 1) & 1u);
      mbar_wait(smem_u32(&bars->w1_full[s]), ph);
      mbar_wait(smem_u32(&bars->d1_empty[s]), ph ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + s * D1_COLS;
      const uint64_t ad = adesc0 + (uint64_t)((ab * p.a_bytes) >> 4), wd = wdesc0 + (uint64_t)((s * p.win_bytes) >> 4);
      if constexpr (CW > 0) {
        constexpr int NKB = (CW + 63) / 64, KS_LAST = (CW - 64 * (NKB - 1)) / 16;
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
          for (int kk = 0; kk < (kb == NKB - 1 ? KS_LAST : 4); ++kk) {
            const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
            const uint64_t bd = wd + (uint64_t)((kb * WINBOX + kk * 32) >> 4);
            const uint64_t a0 = ad + (uint64_t)((kb * ABOX + kk * 32) >> 4);
            umma_elect<__half>(d, a0, bd, idesc1, acc);
            umma_elect<__half>(d + 64, a0 + (uint64_t)((64 * 128) >> 4), bd, idesc1, acc);
          }
      } else {
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int ks = kb == p.nkb - 1 ? p.ks_last : 4;
          for (int kk = 0; kk < ks; ++kk) {
            const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
            const uint64_t bd = wd + (uint64_t)((kb * WINBOX + kk * 32) >> 4);
            const uint64_t a0 = ad + (uint64_t)((kb * ABOX + kk * 32) >> 4);
            umma_elect<__half>(d, a0, bd, idesc1, acc);
            umma_elect<__half>(d + 64, a0 + (uint64_t)((64 * 128) >> 4), bd, idesc1, acc);
          }
        }
      }
      umma_commit_elect(smem_u32(&bars->d1_full[s]));
      umma_commit_elect(smem_u32(&bars->w1_empty[s]));
      if (u1 == nunits - 1) umma_commit_elect(smem_u32(&bars->a_empty[ab]));
      __syncwarp();
      if (++u1 == nunits) { u1 = 0; ++j1; }
    };
    // MMA2 of chunk cc: (tile j2, chunk c2) advance with it
    uint32_t j2 = 0, c2 = 0;
    auto issue2 = [&](uint32_t cc) {
      const uint32_t slot = j2 & 1u, s = cc & 1u;
      if (c2 == 0) mbar_wait(smem_u32(&bars->acc_empty[slot]), ((j2 >> 1) & 1u) ^ 1u);
      mbar_wait(smem_u32(&bars->w2_full[s]), (cc >> 1) & 1u);
      mbar_wait(smem_u32(&bars->op_ready), cc & 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + D2_COL0 + slot * (uint32_t)p.acc_stride;
      const uint64_t wd = wodesc0 + (uint64_t)((s * p.wout_bytes) >> 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_elect<__half>(d, opdesc + (uint64_t)((kk * 32) >> 4), wd + (uint64_t)((kk * 32) >> 4), idesc2, (c2 > 0 || kk > 0) ? 1u : 0u);
      umma_commit_elect(smem_u32(&bars->op_empty));
      umma_commit_elect(smem_u32(&bars->w2_empty[s]));
      if (c2 == (uint32_t)p.nchunk - 1) umma_commit_elect(smem_u32(&bars->acc_full[slot]));
      __syncwarp();
      if (++c2 == (uint32_t)p.nchunk) { c2 = 0; ++j2; }
    };
    if (G > 0) issue1(0);
    if (G > 1) issue1(1);
    for (uint32_t g = 0; g < G; ++g) {
      if (g + 2 < G) issue1(g + 2);
      if (g & 1u) issue2(g >> 1);
    }
  } else if (warp >= WARP_CVB) {
    // =============================== convert: patch rows 128 .. 179 (second MMA, TMEM lanes 64 .. 127) ===============================
    const int q = warp & 3;
    uint32_t g = 0;
    for (TileIter ti(p); ti.valid(); ti.next())
      for (uint32_t u = 0; u < nunits; ++u, ++g) convert_unit(bars, g, tmem_base, q, 64, 128 + (q - 2) * 32, sH, lane);
  } else if (warp >= WARP_DW) {
    // =============================== depthwise 3x3 + gate -> operand box ===============================
    constexpr int BW = 4;
    const int cp = lane, blk = warp - WARP_DW;
    const int by = blk / (TW / BW), bx = blk % (TW / BW);
    const uint32_t win0 = (uint32_t)((BH * by) * PW + BW * bx) * HROW + (uint32_t)cp * 4u;
    uint32_t g = 0, cc = 0;
    for (TileIter ti(p); ti.valid(); ti.next()) {
      for (int ch = 0; ch < p.nchunk; ++ch, ++cc) {
        f2_t acc[2][BH][BW];
#pragma unroll
        for (int set = 0; set < 2; ++set, ++g) {
          f2_t w[9];
          const float* taps = p.dw + ((size_t)(ch * 2 + set) * 9) * UC + cp * 2;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float2 f = __ldg(reinterpret_cast<const float2*>(taps + t * UC));
            w[t] = pack2(f.x, f.y);
          }
          const uint32_t s = g & 1u;
          mbar_wait(smem_u32(&bars->h_full[s]), (g >> 1) & 1u);
          const uint32_t src = sH + s * HSTAGE + win0;
          if (DBG & 4) {
#pragma unroll
            for (int oy = 0; oy < BH; ++oy)
#pragma unroll
              for (int ox = 0; ox < BW; ++ox) acc[set][oy][ox] = w[oy];
          } else
#pragma unroll
          for (int iy = 0; iy < BH + 2; ++iy) {
            f2_t v[BW + 2];
#pragma unroll
            for (int ix = 0; ix < BW + 2; ++ix) v[ix] = ld_h2(src + (uint32_t)(iy * PW + ix) * HROW);
#pragma unroll
            for (int oy = 0; oy < BH; ++oy) {
              const int ky = iy - oy;
              if (ky < 0 || ky > 2) continue;
#pragma unroll
              for (int ox = 0; ox < BW; ++ox) {
                if (ky == 0) acc[set][oy][ox] = mul2(w[0], v[ox]);
                else acc[set][oy][ox] = fma2(w[ky * 3], v[ox], acc[set][oy][ox]);
                acc[set][oy][ox] = fma2(w[ky * 3 + 1], v[ox + 1], acc[set][oy][ox]);
                acc[set][oy][ox] = fma2(w[ky * 3 + 2], v[ox + 2], acc[set][oy][ox]);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        }
        mbar_wait(smem_u32(&bars->op_empty), (cc & 1u) ^ 1u);
#pragma unroll
        for (int oy = 0; oy < BH; ++oy)
#pragma unroll
          for (int ox = 0; ox < BW; ++ox) {
            float gx, gy;
            unpack2(gelu_gate2(acc[0][oy][ox], acc[1][oy][ox]), gx, gy);
            const uint32_t row = (uint32_t)((BH * by + oy) * TW + BW * bx + ox);
            const uint32_t a = sOP + row * 128u + ((((uint32_t)cp >> 2) ^ (row & 7u)) << 4) + ((uint32_t)cp & 3u) * 4u;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(pack_h2(gx, gy)) : "memory");
          }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->op_ready));
      }
    }
  } else {
    // =============================== convert (patch rows 0 .. 127) + epilogue ===============================
    const int q = warp;
    const uint32_t lsw = (uint32_t)(lane & 7);
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    auto epilogue = [&](uint32_t j, int b, int y0, int x0) {
      const uint32_t slot = j & 1u;
      mbar_wait(smem_u32(&bars->acc_full[slot]), (j >> 1) & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + D2_COL0 + slot * (uint32_t)p.acc_stride;
      const uint32_t stg = (p.stg_in_a ? sA + (j & 1u) * p.a_bytes : sStg) + (uint32_t)q * 4096u + (uint32_t)lane * 128u;
      for (int gi = 0; gi < p.ngroups; ++gi) {
        float v[32];
        tmem_ld32(tacc + (uint32_t)(gi * 32), v);
        tmem_ld_wait();
        if (gi == p.ngroups - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[slot]));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(stg + (uint32_t)gi * OPBOX + (((uint32_t)c ^ lsw) << 4), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t box = (p.stg_in_a ? sA + (j & 1u) * p.a_bytes : sStg) + (uint32_t)q * 4096u;
        for (int gi = 0; gi < p.ngroups; ++gi)
          tma_reduce_add_4d(&tmY, box + (uint32_t)gi * OPBOX, gi * 32, x0, y0 + 2 * q, b);   // this warp's rows 2q, 2q+1
        bulk_commit();
      }
      __syncwarp();
    };
    // the staging boxes are free (and, at C = 96, the xn buffer they live in may be reloaded) once the reductions have read them
    auto epilogue_done = [&](uint32_t j) {
      if (lane == 0) {
        bulk_wait_read<0>();
        if (p.stg_in_a) mbar_arrive(smem_u32(&bars->a_empty[j & 1u]));
      }
      __syncwarp();
    };
    uint32_t g = 0, j = 0;
    int pb = 0, py0 = 0, px0 = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      for (uint32_t u = 0; u < nunits; ++u, ++g) {
        convert_unit(bars, g, tmem_base, q, 0, q * 32, sH, lane);
        if (j > 0) {
          if (u == 1) epilogue(j - 1, pb, py0, px0);             // the previous tile's output, once this tile is primed
          if (u == 3) epilogue_done(j - 1);
        }
      }
      pb = ti.img(); py0 = ti.y0(); px0 = ti.x0();
    }
    if (j > 0) { epilogue(j - 1, pb, py0, px0); epilogue_done(j - 1); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct FusedCfg {
  int nkb, ks_last, ngroups, stg_in_a;
  uint32_t a_bytes, win_bytes, wout_bytes, off_a, off_win, off_op, off_wout, off_stg, off_h, off_bars;
  size_t smem;
};

bool configure(int C, int hp, FusedCfg& c) {
  if (C % 16 != 0 || C < 16 || C > 128 || hp % UC != 0 || hp / UC < 2) return false;
  c.nkb = (C + 63) / 64;
  c.ks_last = (C - 64 * (c.nkb - 1)) / 16;
  c.ngroups = (C + 31) / 32;
  c.a_bytes = (uint32_t)c.nkb * ABOX;
  c.win_bytes = (uint32_t)c.nkb * WINBOX;
  c.wout_bytes = (uint32_t)align_up((size_t)C * 128, 1024);
  const uint32_t stg_bytes = (uint32_t)c.ngroups * OPBOX;
  c.stg_in_a = stg_bytes <= c.a_bytes ? 1 : 0;
  size_t off = 0;
  c.off_a = (uint32_t)off; off += 2 * (size_t)c.a_bytes;
  c.off_win = (uint32_t)off; off += 2 * (size_t)c.win_bytes;
  c.off_op = (uint32_t)off; off += OPBOX;
  c.off_wout = (uint32_t)off; off += 2 * (size_t)c.wout_bytes;
  c.off_stg = (uint32_t)off; if (!c.stg_in_a) off += stg_bytes;
  c.off_h = (uint32_t)off; off += 2 * (size_t)HSTAGE;
  off = align_up(off, 16);
  c.off_bars = (uint32_t)off; off += sizeof(Bars);
  c.smem = off + 1024;          // alignment slack
  return c.smem <= 227 * 1024;
}

}  // namespace

bool ffn_fused_v1_supported(int C, int hp) {
  FusedCfg c;
  return configure(C, hp, c);
}

int launch_ffn_fused_v1(const FfnFusedArgs& a, cudaStream_t s) {
  FusedCfg c;
  IRB_REQUIRE(configure(a.C, a.hp, c), "ffn_fused: unsupported shape");
  IRB_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0, "ffn_fused: empty input");
  CUtensorMap tA, tY;
  {
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.C * 2 * a.W, (cuuint64_t)a.C * 2 * a.W * a.H};
    cuuint32_t box[4] = {64, PW, TH + 2, 1};
    IRB_TRY(make_tmap(&tA, a.xn, true, 4, d, st, box, true));
  }
  {
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)a.C * 4 * a.W, (cuuint64_t)a.C * 4 * a.W * a.H};
    cuuint32_t box[4] = {32, TW, 2, 1};
    IRB_TRY(make_tmap(&tY, a.x, false, 4, d, st, box, true));
  }
  FusedParams p{};
  p.w_in = reinterpret_cast<const uint8_t*>(a.w_in);
  p.w_out = reinterpret_cast<const uint8_t*>(a.w_out);
  p.dw = a.dw_chunked;
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.hp = a.hp; p.nchunk = a.hp / UC; p.nunits = 2 * p.nchunk;
  p.nkb = c.nkb; p.ks_last = c.ks_last; p.ngroups = c.ngroups; p.stg_in_a = c.stg_in_a;
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.ntiles = p.tiles_x * p.tiles_y * a.B;
  p.acc_stride = (a.C + 31) / 32 * 32;
  p.a_bytes = c.a_bytes; p.win_bytes = c.win_bytes; p.wout_bytes = c.wout_bytes;
  p.off_a = c.off_a; p.off_win = c.off_win; p.off_op = c.off_op; p.off_wout = c.off_wout; p.off_stg = c.off_stg;
  p.off_h = c.off_h; p.off_bars = c.off_bars;
  const int grid = std::max(1, std::min(p.ntiles, 148));
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);     // one CTA per SM: the kernel owns all 512 TMEM columns
  const double pix = (double)a.B * a.H * a.W;
  // algorithmic bytes: xn read (fp16) + x read-modify-write (fp32); flops: project_in + depthwise + project_out
  ProfScope prof(TAG_FFN_FUSED, pix * (2.0 * a.C + 8.0 * a.C), pix * (4.0 * a.hp * a.C + 36.0 * a.hp + 2.0 * a.hp * a.C), s);
  // one opt-in table per KERNEL: keyed by an integer tag, because every instantiation has the same function-pointer type
  // (a table in a generic lambda would be shared by all of them and only the first kernel would ever be opted in)
  auto go = [&](auto kernel, auto tag) -> int {
    static SmemOptIn optin;      // one per (lambda instantiation == tag type)
    (void)tag;
    IRB_TRY(opt_in_smem(kernel, optin));
    kernel<<<grid, Roles<4>::NWARPS * 32, smem, s>>>(tA, tY, p);
    return IR_OK;
  };
#define IRB_GO(DBG_, CW_) go(ffn_fused_kernel<4, DBG_, CW_>, std::integral_constant<int, (DBG_) * 1000 + (CW_)>{})
#ifdef IRB_FUSED_EXPERIMENTS
  static const int dbg = getenv("IRB_FUSED_DBG") ? atoi(getenv("IRB_FUSED_DBG")) : 0;
  switch (dbg) {
    case 0: IRB_TRY(IRB_GO(0, 0)); break;
    case 1: IRB_TRY(IRB_GO(1, 0)); break;
    case 3: IRB_TRY(IRB_GO(3, 0)); break;
    case 4: IRB_TRY(IRB_GO(4, 0)); break;
    case 11: IRB_TRY(IRB_GO(11, 0)); break;
    default: IRB_TRY(IRB_GO(15, 0)); break;
  }
#else
  static const bool generic = getenv("IRB_FFN_GENERIC_ISSUE") != nullptr;      // A/B switch for benchmarks
  if (a.C == 96 && !generic) IRB_TRY(IRB_GO(0, 96));
  else if (a.C == 48 && !generic) IRB_TRY(IRB_GO(0, 48));
  else IRB_TRY(IRB_GO(0, 0));
#endif
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace irb
