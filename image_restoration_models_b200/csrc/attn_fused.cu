// The whole front of the transposed attention (MDTA, restormer.py:111-124) behind norm1 in ONE kernel, second version: the
// depthwise taps read the qkv patch STRAIGHT OUT OF TENSOR MEMORY (dw_tmem.cuh).
//
//     qkv     = dw3x3(W_qkv . xn)                             (qkv 1x1 + qkv_dwconv :114), xn = norm1(x) as an fp16 tensor
//     S[b]   += q . k^T over the pixels, |q_i|^2, |k_j|^2     (Gram + the norms F.normalize needs :121-124)
//     v      -> HBM (fp16)                                    (the only tensor written)
//
// As in the first version (git history: attn_fused.cu before the TMEM-direct rewrite) the 1x1 contraction is recomputed per 8 x 16 pixel tile over the (8+2) x (16+2)
// halo the depthwise conv needs and the 3C-wide qkv tensor never exists in HBM.  What changed:
//
// * the orientation of that contraction: D^T[channel][patch pixel] = W_qkv . xn_patch^T -- the WEIGHTS are the MMA's A
//   operand (M = 128 channels = 128 TMEM lanes), the fp16 xn patch the B operand (N = 192 >= 180 patch pixels = TMEM columns).
//   A depthwise thread owns one channel and pulls the patch pixels it needs with tcgen05.ld: fp32, in registers, in the order
//   the packed FFMA2 taps consume them.  The first version moved D1 through six convert warps (tcgen05.ld -> fp16 -> shared
//   memory) and read it back with 24 LDS + 48 conversions per 72 FFMA2.  Here the qkv patch is never rounded to fp16, never
//   touches shared memory, and the convert warps are gone (scripts/probe_dw_tmem.cu: 72 % of the FFMA2 issue rate with two
//   depthwise warps per scheduler);
// * LayerNorm left the kernel.  Eight LayerNorm warps per CTA (register loads of the fp32 halo rows, two shuffle reductions
//   per row, swizzled fp16 stores) paced the first TMEM-direct build at 7.7 k cycles per tile with the depthwise role
//   switched off (timing experiments, profiles/r02_attn_fused_experiments.json).  xn = norm1(x) now arrives as an fp16 tensor
//   (layernorm_rows_kernel, 0.9 of the HBM roofline, or the producing kernel's epilogue) and ONE bulk-tensor load per K box
//   drops the halo patch into the MMA's SWIZZLE_128B operand boxes; the zero fill outside the image IS the depthwise conv's
//   zero padding (the qkv conv has no bias: xn = 0 gives qkv = 0).
//
// One persistent CTA per SM, bound to one image (the Gram accumulates per image), 10 warps:
//
//   depthwise (8 warps)   warp w owns TMEM lane quarter w & 3 (32 channels) and output columns 8 (w >> 2) .. +7 of the tile.
//                         q / k channels: fp16 rows of the X tile ([channel][pixel], one 16-byte store per output row),
//                         squared norms in a register; v channels: fp16 straight to global memory
//                         (a warp writes 64 contiguous bytes per pixel); at the end TMEM -> the CTA's partial S
//   producer + MMA (2 warps) once: W_qkv (fp16 operand image) into shared memory, where it stays.  Per tile: the xn patch
//                         load of the tile after next, MMA1 in groups of 128 channels (A = 128 rows of W from a row offset,
//                         B = the xn patch, N = 192) into a double-buffered accumulator, and the Gram S += q . k^T of the
//                         PREVIOUS tile (A = the q rows, B = the k rows of the X tile)
//
// Channel groups.  3C = 288 (144) channels are 9 (4.5) units of 32 = 4 n + 1; the accumulator has 128 lanes.  Full groups:
// W rows 128 g .. 128 g + 127 (units 4 g .. 4 g + 3 on quarters 0-3; they hold every q | k channel).  The last group is the
// LONE unit (the last 32 channels of v): tile j puts it on lane quarter j & 3 by starting its 128-row window 32 (j & 3) rows
// early.  Rows of D^T are independent, so lanes fed by other rows (or by rows past the weight image) hold values that are
// never read.  The quarter with three units thus changes from tile to tile, and the others run ahead into the next tile.
//
// fold_kernel then reduces the partials in a fixed order (deterministic, no atomics).
#include "attn_fused.cuh"
#include "common.cuh"
#include "dw_tmem.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace irb {

namespace {

using namespace sm100;
using dwt::f2_t;

constexpr int TH = 8, TW = 16;
constexpr int PW = TW + 2;
constexpr int HPIX = (TH + 2) * PW;            // 180 halo pixels
constexpr int AROWS = 192;                     // patch rows the MMA reads (N); rows 180 .. 191 are never written nor read back
constexpr int ABOX = AROWS * 128;              // one 64-channel box of the xn patch
constexpr int UC = 32;                         // channels per unit (one warp's TMEM lane quarter)

constexpr int DW_WARPS = 8;
constexpr int WARP_MMA = DW_WARPS, WARP_PROD = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_PROD + 1) * 32;
constexpr int D1_COLS = AROWS;                 // accumulator columns per group
constexpr int ND = 2;                          // group accumulators in TMEM
constexpr int S_COL0 = ND * D1_COLS;           // the Gram accumulator (<= 96 columns)
constexpr int W_COL0 = S_COL0 + 96;            // the depthwise taps: 9 columns per group, lane = channel (<= 27 of the 32 spare columns)
constexpr int TMEM_COLS = 512;
#ifndef IRB_PROD_POLL_NS
#define IRB_PROD_POLL_NS 200
#endif
#ifndef IRB_MMA_POLL_NS
#define IRB_MMA_POLL_NS 32
#endif
constexpr int PROD_POLL_NS = IRB_PROD_POLL_NS, MMA_POLL_NS = IRB_MMA_POLL_NS;   // sleep between barrier probes of the one-thread roles
constexpr int NA = 2;                          // xn patch buffers

struct Bars {
  unsigned long long a_full[NA], a_empty[NA];
  unsigned long long w_full;
  unsigned long long d1_full[ND], d1_empty[ND];
  unsigned long long x_ready, x_empty, acc_done;
  uint32_t tmem_base;
};

struct FusedFrontParams {
  const uint8_t* w_qkv;    // fp16 SWIZZLE_128B image [nkb][NP][128 B]
  const float* dw;         // taps [nunits][9][32]
  __half* v;               // [B][H][W][C]
  float* s_part;           // [B][heads][parts][ch][ch]
  float* n_part;           // [B][heads][parts][2][ch]
  int B, H, W, C, heads, parts;
  int tiles_x, tiles_y, tiles_per_img;
  int xrows;
  uint32_t off_a, off_w, off_x, off_red, off_bars;
};

// compile-time geometry of one channel width
template <int CW> struct Geo {
  static constexpr int NKB = (CW + 63) / 64;                 // 64-channel K boxes of xn / W_qkv
  static constexpr int KS_LAST = (CW - 64 * (NKB - 1)) / 16; // K steps (16 channels) in the last box
  static constexpr int NUNITS = (3 * CW + UC - 1) / UC;      // 5 / 9
  static constexpr int NP = NUNITS * UC;                     // rows of the W_qkv image
  static constexpr int NG = CW <= 48 ? 2 : 3;                // MMA1 groups per tile
  static constexpr uint32_t A_BYTES = NKB * ABOX;
  static constexpr uint32_t A_TX = NKB * HPIX * 128;         // bytes one patch load delivers
  static constexpr uint32_t W_BYTES = NKB * NP * 128;
  // first W row of a group
  // first W row of a full group (the last group is the lone unit, see lone_start_row)
  __host__ __device__ static constexpr int start_row(int g) { return g * 128; }
  // the LAST group holds one unit only (the last 32 channels of v: 3C is 4 n + 1 units at both widths).  Tile j places it on
  // lane quarter j & 3 (its 128-row window starts 32 * (j & 3) rows early; the other lanes get rows nobody reads), so that the
  // quarter with three units changes from tile to tile while the others run ahead into the next tile's first group
  __host__ __device__ static constexpr int lone_start_row(int r) { return (NUNITS - 1 - r) * UC; }
};
// the unit lane quarter q reads from group g (-1: nothing new in that quarter)
// the unit lane quarter q reads from group g of `ng` (-1: nothing new in that quarter); r = the lone unit's quarter
__device__ __forceinline__ int unit_of(int g, int ng, int nunits, int q, int r) {
  return g < ng - 1 ? 4 * g + q : (q == r ? nunits - 1 : -1);
}

struct TileIter {
  int t, step, end, tx_n;
  __device__ TileIter(const FusedFrontParams& p) : t(blockIdx.x), step(gridDim.x), end(p.tiles_per_img), tx_n(p.tiles_x) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int y0() const { return (t / tx_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * TW; }
};

__device__ __forceinline__ uint32_t h2_bits(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }

// DBG != 0: timing experiments only (results are garbage; compiled with -DIRB_FUSED_EXPERIMENTS, selected by IRB_AF_DBG):
// 2 depthwise warps without TMEM loads / taps, 8 no MMA1 instructions, 16 no v stores, 32 no X-tile stores / norms
template <int CW, int DBG>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const FusedFrontParams p) {
  using G = Geo<CW>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  Bars* bars = reinterpret_cast<Bars*>(gbase + p.off_bars);
  const uint32_t sA = base + p.off_a, sW = base + p.off_w, sX = base + p.off_x;
  float* red = reinterpret_cast<float*>(gbase + p.off_red);    // [2 column halves][2C] squared-norm partials
  const uint32_t xbox = (uint32_t)p.xrows * 128u;

  // the warp index through a warp reduction: its result lives in a UNIFORM register, so the role branches below are uniform
  // branches and the code under them may use the uniform datapath (memory descriptors, TMEM addresses) without R2UR
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));
  const int b = blockIdx.y, part = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->w_full), 1);
    for (int s = 0; s < ND; ++s) {
      mbar_init(smem_u32(&bars->d1_full[s]), 1);
      mbar_init(smem_u32(&bars->d1_empty[s]), DW_WARPS);
    }
    mbar_init(smem_u32(&bars->x_ready), DW_WARPS);
    mbar_init(smem_u32(&bars->x_empty), 1);
    mbar_init(smem_u32(&bars->acc_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * CW; i += NTHREADS) red[i] = 0.f;
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
  pdl_sync();   // set-up done under the previous kernel's tail; from here on global memory is ours (common.cuh)

  if (warp == WARP_PROD) {
    // =============================== producer: W_qkv once, the xn patches ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      const uint32_t fw = smem_u32(&bars->w_full);
      mbar_expect_tx(fw, G::W_BYTES);
      constexpr uint32_t CHUNK = 8192;
      for (uint32_t o = 0; o < G::W_BYTES; o += CHUNK)
        bulk_load(sW + o, p.w_qkv + o, (G::W_BYTES - o) < CHUNK ? (G::W_BYTES - o) : CHUNK, fw);
      uint32_t ja = 0;
      for (TileIter ta(p); ta.valid(); ta.next(), ++ja) {
        const uint32_t buf = ja % NA, fb = smem_u32(&bars->a_full[buf]);
        // the buffer is free once the last group's MMAs of the tile that used it have completed
        mbar_wait_poll<PROD_POLL_NS>(smem_u32(&bars->a_empty[buf]), ((ja / NA) & 1u) ^ 1u);
        mbar_expect_tx(fb, G::A_TX);
#pragma unroll
        for (int kb = 0; kb < G::NKB; ++kb)
          tma_load_4d(&tmA, fb, sA + buf * G::A_BYTES + (uint32_t)kb * ABOX, kb * 64, ta.x0() - 1, ta.y0() - 1, b);
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== the MMA issue loop ===============================
    const uint32_t idesc1 = make_idesc<__half>(AROWS);
    const uint32_t idescS = make_idesc<__half>(CW);
    const uint64_t wdesc = sw128_desc(sW), xdesc = sw128_desc(sX);
    mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->w_full), 0);
    // MMA1 of one group: D^T[128 channels][192 patch pixels] = W[start_row ..][K] . xn_patch[192][K]^T
    auto issue_group = [&](auto grp_tag, uint32_t gg, uint32_t jg) {
      constexpr int grp = decltype(grp_tag)::value;
      const uint32_t sd = gg & 1u, ab = jg % NA;
      tc_fence_after();
      const uint32_t d = tmem_base + sd * D1_COLS;
      const uint64_t pdesc = sw128_desc(sA + ab * G::A_BYTES);
      // the lone unit's window moves with the tile (the only runtime term of the descriptors)
      const uint64_t wbase = wdesc + (grp == G::NG - 1 ? (uint64_t)((G::lone_start_row((int)(jg & 3u)) * 128) >> 4)
                                                       : (uint64_t)((G::start_row(grp) * 128) >> 4));
#pragma unroll
      for (int kb = 0; kb < G::NKB; ++kb) {
#pragma unroll
        for (int kk = 0; kk < (kb == G::NKB - 1 ? G::KS_LAST : 4); ++kk) {
          if (DBG & 8) continue;
          // descriptors: base + (byte offset >> 4); the whole offset is a compile-time constant
          const uint64_t wd = wbase + (uint64_t)((kb * G::NP * 128 + kk * 32) >> 4);
          const uint64_t pd = pdesc + (uint64_t)((kb * ABOX + kk * 32) >> 4);
          umma_elect<__half>(d, wd, pd, idesc1, (kb > 0 || kk > 0) ? 1u : 0u);
        }
      }
      umma_commit_elect(smem_u32(&bars->d1_full[sd]));
      if (grp == G::NG - 1) umma_commit_elect(smem_u32(&bars->a_empty[ab]));
      __syncwarp();
    };
    auto gram = [&](uint32_t jx) {
      mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->x_ready), jx & 1u);          // the depthwise warps have written tile jx's q | k rows
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
        if (grp == 0) mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->a_full[j % NA]), (j / NA) & 1u);
        mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->d1_empty[gg & 1u]), ((gg >> 1) & 1u) ^ 1u);
        issue_group(grp_tag, gg, j);
        ++gg;
      };
      // group 0 of this tile goes out BEFORE the Gram of the previous one: its accumulator is ready when the depthwise
      // warps come back from the previous tile's last group
      one(std::integral_constant<int, 0>{});
      if (j > 0) gram(j - 1);
      one(std::integral_constant<int, 1>{});
      if (G::NG > 2) one(std::integral_constant<int, G::NG - 1>{});
    }
    if (j > 0) gram(j - 1);
    umma_commit_elect(smem_u32(&bars->acc_done));
  } else {
    // =============================== depthwise 3x3 from tensor memory -> X tile (q, k) / global (v) ===============================
    const int q = warp & 3, h = warp >> 2;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(8 * h);
    // The nine taps of every unit this lane quarter will ever read, parked in the spare TENSOR-MEMORY columns with
    // lane = channel: the kernel's shared memory leaves the L1 no room, and a tap fetch from L2 in front of every group sat
    // on the critical path of the lock-step phases (clock64 instrumentation: ~700 cycles per group outside the waits and the
    // units).  Both warps of a quarter store the same values; each reads them back behind its own tcgen05.wait::st.
    const uint32_t wlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)W_COL0;
#pragma unroll
    for (int g = 0; g < G::NG; ++g) {
      const int u = g < G::NG - 1 ? 4 * g + q : G::NUNITS - 1;
      uint32_t r[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) r[t] = __float_as_uint(__ldg(p.dw + (size_t)(u * 9 + t) * UC + lane));
      tmem_st8(wlane + (uint32_t)(9 * g), r);
      tmem_st1(wlane + (uint32_t)(9 * g + 8), r[8]);
    }
    tmem_st_wait();
    uint32_t gg = 0, j = 0;
    long long t_wait[3] = {0, 0, 0}, t_unit[3] = {0, 0, 0}, t_xw = 0, t_all = 0, t_arr = 0, t_pre = 0, t_mark = 0;   // DBG & 64: where a depthwise warp's time goes
    if (DBG & 64) t_all = clock64();
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int y0 = ti.y0(), x0 = ti.x0();
      const int vy = min(TH, p.H - y0), vx = max(0, min(8, p.W - (x0 + 8 * h)));   // valid output rows / columns of this half
      const bool partial = vy < TH || vx < 8;
      // The Gram MMA of the previous tile must have read the X tile before this tile's first q | k row is stored: the wait
      // sits in front of that store (first output row of the tile's first q | k unit), behind the taps of three patch rows
      bool xwait = true;
#pragma unroll 1
      for (int grp = 0; grp < G::NG; ++grp, ++gg) {
        const int u = unit_of(grp, G::NG, G::NUNITS, q, (int)(j & 3u));
        const bool active = u >= 0 && u < G::NUNITS && !(DBG & 2);
        const int ch = u * UC + lane;                        // this thread's qkv channel
        f2_t w[9];
        if (active) {
          uint32_t r[9];
          tmem_ld8(wlane + (uint32_t)(9 * grp), r);
          tmem_ld1(wlane + (uint32_t)(9 * grp + 8), r[8]);
          tmem_ld_wait9(r);
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = dwt::pack2u(r[t], r[t]);
        }
        const uint32_t sd = gg & 1u;
        long long tw0 = 0;
        if (DBG & 64) { tw0 = clock64(); if (t_mark) t_pre += tw0 - t_mark; }
        mbar_wait_spin(smem_u32(&bars->d1_full[sd]), (gg >> 1) & 1u);
        tc_fence_after();
        if (DBG & 64) { const long long t1 = clock64(); t_wait[grp] += t1 - tw0; tw0 = t1; }
        if (active) {
          const uint32_t taddr = tlane + sd * D1_COLS;
          // PARTIAL (compile time): tiles that cross the right / bottom image edge mask their outputs; every other tile
          // runs the variant without a single predicate or select (the masks and the per-store address arithmetic were as
          // many integer instructions as the v unit has FFMA2 -- and IMAD shares the FMA pipe)
          auto run_unit = [&](auto partial_tag) {
            constexpr bool PARTIAL = decltype(partial_tag)::value;
            if (u * UC < 2 * CW) {
              // ---- q | k unit: row `ch` of the X tile, 8 fp16 pixels per output row; squared norms ----
              f2_t n = dwt::pack2(0.f, 0.f);
              const uint32_t xrow = sX + (uint32_t)ch * 128u + (uint32_t)(h << 4), sw = ((uint32_t)ch & 7u) << 4;
              const uint32_t xr0 = xrow ^ sw, xr1 = (xrow + xbox) ^ sw;      // X boxes of output rows 0-3 / 4-7 (xbox is a multiple of 1024)
              dwt::unit(taddr, w, [&](int oy, const f2_t (&acc)[4]) {
                if (oy == 0 && xwait) {
                  long long tx0 = 0;
                  if (DBG & 64) tx0 = clock64();
                  mbar_wait_spin(smem_u32(&bars->x_empty), (j & 1u) ^ 1u); xwait = false;
                  if (DBG & 64) t_xw += clock64() - tx0;
                }
                if (DBG & 32) { n = dwt::add2(n, dwt::add2(dwt::add2(acc[0], acc[1]), dwt::add2(acc[2], acc[3]))); return; }
                float a[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) dwt::unpack2(acc[e], a[2 * e], a[2 * e + 1]);
                if (PARTIAL) {
                  // output pixels past the right / bottom image edge still see taps from inside the image: they must not
                  // reach the Gram or the norms
#pragma unroll
                  for (int e = 0; e < 8; ++e) if (oy >= vy || e >= vx) a[e] = 0.f;
                }
                const __half2 h0 = f2h2_sat(a[0], a[1]), h1 = f2h2_sat(a[2], a[3]), h2 = f2h2_sat(a[4], a[5]), h3 = f2h2_sat(a[6], a[7]);
                // pixel (oy, 8h + e) of the tile: X box oy / 4, 16-byte chunk (oy % 4) * 2 + h of the channel's 128-byte row
                // (the chunk bits (oy & 3) << 5 and h << 4 never carry into the swizzle's bits, so they XOR in as a constant)
                sts128u((oy < 4 ? xr0 : xr1) ^ (uint32_t)((oy & 3) << 5),
                        make_uint4(h2_bits(h0), h2_bits(h1), h2_bits(h2), h2_bits(h3)));
                // the norms take the unrounded fp32 values: against the fp16 operands the Gram sees, each term differs by a
                // mean-zero rounding error of 2^-11, which averages out over the H * W pixels of the sum
                if (PARTIAL) {
                  n = dwt::fma2(dwt::pack2(a[0], a[1]), dwt::pack2(a[0], a[1]), n);
                  n = dwt::fma2(dwt::pack2(a[2], a[3]), dwt::pack2(a[2], a[3]), n);
                  n = dwt::fma2(dwt::pack2(a[4], a[5]), dwt::pack2(a[4], a[5]), n);
                  n = dwt::fma2(dwt::pack2(a[6], a[7]), dwt::pack2(a[6], a[7]), n);
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) n = dwt::fma2(acc[e], acc[e], n);
                }
              });
              float n0, n1;
              dwt::unpack2(n, n0, n1);
              red[h * (2 * CW) + ch] += n0 + n1;               // this thread is the only writer of the slot
            } else {
              // ---- v unit: fp16 straight to global memory, a warp writes 64 contiguous bytes per pixel ----
              const bool chv = ch < 3 * CW;                    // C = 48: the last unit is half padding
              unsigned short* dst = reinterpret_cast<unsigned short*>(p.v) +
                                    (((long long)b * p.H + y0) * p.W + x0 + 8 * h) * CW + (ch - 2 * CW);
              const long long rs = (long long)p.W * CW;
              f2_t dummy = dwt::pack2(0.f, 0.f);
              dwt::unit(taddr, w, [&](int oy, const f2_t (&acc)[4]) {
                if (DBG & 16) { dummy = dwt::add2(dummy, dwt::add2(dwt::add2(acc[0], acc[1]), dwt::add2(acc[2], acc[3]))); return; }
                if (chv && (!PARTIAL || oy < vy)) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float lo, hi;
                    dwt::unpack2(acc[e], lo, hi);
                    const uint32_t bits = h2_bits(f2h2_sat(lo, hi));
                    // the pixel offsets (2e) * CW are immediates; the row pointer advances once per output row
                    if (!PARTIAL || 2 * e < vx) dst[(2 * e) * CW] = (unsigned short)(bits & 0xffffu);
                    if (!PARTIAL || 2 * e + 1 < vx) dst[(2 * e + 1) * CW] = (unsigned short)(bits >> 16);
                  }
                }
                dst += rs;
              });
              if (DBG & 16) { float d0, d1; dwt::unpack2(dummy, d0, d1); if (d0 + d1 == 123.456f) red[0] = d0; }
            }
          };
          if (partial) run_unit(std::true_type{});
          else run_unit(std::false_type{});
        }
        if (DBG & 64) { t_mark = clock64(); t_unit[grp] += t_mark - tw0; }
        // the group's accumulator is free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->d1_empty[sd]));
        if (grp == G::NG - 2) {
          // every q | k unit lives in the full groups: the X tile is complete, the Gram can go (the lone v unit is still ahead)
          if (!(DBG & 128)) fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->x_ready));
        }
        if (DBG & 64) { const long long t1 = clock64(); t_arr += t1 - t_mark; t_mark = t1; }
      }
    }
    if ((DBG & 64) && lane == 0 && blockIdx.x == 1 && blockIdx.y == 0)
      printf("af-dbg warp %d tiles %u total %lld | wait g0 %lld g1 %lld g2 %lld | unit g0 %lld g1 %lld g2 %lld | xwait %lld\n", warp, j,
             clock64() - t_all, t_wait[0], t_wait[1], t_wait[2], t_unit[0], t_unit[1], t_unit[2], t_xw);
    if ((DBG & 64) && lane == 0 && blockIdx.x == 1 && blockIdx.y == 0) printf("af-dbg2 warp %d arrive %lld pre %lld\n", warp, t_arr, t_pre);
    // ---- squared-norm partials: the two column halves of every q | k channel ----
    asm volatile("bar.sync 1, %0;" ::"n"(DW_WARPS * 32) : "memory");
    if (tid < 2 * CW) {
      const float sum = red[tid] + red[2 * CW + tid];
      // channel tid of q|k -> n_part[b][head][part][which][c]
      const int which = tid >= CW ? 1 : 0, c = tid - which * CW;
      const int chd = CW / p.heads, head = c / chd, cc = c - head * chd;
      p.n_part[((((long long)b * p.heads + head) * p.parts + part) * 2 + which) * chd + cc] = sum;
    }
    // ---- the CTA's partial Gram: TMEM lane == q channel ----
    if (warp < 4) {
      const bool any = TileIter(p).valid();
      const int chd = CW / p.heads;
      const int i = warp * 32 + lane;
      if (any) {
        mbar_wait_spin(smem_u32(&bars->acc_done), 0);
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
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

struct FusedFrontCfg {
  int xrows;
  uint32_t off_a, off_w, off_x, off_red, off_bars;
  size_t smem;
};

template <int CW>
bool configure_w(int heads, FusedFrontCfg& c) {
  using G = Geo<CW>;
  if (heads <= 0 || CW % heads != 0) return false;
  c.xrows = std::max(2 * CW, 128);                      // A reads 128 rows from row 0 (M = 128), B reads C rows from row C
  size_t off = 0;
  // W first: group 1 reads 128 rows from row 128 whatever the image holds (C = 48: 160 rows); what follows must be mapped
  c.off_w = (uint32_t)off; off += align_up((size_t)G::W_BYTES, 1024);
  c.off_a = (uint32_t)off; off += (size_t)NA * G::A_BYTES;
  c.off_x = (uint32_t)off; off += (size_t)2 * c.xrows * 128;
  c.off_red = (uint32_t)off; off += (size_t)2 * 2 * CW * sizeof(float);
  off = align_up(off, 16);
  c.off_bars = (uint32_t)off; off += sizeof(Bars);
  c.smem = off + 1024;          // alignment slack
  return c.smem <= 227 * 1024;
}

bool configure(int C, int heads, FusedFrontCfg& c) {
  return C == 48 ? configure_w<48>(heads, c) : C == 96 ? configure_w<96>(heads, c) : false;
}

template <int CW, int DBG = 0>
int launch_inst(const CUtensorMap& tA, const FusedFrontParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(attn_fused_kernel<CW, DBG>, optin));
  IRB_CUDA(launch_pdl(attn_fused_kernel<CW, DBG>, grid, dim3(NTHREADS), smem, s, tA, p));
  return IR_OK;
}

}  // namespace

bool attn_fused_supported(int C, int heads) {
  FusedFrontCfg c;
  return configure(C, heads, c);
}

int attn_fused_wrows(int C) { return (3 * C + UC - 1) / UC * UC; }

int attn_fused_parts(int B, int H, int W) {
  const int tiles = cdiv(W, TW) * cdiv(H, TH);
  return std::max(1, std::min(tiles, 148 / std::max(1, B)));
}

int launch_attn_fused(const AttnFusedArgs& a, cudaStream_t s) {
  FusedFrontCfg c;
  IRB_REQUIRE(configure(a.C, a.heads, c), "attn_fused: unsupported shape");
  IRB_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0 && a.B <= 65535, "attn_fused: bad extent");
  IRB_REQUIRE(a.xn != nullptr, "attn_fused: the fp16 norm1(x) tensor is required");
  CUtensorMap tA;
  {
    // the fp16 xn halo patch of a tile: one SWIZZLE_128B box of 64 channels x 18 x 10 pixels per K box
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.C * 2 * a.W, (cuuint64_t)a.C * 2 * a.W * a.H};
    cuuint32_t box[4] = {64, PW, TH + 2, 1};
    IRB_TRY(make_tmap(&tA, a.xn, true, 4, d, st, box, true));
  }
  FusedFrontParams p{};
  p.w_qkv = reinterpret_cast<const uint8_t*>(a.w_qkv); p.dw = a.dw_chunked;
  p.v = reinterpret_cast<__half*>(a.v);
  p.s_part = a.s_part; p.n_part = a.n_part;
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.heads = a.heads;
  p.parts = attn_fused_parts(a.B, a.H, a.W);
  IRB_REQUIRE(p.parts == a.parts, "attn_fused: partial count mismatch");
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.tiles_per_img = p.tiles_x * p.tiles_y;
  p.xrows = c.xrows;
  p.off_a = c.off_a; p.off_w = c.off_w; p.off_x = c.off_x; p.off_red = c.off_red; p.off_bars = c.off_bars;
  dim3 grid(p.parts, a.B, 1);
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);     // one CTA per SM
  const double pix = (double)a.B * a.H * a.W;
  // algorithmic bytes: xn read (fp16) + v write (fp16); flops: qkv 1x1 + depthwise + Gram
  ProfScope prof(TAG_ATTN_FUSED, pix * (2.0 * a.C + 2.0 * a.C),
                 pix * (2.0 * 3 * a.C * a.C + 2.0 * 9 * 3 * a.C + 2.0 * a.C * (a.C / a.heads)), s);
#ifdef IRB_FUSED_EXPERIMENTS
  static const int dbg = getenv("IRB_AF_DBG") ? atoi(getenv("IRB_AF_DBG")) : 0;
  if (a.C == 96) {
    switch (dbg) {
      case 2: return launch_inst<96, 2>(tA, p, grid, smem, s);
      case 8: return launch_inst<96, 8>(tA, p, grid, smem, s);
      case 16: return launch_inst<96, 16>(tA, p, grid, smem, s);
      case 32: return launch_inst<96, 32>(tA, p, grid, smem, s);
      case 48: return launch_inst<96, 48>(tA, p, grid, smem, s);
      case 56: return launch_inst<96, 56>(tA, p, grid, smem, s);
      case 64: return launch_inst<96, 64>(tA, p, grid, smem, s);
      case 128: return launch_inst<96, 128>(tA, p, grid, smem, s);
      case 192: return launch_inst<96, 192>(tA, p, grid, smem, s);
      default: break;
    }
  }
#endif
  return a.C == 48 ? launch_inst<48>(tA, p, grid, smem, s) : launch_inst<96>(tA, p, grid, smem, s);
}

}  // namespace irb
