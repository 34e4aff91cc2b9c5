// The whole gated-dconv feed-forward network (GDFN, restormer.py:88-93 and the residual add :148) in ONE kernel, second
// version: the depthwise taps read the hidden patch STRAIGHT OUT OF TENSOR MEMORY (dw_tmem.cuh).
//
//     x[pixel, :] += W_out . ( gelu(dw3x3(W_in . xn)[pixel, 0:hp]) * dw3x3(W_in . xn)[pixel, hp:2hp] )
//
// xn = LayerNorm(x) (norm2) arrives as an fp16 tensor.  As in the first version (git history) project_in is recomputed per
// 8 x 16 pixel tile over the (8+2) x (16+2) halo the depthwise conv needs and the 2*hp-wide hidden tensor never exists in
// HBM.  What changed is the orientation of project_in: D^T[hidden channel][patch pixel] = W_in . xn_patch^T -- 128 rows of
// W_in are the MMA's A operand (128 TMEM lanes), the fp16 xn patch the B operand (N = 192 >= 180 patch pixels = TMEM
// columns).  A depthwise thread owns ONE hidden channel pair (x1[c], x2[c] sit in the same lane of two accumulators) and
// pulls the patch pixels with tcgen05.ld: fp32, in registers, in the order the packed FFMA2 taps consume them.  The first
// version moved D1 through six convert warps (tcgen05.ld -> fp16 -> a 57 KB shared-memory patch) and read it back with
// 36 LDS + 72 conversions per 144 FFMA2; here the hidden tensor is never rounded, never touches shared memory, the gate
// runs in the registers that hold both depthwise results, and the convert warps are gone.
//
// One persistent CTA per SM, 10 warps.  The pipeline unit is a CHUNK of 128 hidden channels:
//
//   producer + MMA (2 warps) per tile: the xn halo patch as [180 px][128 B] SWIZZLE_128B boxes (4-D bulk-tensor load; the
//                         zero fill outside the image IS the conv's zero padding: project_in has no bias, so xn = 0 gives
//                         hidden = 0); per chunk: the 128 x1 rows and the 128 x2 rows of W_in through a two-slot ring, the
//                         chunk's W_out rows; MMA1 (x1 rows) -> accumulator A, MMA1 (x2 rows) -> accumulator B (the two
//                         alternate, so each is recomputed while the depthwise warps read the other); MMA2:
//                         D2[128 px][C] += gated_chunk . W_out_chunk^T
//   depthwise (8 warps)   warp w owns TMEM lane quarter w & 3 (32 hidden channels) and output columns 8 (w >> 2) .. +7.
//                         Phase 1: the taps of x1 into 64 registers.  Phase 2: the taps of x2, and per finished output row
//                         gelu(x1) * x2 (gdfn_math.cuh, one MUFU per element) -> fp16 -> MMA2's operand boxes
//                         ([pixel][channel]: a warp writes 64 contiguous bytes of one pixel row per store).
//                         After phase 1 of a tile's first chunk: the PREVIOUS tile's epilogue, tcgen05.ld of D2 -> swizzled
//                         staging boxes -> bulk-tensor REDUCTION (x += ...; the residual is never loaded)
//
// Channels-last: x fp32 [B][H][W][C], xn fp16 [B][H][W][C]; every access to either is a bulk-tensor copy.
#include "common.cuh"
#include "dw_tmem.cuh"
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
using gdfn::f2_t;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int PW = TW + 2;                     // halo patch width
constexpr int HPIX = (TH + 2) * PW;            // 180 halo pixels
constexpr int AROWS = 192;                     // patch rows the MMA reads (N); rows 180 .. 191 are never written nor read back
constexpr int ABOX = AROWS * 128;              // one 64-channel box of the xn patch
constexpr int GC = 128;                        // hidden channels per chunk = rows of an MMA1 group
constexpr int WGBOX = GC * 128;                // one 64-channel K box of a group's W_in rows
constexpr int OPBOX = TM * 128;                // one 64-channel K box of MMA2's operand; one 32-channel staging group
constexpr int DW_WARPS = 8;
constexpr int WARP_MMA = DW_WARPS, WARP_PROD = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_PROD + 1) * 32;
constexpr int D1_COLS = AROWS;                 // accumulator A at column 0, B at 192
constexpr int D2_COL0 = 2 * D1_COLS;           // D2[128 px][C] at column 384
constexpr int TMEM_COLS = 512;
constexpr int W_COL0 = D2_COL0 + 96;           // depthwise taps, lane = hidden channel: 9 columns per (chunk, x1 | x2) set
constexpr int W_SETS = (TMEM_COLS - W_COL0) / 9;   // three sets fit the 32 spare columns
#ifndef IRB_PROD_POLL_NS
#define IRB_PROD_POLL_NS 200
#endif
#ifndef IRB_MMA_POLL_NS
#define IRB_MMA_POLL_NS 32
#endif
constexpr int PROD_POLL_NS = IRB_PROD_POLL_NS, MMA_POLL_NS = IRB_MMA_POLL_NS;   // sleep between barrier probes of the one-thread roles

struct Bars {
  unsigned long long a_full[2], a_empty[2];
  unsigned long long w1_full[2], w1_empty[2];
  unsigned long long w2_full, w2_empty;
  unsigned long long d1_full[2], d1_empty[2];
  unsigned long long op_ready, op_empty;
  unsigned long long acc_full, acc_empty;
  unsigned long long xs_full, xs_empty;      // NEXT: the tile's residual rows have landed in / been stored from the staging boxes
  uint32_t tmem_base;
};

struct FusedParams {
  const uint8_t* w_in;     // fp16 SWIZZLE_128B image [nkb][2 hp][128 B]
  const uint8_t* w_out;    // [hp / 64][C][128 B]
  const float* dw;         // taps [hp / 64][2][9][64]
  __half* xn_next;         // NEXT: [B][H][W][C] fp16, LayerNorm of the updated residual stream (the next block's norm1)
  const float* lnw;        // NEXT: that LayerNorm's weight / bias
  const float* lnb;
  int ln_mode;
  int B, H, W;
  int tiles_x, tiles_y, ntiles;
  uint32_t off_a, off_win, off_wout, off_op, off_stg, off_bars;
};

template <int CW> struct Geo {
  static constexpr int NKB = (CW + 63) / 64;                 // 64-channel K boxes of xn / W_in
  static constexpr int KS_LAST = (CW - 64 * (NKB - 1)) / 16; // K steps (16 channels) in the last box
  static constexpr int HP = CW <= 48 ? 128 : 256;            // padded hidden width of one GDFN half
  static constexpr int NC = HP / GC;                         // chunks per tile
  static constexpr int NGRP = (CW + 31) / 32;                // 32-channel groups of the output
  static constexpr int NA = CW <= 48 ? 2 : 1;                // xn patch buffers
  static constexpr uint32_t A_BYTES = NKB * ABOX;
  static constexpr uint32_t A_TX = NKB * HPIX * 128;         // bytes one patch load delivers
  static constexpr uint32_t WIN_BYTES = NKB * WGBOX;         // one group of W_in rows
  static constexpr uint32_t WOUT_BYTES = 2 * CW * 128;       // one chunk of W_out rows (two K boxes)
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

__device__ __forceinline__ uint32_t h2_bits(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}

// DBG != 0: timing experiments only (results are garbage; compiled with -DIRB_FUSED_EXPERIMENTS, selected by IRB_FUSED_DBG):
// 2 depthwise warps without TMEM loads / taps, 4 plain product instead of the GELU gate, 8 no MMA instructions,
// 16 no operand stores, 32 W_in loaded only twice, 64 W_out loaded only once, 128 xn patches loaded only NA times
// NEXT: the epilogue LOADS the tile's residual rows (bulk-tensor load into the staging boxes), adds, stores, and also
// writes LayerNorm(x_new) as the fp16 operand tensor of the next block's attention front: the next block's norm1 pass (4C
// read + 2C written per pixel, one launch) disappears.  Without NEXT the residual is reduced in L2 and never loaded.
template <int CW, int DBG, bool NEXT>
__global__ void __launch_bounds__(NTHREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY, const FusedParams p) {
  using G = Geo<CW>;
  constexpr int NC = G::NC, NA = G::NA;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  Bars* bars = reinterpret_cast<Bars*>(gbase + p.off_bars);
  const uint32_t sA = base + p.off_a, sWin = base + p.off_win, sWout = base + p.off_wout, sOP = base + p.off_op,
                 sStg = base + p.off_stg;

  // the warp index through a warp reduction lives in a UNIFORM register: the role branches become uniform branches and the
  // code under them uses the uniform datapath (memory descriptors, TMEM / barrier addresses) without one R2UR per use
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
      mbar_init(smem_u32(&bars->w1_full[s]), 1);
      mbar_init(smem_u32(&bars->w1_empty[s]), 1);
      mbar_init(smem_u32(&bars->d1_full[s]), 1);
      mbar_init(smem_u32(&bars->d1_empty[s]), DW_WARPS);
    }
    mbar_init(smem_u32(&bars->w2_full), 1);
    mbar_init(smem_u32(&bars->w2_empty), 1);
    mbar_init(smem_u32(&bars->op_ready), DW_WARPS);
    mbar_init(smem_u32(&bars->op_empty), 1);
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_empty), NEXT ? 4 : DW_WARPS);
    mbar_init(smem_u32(&bars->xs_full), 1);
    mbar_init(smem_u32(&bars->xs_empty), 4);
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
  pdl_sync();   // set-up done under the previous kernel's tail; from here on global memory is ours (common.cuh)

  uint32_t ntl = 0;
  for (TileIter ti(p); ti.valid(); ti.next()) ++ntl;

  if (warp == WARP_PROD) {
    // =============================== producer: xn patches, W_in groups, W_out chunks ===============================
    // One thread, blocking waits, loads issued in the order their slots retire: the x1 slot of the ring with MMA1 (x1) of
    // the chunk, the x2 slot with MMA1 (x2), the patch buffer with the tile's last MMA1, the W_out buffer with MMA2 of the
    // previous chunk.  Every load then has more than a chunk (~2 us) to land.
    if (lane == 0) {
      const uint32_t NGT = ntl * 2u * NC, NCT = ntl * NC;      // W_in groups / chunks this CTA processes
      TileIter ta(p);
      uint32_t ja = 0, gl = 0, cl = 0;
      auto load_a = [&]() {
        const uint32_t buf = ja % NA, fb = smem_u32(&bars->a_full[buf]);
        mbar_wait_poll<PROD_POLL_NS>(smem_u32(&bars->a_empty[buf]), ((ja / NA) & 1u) ^ 1u);
        if ((DBG & 128) && ja >= (uint32_t)NA) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, G::A_TX);
#pragma unroll
          for (int kb = 0; kb < G::NKB; ++kb)
            tma_load_4d(&tmA, fb, sA + buf * G::A_BYTES + (uint32_t)kb * ABOX, kb * 64, ta.x0() - 1, ta.y0() - 1, ta.img());
        }
        ta.next();
        ++ja;
      };
      auto load_win = [&]() {           // group gl = (chunk c, set) of its tile
        const uint32_t s = gl & 1u, fb = smem_u32(&bars->w1_full[s]);
        const uint32_t gi = gl % (2u * NC), c = gi >> 1, set = gi & 1u;
        mbar_wait_poll<PROD_POLL_NS>(smem_u32(&bars->w1_empty[s]), ((gl >> 1) & 1u) ^ 1u);
        if ((DBG & 32) && gl >= 2) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, G::WIN_BYTES);
          const size_t row0 = (size_t)set * G::HP + (size_t)c * GC;
#pragma unroll
          for (int kb = 0; kb < G::NKB; ++kb)
            bulk_load(sWin + s * G::WIN_BYTES + (uint32_t)kb * WGBOX, p.w_in + ((size_t)kb * 2 * G::HP + row0) * 128, WGBOX, fb);
        }
        ++gl;
      };
      auto load_wout = [&]() {          // chunk cl
        const uint32_t fb = smem_u32(&bars->w2_full);
        mbar_wait_poll<PROD_POLL_NS>(smem_u32(&bars->w2_empty), (cl & 1u) ^ 1u);
        if ((DBG & 64) && cl >= 1) {
          mbar_arrive(fb);
        } else {
          mbar_expect_tx(fb, G::WOUT_BYTES);
          bulk_load(sWout, p.w_out + (size_t)(cl % NC) * G::WOUT_BYTES, G::WOUT_BYTES, fb);
        }
        ++cl;
      };
      // NEXT: residual rows of tile jx into the staging boxes, once the previous tile's stores have read them
      TileIter tx(p);
      uint32_t jx = 0;
      auto load_x = [&]() {
        const uint32_t fb = smem_u32(&bars->xs_full);
        mbar_wait_poll<PROD_POLL_NS>(smem_u32(&bars->xs_empty), (jx & 1u) ^ 1u);
        mbar_expect_tx(fb, (uint32_t)G::NGRP * OPBOX);
        for (int gi = 0; gi < G::NGRP; ++gi)
          for (int q = 0; q < 4; ++q)
            tma_load_4d(&tmY, fb, sStg + (uint32_t)gi * OPBOX + (uint32_t)q * 4096u, gi * 32, tx.x0(), tx.y0() + 2 * q, tx.img());
        tx.next();
        ++jx;
      };
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      if (NEXT) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
      for (int k = 0; k < NA && ta.valid(); ++k) load_a();
      for (int k = 0; k < 2 && gl < NGT; ++k) load_win();
      for (uint32_t cc = 0; cc < NCT; ++cc) {
        if (gl < NGT) load_win();
        if (gl < NGT) load_win();
        if (cc % NC == NC - 1 && ta.valid()) load_a();
        load_wout();
        if (NEXT && cc % NC == 0) load_x();                  // (tile jx's epilogue runs a tile later)
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc1 = make_idesc<__half>(AROWS), idesc2 = make_idesc<__half>(CW);
    const uint64_t adesc0 = sw128_desc(sA), wdesc0 = sw128_desc(sWin), opdesc = sw128_desc(sOP), wodesc = sw128_desc(sWout);
    uint32_t gg = 0;
    // MMA1 of group gg (tile j, first group of the tile when `first`, last when `last`) into accumulator gg & 1
    auto issue1 = [&](uint32_t j, bool first, bool last) {
      const uint32_t s = gg & 1u, ph = (gg >> 1) & 1u, ab = j % NA;
      if (first) mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->a_full[ab]), (j / NA) & 1u);
      mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->w1_full[s]), ph);
      mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->d1_empty[s]), ph ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + s * D1_COLS;
      const uint64_t pd0 = adesc0 + (uint64_t)((ab * G::A_BYTES) >> 4), wd0 = wdesc0 + (uint64_t)((s * G::WIN_BYTES) >> 4);
#pragma unroll
      for (int kb = 0; kb < G::NKB; ++kb)
#pragma unroll
        for (int kk = 0; kk < (kb == G::NKB - 1 ? G::KS_LAST : 4); ++kk) {
          if (DBG & 8) continue;
          umma_elect<__half>(d, wd0 + (uint64_t)((kb * WGBOX + kk * 32) >> 4), pd0 + (uint64_t)((kb * ABOX + kk * 32) >> 4), idesc1,
                             (kb > 0 || kk > 0) ? 1u : 0u);
        }
      umma_commit_elect(smem_u32(&bars->d1_full[s]));
      umma_commit_elect(smem_u32(&bars->w1_empty[s]));
      if (last) umma_commit_elect(smem_u32(&bars->a_empty[ab]));
      __syncwarp();
      ++gg;
    };
    // MMA2 of chunk cc (chunk c of tile j)
    auto issue2 = [&](uint32_t cc) {
      const uint32_t c = cc % NC, j = cc / NC;
      if (c == 0) mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->acc_empty), (j & 1u) ^ 1u);
      mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->w2_full), cc & 1u);
      mbar_wait_poll<MMA_POLL_NS>(smem_u32(&bars->op_ready), cc & 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + D2_COL0;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (DBG & 8) continue;
          umma_elect<__half>(d, opdesc + (uint64_t)((kb * OPBOX + kk * 32) >> 4), wodesc + (uint64_t)((kb * CW * 128 + kk * 32) >> 4),
                             idesc2, (c > 0 || kb > 0 || kk > 0) ? 1u : 0u);
        }
      umma_commit_elect(smem_u32(&bars->op_empty));
      umma_commit_elect(smem_u32(&bars->w2_empty));
      if (c == NC - 1) umma_commit_elect(smem_u32(&bars->acc_full));
      __syncwarp();
    };
    // Issue order per chunk: MMA1 (x1) as soon as accumulator A retires (end of the previous chunk's phase 1), MMA1 (x2) as
    // soon as B retires (end of its phase 2) -- the depthwise warps wait for nothing else -- and only then MMA2 of the
    // previous chunk, whose operand became ready at that same moment and whose result nobody needs for another phase.
    uint32_t cc = 0;
    for (uint32_t j = 0; j < ntl; ++j)
      for (int c = 0; c < NC; ++c, ++cc) {
        issue1(j, c == 0, false);                            // x1 rows -> accumulator A
        issue1(j, false, c == NC - 1);                       // x2 rows -> accumulator B
        if (cc > 0) issue2(cc - 1);
      }
    if (cc > 0) issue2(cc - 1);
  } else {
    // =============================== depthwise 3x3 + gate from tensor memory; epilogue ===============================
    const int q = warp & 3, h = warp >> 2;
    const int lc = q * 32 + lane;                            // hidden channel inside a chunk
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(8 * h);
    // operand box addressing of this thread's channel: K box lc / 64, 16-byte chunk (k / 8) ^ (pixel & 7) of the pixel's row
    const uint32_t k = (uint32_t)lc & 63u, c16 = k >> 3;
    const uint32_t opbase = sOP + ((uint32_t)lc >> 6) * OPBOX + (k & 7u) * 2u + (uint32_t)(8 * h) * 128u;
    uint32_t xo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) xo[e] = opbase + (((c16 ^ (uint32_t)e) << 4) + (uint32_t)e * 128u);
    const uint32_t lsw = (uint32_t)(lane & 7);
    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    // The taps this thread's channel needs in every phase, parked in the spare tensor-memory columns (lane = channel): a tap
    // fetch through the L1 this kernel's shared memory leaves no room for sat in front of every lock-step phase.  Three of
    // the 2 NC sets fit: all of them at C = 48.  C = 96 (four sets) keeps the global loads -- three sets from tensor memory and
    // one from global memory measured SLOWER than four from global memory (0.81 against 0.75 ms), and eight taps of every set in
    // tensor memory with the ninth in a register no faster (0.74 against 0.73 ms): at C = 96 the tap fetch is not what a phase waits for.
    const uint32_t wlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)W_COL0;
    auto tap_ptr = [&](int set) {     // set = 2 * chunk + (0: x1, 1: x2); layout [64-channel chunk][x1 | x2][9][64]
      return p.dw + ((size_t)(((set >> 1) * 2 + (lc >> 6)) * 2 + (set & 1)) * 9) * 64 + (lc & 63);
    };
    constexpr bool W_TMEM = 2 * NC <= W_SETS;
#pragma unroll
    for (int set = 0; W_TMEM && set < 2 * NC; ++set) {
      uint32_t r[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) r[t] = __float_as_uint(__ldg(tap_ptr(set) + t * 64));
      tmem_st8(wlane + (uint32_t)(9 * set), r);
      tmem_st1(wlane + (uint32_t)(9 * set + 8), r[8]);
    }
    if (W_TMEM) tmem_st_wait();
    auto load_taps = [&](int set, f2_t (&w)[9]) {
      if (W_TMEM) {
        uint32_t r[9];
        tmem_ld8(wlane + (uint32_t)(9 * set), r);
        tmem_ld1(wlane + (uint32_t)(9 * set + 8), r[8]);
        tmem_ld_wait9(r);
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = dwt::pack2u(r[t], r[t]);
      } else {
        const float* taps = tap_ptr(set);
#pragma unroll
        for (int t = 0; t < 9; ++t) { const float f = __ldg(taps + t * 64); w[t] = gdfn::pack2(f, f); }
      }
    };

    // D2 of tile j -> staging -> x += : lanes = the 32 pixels of tile rows 2q, 2q+1; this warp's 32-channel groups
    auto epilogue = [&](uint32_t j, int b, int y0, int x0) {
      if (NEXT) {
        // ---- x_new = x + ffn for this quarter's 32 pixels (all C channels per lane), stored back; norm1 of the next block ----
        if (h != (int)(j & 1u)) return;                      // the two warps of a quarter take turns
        mbar_wait_spin(smem_u32(&bars->acc_full), j & 1u);
        tc_fence_after();
        mbar_wait_spin(smem_u32(&bars->xs_full), j & 1u);    // the residual rows have landed in the staging boxes
        float xr[G::NGRP * 32];
#pragma unroll
        for (int gi = 0; gi < G::NGRP; ++gi) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + D2_COL0 + (uint32_t)(gi * 32), v);
          tmem_ld_wait();
          const uint32_t stg = sStg + (uint32_t)gi * OPBOX + (uint32_t)q * 4096u + (uint32_t)lane * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = stg + (((uint32_t)c ^ lsw) << 4);
            const float4 x4 = lds128(a);
            const float4 o = make_float4(x4.x + v[4 * c], x4.y + v[4 * c + 1], x4.z + v[4 * c + 2], x4.w + v[4 * c + 3]);
            sts128(a, o);
            xr[gi * 32 + 4 * c] = o.x; xr[gi * 32 + 4 * c + 1] = o.y; xr[gi * 32 + 4 * c + 2] = o.z; xr[gi * 32 + 4 * c + 3] = o.w;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty));
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          for (int gi = 0; gi < G::NGRP; ++gi)
            tma_store_4d(&tmY, sStg + (uint32_t)gi * OPBOX + (uint32_t)q * 4096u, gi * 32, x0, y0 + 2 * q, b);
          bulk_commit();
        }
        // LayerNorm over the lane's CW channels (two-pass statistics in registers, restormer.py:25-57)
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < CW; ++i) sum += xr[i];
        const float mu = sum * (1.0f / CW);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < CW; ++i) { const float d = xr[i] - mu; ss = fmaf(d, d, ss); }
        const float rstd = 1.0f / sqrtf(ss * (1.0f / CW) + 1e-5f);
        const bool wb = p.ln_mode == LN_WITHBIAS;
        const float sub = wb ? mu : 0.f;
        const int py = y0 + 2 * q + (lane >> 4), px = x0 + (lane & 15);
        if (py < p.H && px < p.W) {
          uint4* dst = reinterpret_cast<uint4*>(p.xn_next + (((long long)b * p.H + py) * p.W + px) * CW);
#pragma unroll
          for (int c8 = 0; c8 < CW / 8; ++c8) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.lnw) + 2 * c8), w1 = __ldg(reinterpret_cast<const float4*>(p.lnw) + 2 * c8 + 1);
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (wb) { b0 = __ldg(reinterpret_cast<const float4*>(p.lnb) + 2 * c8); b1 = __ldg(reinterpret_cast<const float4*>(p.lnb) + 2 * c8 + 1); }
            const float* xv = xr + 8 * c8;
            uint4 u;
            u.x = h2_bits(f2h2_sat(fmaf((xv[0] - sub) * rstd, w0.x, b0.x), fmaf((xv[1] - sub) * rstd, w0.y, b0.y)));
            u.y = h2_bits(f2h2_sat(fmaf((xv[2] - sub) * rstd, w0.z, b0.z), fmaf((xv[3] - sub) * rstd, w0.w, b0.w)));
            u.z = h2_bits(f2h2_sat(fmaf((xv[4] - sub) * rstd, w1.x, b1.x), fmaf((xv[5] - sub) * rstd, w1.y, b1.y)));
            u.w = h2_bits(f2h2_sat(fmaf((xv[6] - sub) * rstd, w1.z, b1.z), fmaf((xv[7] - sub) * rstd, w1.w, b1.w)));
            dst[c8] = u;
          }
        }
        if (lane == 0) {
          bulk_wait_read<0>();                               // the stores have read the staging boxes
          mbar_arrive(smem_u32(&bars->xs_empty));
        }
        __syncwarp();
        return;
      }
      mbar_wait_spin(smem_u32(&bars->acc_full), j & 1u);
      tc_fence_after();
      if (lane == 0) bulk_wait_read<0>();                    // the previous tile's reductions have read the staging boxes
      __syncwarp();
      const int g0 = h == 0 ? 0 : (G::NGRP + 1) / 2, g1 = h == 0 ? (G::NGRP + 1) / 2 : G::NGRP;
      for (int gi = g0; gi < g1; ++gi) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + D2_COL0 + (uint32_t)(gi * 32), v);
        tmem_ld_wait();
        const uint32_t stg = sStg + (uint32_t)gi * OPBOX + (uint32_t)q * 4096u + (uint32_t)lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(stg + (((uint32_t)c ^ lsw) << 4), make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty));
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        for (int gi = g0; gi < g1; ++gi)
          tma_reduce_add_4d(&tmY, sStg + (uint32_t)gi * OPBOX + (uint32_t)q * 4096u, gi * 32, x0, y0 + 2 * q, b);
        bulk_commit();
      }
      __syncwarp();
    };

    uint32_t cc = 0, j = 0;
    int pb = 0, py0 = 0, px0 = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
#pragma unroll 1
      for (int c = 0; c < NC; ++c, ++cc) {
        const uint32_t ph = cc & 1u;
        f2_t r1[8][4];
        // ---- phase 1: depthwise taps of x1 (accumulator A) into registers ----
        {
          f2_t w[9];
          load_taps(2 * c, w);
          mbar_wait_spin(smem_u32(&bars->d1_full[0]), ph);
          tc_fence_after();
          if (!(DBG & 2)) {
            dwt::unit(tlane, w, [&](int oy, const f2_t (&acc)[4]) {
#pragma unroll
              for (int e = 0; e < 4; ++e) r1[oy][e] = acc[e];
            });
          } else {
#pragma unroll
            for (int oy = 0; oy < 8; ++oy)
#pragma unroll
              for (int e = 0; e < 4; ++e) r1[oy][e] = w[e];
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->d1_empty[0]));
        }
        // the previous tile's output, while this tile's x2 accumulator is (long) ready and its MMA2 has long retired
        // (NEXT: behind phase 2 instead -- the epilogue then holds the pixel's whole row in registers, r1 must be dead)
        if (!NEXT && c == 0 && j > 0) epilogue(j - 1, pb, py0, px0);
        // ---- phase 2: depthwise taps of x2 (accumulator B), gate, fp16 operand rows ----
        {
          f2_t w[9];
          load_taps(2 * c + 1, w);
          mbar_wait_spin(smem_u32(&bars->d1_full[1]), ph);
          tc_fence_after();
          mbar_wait_spin(smem_u32(&bars->op_empty), ph ^ 1u);     // MMA2 of the previous chunk has read the operand boxes
          auto emit2 = [&](int oy, const f2_t (&acc)[4]) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const f2_t g = (DBG & 4) ? gdfn::mul2(r1[oy][e], acc[e]) : gdfn::gelu_gate2e(r1[oy][e], acc[e]);
              float lo, hi;
              gdfn::unpack2(g, lo, hi);
              const uint32_t bits = h2_bits(f2h2_sat(lo, hi));
              if (!(DBG & 16)) {
                sts16(xo[2 * e] + (uint32_t)(oy * 2048), bits & 0xffffu);
                sts16(xo[2 * e + 1] + (uint32_t)(oy * 2048), bits >> 16);
              } else if (lo == 123.456f) {
                sts16(xo[0], bits);
              }
            }
          };
          if (!(DBG & 2)) {
            dwt::unit(tlane + D1_COLS, w, emit2);
          } else {
#pragma unroll
            for (int oy = 0; oy < 8; ++oy) { f2_t a4[4] = {w[0], w[1], w[2], w[3]}; emit2(oy, a4); }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->d1_empty[1]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->op_ready));
        }
        if (NEXT && c == 0 && j > 0) epilogue(j - 1, pb, py0, px0);
      }
      pb = ti.img(); py0 = ti.y0(); px0 = ti.x0();
    }
    if (j > 0) epilogue(j - 1, pb, py0, px0);
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

struct FusedCfg {
  uint32_t off_a, off_win, off_wout, off_op, off_stg, off_bars;
  size_t smem;
};

template <int CW>
bool configure_w(int hp, FusedCfg& c) {
  using G = Geo<CW>;
  if (hp != G::HP) return false;
  size_t off = 0;
  c.off_a = (uint32_t)off; off += (size_t)G::NA * G::A_BYTES;
  c.off_win = (uint32_t)off; off += 2 * (size_t)G::WIN_BYTES;
  c.off_wout = (uint32_t)off; off += align_up((size_t)G::WOUT_BYTES, 1024);
  c.off_op = (uint32_t)off; off += 2 * (size_t)OPBOX;
  c.off_stg = (uint32_t)off; off += (size_t)G::NGRP * OPBOX;
  off = align_up(off, 16);
  c.off_bars = (uint32_t)off; off += sizeof(Bars);
  c.smem = off + 1024;          // alignment slack
  return c.smem <= 227 * 1024;
}

bool configure(int C, int hp, FusedCfg& c) {
  return C == 48 ? configure_w<48>(hp, c) : C == 96 ? configure_w<96>(hp, c) : false;
}

template <int CW, int DBG = 0, bool NEXT = false>
int launch_inst(const CUtensorMap& tA, const CUtensorMap& tY, const FusedParams& p, int grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(ffn_fused_kernel<CW, DBG, NEXT>, optin));
  IRB_CUDA(launch_pdl(ffn_fused_kernel<CW, DBG, NEXT>, dim3(grid), dim3(NTHREADS), smem, s, tA, tY, p));
  return IR_OK;
}

}  // namespace

bool ffn_fused_supported(int C, int hp) {
  FusedCfg c;
  return configure(C, hp, c);
}

int launch_ffn_fused(const FfnFusedArgs& a, cudaStream_t s) {
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
  p.xn_next = reinterpret_cast<__half*>(a.xn_next); p.lnw = a.ln_w_next; p.lnb = a.ln_b_next; p.ln_mode = a.ln_mode_next;
  IRB_REQUIRE(a.xn_next == nullptr || (a.ln_w_next != nullptr && (a.ln_mode_next == LN_BIASFREE ||
              (a.ln_mode_next == LN_WITHBIAS && a.ln_b_next != nullptr))), "ffn_fused: bad LayerNorm of the next block");
  p.B = a.B; p.H = a.H; p.W = a.W;
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.ntiles = p.tiles_x * p.tiles_y * a.B;
  p.off_a = c.off_a; p.off_win = c.off_win; p.off_wout = c.off_wout; p.off_op = c.off_op; p.off_stg = c.off_stg;
  p.off_bars = c.off_bars;
  const int grid = std::max(1, std::min(p.ntiles, 148));
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);     // one CTA per SM: the kernel owns all 512 TMEM columns
  const double pix = (double)a.B * a.H * a.W;
  // algorithmic bytes: xn read (fp16) + x read-modify-write (fp32); flops: project_in + depthwise + project_out
  ProfScope prof(TAG_FFN_FUSED, pix * (2.0 * a.C + 8.0 * a.C + (a.xn_next ? 2.0 * a.C : 0.0)), pix * (4.0 * a.hp * a.C + 36.0 * a.hp + 2.0 * a.hp * a.C), s);
#ifdef IRB_FUSED_EXPERIMENTS
  static const int dbg = getenv("IRB_FUSED_DBG") ? atoi(getenv("IRB_FUSED_DBG")) : 0;
  if (a.C == 96) {
    switch (dbg) {
      case 2: return launch_inst<96, 2>(tA, tY, p, grid, smem, s);
      case 4: return launch_inst<96, 4>(tA, tY, p, grid, smem, s);
      case 8: return launch_inst<96, 8>(tA, tY, p, grid, smem, s);
      case 16: return launch_inst<96, 16>(tA, tY, p, grid, smem, s);
      case 20: return launch_inst<96, 20>(tA, tY, p, grid, smem, s);
      case 22: return launch_inst<96, 22>(tA, tY, p, grid, smem, s);
      case 32: return launch_inst<96, 32>(tA, tY, p, grid, smem, s);
      case 96: return launch_inst<96, 96>(tA, tY, p, grid, smem, s);
      case 224: return launch_inst<96, 224>(tA, tY, p, grid, smem, s);
      case 246: return launch_inst<96, 246>(tA, tY, p, grid, smem, s);
      default: break;
    }
  }
#endif
  if (a.xn_next != nullptr)
    return a.C == 48 ? launch_inst<48, 0, true>(tA, tY, p, grid, smem, s) : launch_inst<96, 0, true>(tA, tY, p, grid, smem, s);
  return a.C == 48 ? launch_inst<48>(tA, tY, p, grid, smem, s) : launch_inst<96>(tA, tY, p, grid, smem, s);
}

}  // namespace irb
