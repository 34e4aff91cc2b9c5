// Fused front of the transposed attention (MDTA, restormer.py:111-124): the depthwise 3x3 over the 3C qkv channels
// (qkv_dwconv :106,:114), the q / k / v split (:115), the Gram matrix q.k^T reduced over H*W (:124) and the squared row
// norms that F.normalize needs (:121-122) in ONE kernel.  q and k never reach HBM (the two-kernel version wrote them and
// read them back: 16C of its 32C bytes per pixel); only v is stored.
//
// One persistent CTA per SM, bound to one image (the Gram accumulates per image); a tile is an 8 x 16 pixel patch:
//
//   H-producer (1 thread)  per 32-channel chunk of qkv: 4-D bulk-tensor load of the (8+2) x (16+2) halo patch (TMA
//                          zero-fill == the conv's zero padding) + the chunk's 3x3 taps
//   dw warps (8)           thread = 2 channels x a 2 x 4 pixel block, packed FFMA2.  q / k chunks are written TRANSPOSED
//                          ([channel][pixel], SWIZZLE_128B, pixels = the MMA's K dimension) into the X tile and their
//                          squares accumulate in registers; v chunks go to a [pixel][channel] staging box -> TMA store
//   MMA (1 thread)         S[i][j] += sum_p q_i[p] k_j[p]: tcgen05.mma with A = the q rows and B = the k rows of the SAME
//                          X tile (M = 128 >= C, N = C, K = 128 pixels), accumulating in TMEM over every tile of the CTA
//   epilogue (4 warps)     once, at the end: TMEM -> per-CTA partial S (per head: the diagonal ch x ch blocks) and the
//                          norm partials, in the layout the softmax/fold kernel already reduces (deterministic order)
#include "attn_front.cuh"
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

#include <type_traits>

// Barrier waits of the depthwise warps and of the MMA issuer.  -DIRB_TAIL_SPIN: spinning / short-poll waits as in the fused
// kernels (a wait with a suspend-time hint that outlasts the hardware window wakes up late); default: suspend-hint waits.
#ifdef IRB_TAIL_SPIN
#define IRB_DW_WAIT(bar, ph) sm100::mbar_wait_spin(bar, ph)
#define IRB_MMA_WAIT(bar, ph) sm100::mbar_wait_poll<32>(bar, ph)
#else
#define IRB_DW_WAIT(bar, ph) sm100::mbar_wait(bar, ph)
#define IRB_MMA_WAIT(bar, ph) sm100::mbar_wait(bar, ph)
#endif

namespace irb {

namespace {

using namespace sm100;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int HPIX = (TH + 2) * (TW + 2);
constexpr int KC = 32;                          // channels per chunk
constexpr int CP = 16, BH = 2, BW = 4;          // dw thread: channel pair x 2x4 pixel block; 16 x 16 = 256 threads
constexpr int EPI_WARPS = 4, DW_WARPS = 8;
constexpr int WARP_H = EPI_WARPS + DW_WARPS, WARP_MMA = WARP_H + 1;
constexpr int NTHREADS = (WARP_MMA + 1) * 32;
constexpr int NST = 4;                          // patch stages (at most; p.nst are used)
constexpr int HDR = 1024;
constexpr int TAPB = 9 * KC * 4;                // bytes of one chunk's taps

struct Bars {
  unsigned long long h_full[NST], h_empty[NST];
  unsigned long long x_ready, x_empty, acc_done;
  uint32_t tmem_base;
};

struct FrontParams {
  const float* dw;         // taps per chunk [nchunk][9][32]
  float* s_part;           // [B][heads][parts][ch][ch]
  float* n_part;           // [B][heads][parts][2][ch]
  int B, H, W, C, heads, parts;   // C, heads: of ONE channel group (blockIdx.z); a group is a whole number of heads
  int Ct, heads_total, ngroups;   // the tensor's channel count / heads; groups = Ct / C (1 for C = 48 / 96, C = 96 above that)
  int nqk, nv;             // chunks of q|k (2C/32) and of v (ceil(C/32))
  int tiles_x, tiles_y, tiles_per_img;
  int xrows;               // rows allocated per X box (>= 128)
  int nst;                 // patch stages in use (2..NST)
  int tmem_cols;
  uint32_t stage_bytes, off_stage, off_x, off_v, off_red;
};

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <typename T> __device__ __forceinline__ f2_t ld2(uint32_t a);
template <> __device__ __forceinline__ f2_t ld2<float>(uint32_t a) {
  f2_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
template <> __device__ __forceinline__ f2_t ld2<__half>(uint32_t a) {
  uint32_t t;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(t) : "r"(a));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&t));
  return pack2(f.x, f.y);
}

struct TileIter {
  int t, step, end, tx_n;
  __device__ TileIter(const FrontParams& p) : t(blockIdx.x), step(gridDim.x), end(p.tiles_per_img), tx_n(p.tiles_x) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int y0() const { return (t / tx_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * TW; }
};

// depthwise 3x3 of one chunk for this thread's 2 channels x (2 x 4) pixels
template <typename TH_>
__device__ __forceinline__ void dw_chunk(uint32_t patch, uint32_t taps, f2_t (&acc)[BH][BW]) {
  constexpr int PXB = KC * (int)sizeof(TH_);           // bytes per pixel in the patch
  f2_t w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = ld2<float>(taps + (uint32_t)(t * KC) * 4u);
#pragma unroll
  for (int iy = 0; iy < BH + 2; ++iy) {
    f2_t v[BW + 2];
#pragma unroll
    for (int ix = 0; ix < BW + 2; ++ix) v[ix] = ld2<TH_>(patch + (uint32_t)(iy * (TW + 2) + ix) * PXB);
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

// TH_: element type of qkv == the Gram's operand type.  NQK: q|k chunks per tile (2C/32: 3 for C = 48, 6 for C = 96).
// VH: v is stored as fp16 (always with fp16 qkv; with fp32 qkv when the attention-output contraction takes fp16 operands:
// the tensor core would round v to a 10-bit mantissa anyway, and the fp16 tensor is half the bytes)
template <typename TH_, int NQK, bool VH>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_front_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmV, const FrontParams p) {
  constexpr bool F32 = std::is_same<TH_, float>::value;
  constexpr int ES = (int)sizeof(TH_);
  constexpr int PXB = KC * ES;                         // patch bytes per pixel (128 fp32 / 64 fp16)
  constexpr int NB = TM * ES / 128;                    // X boxes (128-byte K atoms) per tile: 4 fp32 / 2 fp16
  constexpr int VBOX = TM * KC * ES;                   // v staging box bytes
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  Bars* bars = reinterpret_cast<Bars*>(gbase);
  const uint32_t sST = base + p.off_stage, sX = base + p.off_x, sV = base + p.off_v;
  float* red = reinterpret_cast<float*>(gbase + p.off_red);
  const uint32_t xbox = (uint32_t)p.xrows * 128u;      // bytes of one X box

  // the warp index through a warp reduction lives in a UNIFORM register: the role branches become uniform branches and the
  // code under them uses the uniform datapath (memory descriptors, TMEM / barrier addresses) without one R2UR per use
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));
  const int b = blockIdx.y, part = blockIdx.x, grp = blockIdx.z;
  const int nchunk = p.nqk + p.nv;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bars->h_full[s]), 1);
      mbar_init(smem_u32(&bars->h_empty[s]), DW_WARPS);
    }
    mbar_init(smem_u32(&bars->x_ready), DW_WARPS);
    mbar_init(smem_u32(&bars->x_empty), 1);
    mbar_init(smem_u32(&bars->acc_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_sync();   // set-up done under the previous kernel's tail; from here on global memory is ours (common.cuh)

  if (warp == WARP_H) {
    // =============================== patch / tap producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
      uint32_t s = 0, ph = 0;                              // stage ring position and its phase bit
      for (TileIter ti(p); ti.valid(); ti.next()) {
        const int y0 = ti.y0(), x0 = ti.x0();
        for (int ch = 0; ch < nchunk; ++ch) {
          mbar_wait(smem_u32(&bars->h_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->h_full[s]);
          const uint32_t st = sST + s * p.stage_bytes;
          mbar_expect_tx(fb, (uint32_t)(HPIX * PXB + TAPB));
          // channel of the chunk in the 3*Ct-wide qkv tensor: one group -> q | k | v are consecutive chunks; several
          // groups -> this group's C channels of q, of k, of v (C % 32 == 0 there)
          int chan = ch * KC;
          if (p.ngroups > 1) {
            const int hq = p.nqk / 2;
            chan = ch < hq ? grp * p.C + ch * KC
                 : ch < p.nqk ? p.Ct + grp * p.C + (ch - hq) * KC
                              : 2 * p.Ct + grp * p.C + (ch - p.nqk) * KC;
          }
          tma_load_4d(&tmH, fb, st, chan, x0 - 1, y0 - 1, b);
          bulk_load(st + HPIX * PXB, reinterpret_cast<const uint8_t*>(p.dw) + (size_t)(chan / KC) * TAPB, TAPB, fb);
          if (++s == (uint32_t)p.nst) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== Gram MMA ===============================
    const uint32_t idesc = make_idesc<TH_>(p.C);
    uint32_t j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      IRB_MMA_WAIT(smem_u32(&bars->x_ready), j & 1u);
      tc_fence_after();
      {
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t qa = sX + (uint32_t)a * xbox + kk * 32;              // q rows start at row 0
            const uint32_t ka = qa + (uint32_t)p.C * 128u;                      // k rows start at row C
            umma_elect<TH_>(tmem_base, sw128_desc(qa), sw128_desc(ka), idesc, (j > 0 || a > 0 || kk > 0) ? 1u : 0u);
          }
        umma_commit_elect(smem_u32(&bars->x_empty));
      }
      __syncwarp();
    }
    umma_commit_elect(smem_u32(&bars->acc_done));
  } else if (warp >= EPI_WARPS) {
    // =============================== depthwise 3x3 -> X tile (q, k) / v staging ===============================
    const int ctid = tid - EPI_WARPS * 32;
    const int cp = ctid % CP, blk = ctid / CP;
    const int by = blk / (TW / BW), bx = blk % (TW / BW);
    const uint32_t win0 = (uint32_t)((BH * by) * (TW + 2) + BW * bx) * PXB + (uint32_t)cp * (2 * ES);
    f2_t nrm[NQK];
#pragma unroll
    for (int i = 0; i < NQK; ++i) nrm[i] = pack2(0.f, 0.f);
    uint32_t s = 0, ph = 0, j = 0, vc = 0;
    const int dwi = warp - EPI_WARPS;                       // this warp's pixels: rows 2(dwi/2)..+1, columns 8(dwi%2)..+7
    const uint32_t vwarp = sV + (uint32_t)dwi * (2u * 16u * PXB);
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
      // the MMA of the previous tile has read the X tile (it ran while this CTA produced that tile's v chunks)
      IRB_DW_WAIT(smem_u32(&bars->x_empty), (j & 1u) ^ 1u);
#pragma unroll
      for (int ch = 0; ch < NQK; ++ch) {
        IRB_DW_WAIT(smem_u32(&bars->h_full[s]), ph);
        const uint32_t st = sST + s * p.stage_bytes;
        f2_t acc[BH][BW];
        dw_chunk<TH_>(st + win0, st + HPIX * PXB + (uint32_t)cp * 8u, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        if (++s == (uint32_t)p.nst) { s = 0; ph ^= 1u; }
        // transposed store: rows = channels ch*32 + 2cp (+1), columns = the block's pixels (4 consecutive per row)
        const uint32_t r0 = (uint32_t)(ch * KC + 2 * cp);
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
          if constexpr (F32) {
            a0 = to_tf32(a0); b0 = to_tf32(b0); c0 = to_tf32(c0); d0 = to_tf32(d0);
            a1 = to_tf32(a1); b1 = to_tf32(b1); c1 = to_tf32(c1); d1 = to_tf32(d1);
            // pixel p = (2by+oy)*16 + 4bx: box p/32 = by, 16-byte chunk (p%32)/4 = 4oy + bx
            const uint32_t boxa = sX + (uint32_t)by * xbox;
            const uint32_t c16 = (uint32_t)(4 * oy + bx);
            sts128(boxa + r0 * 128u + ((c16 ^ (r0 & 7u)) << 4), make_float4(a0, b0, c0, d0));
            sts128(boxa + (r0 + 1) * 128u + ((c16 ^ ((r0 + 1) & 7u)) << 4), make_float4(a1, b1, c1, d1));
          } else {
            // fp16: box p/64 = by/2, byte offset in the 128-byte row = (p%64)*2 = ((by&1)*32 + oy*16 + 4bx)*2
            const __half2 h01 = f2h2_sat(a0, b0), h23 = f2h2_sat(c0, d0);
            const __half2 g01 = f2h2_sat(a1, b1), g23 = f2h2_sat(c1, d1);
            // the norms use the rounded operands, like the MMA
            a0 = __low2float(h01); b0 = __high2float(h01); c0 = __low2float(h23); d0 = __high2float(h23);
            a1 = __low2float(g01); b1 = __high2float(g01); c1 = __low2float(g23); d1 = __high2float(g23);
            const uint32_t boxa = sX + (uint32_t)(by >> 1) * xbox;
            const uint32_t off = (uint32_t)(((by & 1) * 32 + oy * 16 + 4 * bx) * 2);
            const uint32_t c16 = off >> 4, sub = off & 15u;
            uint2 t0, t1;
            t0.x = *reinterpret_cast<const uint32_t*>(&h01); t0.y = *reinterpret_cast<const uint32_t*>(&h23);
            t1.x = *reinterpret_cast<const uint32_t*>(&g01); t1.y = *reinterpret_cast<const uint32_t*>(&g23);
            sts64u(boxa + r0 * 128u + ((c16 ^ (r0 & 7u)) << 4) + sub, t0);
            sts64u(boxa + (r0 + 1) * 128u + ((c16 ^ ((r0 + 1) & 7u)) << 4) + sub, t1);
          }
          // squared row norms of the (rounded) operands; pixels outside the image are exact zeros
          f2_t n = nrm[ch];
          n = fma2(pack2(a0, a1), pack2(a0, a1), n); n = fma2(pack2(b0, b1), pack2(b0, b1), n);
          n = fma2(pack2(c0, c1), pack2(c0, c1), n); n = fma2(pack2(d0, d1), pack2(d0, d1), n);
          nrm[ch] = n;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->x_ready));
      // ---- v chunks: each warp stages its own 2 x 8 pixel region ([pixel][channel]) and stores it with its own
      //      bulk-tensor copy: no block-wide barrier on the path ----
      for (int ch = 0; ch < p.nv; ++ch, ++vc) {
        IRB_DW_WAIT(smem_u32(&bars->h_full[s]), ph);
        const uint32_t st = sST + s * p.stage_bytes;
        f2_t acc[BH][BW];
        dw_chunk<TH_>(st + win0, st + HPIX * PXB + (uint32_t)cp * 8u, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        if (++s == (uint32_t)p.nst) { s = 0; ph ^= 1u; }
        const uint32_t vb = vwarp + (vc & 1u) * (16u * PXB);
        if (lane == 0) bulk_wait_read<1>();                  // the store that last read this buffer has drained it
        __syncwarp();
#pragma unroll
        for (int oy = 0; oy < BH; ++oy)
#pragma unroll
          for (int ox = 0; ox < BW; ++ox) {
            float gx, gy;
            unpack2(acc[oy][ox], gx, gy);
            const uint32_t lp = (uint32_t)(oy * 8 + (lane >> 4) * 4 + ox);     // pixel inside the warp's 2 x 8 region
            if constexpr (F32 && !VH) {
              // rounded to tf32: the attention-output contraction feeds v to the tensor core straight from its TMA box
              uint2 t = make_uint2(__float_as_uint(to_tf32(gx)), __float_as_uint(to_tf32(gy)));
              sts64u(vb + lp * 128u + ((((uint32_t)cp >> 1) ^ (lp & 7u)) << 4) + ((uint32_t)cp & 1u) * 8u, t);
            } else {
              const __half2 h = f2h2_sat(gx, gy);
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(vb + lp * 64u + (uint32_t)cp * 4u),
                           "r"(*reinterpret_cast<const uint32_t*>(&h)) : "memory");
            }
          }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmV, vb, grp * p.C + ch * KC, x0 + 8 * (dwi & 1), y0 + 2 * (dwi >> 1), b);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();
    // ---- norm partials: reduce the 16 pixel blocks of every channel in a fixed order ----
#pragma unroll
    for (int ch = 0; ch < NQK; ++ch) {
      float n0, n1;
      unpack2(nrm[ch], n0, n1);
      red[blk * (NQK * KC) + ch * KC + 2 * cp] = n0;
      red[blk * (NQK * KC) + ch * KC + 2 * cp + 1] = n1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DW_WARPS * 32) : "memory");
    if (ctid < 2 * p.C) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sum += red[k * (NQK * KC) + ctid];
      // channel ctid of q|k -> n_part[b][head][part][which][c]
      const int which = ctid >= p.C ? 1 : 0, c = ctid - which * p.C;
      const int chd = p.C / p.heads, head = c / chd, cc = c - head * chd;
      p.n_part[((((long long)b * p.heads_total + grp * p.heads + head) * p.parts + part) * 2 + which) * chd + cc] = sum;
    }
  } else {
    // =============================== epilogue: the CTA's partial Gram ===============================
    TileIter ti(p);
    const bool any = ti.valid();
    const int chd = p.C / p.heads;
    const int i = warp * 32 + lane;                         // TMEM lane == q channel
    if (any) {
      mbar_wait(smem_u32(&bars->acc_done), 0);
      tc_fence_after();
    }
    const int head = i / chd, ii = i - head * chd;
    for (int c0 = 0; c0 < p.C; c0 += 32) {
      float v[32];
      if (any) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
      if (i < p.C) {
        float* dst = p.s_part + ((((long long)b * p.heads_total + grp * p.heads + head) * p.parts + part) * chd + ii) * chd;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int jn = c0 + e - head * chd;                // column inside this head's diagonal block
          if (c0 + e < p.C && jn >= 0 && jn < chd) dst[jn] = v[e];
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

struct FrontCfg { int xrows, nst; uint32_t stage_bytes, off_stage, off_x, off_v, off_red; size_t smem; };

// channels one CTA takes: the whole tensor at C = 48 / 96, groups of 96 channels (whole heads) above that
int group_channels(int C, int heads) {
  if (heads <= 0 || C % heads != 0) return 0;
  if (C == 48 || C == 96) return C;
  const int chd = C / heads;
  return (C % 96 == 0 && 96 % chd == 0) ? 96 : 0;
}

bool configure(int Ct, int heads_total, bool half, FrontCfg& c) {
  const int C = group_channels(Ct, heads_total);        // NQK instantiations (2C/32 = 3, 6); M = 128 >= C
  if (C == 0) return false;
  const int es = half ? 2 : 4;
  const int nb = TM * es / 128;
  c.xrows = std::max(2 * C, 128);                       // A reads 128 rows from row 0 (M = 128), B reads C rows from row C
  const size_t stage = ((size_t)HPIX * KC * es + TAPB + 1023) / 1024 * 1024;
  for (c.nst = NST; c.nst >= 2; --c.nst) {
    size_t off = HDR;
    c.off_stage = (uint32_t)off; off += c.nst * stage;
    c.off_x = (uint32_t)off; off += (size_t)nb * c.xrows * 128;
    c.off_v = (uint32_t)off; off += (size_t)2 * TM * KC * es;
    c.off_red = (uint32_t)off; off += (size_t)16 * 2 * C * 4;
    c.stage_bytes = (uint32_t)stage;
    c.smem = off + 1024;
    if (c.smem <= 227 * 1024) return true;
  }
  return false;
}

template <typename TH_, int NQK, bool VH>
int launch_inst(const CUtensorMap& tH, const CUtensorMap& tV, const FrontParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(attn_front_kernel<TH_, NQK, VH>, optin));
  IRB_CUDA(launch_pdl(attn_front_kernel<TH_, NQK, VH>, grid, dim3(NTHREADS), smem, s, tH, tV, p));
  return IR_OK;
}

}  // namespace

bool attn_front_supported(int C, int heads, bool half) {
  FrontCfg c;
  return configure(C, heads, half, c);
}

int attn_front_parts(int B, int H, int W, int C, int heads) {
  const int tiles = cdiv(W, TW) * cdiv(H, TH);
  const int cg = group_channels(C, heads);
  const int groups = cg > 0 ? C / cg : 1;
  return std::max(1, std::min(tiles, 148 / std::max(1, B * groups)));
}

int launch_attn_front(const AttnFrontArgs& a, cudaStream_t s) {
  FrontCfg c;
  IRB_REQUIRE(configure(a.C, a.heads, a.half != 0, c), "attn_front: unsupported shape");
  IRB_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0 && a.B <= 65535, "attn_front: bad extent");
  const int es = a.half ? 2 : 4;
  CUtensorMap tH, tV;
  {
    cuuint64_t d[4] = {(cuuint64_t)3 * a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)3 * a.C * es, (cuuint64_t)3 * a.C * es * a.W, (cuuint64_t)3 * a.C * es * a.W * a.H};
    cuuint32_t box[4] = {KC, TW + 2, TH + 2, 1};
    IRB_TRY(make_tmap(&tH, a.qkv, a.half != 0, 4, d, st, box, false));
  }
  {
    const bool vh = a.half || a.v_half;
    const int ves = vh ? 2 : 4;
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * ves, (cuuint64_t)a.C * ves * a.W, (cuuint64_t)a.C * ves * a.W * a.H};
    cuuint32_t box[4] = {KC, 8, 2, 1};                 // one dw warp's region
    IRB_TRY(make_tmap(&tV, a.v, vh, 4, d, st, box, !vh));   // fp32 staging rows are 128 B (swizzled), fp16 64 B
  }
  FrontParams p{};
  p.dw = a.dw_chunked; p.s_part = a.s_part; p.n_part = a.n_part;
  const int cg = group_channels(a.C, a.heads);
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = cg; p.heads = cg / (a.C / a.heads);
  p.Ct = a.C; p.heads_total = a.heads; p.ngroups = a.C / cg;
  p.parts = attn_front_parts(a.B, a.H, a.W, a.C, a.heads);
  IRB_REQUIRE(p.parts == a.parts, "attn_front: partial count mismatch");
  p.nqk = 2 * cg / KC; p.nv = cdiv(cg, KC);
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.tiles_per_img = p.tiles_x * p.tiles_y;
  p.xrows = c.xrows; p.nst = c.nst;
  int cols = 32; while (cols < cg) cols <<= 1;
  p.tmem_cols = cols;
  p.stage_bytes = c.stage_bytes; p.off_stage = c.off_stage; p.off_x = c.off_x; p.off_v = c.off_v; p.off_red = c.off_red;
  dim3 grid(p.parts, a.B, p.ngroups);
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);
  const double pix = (double)a.B * a.H * a.W;
  ProfScope prof(TAG_ATTN_FRONT, pix * a.C * (3.0 * es + (a.half || a.v_half ? 2.0 : 4.0)),
                 pix * (2.0 * 9 * 3 * a.C + 2.0 * a.C * (a.C / a.heads)), s);
  if (cg == 48)
    return a.half ? launch_inst<__half, 3, true>(tH, tV, p, grid, smem, s)
         : a.v_half ? launch_inst<float, 3, true>(tH, tV, p, grid, smem, s) : launch_inst<float, 3, false>(tH, tV, p, grid, smem, s);
  return a.half ? launch_inst<__half, 6, true>(tH, tV, p, grid, smem, s)
       : a.v_half ? launch_inst<float, 6, true>(tH, tV, p, grid, smem, s) : launch_inst<float, 6, false>(tH, tV, p, grid, smem, s);
}

}  // namespace irb
