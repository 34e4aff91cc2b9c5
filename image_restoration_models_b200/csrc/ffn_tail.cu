// Fused tail of the gated-dconv feed-forward network (GDFN, restormer.py:88-93 and the residual add :148):
//
//     x[pixel, :] += W_out . ( gelu(dw3x3(hidden)[pixel, 0:hp]) * dw3x3(hidden)[pixel, hp:2hp] )
//
// i.e. FeedForward.dwconv (:83,:90), the gate (:91), project_out (:86,:92) and `x + ffn(...)` in ONE kernel: the
// gated tensor never exists in HBM (it was written and re-read by the two-kernel version: 8*hp bytes per pixel,
// a quarter of the whole block's traffic), and the depthwise convolution becomes the producer stage of a tcgen05
// contraction.  One persistent CTA per SM; a tile is an 8 x 16 pixel patch (= the 128 rows of one MMA):
//
//   H-producer (1 thread)  per hidden-channel chunk (32 fp32 / 64 fp16 channels = 128 B per pixel): two 4-D bulk-tensor
//                          loads of the (8+2) x (16+2) halo patch (x1 and x2 halves; TMA zero-fills the image border,
//                          which IS the conv's zero padding), plus the chunk's project_out rows and 3x3 taps
//   dw warps (8)           thread = 2 channels x a 2x4 / 4x4 pixel block: 3x3 taps from the smem halo patch, exact-erf
//                          GELU gate, result stored as the chunk's [128 px][128 B] SWIZZLE_128B operand box
//   MMA (1 thread)         tcgen05.mma (tf32 / f16) of the operand box with the chunk's W_out rows, accumulating over
//                          the hp/32 (hp/64) chunks in TMEM (double buffered across tiles)
//   epilogue (4 warps)     tcgen05.ld + residual (TMA-loaded patch of x, updated in place in smem) -> TMA store
#include "common.cuh"
#include "ffn_tail.cuh"
#include "gdfn_math.cuh"
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
using namespace gdfn;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int HPIX = (TH + 2) * (TW + 2);      // 180 halo pixels
constexpr int HBOX = HPIX * 128;               // bytes of one halo box (one half of one chunk)
constexpr int OPBOX = TM * 128;                // operand / residual box
constexpr int WBOX = 32 * 128;
constexpr int EPI_WARPS = 4;
constexpr int NST = 2, NOP = 2, MAX_RB = 4;
constexpr int HDR = 1024;

struct Bars {
  unsigned long long h_full[NST], h_empty[NST];
  unsigned long long op_ready[NOP], op_empty[NOP];
  unsigned long long acc_full[2], acc_empty[2];
  unsigned long long r_full[MAX_RB], r_empty[MAX_RB];
  uint32_t tmem_base;
};

struct FfnTailParams {
  const uint8_t* w;        // project_out, SWIZZLE_128B image [hp/KC][C][128 B]
  const float* dw;         // depthwise taps per chunk: [hp/KC][2 halves][9 taps][KC]
  const float* bias;       // project_out bias [C] or null
  int B, H, W, C, hp, nchunk;
  int tiles_x, tiles_y, ntiles;
  int RB, nacc, acc_stride, tmem_cols;
  uint32_t stage_bytes, off_stage, off_op, off_r;
  int round_out;
};

// A dw thread owns 2 channels x a BH x BW pixel block of the patch: every staged value is read (BH+2)(BW+2)/(BH BW)
// times and the taps once per block.  CP channel pairs per chunk x (8/BH)(16/BW) blocks = 256 dw threads.
template <typename TH_> struct Geo;
template <> struct Geo<float>  { static constexpr int CP = 16, KC = 32, EB = 8, BH = 2, BW = 4, DW_WARPS = 8; };   // EB: bytes of 2 channels
template <> struct Geo<__half> { static constexpr int CP = 32, KC = 64, EB = 4, BH = 4, BW = 4, DW_WARPS = 8; };

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
  int t, step, end, tx_n, ty_n;
  __device__ TileIter(const FfnTailParams& p) : t(blockIdx.x), step(gridDim.x), end(p.ntiles), tx_n(p.tiles_x), ty_n(p.tiles_y) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int img() const { return t / (tx_n * ty_n); }
  __device__ int y0() const { return ((t / tx_n) % ty_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * TW; }
};

// TH_: element type of the hidden tensor == tensor-core operand type (float -> tf32, __half -> f16)
template <typename TH_>
__global__ void __launch_bounds__((EPI_WARPS + Geo<TH_>::DW_WARPS + 3) * 32, 1)
ffn_tail_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmY, const FfnTailParams p) {
  using G = Geo<TH_>;
  constexpr int KC = G::KC, CP = G::CP, DW_WARPS = G::DW_WARPS;
  constexpr int WARP_H = EPI_WARPS + DW_WARPS, WARP_MMA = WARP_H + 1, WARP_R = WARP_H + 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (base - smem_u32(smem_raw)));
  const uint32_t sST = base + p.off_stage, sOP = base + p.off_op, sR = base + p.off_r;
  const uint32_t wbytes = (uint32_t)p.C * 128u, dwbytes = 2u * 9u * KC * 4u;

  // the warp index through a warp reduction lives in a UNIFORM register: the role branches become uniform branches and the
  // code under them uses the uniform datapath (memory descriptors, TMEM / barrier addresses) without one R2UR per use
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bars->h_full[s]), 1);
      mbar_init(smem_u32(&bars->h_empty[s]), DW_WARPS + 1);      // the dw warps have read the patch + the MMA has read W
    }
    for (int s = 0; s < NOP; ++s) {
      mbar_init(smem_u32(&bars->op_ready[s]), DW_WARPS);
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
    // =============================== hidden / weight producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmH)) : "memory");
      uint32_t it = 0;
      for (TileIter ti(p); ti.valid(); ti.next()) {
        const int b = ti.img(), y0 = ti.y0(), x0 = ti.x0();
        for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
          const uint32_t s = it % NST, ph = (it / NST) & 1u;
          mbar_wait(smem_u32(&bars->h_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->h_full[s]);
          const uint32_t st = sST + s * p.stage_bytes;
          mbar_expect_tx(fb, 2u * HBOX + wbytes + dwbytes);
          tma_load_4d(&tmH, fb, st, ch * KC, x0 - 1, y0 - 1, b);
          tma_load_4d(&tmH, fb, st + HBOX, p.hp + ch * KC, x0 - 1, y0 - 1, b);
          bulk_load(st + 2 * HBOX, p.w + (size_t)ch * wbytes, wbytes, fb);
          bulk_load(st + 2 * HBOX + wbytes, reinterpret_cast<const uint8_t*>(p.dw) + (size_t)ch * dwbytes, dwbytes, fb);
        }
      }
    }
  } else if (warp == WARP_R) {
    // =============================== residual producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
      const int ngroups = (p.C + 31) / 32;
      uint32_t gc = 0;
      for (TileIter ti(p); ti.valid(); ti.next()) {
        const int b = ti.img(), y0 = ti.y0(), x0 = ti.x0();
        for (int g = 0; g < ngroups; ++g, ++gc) {
          const uint32_t s = gc % (uint32_t)p.RB, ph = (gc / (uint32_t)p.RB) & 1u;
          mbar_wait(smem_u32(&bars->r_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->r_full[s]);
          mbar_expect_tx(fb, OPBOX);
          tma_load_4d(&tmR, fb, sR + s * OPBOX, g * 32, x0, y0, b);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = make_idesc<TH_>(p.C);
    uint32_t it = 0, j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const uint32_t slot = p.nacc == 2 ? (j & 1u) : 0u;
      const uint32_t use = p.nacc == 2 ? (j >> 1) : j;
      IRB_MMA_WAIT(smem_u32(&bars->acc_empty[slot]), (use & 1u) ^ 1u);
      tc_fence_after();
      for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
        const uint32_t s = it % NST, o = it % NOP;
        IRB_MMA_WAIT(smem_u32(&bars->h_full[s]), (it / NST) & 1u);          // the chunk's W_out rows (already there)
        IRB_MMA_WAIT(smem_u32(&bars->op_ready[o]), (it / NOP) & 1u);
        tc_fence_after();
        {
          const uint32_t a_addr = sOP + o * OPBOX;
          const uint32_t w_addr = sST + s * p.stage_bytes + 2 * HBOX;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_elect<TH_>(tmem_base + slot * (uint32_t)p.acc_stride, sw128_desc(a_addr + kk * 32),
                            sw128_desc(w_addr + kk * 32), idesc, (ch > 0 || kk > 0) ? 1u : 0u);
          umma_commit_elect(smem_u32(&bars->op_empty[o]));
          umma_commit_elect(smem_u32(&bars->h_empty[s]));
          if (ch == p.nchunk - 1) umma_commit_elect(smem_u32(&bars->acc_full[slot]));
        }
        __syncwarp();
      }
    }
  } else if (warp >= EPI_WARPS) {
    // =============================== depthwise 3x3 + gate -> operand box ===============================
    constexpr int BH = G::BH, BW = G::BW;
    const int ctid = tid - EPI_WARPS * 32;
    const int cp = ctid % CP, blk = ctid / CP;
    const int by = blk / (TW / BW), bx = blk % (TW / BW);
    const uint32_t win0 = (uint32_t)((BH * by) * (TW + 2) + BW * bx) * 128u + (uint32_t)cp * G::EB;
    uint32_t it = 0;
    for (TileIter ti(p); ti.valid(); ti.next()) {
      for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
        const uint32_t s = it % NST, o = it % NOP;
        IRB_DW_WAIT(smem_u32(&bars->h_full[s]), (it / NST) & 1u);
        const uint32_t st = sST + s * p.stage_bytes;
        const uint32_t dws = st + 2 * HBOX + wbytes + (uint32_t)cp * 8u;
        f2_t acc[2][BH][BW];
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          f2_t w[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = ld2<float>(dws + (uint32_t)((set * 9 + t) * KC) * 4u);
          const uint32_t src = st + set * HBOX + win0;
#pragma unroll
          for (int iy = 0; iy < BH + 2; ++iy) {
            f2_t v[BW + 2];
#pragma unroll
            for (int ix = 0; ix < BW + 2; ++ix) v[ix] = ld2<TH_>(src + (uint32_t)(iy * (TW + 2) + ix) * 128u);
#pragma unroll
            for (int oy = 0; oy < BH; ++oy) {
              const int ky = iy - oy;
              if (ky < 0 || ky > 2) continue;
#pragma unroll
              for (int ox = 0; ox < BW; ++ox) {
                // the first tap that reaches an output (ky = 0, kx = 0) initialises it: no zero-fill pass
                if (ky == 0) acc[set][oy][ox] = mul2(w[0], v[ox]);
                else acc[set][oy][ox] = fma2(w[ky * 3], v[ox], acc[set][oy][ox]);
                acc[set][oy][ox] = fma2(w[ky * 3 + 1], v[ox + 1], acc[set][oy][ox]);
                acc[set][oy][ox] = fma2(w[ky * 3 + 2], v[ox + 2], acc[set][oy][ox]);
              }
            }
          }
        }
        // the patch (and the taps) are consumed: release the stage as far as this warp is concerned
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->h_empty[s]));
        IRB_DW_WAIT(smem_u32(&bars->op_empty[o]), ((it / NOP) & 1u) ^ 1u);
        const uint32_t ob = sOP + o * OPBOX;
#pragma unroll
        for (int oy = 0; oy < BH; ++oy)
#pragma unroll
          for (int ox = 0; ox < BW; ++ox) {
            float gx, gy;
            unpack2(gelu_gate2e(acc[0][oy][ox], acc[1][oy][ox]), gx, gy);
            const uint32_t row = (uint32_t)((BH * by + oy) * TW + BW * bx + ox);
            if constexpr (std::is_same<TH_, float>::value) {
              // 2 tf32 values = 8 bytes: 16-byte chunk cp/2 of the row, swizzled by the row
              uint2 t = make_uint2(__float_as_uint(to_tf32(gx)), __float_as_uint(to_tf32(gy)));
              sts64u(ob + row * 128u + ((((uint32_t)cp >> 1) ^ (row & 7u)) << 4) + ((uint32_t)cp & 1u) * 8u, t);
            } else {
              const __half2 h = f2h2_sat(gx, gy);
              const uint32_t a = ob + row * 128u + ((((uint32_t)cp >> 2) ^ (row & 7u)) << 4) + ((uint32_t)cp & 3u) * 4u;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(*reinterpret_cast<const uint32_t*>(&h)) : "memory");
            }
          }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->op_ready[o]));
      }
    }
  } else {
    // =============================== epilogue: + residual, store ===============================
    const int q = warp;
    const uint32_t lsw = (uint32_t)(lane & 7);
    const int ngroups = (p.C + 31) / 32;
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    uint32_t j = 0, gc = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int b = ti.img(), y0 = ti.y0(), x0 = ti.x0();
      const uint32_t slot = p.nacc == 2 ? (j & 1u) : 0u;
      const uint32_t use = p.nacc == 2 ? (j >> 1) : j;
      mbar_wait(smem_u32(&bars->acc_full[slot]), use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.acc_stride;
      for (int g = 0; g < ngroups; ++g, ++gc) {
        const uint32_t s = gc % (uint32_t)p.RB, ph = (gc / (uint32_t)p.RB) & 1u;
        mbar_wait(smem_u32(&bars->r_full[s]), ph);
        const uint32_t box = sR + s * OPBOX + (uint32_t)q * WBOX;
        const uint32_t myrow = box + (uint32_t)lane * 128u;
        float v[32];
        tmem_ld32(tacc + (uint32_t)(g * 32), v);
        tmem_ld_wait();
        if (g == ngroups - 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] += (g * 32 + e < p.C) ? __ldg(p.bias + g * 32 + e) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t a = myrow + (((uint32_t)c ^ lsw) << 4);
          const float4 rr = lds128(a);
          sts128(a, make_float4(rr.x + v[4 * c], rr.y + v[4 * c + 1], rr.z + v[4 * c + 2], rr.w + v[4 * c + 3]));
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmY, box, g * 32, x0, y0 + 2 * q, b);     // this warp's 32 rows = patch rows 2q, 2q+1
          bulk_commit();
          if (gc > 0) {
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

template <typename TH_>
int launch_inst(const CUtensorMap& tH, const CUtensorMap& tR, const CUtensorMap& tY, const FfnTailParams& p, int grid,
                size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(ffn_tail_kernel<TH_>, optin));
  IRB_CUDA(launch_pdl(ffn_tail_kernel<TH_>, dim3(grid), dim3((EPI_WARPS + Geo<TH_>::DW_WARPS + 3) * 32), smem, s, tH, tR, tY, p));
  return IR_OK;
}

struct TailCfg { int RB; uint32_t stage_bytes, off_stage, off_op, off_r; size_t smem; };

bool configure(int C, int hp, bool half, TailCfg& c) {
  const int kc = half ? 64 : 32;
  if (C % 16 != 0 || C > 256 || C < 16 || hp % kc != 0) return false;
  const size_t budget = 227 * 1024 - 1024;
  const size_t stage = ((size_t)2 * HBOX + (size_t)C * 128 + (size_t)2 * 9 * kc * 4 + 1023) / 1024 * 1024;
  size_t off = HDR;
  c.off_stage = (uint32_t)off; off += NST * stage;
  c.off_op = (uint32_t)off; off += (size_t)NOP * OPBOX;
  if (off + 2 * OPBOX > budget) return false;
  c.RB = (int)std::min<size_t>(MAX_RB, (budget - off) / OPBOX);
  c.off_r = (uint32_t)off; off += (size_t)c.RB * OPBOX;
  c.stage_bytes = (uint32_t)stage;
  c.smem = off + 1024;
  return true;
}

}  // namespace

bool ffn_tail_supported(int C, int hp, bool half) {
  TailCfg c;
  return configure(C, hp, half, c);
}

int ffn_tail_kc(bool half) { return half ? 64 : 32; }

int launch_ffn_tail(const FfnTailArgs& a, cudaStream_t s) {
  TailCfg c;
  IRB_REQUIRE(configure(a.C, a.hp, a.half != 0, c), "ffn_tail: unsupported shape");
  IRB_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0, "ffn_tail: empty input");
  const int es = a.half ? 2 : 4, kc = a.half ? 64 : 32;
  CUtensorMap tH, tR, tY;
  {
    cuuint64_t d[4] = {(cuuint64_t)2 * a.hp, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)2 * a.hp * es, (cuuint64_t)2 * a.hp * es * a.W, (cuuint64_t)2 * a.hp * es * a.W * a.H};
    cuuint32_t box[4] = {(cuuint32_t)kc, TW + 2, TH + 2, 1};
    IRB_TRY(make_tmap(&tH, a.hidden, a.half != 0, 4, d, st, box, false));
  }
  {
    cuuint64_t d[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t st[3] = {(cuuint64_t)a.C * 4, (cuuint64_t)a.C * 4 * a.W, (cuuint64_t)a.C * 4 * a.W * a.H};
    cuuint32_t boxr[4] = {32, TW, TH, 1}, boxy[4] = {32, TW, 2, 1};
    IRB_TRY(make_tmap(&tR, a.x, false, 4, d, st, boxr, true));
    IRB_TRY(make_tmap(&tY, a.x, false, 4, d, st, boxy, true));
  }
  FfnTailParams p{};
  p.w = reinterpret_cast<const uint8_t*>(a.w_out); p.dw = a.dw_chunked; p.bias = a.bias;
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.hp = a.hp; p.nchunk = a.hp / kc;
  p.tiles_x = cdiv(a.W, TW); p.tiles_y = cdiv(a.H, TH); p.ntiles = p.tiles_x * p.tiles_y * a.B;
  p.RB = c.RB;
  p.acc_stride = (a.C + 31) / 32 * 32;
  p.nacc = 2 * p.acc_stride <= 512 ? 2 : 1;
  int cols = 32; while (cols < p.nacc * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.stage_bytes = c.stage_bytes; p.off_stage = c.off_stage; p.off_op = c.off_op; p.off_r = c.off_r;
  const int grid = std::max(1, std::min(p.ntiles, 148));
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);
  const double pix = (double)a.B * a.H * a.W;
  ProfScope prof(TAG_FFN_TAIL, pix * (2.0 * a.hp * es + 8.0 * a.C), pix * (36.0 * a.hp + 2.0 * a.hp * a.C), s);
  return a.half ? launch_inst<__half>(tH, tR, tY, p, grid, smem, s) : launch_inst<float>(tH, tR, tY, p, grid, smem, s);
}

}  // namespace irb
