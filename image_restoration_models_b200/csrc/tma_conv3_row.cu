// Row-strip 3x3 convolution (stride 1, zero padding 1) for narrow inputs (Cin <= 64): DnCNN's body layers
// (network_dncnn.py:63-68, conv + folded BatchNorm + ReLU) and Restormer's first Downsample (restormer.py:173-179).
//
// tma_conv3.cu loads every patch nine times, once per tap.  Here an image row is loaded ONCE: a tile is 128 consecutive
// pixels of one image row, the CTA walks down a column of tiles, and shared memory holds a ring of row strips
// [130 pixels][128 B] (SWIZZLE_128B, one box per 32 channels, TMA zero fill = the padding).  The operand of tap (dy, dx)
// is simply a DESCRIPTOR into strip y+dy whose start address is shifted by dx+1 rows: tcgen05 applies the 128-byte
// swizzle to absolute shared-memory address bits, so a row-shifted descriptor reads exactly what TMA wrote
// (scripts/probe_desc.py measures this on the device; the descriptor's base-offset field stays 0).  Each strip serves
// 3 output rows x 3 dx shifts; activations cross L2 -> SM once instead of nine times.
//
//   A-producer (1 thread)  one 4-D bulk-tensor load per (row, 32-channel box)
//   W-producer (1 thread)  the [N][128 B] weight box of every (tap, box), streamed per output row
//   transform (4 warps)    fp32 -> tf32 rounding in place, or fp32 -> fp16 into the operand row ring
//   MMA (1 thread)         9 taps x boxes x K-steps per output row into a double-buffered TMEM accumulator
//   epilogue (4 warps)     bias + ReLU + plain rows, or the PixelUnshuffle scatter
#include "common.cuh"
#include "sm100.cuh"
#include "tc_gemm.cuh"
#include "tmap.cuh"

namespace irb {

namespace {

using namespace sm100;

constexpr int TM = 128;
constexpr int SPIX = TM + 2;                 // strip pixels (one halo pixel on each side)
constexpr int SBYTES = SPIX * 128;           // bytes TMA writes per box
constexpr int ROWB = 17 * 1024;              // box pitch (1024-aligned)
constexpr int EPI_WARPS = 4, XF_WARPS = 4;
constexpr int WARP_A = 8, WARP_W = 9, WARP_MMA = 10;
constexpr int NTHREADS = 11 * 32;
constexpr int MAX_R = 8, MAX_W = 16;
constexpr int HDR = 1024;

struct Bars {
  unsigned long long raw_full[MAX_R], raw_empty[MAX_R];     // fp32 rows as TMA wrote them
  unsigned long long op_ready[MAX_R], op_empty[MAX_R];      // operand rows (fp32 mode: the same slots, rounded in place)
  unsigned long long w_full[MAX_W], w_empty[MAX_W];
  unsigned long long acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct RowParams {
  const uint8_t* w; float* y; int ldy; const float* bias; int relu;
  int B, H, W, Cin, N, n_valid, o_mode;
  int nkb, nob;              // raw boxes (32 fp32 channels) / operand boxes per row
  int RR, RO, NW;            // ring depths: raw rows, operand rows (fp32: RO == RR, same memory), weight boxes
  int w_resident;            // all 9 * nob weight boxes stay in shared memory (NW == 9 * nob, loaded once)
  int strips, segs, seg, nitems;
  int acc_stride, tmem_cols;
  uint32_t off_raw, off_op, off_w, off_stg, wstage;
};

struct Item { int b, x0, y_lo, y_hi; };
__device__ __forceinline__ Item item_of(const RowParams& p, int it) {
  Item r;
  const int sx = it % p.strips, sy = (it / p.strips) % p.segs;
  r.b = it / (p.strips * p.segs);
  r.x0 = sx * TM;
  r.y_lo = sy * p.seg;
  r.y_hi = min(p.H, r.y_lo + p.seg);
  return r;
}

// The MMA issue loop, run by the whole (converged) MMA warp with one elected lane issuing.  ncu showed this thread -- not the tensor core, the loads or the epilogue --
// pacing the kernel: 72 small MMAs (N = 64, K = 8) per tile, each costing ~10 issue slots of descriptor arithmetic and
// lane election at ~9 cycles apiece.  Everything that can be a compile-time constant is one: boxes per row (NOB), K steps
// of the last box (NKL), resident weights (RES), the nine taps; descriptors are a constant plus (address >> 4).
template <typename TOp, int NOB, int NKL, bool RES>
__device__ __forceinline__ void mma_loop(const RowParams& p, Bars* bars, uint32_t tmem_base, uint32_t sOp, uint32_t sW,
                                         uint32_t op_row) {
  constexpr bool OPRING = sizeof(TOp) == 2;
  const uint32_t idesc = make_idesc<TOp>(p.N);
  const uint32_t R = (uint32_t)p.RO;
  const uint64_t dhi = sw128_desc(0);
  const uint32_t wst = p.wstage >> 4;                      // weight box pitch in descriptor units
  uint32_t pos = 0, phs = 0;                               // ring slot / phase of the current item's row 0
  uint32_t sw = 0, pw = 0, j = 0;
  if (RES && (int)blockIdx.x < p.nitems) { mbar_wait(smem_u32(&bars->w_full[0]), 0); tc_fence_after(); }
  unsigned long long* rel = OPRING ? bars->op_empty : bars->raw_empty;
  const uint64_t wd0 = dhi | (uint64_t)(sW >> 4);
  for (int it = blockIdx.x; it < p.nitems; it += gridDim.x) {
    const Item im = item_of(p, it);
    const int n = im.y_hi - im.y_lo;
    uint32_t s0 = pos, p0 = phs;
    uint32_t s1 = s0 + 1, p1 = p0; if (s1 == R) { s1 = 0; p1 ^= 1u; }
    uint32_t s2 = s1 + 1, p2 = p1; if (s2 == R) { s2 = 0; p2 ^= 1u; }
    mbar_wait(smem_u32(&bars->op_ready[s0]), p0);
    mbar_wait(smem_u32(&bars->op_ready[s1]), p1);
    for (int i = 0; i < n; ++i, ++j) {
      mbar_wait(smem_u32(&bars->op_ready[s2]), p2);
      const uint32_t slot = j & 1u;
      mbar_wait(smem_u32(&bars->acc_empty[slot]), ((j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dacc = tmem_base + slot * (uint32_t)p.acc_stride;
      const uint64_t rd[3] = {dhi | (uint64_t)((sOp + s0 * op_row) >> 4), dhi | (uint64_t)((sOp + s1 * op_row) >> 4),
                              dhi | (uint64_t)((sOp + s2 * op_row) >> 4)};
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int ob = 0; ob < NOB; ++ob) {
          constexpr int dummy = 0; (void)dummy;
          const uint64_t ad = rd[tap / 3] + (uint64_t)(((tap % 3) * 128 + ob * ROWB) >> 4);
          uint64_t wd;
          if (RES) wd = wd0 + (uint64_t)((tap * NOB + ob) * wst);
          else {
            mbar_wait(smem_u32(&bars->w_full[sw]), pw);
            tc_fence_after();
            wd = wd0 + (uint64_t)(sw * wst);
          }
#pragma unroll
          for (int kk = 0; kk < (ob == NOB - 1 ? NKL : 4); ++kk)
            umma_elect<TOp>(dacc, ad + (uint64_t)(2 * kk), wd + (uint64_t)(2 * kk), idesc, (tap | ob | kk) != 0 ? 1u : 0u);
          if (!RES) {
            umma_commit_elect(smem_u32(&bars->w_empty[sw]));
            if (++sw == (uint32_t)p.NW) { sw = 0; pw ^= 1u; }
          }
        }
      }
      umma_commit_elect(smem_u32(&bars->acc_full[slot]));
      umma_commit_elect(smem_u32(&rel[s0]));                      // row i is not needed any more
      if (i == n - 1) {                                     // the item's last two rows as well
        umma_commit_elect(smem_u32(&rel[s1]));
        umma_commit_elect(smem_u32(&rel[s2]));
      }
      s0 = s1; p0 = p1; s1 = s2; p1 = p2;
      if (++s2 == R) { s2 = 0; p2 ^= 1u; }
    }
    pos = s2; phs = p2;                                     // row n + 2 of this item == row 0 of the next one
  }
}

template <typename TOp>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3_row_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmY, const RowParams p) {
  constexpr bool OPRING = sizeof(TOp) == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (base - smem_u32(smem_raw)));
  const uint32_t sRaw = base + p.off_raw, sOp = base + p.off_op, sW = base + p.off_w, sStg = base + p.off_stg;
  const uint32_t raw_row = (uint32_t)p.nkb * ROWB, op_row = (uint32_t)p.nob * ROWB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwb = 9 * p.nob;                     // weight boxes per output row

  if (tid == 0) {
    for (int s = 0; s < MAX_R; ++s) {
      mbar_init(smem_u32(&bars->raw_full[s]), 1);
      mbar_init(smem_u32(&bars->raw_empty[s]), OPRING ? XF_WARPS * 32 : 1);
      mbar_init(smem_u32(&bars->op_ready[s]), XF_WARPS * 32);
      mbar_init(smem_u32(&bars->op_empty[s]), 1);
    }
    for (int s = 0; s < MAX_W; ++s) {
      mbar_init(smem_u32(&bars->w_full[s]), 1);
      mbar_init(smem_u32(&bars->w_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), EPI_WARPS * 32);
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

  if (warp == WARP_A) {
    // =============================== row producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      uint32_t s = 0, ph = 0;
      for (int it = blockIdx.x; it < p.nitems; it += gridDim.x) {
        const Item im = item_of(p, it);
        for (int r = im.y_lo - 1; r <= im.y_hi; ++r) {          // rows -1 and H are all zero fill
          mbar_wait(smem_u32(&bars->raw_empty[s]), ph ^ 1u);
          const uint32_t fb = smem_u32(&bars->raw_full[s]);
          mbar_expect_tx(fb, (uint32_t)(p.nkb * SBYTES));
          for (int cb = 0; cb < p.nkb; ++cb)
            tma_load_4d(&tmA, fb, sRaw + s * raw_row + cb * ROWB, cb * 32, im.x0 - 1, r, im.b);
          if (++s == (uint32_t)p.RR) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == WARP_W) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const uint32_t wbytes = (uint32_t)p.N * 128u;
      if (p.w_resident) {
        if ((int)blockIdx.x < p.nitems) {
          const uint32_t fb = smem_u32(&bars->w_full[0]);
          mbar_expect_tx(fb, wbytes * (uint32_t)nwb);
          for (int wb = 0; wb < nwb; ++wb) bulk_load(sW + wb * p.wstage, p.w + (size_t)wb * wbytes, wbytes, fb);
        }
      } else
      for (int it = blockIdx.x; it < p.nitems; it += gridDim.x) {
        const Item im = item_of(p, it);
        for (int y = im.y_lo; y < im.y_hi; ++y)
          for (int wb = 0; wb < nwb; ++wb) {
            mbar_wait(smem_u32(&bars->w_empty[s]), ph ^ 1u);
            const uint32_t fb = smem_u32(&bars->w_full[s]);
            mbar_expect_tx(fb, wbytes);
            bulk_load(sW + s * p.wstage, p.w + (size_t)wb * wbytes, wbytes, fb);
            if (++s == (uint32_t)p.NW) { s = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    {
      const int nk_last = ((p.Cin - (p.nob - 1) * (OPRING ? 64 : 32)) * (int)sizeof(TOp) + 31) / 32;
      const int key = (p.nob - 1) * 8 + (nk_last - 1) * 2 + (p.w_resident ? 1 : 0);
      switch (key) {
#define IRB_ROW_CASE(NOB, NKL, RES) \
        case ((NOB) - 1) * 8 + ((NKL) - 1) * 2 + (RES): mma_loop<TOp, NOB, NKL, (RES) != 0>(p, bars, tmem_base, sOp, sW, op_row); break;
        IRB_ROW_CASE(1, 1, 0) IRB_ROW_CASE(1, 1, 1) IRB_ROW_CASE(1, 2, 0) IRB_ROW_CASE(1, 2, 1)
        IRB_ROW_CASE(1, 3, 0) IRB_ROW_CASE(1, 3, 1) IRB_ROW_CASE(1, 4, 0) IRB_ROW_CASE(1, 4, 1)
        IRB_ROW_CASE(2, 1, 0) IRB_ROW_CASE(2, 1, 1) IRB_ROW_CASE(2, 2, 0) IRB_ROW_CASE(2, 2, 1)
        IRB_ROW_CASE(2, 3, 0) IRB_ROW_CASE(2, 3, 1) IRB_ROW_CASE(2, 4, 0) IRB_ROW_CASE(2, 4, 1)
#undef IRB_ROW_CASE
        default: __trap();
      }
    }
  } else if (warp >= EPI_WARPS) {
    // =============================== operand rounding ===============================
    const int t = tid - EPI_WARPS * 32;
    uint32_t sr = 0, pr = 0, so = 0, po = 0;
    for (int it = blockIdx.x; it < p.nitems; it += gridDim.x) {
      const Item im = item_of(p, it);
      const int nrows = im.y_hi - im.y_lo + 2;
      for (int rr = 0; rr < nrows; ++rr) {
        mbar_wait(smem_u32(&bars->raw_full[sr]), pr);
        if constexpr (!OPRING) {
          for (int px = t; px < SPIX; px += XF_WARPS * 32) {
            const uint32_t sw7 = (uint32_t)(px & 7);
            for (int cb = 0; cb < p.nkb; ++cb) {
              const uint32_t row = sRaw + sr * raw_row + cb * ROWB + (uint32_t)px * 128u;
              float4 x[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) x[c] = lds128(row + (((uint32_t)c ^ sw7) << 4));
#pragma unroll
              for (int c = 0; c < 8; ++c)
                sts128(row + (((uint32_t)c ^ sw7) << 4),
                       make_float4(to_tf32(x[c].x), to_tf32(x[c].y), to_tf32(x[c].z), to_tf32(x[c].w)));
            }
          }
          fence_async_smem();
          mbar_arrive(smem_u32(&bars->op_ready[sr]));      // fp32 mode: operand slot == raw slot
        } else {
          mbar_wait(smem_u32(&bars->op_empty[so]), po ^ 1u);
          for (int px = t; px < SPIX; px += XF_WARPS * 32) {
            const uint32_t sw7 = (uint32_t)(px & 7);
            for (int ob = 0; ob < p.nob; ++ob) {
              float4 x[16];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (2 * ob + h < p.nkb) {
                  const uint32_t row = sRaw + sr * raw_row + (2 * ob + h) * ROWB + (uint32_t)px * 128u;
#pragma unroll
                  for (int c = 0; c < 8; ++c) x[h * 8 + c] = lds128(row + (((uint32_t)c ^ sw7) << 4));
                } else {
#pragma unroll
                  for (int c = 0; c < 8; ++c) x[h * 8 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
              const uint32_t orow = sOp + so * op_row + ob * ROWB + (uint32_t)px * 128u;
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                uint4 u;
                __half2* hh = reinterpret_cast<__half2*>(&u);
                hh[0] = f2h2_sat(x[2 * c8].x, x[2 * c8].y); hh[1] = f2h2_sat(x[2 * c8].z, x[2 * c8].w);
                hh[2] = f2h2_sat(x[2 * c8 + 1].x, x[2 * c8 + 1].y); hh[3] = f2h2_sat(x[2 * c8 + 1].z, x[2 * c8 + 1].w);
                sts128u(orow + (((uint32_t)c8 ^ sw7) << 4), u);
              }
            }
          }
          mbar_arrive(smem_u32(&bars->raw_empty[sr]));
          fence_async_smem();
          mbar_arrive(smem_u32(&bars->op_ready[so]));
          if (++so == (uint32_t)p.RO) { so = 0; po ^= 1u; }
        }
        if (++sr == (uint32_t)p.RR) { sr = 0; pr ^= 1u; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp;
    const int ngroups = (p.N + 31) / 32;
    const uint32_t lsw = (uint32_t)(lane & 7);
    uint32_t j = 0, gc = 0;
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    for (int it = blockIdx.x; it < p.nitems; it += gridDim.x) {
      const Item im = item_of(p, it);
      const int x = im.x0 + q * 32 + lane;                     // TMEM lane == pixel of the row tile
      for (int y = im.y_lo; y < im.y_hi; ++y, ++j) {
        const uint32_t slot = j & 1u;
        mbar_wait(smem_u32(&bars->acc_full[slot]), (j >> 1) & 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.acc_stride;
        for (int g = 0; g < ngroups; ++g) {
          float v[32];
          tmem_ld32(tacc + (uint32_t)(g * 32), v);
          tmem_ld_wait();
          if (g == ngroups - 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
          const int c0 = g * 32;
          if (p.o_mode == O_NHWC) {
            // plain rows: this warp's 32 pixels x 32 channels go through a swizzled staging box to ONE bulk-tensor store
            // (full 128-byte lines; thread-per-pixel STGs would write 16 bytes into 32 different lines each)
            if (p.bias) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] += (c0 + e < p.n_valid) ? __ldg(p.bias + c0 + e) : 0.f;
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            const uint32_t box = sStg + (uint32_t)(q * 2 + (int)(gc & 1u)) * 4096u;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              sts128(box + (uint32_t)lane * 128u + (((uint32_t)c ^ lsw) << 4),
                     make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmY, box, c0, im.x0 + q * 32, y, im.b);      // columns >= N and pixels >= W are clipped
              bulk_commit();
            }
            ++gc;
            continue;
          }
          if (x >= p.W) continue;
          {
            // PixelUnshuffle: out[c*4 + 2*(y&1) + (x&1), y/2, x/2] = conv[c, y, x]   (restormer.py:176)
            float* dst = p.y + (((long long)im.b * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * p.ldy + (y & 1) * 2 + (x & 1);
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < p.n_valid) dst[(c0 + e) * 4] = v[e];
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

struct RowCfg { int nkb, nob, RR, RO, NW, w_resident; uint32_t off_raw, off_op, off_w, off_stg, wstage; size_t smem; };

bool configure(int cin, int n, bool half, RowCfg& c) {
  if (cin % 4 != 0 || cin < 4 || cin > 64 || n % 16 != 0 || n < 16 || n > 256) return false;
  if (half && cin % 8 != 0) return false;
  const size_t budget = 227 * 1024 - 1024 - 32 * 1024;      // minus the epilogue's staging boxes (4 warps x 2 x 4 KB)
  c.nkb = (cin + 31) / 32;
  c.nob = half ? (cin + 63) / 64 : c.nkb;
  c.wstage = (uint32_t)(((size_t)n * 128 + 1023) / 1024 * 1024);
  // weights: resident when all 9 * nob boxes fit next to four operand rows (+ the fp16 path's two raw rows); otherwise
  // streamed per output row through as deep a ring as fits (the stream is latency-bound: bytes in flight matter)
  const size_t rows_min = half ? (size_t)2 * c.nkb * ROWB + (size_t)4 * c.nob * ROWB : (size_t)4 * c.nkb * ROWB;
  const int nwb = 9 * c.nob;
  if (HDR + rows_min + (size_t)nwb * c.wstage <= budget && nwb <= 64) { c.w_resident = 1; c.NW = nwb; }
  else {
    c.w_resident = 0;
    if (HDR + rows_min + 2 * (size_t)c.wstage > budget) return false;
    c.NW = (int)std::min<size_t>(MAX_W, (budget - HDR - rows_min) / c.wstage);
  }
  size_t off = HDR;
  c.off_w = (uint32_t)off; off += (size_t)c.NW * c.wstage;
  if (half) {
    c.RR = 2;
    c.off_raw = (uint32_t)off; off += (size_t)c.RR * c.nkb * ROWB;
    if (off + (size_t)4 * c.nob * ROWB > budget) return false;
    c.RO = (int)std::min<size_t>(MAX_R, (budget - off) / ((size_t)c.nob * ROWB));
    c.off_op = (uint32_t)off; off += (size_t)c.RO * c.nob * ROWB;
  } else {
    if (off + (size_t)4 * c.nkb * ROWB > budget) return false;
    c.RR = (int)std::min<size_t>(MAX_R, (budget - off) / ((size_t)c.nkb * ROWB));
    c.RO = c.RR;
    c.off_raw = c.off_op = (uint32_t)off; off += (size_t)c.RR * c.nkb * ROWB;
  }
  c.off_stg = (uint32_t)off; off += 32 * 1024;
  c.smem = off + 1024;
  return true;
}

template <typename TOp>
int launch_inst(const CUtensorMap& tA, const CUtensorMap& tY, const RowParams& p, int grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(conv3_row_kernel<TOp>, optin));
  IRB_CUDA(launch_pdl(conv3_row_kernel<TOp>, dim3(grid), dim3(NTHREADS), smem, s, tA, tY, p));
  return IR_OK;
}

}  // namespace

// Same weight layout as tma_conv3.cu; worth it when the rows are wide enough to fill the 128-pixel tiles.
bool conv3_row_supported(int cin, int cout_p, bool half) {
  RowCfg c;
  return configure(cin, cout_p, half, c);
}

int launch_conv3_row(const float* in, int ld_in, int cin, const void* w_packed, const float* bias, int relu, int cout_p,
                     int cout_valid, int B, int H, int W, float* out, int ld_out, int o_mode, bool half, cudaStream_t s) {
  RowCfg c;
  IRB_REQUIRE(configure(cin, cout_p, half, c), "conv3_row: unsupported shape");
  IRB_REQUIRE(o_mode == O_UNSHUFFLE || o_mode == O_NHWC, "conv3_row: plain rows or the PixelUnshuffle scatter");
  IRB_REQUIRE(o_mode == O_NHWC || (bias == nullptr && !relu), "conv3_row: bias / ReLU belong to the plain-row epilogue");
  IRB_REQUIRE(o_mode != O_UNSHUFFLE || (H % 2 == 0 && W % 2 == 0), "conv3_row: unshuffle needs even H, W");
  IRB_REQUIRE(ld_in % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && ld_out % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0,
              "conv3_row: 16-byte alignment");
  CUtensorMap tA;
  {
    cuuint64_t d[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)ld_in * 4, (cuuint64_t)ld_in * 4 * W, (cuuint64_t)ld_in * 4 * W * H};
    cuuint32_t box[4] = {32, SPIX, 1, 1};
    IRB_TRY(make_tmap(&tA, in, false, 4, d, st, box, true));
  }
  CUtensorMap tY = tA;
  if (o_mode == O_NHWC) {
    IRB_REQUIRE(cout_p == cout_valid, "conv3_row: plain rows store all computed channels");
    cuuint64_t d[4] = {(cuuint64_t)cout_valid, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)ld_out * 4, (cuuint64_t)ld_out * 4 * W, (cuuint64_t)ld_out * 4 * W * H};
    cuuint32_t box[4] = {32, 32, 1, 1};
    IRB_TRY(make_tmap(&tY, out, false, 4, d, st, box, true));
  }
  RowParams p{};
  p.w = reinterpret_cast<const uint8_t*>(w_packed); p.y = out; p.ldy = ld_out; p.bias = bias; p.relu = relu;
  p.B = B; p.H = H; p.W = W; p.Cin = cin; p.N = cout_p; p.n_valid = cout_valid; p.o_mode = o_mode;
  p.nkb = c.nkb; p.nob = c.nob; p.RR = c.RR; p.RO = c.RO; p.NW = c.NW; p.w_resident = c.w_resident;
  p.strips = cdiv(W, TM);
  // row segments: enough work items for two waves of CTAs, but at least 8 rows each (2 halo rows are loaded per item)
  int seg = H;
  while ((long long)B * p.strips * cdiv(H, seg) < 2 * 148 && seg > 8) seg = cdiv(seg, 2);
  p.seg = seg; p.segs = cdiv(H, seg); p.nitems = B * p.strips * p.segs;
  p.acc_stride = (cout_p + 31) / 32 * 32;
  int cols = 32; while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.off_raw = c.off_raw; p.off_op = c.off_op; p.off_w = c.off_w; p.off_stg = c.off_stg; p.wstage = c.wstage;
  const int grid = std::max(1, std::min(p.nitems, 148));
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);
  const double pix = (double)B * H * W;
  ProfScope prof(TAG_CONV3, pix * 4.0 * (cin + cout_p), 2.0 * pix * 9.0 * cin * cout_p, s);
  return half ? launch_inst<__half>(tA, tY, p, grid, smem, s) : launch_inst<float>(tA, tY, p, grid, smem, s);
}

}  // namespace irb
