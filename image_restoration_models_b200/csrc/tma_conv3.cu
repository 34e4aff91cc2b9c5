// 3x3 convolution (stride 1, zero padding 1) as a TMA-fed implicit GEMM on tcgen05, with PixelUnshuffle / PixelShuffle
// folded into the store: Downsample (restormer.py:173-179) and Upsample (:182-189).
//
//     y[pixel, n] = sum_{tap, c} x[pixel + tap, c] * W[n, c, tap]        M = an 8 x 16 pixel patch, N <= 256 per CTA
//
// The first-generation kernel gathered the nine shifted pixels through registers (9 loads per input element).  Here
// the im2col never exists: for every (tap, 32-channel box) the producer issues ONE 4-D bulk-tensor load of the patch
// shifted by the tap -- [128 px][128 B], already in the tensor core's SWIZZLE_128B layout, with TMA's out-of-bounds
// zero fill playing the role of the padding -- plus one bulk copy of the matching [N][128 B] weight box (weights are
// streamed: K = 9 Cin does not fit shared memory).  Four transform warps round the fp32 activations to the operand
// type (tf32 in place, or fp16 into a second ring), one thread issues the MMAs, four epilogue warps scatter the
// accumulator to its PixelUnshuffle / PixelShuffle position.
#include "common.cuh"
#include "sm100.cuh"
#include "tc_gemm.cuh"
#include "tmap.cuh"

#include <type_traits>
#include <cstdlib>

namespace irb {

namespace {

using namespace sm100;

constexpr int TH = 8, TW = 16, TM = TH * TW;
constexpr int BOX = TM * 128;
// Patch mode (tf32 operands): the (8+2) x (14+2) halo patch of a 32-channel box is loaded ONCE -- [160 px][128 B], flattened
// with a pitch of 16 pixels -- and tap (dy, dx) is a descriptor whose start address is shifted by dy * 16 + dx rows
// (tcgen05 swizzles on absolute address bits: tma_conv3_row.cu).  MMA row i is then patch position (i / 16, i % 16): the
// 8 x 14 positions with i % 16 < 14 are the tile's outputs, the other 16 lanes compute values nobody stores.  Activations
// cross L2 -> SM 1.4 times instead of 9, and the tf32 rounding pass runs once per box instead of once per (tap, box).
constexpr int PTW = 14, PPW = 16, PPH = TH + 2;
constexpr int PTX = PPH * PPW * 128;           // bytes one patch load delivers (160 rows)
constexpr int PBOX = 21 * 1024;                // stage pitch: the last tap's window ends at row 2 * 16 + 2 + 127 = 161
constexpr int EPI_WARPS = 4, XF_WARPS = 4;
constexpr int WARP_A = 8, WARP_MMA = 9;
constexpr int NTHREADS = 10 * 32;
constexpr int MAX_S = 8, MAX_OP = 4, MAX_W = 16;
constexpr int HDR = 1024;

struct Bars {
  unsigned long long a_full[MAX_S], a_empty[MAX_S], a_ready[MAX_S];
  unsigned long long op_ready[MAX_OP], op_empty[MAX_OP];
  unsigned long long w_full[MAX_W], w_empty[MAX_W];
  unsigned long long acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct Conv3Params {
  const uint8_t* w;        // [9 * nob_t boxes][N][128 B] swizzled operand image (PackMat kind 2, fmt 3 / 4)
  float* y; int ldy;
  const float* bias; int relu;          // plain-row epilogue only (DnCNN body: conv + bias [+ folded BN] + ReLU)
  int B, H, W, Cin, N, n_valid, o_mode;
  int nc;                  // output columns per CTA
  int nkb_t, nob_t;        // raw (32 fp32 channel) boxes / operand boxes per tap
  int S, SOP, NW;
  int tiles_x, tiles_y, ntiles;
  int nacc, acc_stride, tmem_cols;
  uint32_t off_a, off_op, off_w, wstage;
  int patch;               // 1: patch mode (see PTW)
};

struct TileIter {
  int t, step, end, tx_n, ty_n, tw;
  __device__ TileIter(const Conv3Params& p)
      : t(blockIdx.x), step(gridDim.x), end(p.ntiles), tx_n(p.tiles_x), ty_n(p.tiles_y), tw(p.patch ? PTW : TW) {}
  __device__ bool valid() const { return t < end; }
  __device__ void next() { t += step; }
  __device__ int img() const { return t / (tx_n * ty_n); }
  __device__ int y0() const { return ((t / tx_n) % ty_n) * TH; }
  __device__ int x0() const { return (t % tx_n) * tw; }
};

template <typename TOp>
__global__ void __launch_bounds__(NTHREADS, 1)
tma_conv3_kernel(const __grid_constant__ CUtensorMap tmA, const Conv3Params p) {
  constexpr bool OPRING = sizeof(TOp) == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  Bars* bars = reinterpret_cast<Bars*>(smem_raw + (base - smem_u32(smem_raw)));
  const uint32_t sA = base + p.off_a, sOP = base + p.off_op, sW = base + p.off_w;
  // the warp index through a warp reduction lives in a UNIFORM register: the role branches become uniform branches and the
  // code under them uses the uniform datapath (memory descriptors, TMEM / barrier addresses) without one R2UR per use
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)__reduce_or_sync(0xffffffffu, (unsigned)(tid >> 5));
  const int n0 = blockIdx.y * p.nc;
  const int ncur = min(p.nc, p.N - n0);
  const int nraw = 9 * p.nkb_t;                      // raw boxes per tile
  const int nop = OPRING ? 9 * p.nob_t : nraw;       // operand (== weight) boxes per tile

  if (tid == 0) {
    for (int s = 0; s < MAX_S; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), OPRING ? XF_WARPS * 32 : 1);
      mbar_init(smem_u32(&bars->a_ready[s]), XF_WARPS * 32);
    }
    for (int s = 0; s < MAX_OP; ++s) {
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
    // =============================== activation / weight producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      uint32_t sa = 0, pa = 0, sw = 0, pw = 0;
      const uint32_t wbytes = (uint32_t)ncur * 128u;
      if (!OPRING && p.patch) {
        for (TileIter ti(p); ti.valid(); ti.next()) {
          const int b = ti.img(), y0 = ti.y0(), x0 = ti.x0();
          for (int cb = 0; cb < p.nkb_t; ++cb) {
            mbar_wait(smem_u32(&bars->a_empty[sa]), pa ^ 1u);
            const uint32_t fa = smem_u32(&bars->a_full[sa]);
            mbar_expect_tx(fa, PTX);
            tma_load_4d(&tmA, fa, sA + sa * PBOX, cb * 32, x0 - 1, y0 - 1, b);
            if (++sa == (uint32_t)p.S) { sa = 0; pa ^= 1u; }
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(smem_u32(&bars->w_empty[sw]), pw ^ 1u);
              const uint32_t fb = smem_u32(&bars->w_full[sw]);
              mbar_expect_tx(fb, wbytes);
              bulk_load(sW + sw * p.wstage, p.w + ((size_t)(tap * p.nkb_t + cb) * p.N + n0) * 128, wbytes, fb);
              if (++sw == (uint32_t)p.NW) { sw = 0; pw ^= 1u; }
            }
          }
        }
      } else
      for (TileIter ti(p); ti.valid(); ti.next()) {
        const int b = ti.img(), y0 = ti.y0(), x0 = ti.x0();
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int cb = 0; cb < p.nkb_t; ++cb) {
            // weight box: one per operand box (fp16: every second raw box starts a new one)
            if (!OPRING || (cb & 1) == 0) {
              const int wb = OPRING ? tap * p.nob_t + (cb >> 1) : tap * p.nkb_t + cb;
              mbar_wait(smem_u32(&bars->w_empty[sw]), pw ^ 1u);
              const uint32_t fb = smem_u32(&bars->w_full[sw]);
              mbar_expect_tx(fb, wbytes);
              bulk_load(sW + sw * p.wstage, p.w + ((size_t)wb * p.N + n0) * 128, wbytes, fb);
              if (++sw == (uint32_t)p.NW) { sw = 0; pw ^= 1u; }
            }
            mbar_wait(smem_u32(&bars->a_empty[sa]), pa ^ 1u);
            const uint32_t fb = smem_u32(&bars->a_full[sa]);
            mbar_expect_tx(fb, BOX);
            tma_load_4d(&tmA, fb, sA + sa * BOX, cb * 32, x0 + dx, y0 + dy, b);
            if (++sa == (uint32_t)p.S) { sa = 0; pa ^= 1u; }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================== MMA issuer ===============================
    // the whole converged warp runs the loop and one elected lane issues; ring positions advance by compare-and-wrap
    // and descriptors are a constant plus (address >> 4): this warp's instruction stream, not the tensor core, paces a
    // convolution made of many small MMAs (ncu, see tma_conv3_row.cu)
    const uint32_t idesc = make_idesc<TOp>(ncur);
    const uint64_t dhi = sw128_desc(0);
    const int nbox_t = OPRING ? p.nob_t : p.nkb_t;                       // operand boxes per tap
    constexpr int OC = OPRING ? 64 : 32;
    const int nk_last = ((p.Cin - (nbox_t - 1) * OC) * (int)sizeof(TOp) + 31) / 32;
    const uint32_t ring = (uint32_t)(OPRING ? p.SOP : p.S);
    const uint32_t abase = OPRING ? sOP : sA;
    unsigned long long* a_rel = OPRING ? bars->op_empty : bars->a_empty;
    unsigned long long* a_rdy = OPRING ? bars->op_ready : bars->a_ready;
    uint32_t so = 0, po = 0, sw = 0, pw = 0, j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const uint32_t slot = p.nacc == 2 ? (j & 1u) : 0u;
      const uint32_t use = p.nacc == 2 ? (j >> 1) : j;
      mbar_wait(smem_u32(&bars->acc_empty[slot]), (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dacc = tmem_base + slot * (uint32_t)p.acc_stride;
      uint32_t acc = 0u;
      if (!OPRING && p.patch) {
        for (int ob = 0; ob < nbox_t; ++ob) {
          const int nk = ob == nbox_t - 1 ? nk_last : 4;
          mbar_wait(smem_u32(&a_rdy[so]), po);
          tc_fence_after();
          const uint64_t a0 = dhi | (uint64_t)((abase + so * PBOX) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(smem_u32(&bars->w_full[sw]), pw);
            tc_fence_after();
            const uint64_t ad = a0 + (uint64_t)((((tap / 3) * PPW + tap % 3) * 128) >> 4);
            const uint64_t wd = dhi | (uint64_t)((sW + sw * p.wstage) >> 4);
            for (int kk = 0; kk < nk; ++kk) {
              umma_elect<TOp>(dacc, ad + (uint64_t)(2 * kk), wd + (uint64_t)(2 * kk), idesc, acc);
              acc = 1u;
            }
            umma_commit_elect(smem_u32(&bars->w_empty[sw]));
            if (++sw == (uint32_t)p.NW) { sw = 0; pw ^= 1u; }
          }
          umma_commit_elect(smem_u32(&a_rel[so]));
          if (++so == ring) { so = 0; po ^= 1u; }
        }
      } else
      for (int tap = 0; tap < 9; ++tap) {
        for (int ob = 0; ob < nbox_t; ++ob) {
          const int nk = ob == nbox_t - 1 ? nk_last : 4;
          mbar_wait(smem_u32(&a_rdy[so]), po);
          mbar_wait(smem_u32(&bars->w_full[sw]), pw);
          tc_fence_after();
          const uint64_t ad = dhi | (uint64_t)((abase + so * BOX) >> 4);
          const uint64_t wd = dhi | (uint64_t)((sW + sw * p.wstage) >> 4);
          for (int kk = 0; kk < nk; ++kk) {
            umma_elect<TOp>(dacc, ad + (uint64_t)(2 * kk), wd + (uint64_t)(2 * kk), idesc, acc);
            acc = 1u;
          }
          umma_commit_elect(smem_u32(&a_rel[so]));
          umma_commit_elect(smem_u32(&bars->w_empty[sw]));
          if (++so == ring) { so = 0; po ^= 1u; }
          if (++sw == (uint32_t)p.NW) { sw = 0; pw ^= 1u; }
        }
      }
      umma_commit_elect(smem_u32(&bars->acc_full[slot]));
      __syncwarp();
    }
  } else if (warp >= EPI_WARPS) {
    // =============================== operand rounding ===============================
    // thread r owns pixel row r of every box: 8 conflict-free 16-byte chunks under the 128-byte swizzle
    const int r = tid - EPI_WARPS * 32;
    const uint32_t rsw = (uint32_t)(r & 7);
    uint32_t sa = 0, pa = 0, so = 0, po = 0;
    for (TileIter ti(p); ti.valid(); ti.next()) {
      if constexpr (!OPRING) {
        if (p.patch) {
          for (int i = 0; i < p.nkb_t; ++i) {
            mbar_wait(smem_u32(&bars->a_full[sa]), pa);
            for (int rr = r; rr < PPH * PPW; rr += XF_WARPS * 32) {
              const uint32_t row = sA + sa * PBOX + (uint32_t)rr * 128u, sw7 = (uint32_t)(rr & 7);
              float4 x[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) x[c] = lds128(row + (((uint32_t)c ^ sw7) << 4));
#pragma unroll
              for (int c = 0; c < 8; ++c)
                sts128(row + (((uint32_t)c ^ sw7) << 4),
                       make_float4(to_tf32(x[c].x), to_tf32(x[c].y), to_tf32(x[c].z), to_tf32(x[c].w)));
            }
            fence_async_smem();
            mbar_arrive(smem_u32(&bars->a_ready[sa]));
            if (++sa == (uint32_t)p.S) { sa = 0; pa ^= 1u; }
          }
          continue;
        }
        for (int i = 0; i < nraw; ++i) {
          mbar_wait(smem_u32(&bars->a_full[sa]), pa);
          const uint32_t row = sA + sa * BOX + (uint32_t)r * 128u;
          float4 x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = lds128(row + (((uint32_t)c ^ rsw) << 4));
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(row + (((uint32_t)c ^ rsw) << 4),
                   make_float4(to_tf32(x[c].x), to_tf32(x[c].y), to_tf32(x[c].z), to_tf32(x[c].w)));
          fence_async_smem();
          mbar_arrive(smem_u32(&bars->a_ready[sa]));
          if (++sa == (uint32_t)p.S) { sa = 0; pa ^= 1u; }
        }
      } else {
        for (int tap = 0; tap < 9; ++tap) {
          for (int ob = 0; ob < p.nob_t; ++ob) {
            float4 x[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (2 * ob + h < p.nkb_t) {
                mbar_wait(smem_u32(&bars->a_full[sa]), pa);
                const uint32_t row = sA + sa * BOX + (uint32_t)r * 128u;
#pragma unroll
                for (int c = 0; c < 8; ++c) x[h * 8 + c] = lds128(row + (((uint32_t)c ^ rsw) << 4));
                mbar_arrive(smem_u32(&bars->a_empty[sa]));
                if (++sa == (uint32_t)p.S) { sa = 0; pa ^= 1u; }
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) x[h * 8 + c] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
            mbar_wait(smem_u32(&bars->op_empty[so]), po ^ 1u);
            const uint32_t orow = sOP + so * BOX + (uint32_t)r * 128u;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              uint4 t;
              __half2* hh = reinterpret_cast<__half2*>(&t);
              hh[0] = f2h2_sat(x[2 * c8].x, x[2 * c8].y); hh[1] = f2h2_sat(x[2 * c8].z, x[2 * c8].w);
              hh[2] = f2h2_sat(x[2 * c8 + 1].x, x[2 * c8 + 1].y); hh[3] = f2h2_sat(x[2 * c8 + 1].z, x[2 * c8 + 1].w);
              sts128u(orow + (((uint32_t)c8 ^ rsw) << 4), t);
            }
            fence_async_smem();
            mbar_arrive(smem_u32(&bars->op_ready[so]));
            if (++so == (uint32_t)p.SOP) { so = 0; po ^= 1u; }
          }
        }
      }
    }
  } else {
    // =============================== epilogue: PixelUnshuffle / PixelShuffle scatter ===============================
    const int q = warp;
    const int ngroups = (ncur + 31) / 32;
    uint32_t j = 0;
    for (TileIter ti(p); ti.valid(); ti.next(), ++j) {
      const int b = ti.img();
      const int y = ti.y0() + 2 * q + (lane >> 4), x = ti.x0() + (lane & 15);     // TMEM lane 32q + lane == patch pixel
      const bool live = y < p.H && x < p.W && (!p.patch || (lane & 15) < PTW);
      const uint32_t slot = p.nacc == 2 ? (j & 1u) : 0u;
      const uint32_t use = p.nacc == 2 ? (j >> 1) : j;
      mbar_wait(smem_u32(&bars->acc_full[slot]), use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.acc_stride;
      for (int g = 0; g < ngroups; ++g) {
        float v[32];
        tmem_ld32(tacc + (uint32_t)(g * 32), v);
        tmem_ld_wait();
        if (g == ngroups - 1) { tc_fence_before(); mbar_arrive(smem_u32(&bars->acc_empty[slot])); }
        if (!live) continue;
        const int c0 = n0 + g * 32;
        if (p.o_mode == O_NHWC) {
          // plain rows: this thread's pixel, 32 consecutive channels (one full 128-byte line)
          if (p.bias) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += (c0 + e < p.n_valid) ? __ldg(p.bias + c0 + e) : 0.f;
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          float* dst = p.y + (((long long)b * p.H + y) * p.W + x) * p.ldy + c0;
          if (c0 + 32 <= p.n_valid) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              *reinterpret_cast<float4*>(dst + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c0 + e < p.n_valid) dst[e] = v[e];
          }
        } else if (p.o_mode == O_UNSHUFFLE) {
          // out[c*4 + 2*(y&1) + (x&1), y/2, x/2] = conv[c, y, x]   (restormer.py:176)
          float* dst = p.y + (((long long)b * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * p.ldy + (y & 1) * 2 + (x & 1);
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < p.n_valid) dst[(c0 + e) * 4] = v[e];
        } else {
          // out[c/4, 2y + (c%4)/2, 2x + c%2] = conv[c, y, x]   (restormer.py:186): 8 consecutive output channels per sub-pixel
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float* dst = p.y + (((long long)b * (2 * p.H) + 2 * y + (qd >> 1)) * (2 * p.W) + 2 * x + (qd & 1)) * p.ldy + (c0 >> 2);
            if (c0 + 32 <= p.n_valid) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[qd], v[qd + 4], v[qd + 8], v[qd + 12]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(v[qd + 16], v[qd + 20], v[qd + 24], v[qd + 28]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (c0 + 4 * e + qd < p.n_valid) dst[e] = v[4 * e + qd];
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

struct Conv3Cfg { int nc, nchunks, S, SOP, NW, patch; uint32_t off_a, off_op, off_w, wstage; size_t smem; };

bool configure(int cin, int n, bool half, Conv3Cfg& c) {
  static const bool no_patch = getenv("IRB_NO_CONV_PATCH") != nullptr;       // A/B switch for benchmarks
  c.patch = !half && !no_patch;
  const size_t abox = c.patch ? PBOX : BOX;
  if (cin % 4 != 0 || cin < 4 || n % 16 != 0 || n < 16) return false;
  if (half && cin % 8 != 0) return false;
  const size_t budget = 227 * 1024 - 1024;
  c.nchunks = (n + 255) / 256;
  c.nc = ((n + c.nchunks - 1) / c.nchunks + 31) / 32 * 32;
  c.nchunks = (n + c.nc - 1) / c.nc;
  c.wstage = (uint32_t)((size_t)c.nc * 128 + 1023) / 1024 * 1024;
  c.NW = 3; c.SOP = half ? 2 : 0;
  if (c.patch) {
    // a patch box serves nine weight boxes: three or four patch stages are plenty, and the rest of shared memory goes to the
    // weight ring -- the weight stream is latency-bound (every CTA pulls the same boxes out of L2), so bytes in flight are
    // what counts (three 12 KB stages paced the 192 -> 96 Downsample at 0.4 us per box)
    const size_t rest = budget - HDR;
    int S = 4;
    if (rest < (size_t)S * PBOX + 3 * (size_t)c.wstage) S = 3;
    if (rest < (size_t)S * PBOX + 3 * (size_t)c.wstage) return false;
    c.NW = (int)std::min<size_t>(MAX_W, (rest - (size_t)S * PBOX) / c.wstage);
    c.S = S;
    size_t off = HDR;
    c.off_w = (uint32_t)off; off += (size_t)c.NW * c.wstage;
    c.off_op = (uint32_t)off;
    c.off_a = (uint32_t)off; off += (size_t)c.S * PBOX;
    c.smem = off + 1024;
    return true;
  }
  size_t off = HDR;
  c.off_w = (uint32_t)off; off += (size_t)c.NW * c.wstage;
  c.off_op = (uint32_t)off; off += (size_t)c.SOP * BOX;
  if (off + 3 * abox > budget) return false;
  c.S = (int)std::min<size_t>(MAX_S, (budget - off) / abox);
  c.off_a = (uint32_t)off; off += (size_t)c.S * abox;
  c.smem = off + 1024;
  return true;
}

template <typename TOp>
int launch_inst(const CUtensorMap& tA, const Conv3Params& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(tma_conv3_kernel<TOp>, optin));
  IRB_CUDA(launch_pdl(tma_conv3_kernel<TOp>, grid, dim3(NTHREADS), smem, s, tA, p));
  return IR_OK;
}

}  // namespace

bool tma_conv3_supported(int cin, int cout_p, bool half) {
  Conv3Cfg c;
  return configure(cin, cout_p, half, c);
}

// K pitch of the packed weights: per tap, the channels padded to whole operand boxes
int tma_conv3_kpt(int cin, bool half) { const int oc = half ? 64 : 32; return (cin + oc - 1) / oc * oc; }

int launch_conv3_tma(const float* in, int ld_in, int cin, const void* w_packed, const float* bias, int relu, int cout_p,
                     int cout_valid, int B, int H, int W, float* out, int ld_out, int o_mode, bool half, cudaStream_t s) {
  Conv3Cfg c;
  IRB_REQUIRE(configure(cin, cout_p, half, c), "conv3_tma: unsupported shape");
  IRB_REQUIRE(o_mode == O_UNSHUFFLE || o_mode == O_SHUFFLE || o_mode == O_NHWC, "conv3_tma: plain rows or the shuffle scatter");
  IRB_REQUIRE(o_mode == O_NHWC || (bias == nullptr && !relu), "conv3_tma: bias / ReLU belong to the plain-row epilogue");
  IRB_REQUIRE(o_mode != O_UNSHUFFLE || (H % 2 == 0 && W % 2 == 0), "conv3_tma: unshuffle needs even H, W");
  IRB_REQUIRE(ld_in % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && ld_out % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15u) == 0,
              "conv3_tma: 16-byte alignment");
  CUtensorMap tA;
  {
    cuuint64_t d[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)ld_in * 4, (cuuint64_t)ld_in * 4 * W, (cuuint64_t)ld_in * 4 * W * H};
    cuuint32_t box[4] = {32, (cuuint32_t)(c.patch ? PPW : TW), (cuuint32_t)(c.patch ? PPH : TH), 1};
    IRB_TRY(make_tmap(&tA, in, false, 4, d, st, box, true));
  }
  Conv3Params p{};
  p.w = reinterpret_cast<const uint8_t*>(w_packed); p.y = out; p.ldy = ld_out; p.bias = bias; p.relu = relu;
  p.B = B; p.H = H; p.W = W; p.Cin = cin; p.N = cout_p; p.n_valid = cout_valid; p.o_mode = o_mode;
  p.nc = c.nc; p.nkb_t = (cin + 31) / 32; p.nob_t = (cin + 63) / 64;
  p.S = c.S; p.SOP = c.SOP; p.NW = c.NW;
  p.patch = c.patch;
  p.tiles_x = cdiv(W, c.patch ? PTW : TW); p.tiles_y = cdiv(H, TH); p.ntiles = p.tiles_x * p.tiles_y * B;
  p.acc_stride = (c.nc + 31) / 32 * 32;
  p.nacc = 2 * p.acc_stride <= 512 ? 2 : 1;
  int cols = 32; while (cols < p.nacc * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.off_a = c.off_a; p.off_op = c.off_op; p.off_w = c.off_w; p.wstage = c.wstage;
  dim3 grid(std::max(1, std::min(p.ntiles, 148 / c.nchunks)), c.nchunks, 1);
  const size_t smem = std::max<size_t>(c.smem, 120 * 1024);
  const double pix = (double)B * H * W;
  ProfScope prof(TAG_CONV3, pix * 4.0 * (cin + cout_p), 2.0 * pix * 9.0 * cin * cout_p, s);
  return half ? launch_inst<__half>(tA, p, grid, smem, s) : launch_inst<float>(tA, p, grid, smem, s);
}

}  // namespace irb
