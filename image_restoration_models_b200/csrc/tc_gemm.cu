// tcgen05 (5th-gen tensor core) contraction for the 1x1 convolutions of the Restormer block:
//
//     y[pixel, n] = sum_k A(pixel, k) * W[n, k]      M = 128 pixels per tile, N <= 256 per CTA, fp32 accum in TMEM
//
// Replaces (paths relative to /root/reference/src/restormer/restormer.py): Attention.qkv :105 with the
// LayerNorm :37-39/:54-57 fused as a prologue, the attention output (attn@v -> project_out :127-131, folded
// into one per-image matrix) with the residual add :147, FeedForward.project_in :82 (+LayerNorm),
// FeedForward.project_out :86 with the residual add :148, reduce_chan_level{2,3} :223,228 with the channel
// concat :260,:265 as a two-source K loop, and skip_conv :240.
//
// Everything on this path is HBM-bound (K is 48..384, arithmetic intensity 8..40 FLOP/B), so the kernel is a
// streaming pipeline with a tensor-core stage in the middle.  One persistent CTA per SM, 13 warps, three roles:
//
//   producers (8 warps)  coalesced 16-byte global loads of the A rows with several passes in flight, LayerNorm
//                        statistics in registers (two-pass, shuffles), rounding to the operand type, stores into
//                        a ring of smem stages in the UMMA canonical K-major no-swizzle layout
//                        [K/epc][128 (+1 pad row)][16 bytes]; full[s] mbarrier <- all producer threads
//   MMA (1 warp)         one lane issues tcgen05.mma (M=128, N=NC, K=32 bytes per instruction) from the stage and
//                        the CTA-resident weight chunk (staged once, same canonical layout, pre-packed in HBM);
//                        tcgen05.commit -> empty[s] (stage reusable) and tmem_full[a] (tile accumulated)
//   epilogue (4 warps)   tcgen05.ld of the warp's 32 TMEM lanes, transpose through a private padded smem tile,
//                        128-byte-coalesced stores with bias / residual fused; arrive tmem_empty[a]
//
// The accumulator is double buffered in TMEM (2 x NC columns), so the epilogue of tile t overlaps the loads and
// MMAs of tile t+1.
#include "common.cuh"
#include "tc_gemm.cuh"

namespace irb {

namespace {

constexpr int TM = 128;                 // pixels per tile == UMMA M
constexpr int EPI_WARPS = 4;
constexpr int PROD_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int PGROUPS = 1;                      // producer groups (each fills alternate stages); measured on B200: one group of
                                                // 8 warps is faster than two groups of 4 (fewer serial load rounds per stage)
constexpr int PG_THREADS = PROD_THREADS / PGROUPS;
constexpr int NTHREADS = EPI_THREADS + PROD_THREADS + 32;
constexpr int MAX_STAGES = 4;
constexpr int STG_LD = 36;              // floats per staging row (32 + 4 pad -> conflict-free)
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int HDR_BYTES = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// Spin on an mbarrier phase.  A wrong phase would hang the GPU box, so the spin is bounded and traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (it > (1u << 26)) __trap();
  }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: LBO = byte step between the two 16-byte K-chunks of one
  // MMA, SBO = byte step between 8-row groups; version = 1 (sm_100).
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

template <typename TOp>
__device__ __forceinline__ uint32_t make_idesc(int n) {
  // c_format F32 (bits 4-5 = 1); a/b format at bits 7-9 / 10-12: TF32 = 2, F16 = 0; both K-major;
  // N>>3 at bit 17, M>>4 at bit 24
#ifdef IRB_BF16_BUILD
  const uint32_t fmt = sizeof(TOp) == 4 ? 2u : 1u;      // TF32 = 2, BF16 = 1; accumulator F32
#else
  const uint32_t fmt = sizeof(TOp) == 4 ? 2u : 0u;      // TF32 = 2, F16 = 0; accumulator F32
#endif
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

template <typename TOp>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  if constexpr (sizeof(TOp) == 4) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Header {           // first HDR_BYTES of dynamic shared memory
  unsigned long long full[MAX_STAGES];
  unsigned long long empty[MAX_STAGES];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Header) <= HDR_BYTES, "header too large");

// ---- element traits: TA = element type of A in global memory, TOp = tensor-core operand type in smem ---------
template <typename TA> struct GVec;      // one 16-byte global vector
template <> struct GVec<float> { static constexpr int N = 4; };
template <> struct GVec<__half> { static constexpr int N = 8; };

// 16 raw bytes -> GVec<TA>::N floats
template <typename TA>
__device__ __forceinline__ void unpack_vec(const uint4& t, float* out);
template <>
__device__ __forceinline__ void unpack_vec<float>(const uint4& t, float* out) {
  out[0] = __uint_as_float(t.x); out[1] = __uint_as_float(t.y);
  out[2] = __uint_as_float(t.z); out[3] = __uint_as_float(t.w);
}
template <>
__device__ __forceinline__ void unpack_vec<__half>(const uint4& t, float* out) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
}

// store EPC = 16/sizeof(TOp) consecutive K elements of one row as one 16-byte smem chunk
template <typename TOp>
__device__ __forceinline__ void store_chunk(uint8_t* dst, const float* v);
template <>
__device__ __forceinline__ void store_chunk<float>(uint8_t* dst, const float* v) {
  *reinterpret_cast<float4*>(dst) = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
}
template <>
__device__ __forceinline__ void store_chunk<__half>(uint8_t* dst, const float* v) {
  uint4 t;
  __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = f2h2_sat(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(dst) = t;
}

// ---------------------------------------------------------------------------------------------------
// Producer: rows [p0, p0+128) of image b, channels [k0, k0+kc) -> one smem stage [kc/EPC][a_rows_ld][16 B].
// LPP consecutive lanes cooperate on one pixel row; each lane owns UPL "units" (one unit = one smem chunk =
// EPC channels).  UNR row-passes are loaded before any is consumed, to keep enough bytes in flight.
// ---------------------------------------------------------------------------------------------------
// MODE: 0 plain rows (optionally two concatenated sources), 1 LayerNorm prologue, 2 implicit-GEMM 3x3 gather
template <typename TA, typename TOp, int MODE, int UPL, int UNR>
__device__ __forceinline__ void produce_stage(const TcGemmParams& p, uint8_t* __restrict__ sA, int a_rows_ld,
                                              long long rowbase, int p0, int valid, int k0, int kc, int lpp,
                                              bool do_ln, int ptid) {
  constexpr int EPC = 16 / (int)sizeof(TOp);            // elements per smem chunk
  constexpr int VPU = EPC / GVec<TA>::N;                // global vectors per unit (1 or 2)
  static_assert(VPU >= 1, "operand type must not be wider than the global type");
  const TA* a1 = reinterpret_cast<const TA*>(p.a1);
  const TA* a2 = reinterpret_cast<const TA*>(p.a2);
  const int q = ptid % lpp;                             // ptid: thread index inside the producer group
  const int rsub = ptid / lpp;
  const int pp = PG_THREADS / lpp;                      // pixel rows per pass
  const int units = kc / EPC;                           // units per row in this chunk
  // implicit-GEMM 3x3: the unit's K range lies inside one filter tap (cin % EPC == 0)
  int udy[UPL], udx[UPL], uc[UPL];
  if constexpr (MODE == 2) {
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      const int k = k0 + (q + i * lpp) * EPC;
      const int tap = k / p.k1;
      uc[i] = k - tap * p.k1;
      udy[i] = tap / 3 - 1;
      udx[i] = tap - (tap / 3) * 3 - 1;
    }
  }
  // LayerNorm affine parameters of this lane's channels: loaded once per stage, not per row
  float gw[UPL][EPC], gb[UPL][EPC];
  if constexpr (MODE == 1) {
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      const int unit = q + i * lpp;
#pragma unroll
      for (int e4 = 0; e4 < EPC; e4 += 4) {
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (unit < units) {
          g4 = __ldg(reinterpret_cast<const float4*>(p.ln_w + unit * EPC + e4));
          if (p.ln_mode == LN_WITHBIAS) b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + unit * EPC + e4));
        }
        gw[i][e4] = g4.x; gw[i][e4 + 1] = g4.y; gw[i][e4 + 2] = g4.z; gw[i][e4 + 3] = g4.w;
        gb[i][e4] = b4.x; gb[i][e4 + 1] = b4.y; gb[i][e4 + 2] = b4.z; gb[i][e4 + 3] = b4.w;
      }
    }
  }
  for (int r0 = 0; r0 < TM; r0 += pp * UNR) {
    // Loads are UNCONDITIONAL (out-of-range rows / taps / units read a clamped, valid address and are zeroed
    // afterwards): a load under `if` makes the compiler funnel every load through one register set, which
    // serialises them (one outstanding load per thread).  Raw 16-byte vectors, converted after all are issued.
    uint4 raw[UNR][UPL][VPU];
    bool ok[UNR][UPL];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int r = r0 + u * pp + rsub;
      const bool live = r < valid;                      // valid <= TM
      const int rc = live ? r : 0;                      // clamped row (row 0 of the tile always exists)
      int py = 0, px = 0;
      if constexpr (MODE == 2) { py = (p0 + rc) / p.W; px = (p0 + rc) - py * p.W; }
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        const int unit = q + i * lpp;
        const int uu = unit < units ? unit : 0;
        bool good = live && unit < units;
        const TA* src;
        if constexpr (MODE == 2) {
          const int yy = py + udy[i], xx = px + udx[i];
          const bool inside = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
          good = good && inside;
          const int yc = inside ? yy : py, xc = inside ? xx : px;
          src = a1 + (rowbase + (long long)yc * p.W + xc) * (long long)p.lda1 + (unit < units ? uc[i] : 0);
        } else {
          const int k = k0 + uu * EPC;
          src = (k < p.k1) ? a1 + (rowbase + p0 + rc) * (long long)p.lda1 + k
                           : a2 + (rowbase + p0 + rc) * (long long)p.lda2 + (k - p.k1);
        }
        ok[u][i] = good;
#pragma unroll
        for (int g = 0; g < VPU; ++g) raw[u][i][g] = __ldg(reinterpret_cast<const uint4*>(src + g * GVec<TA>::N));
      }
    }
    float v[UNR][UPL][EPC];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int i = 0; i < UPL; ++i)
#pragma unroll
        for (int g = 0; g < VPU; ++g) {
          float t[GVec<TA>::N];
          unpack_vec<TA>(raw[u][i][g], t);
#pragma unroll
          for (int e = 0; e < GVec<TA>::N; ++e) v[u][i][g * GVec<TA>::N + e] = ok[u][i] ? t[e] : 0.f;
        }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int r = r0 + u * pp + rsub;
      if constexpr (MODE == 1) {
        // population variance about the mean, eps inside the sqrt (restormer.py:38,55-56); two-pass in registers
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i)
#pragma unroll
          for (int e = 0; e < EPC; ++e) s += v[u][i][e];
        for (int o = lpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv_k = 1.0f / (float)kc;
        const float mu = s * inv_k;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i) {
          if (q + i * lpp < units) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) { const float d = v[u][i][e] - mu; ss = fmaf(d, d, ss); }
          }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss * inv_k + 1e-5f);
        const float sub = (p.ln_mode == LN_WITHBIAS) ? mu : 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i) {
          const int unit = q + i * lpp;
          if (unit < units) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) v[u][i][e] = fmaf((v[u][i][e] - sub) * rstd, gw[i][e], gb[i][e]);
          }
        }
      }
      if (r < TM) {
#pragma unroll
        for (int i = 0; i < UPL; ++i) {
          const int unit = q + i * lpp;
          if (unit < units) store_chunk<TOp>(sA + ((size_t)unit * a_rows_ld + r) * 16, v[u][i]);
        }
      }
    }
  }
}

template <typename TY>
__device__ __forceinline__ void store_out4(TY* dst, const float4& o);
template <>
__device__ __forceinline__ void store_out4<float>(float* dst, const float4& o) { *reinterpret_cast<float4*>(dst) = o; }
template <>
__device__ __forceinline__ void store_out4<__half>(__half* dst, const float4& o) {
  uint2 t;
  __half2* h = reinterpret_cast<__half2*>(&t);
  h[0] = f2h2_sat(o.x, o.y);
  h[1] = f2h2_sat(o.z, o.w);
  *reinterpret_cast<uint2*>(dst) = t;
}

// ---------------------------------------------------------------------------------------------------
// Epilogue helpers.  One warp drains 32 rows x NCOLS columns: TMEM -> registers -> padded smem tile (transpose)
// -> rows of NCOLS*4 contiguous bytes in global memory.  All lane/row arithmetic is compile-time (shifts).
// ---------------------------------------------------------------------------------------------------
template <typename TY>
struct EpiCtx {
  const float* r; int ldr;
  TY* y; int ldy;
  const float* bias;
  long long row0;        // global row of this warp's first TMEM lane
  int rows_valid;        // rows of this warp inside the image (may be <= 0)
  int lane;
  float* stg;            // this warp's private staging tile [32][STG_LD]
  int relu; float sign;  // y = r + sign * act(acc + bias)
  // PixelUnshuffle / PixelShuffle scatter (o_mode != O_NHWC): image index, extent, pixel index of the warp's first row
  int o_mode, b, H, W, pix0, n_valid;
};

// residual of the column group starting at global column n, fetched ahead of use.  Loads are unconditional (rows past
// the image are clamped to the warp's last valid row and never stored), so all CPR loads are in flight together.
template <int NCOLS, typename TY>
__device__ __forceinline__ void fetch_residual(const EpiCtx<TY>& ec, int n, float4* rr) {
  constexpr int CPR = NCOLS / 4, RPI = 32 / CPR;
  const int rsub = ec.lane / CPR, c4 = (ec.lane % CPR) * 4;     // CPR is a power of two: shifts
  const int last = ec.rows_valid - 1;                            // >= 0: warps without valid rows skip the epilogue
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const int row = min(it * RPI + rsub, last);
    rr[it] = *reinterpret_cast<const float4*>(ec.r + (ec.row0 + row) * ec.ldr + n + c4);
  }
}

template <typename TY, int NCOLS, bool HAS_R>
__device__ __forceinline__ void epi_group(const EpiCtx<TY>& ec, uint32_t taddr, int n, const float4* rr) {
  constexpr int CPR = NCOLS / 4, RPI = 32 / CPR;
  float v[32];
  __syncwarp();                           // tcgen05.ld is .sync.aligned: the warp must be converged
  if (NCOLS == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int jj = 0; jj < CPR; ++jj)
    *reinterpret_cast<float4*>(ec.stg + ec.lane * STG_LD + jj * 4) =
        make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
  __syncwarp();
  const int rsub = ec.lane / CPR, c4 = (ec.lane % CPR) * 4;
  const float* sbase = ec.stg + rsub * STG_LD + c4;
  // straight-line code: all smem reads first (distinct registers), then the arithmetic, then the stores
  float4 o[CPR];
#pragma unroll
  for (int it = 0; it < CPR; ++it) o[it] = *reinterpret_cast<const float4*>(sbase + it * RPI * STG_LD);
  if (ec.bias) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(ec.bias + n + c4));
#pragma unroll
    for (int it = 0; it < CPR; ++it) { o[it].x += bb.x; o[it].y += bb.y; o[it].z += bb.z; o[it].w += bb.w; }
  }
  if (ec.relu) {
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
      o[it].x = fmaxf(o[it].x, 0.f); o[it].y = fmaxf(o[it].y, 0.f); o[it].z = fmaxf(o[it].z, 0.f); o[it].w = fmaxf(o[it].w, 0.f);
    }
  }
  if (HAS_R) {
#pragma unroll
    for (int it = 0; it < CPR; ++it) {
      o[it].x = fmaf(ec.sign, o[it].x, rr[it].x); o[it].y = fmaf(ec.sign, o[it].y, rr[it].y);
      o[it].z = fmaf(ec.sign, o[it].z, rr[it].z); o[it].w = fmaf(ec.sign, o[it].w, rr[it].w);
    }
  }
  TY* yrow = ec.y + (ec.row0 + rsub) * ec.ldy + n + c4;
  const long long step = (long long)RPI * ec.ldy;
  if (ec.rows_valid >= 32) {              // whole 32-row block inside the image: no per-row predicate
#pragma unroll
    for (int it = 0; it < CPR; ++it) store_out4<TY>(yrow + it * step, o[it]);
  } else {
#pragma unroll
    for (int it = 0; it < CPR; ++it)
      if (it * RPI + rsub < ec.rows_valid) store_out4<TY>(yrow + it * step, o[it]);
  }
  __syncwarp();
}

// fp16 output, 64 columns at a time: the accumulators are converted to half BEFORE the transpose, so a staged row is
// 128 bytes and every shared-memory read / global store moves 8 channels (half the instructions of the fp32 path).
__device__ __forceinline__ void epi_group_h64(const EpiCtx<__half>& ec, uint32_t taddr, int n) {
  uint4* srow = reinterpret_cast<uint4*>(ec.stg + ec.lane * STG_LD);
  __syncwarp();
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {            // two 32-column halves, one register set
    float v[32];
    tmem_ld32(taddr + hf * 32, v);
    tmem_ld_wait();
    if (ec.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(ec.bias + n + hf * 32 + j);
    }
    if (ec.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      uint4 t;
      __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = f2h2_sat(v[8 * jj + 2 * e], v[8 * jj + 2 * e + 1]);
      srow[hf * 4 + jj] = t;
    }
  }
  __syncwarp();
  const int rsub = ec.lane >> 3, c8 = (ec.lane & 7) * 8;       // 8 chunks of 8 halfs per row, 4 rows per iteration
  const float* sbase = ec.stg + rsub * STG_LD + (ec.lane & 7) * 4;
  uint4 o[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) o[it] = *reinterpret_cast<const uint4*>(sbase + it * 4 * STG_LD);
  __half* yrow = ec.y + (ec.row0 + rsub) * ec.ldy + n + c8;
  const long long step = 4LL * ec.ldy;
  if (ec.rows_valid >= 32) {
#pragma unroll
    for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(yrow + it * step) = o[it];
  } else {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      if (it * 4 + rsub < ec.rows_valid) *reinterpret_cast<uint4*>(yrow + it * step) = o[it];
  }
  __syncwarp();
}

// Same drain, but each conv output channel lands at its PixelUnshuffle / PixelShuffle position
// (restormer.py:176,186).  The scattered 4-byte stores cost little: these convolutions have K = 9*Cin.
template <typename TY, int NCOLS>
__device__ __forceinline__ void epi_group_scatter(const EpiCtx<TY>& ec, uint32_t taddr, int n) {
  constexpr int CPR = NCOLS / 4, RPI = 32 / CPR;
  float v[32];
  __syncwarp();
  if (NCOLS == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int jj = 0; jj < CPR; ++jj)
    *reinterpret_cast<float4*>(ec.stg + ec.lane * STG_LD + jj * 4) =
        make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
  __syncwarp();
  const int rsub = ec.lane / CPR, c4 = (ec.lane % CPR) * 4;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ec.bias) bb = __ldg(reinterpret_cast<const float4*>(ec.bias + n + c4));
  const float* sbase = ec.stg + rsub * STG_LD + c4;
#pragma unroll
  for (int it = 0; it < CPR; ++it) {
    const int row = it * RPI + rsub;
    if (row < ec.rows_valid) {
      const float4 o4 = *reinterpret_cast<const float4*>(sbase + it * RPI * STG_LD);
      float o[4] = {o4.x + bb.x, o4.y + bb.y, o4.z + bb.z, o4.w + bb.w};
      const int pix = ec.pix0 + row;
      const int y = pix / ec.W, x = pix - y * ec.W;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int nn = n + c4 + e;
        if (nn >= ec.n_valid) continue;
        float val = ec.relu ? fmaxf(o[e], 0.f) : o[e];
        long long idx;
        if (ec.o_mode == O_UNSHUFFLE) {
          const long long opix = ((long long)ec.b * (ec.H >> 1) + (y >> 1)) * (ec.W >> 1) + (x >> 1);
          idx = opix * ec.ldy + nn * 4 + (y & 1) * 2 + (x & 1);
        } else {
          const int qd = nn & 3;
          const long long opix = ((long long)ec.b * (2 * ec.H) + 2 * y + (qd >> 1)) * (2 * ec.W) + 2 * x + (qd & 1);
          idx = opix * ec.ldy + (nn >> 2);
        }
        if constexpr (sizeof(TY) == 4) ec.y[idx] = val; else ec.y[idx] = __float2half_rn(val);
      }
    }
  }
  __syncwarp();
}

template <typename TA, typename TOp, typename TY, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(const TcGemmParams p) {   // 13 warps are allocated as 16: 128 registers per thread
  constexpr int EPC = 16 / (int)sizeof(TOp);
  extern __shared__ __align__(128) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  float* stg = reinterpret_cast<float*>(smem + HDR_BYTES);
  uint8_t* sW = smem + HDR_BYTES + STG_BYTES;
  const int a_rows_ld = TM + p.a_pad;                       // rows per 16-byte K-chunk slab of A (pad breaks conflicts)
  const size_t a_stage_bytes = (size_t)(p.KC / EPC) * a_rows_ld * 16;
  // streamed weights: each stage also carries the [KC/EPC][nc][16 B] slice of the weight chunk
  const size_t stage_bytes = a_stage_bytes + (p.w_stream ? (size_t)(p.KC / EPC) * p.NC * 16 : 0);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * p.NC;
  const int nc = min(p.NC, p.N - n0);
  uint8_t* sA0 = sW + (p.w_stream ? 0 : (size_t)p.NC * p.K * sizeof(TOp));

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&hdr->full[s]), PG_THREADS);
      mbar_init(smem_u32(&hdr->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&hdr->tmem_full[a]), 1);
      mbar_init(smem_u32(&hdr->tmem_empty[a]), EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_sync();   // set-up done under the previous kernel's tail; from here on global memory is ours (common.cuh)

  // tile range of this CTA: contiguous; per-image weights -> the CTA stays inside image blockIdx.z
  long long t_begin, t_end;
  if (p.w_bstride != 0) {
    t_begin = (long long)blockIdx.z * p.tiles_per_img + (long long)p.tiles_per_img * blockIdx.x / gridDim.x;
    t_end = (long long)blockIdx.z * p.tiles_per_img + (long long)p.tiles_per_img * (blockIdx.x + 1) / gridDim.x;
  } else {
    t_begin = (long long)p.ntiles * blockIdx.x / gridDim.x;
    t_end = (long long)p.ntiles * (blockIdx.x + 1) / gridDim.x;
  }
  const int nchunks = (p.K + p.KC - 1) / p.KC;
  const uint32_t a_lbo = (uint32_t)a_rows_ld * 16u;
  const uint32_t w_lbo = (uint32_t)nc * 16u;

  if (warp >= EPI_WARPS && warp < EPI_WARPS + PROD_WARPS) {
    // =============================== producers ===============================
    const int ptid = tid - EPI_THREADS;
    const int grp = ptid / PG_THREADS, gtid = ptid - grp * PG_THREADS;
    const uint4* wg_all = reinterpret_cast<const uint4*>(reinterpret_cast<const TOp*>(p.w) +
                                                         (long long)blockIdx.z * p.w_bstride);
    if (!p.w_stream) {
      // stage this CTA's weight chunk once: global [K/EPC][N][16 B] (+ image stride) -> smem, one block per sub-chunk
      // sub-chunk i occupies [K/EPC][ns_i][16 B] at 16-byte offset (K/EPC) * i * NS (earlier sub-chunks are full)
      const int kq_n = p.K / EPC;
      for (int sub = 0; sub < p.nsub; ++sub) {
        const int c_lo = sub * p.NS;
        const int ns = min(p.NS, nc - c_lo);
        if (ns <= 0) break;
        uint4* ws = reinterpret_cast<uint4*>(sW) + (size_t)kq_n * c_lo;
        const int total = kq_n * ns;
        for (int idx = ptid; idx < total; idx += PROD_THREADS) {
          const int kq = idx / ns, n = idx - kq * ns;
          ws[idx] = __ldg(wg_all + (size_t)kq * p.N + n0 + c_lo + n);
        }
      }
      // the first MMA is released by ONE group's arrivals: order every producer's weight stores before them
      fence_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(PROD_THREADS) : "memory");
    }
    const bool do_ln = p.ln_mode != LN_NONE;
    uint32_t item = 0;
    for (long long tile = t_begin; tile < t_end; ++tile) {
      const int b = (int)(tile / p.tiles_per_img);
      const int p0 = (int)(tile - (long long)b * p.tiles_per_img) * TM;
      const int valid = min(TM, p.HW - p0);
      const long long rowbase = (long long)b * p.HW;
      for (int ch = 0; ch < nchunks; ++ch, ++item) {
        if ((int)(item % PGROUPS) != grp) continue;       // the other group's stage
        const int s = item % p.stages;
        const uint32_t ph = (item / p.stages) & 1u;
        mbar_wait(smem_u32(&hdr->empty[s]), ph ^ 1u);
        uint8_t* sA = sA0 + (size_t)s * stage_bytes;
        const int k0 = ch * p.KC, kc = min(p.KC, p.K - k0);
        if constexpr (EPC == 8) {
          // 8-element chunks: at most 2 units per lane and 2 row passes in flight (register budget)
          if (p.unr >= 2) produce_stage<TA, TOp, MODE, 2, 2>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
          else            produce_stage<TA, TOp, MODE, 2, 1>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
        } else if (p.upl <= 2) {
          if (p.unr >= 4)      produce_stage<TA, TOp, MODE, 2, 4>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
          else if (p.unr == 2) produce_stage<TA, TOp, MODE, 2, 2>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
          else                 produce_stage<TA, TOp, MODE, 2, 1>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
        } else {
          if (p.unr >= 4)      produce_stage<TA, TOp, MODE, 3, 4>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
          else if (p.unr == 2) produce_stage<TA, TOp, MODE, 3, 2>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
          else                 produce_stage<TA, TOp, MODE, 3, 1>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln, gtid);
        }
        if (p.w_stream) {
          // weight slice of this K-chunk: global [K/EPC][N][16 B] -> stage [kc/EPC][nc][16 B]; 4 loads in flight
          uint4* ws = reinterpret_cast<uint4*>(sA + a_stage_bytes);
          const uint4* wg = wg_all + (size_t)(k0 / EPC) * p.N + n0;
          const int kqn = kc / EPC;
          for (int n = gtid; n < nc; n += PG_THREADS) {
            for (int kq = 0; kq < kqn; kq += 4) {
              uint4 t[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (kq + u < kqn) t[u] = __ldg(wg + (size_t)(kq + u) * p.N + n);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (kq + u < kqn) ws[(kq + u) * nc + n] = t[u];
            }
          }
        }
        fence_async_smem();               // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(smem_u32(&hdr->full[s]));
      }
    }
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    // =============================== MMA issuer ===============================
    uint32_t item = 0, j = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++j) {
      const uint32_t a = j % (uint32_t)p.nacc;
      mbar_wait(smem_u32(&hdr->tmem_empty[a]), ((j / (uint32_t)p.nacc) & 1u) ^ 1u);
      tc_fence_after();
      for (int ch = 0; ch < nchunks; ++ch, ++item) {
        const int s = item % p.stages;
        mbar_wait(smem_u32(&hdr->full[s]), (item / p.stages) & 1u);
        tc_fence_after();
        if (lane == 0) {
          const int k0 = ch * p.KC, kc = min(p.KC, p.K - k0);
          const uint32_t a_addr = smem_u32(sA0 + (size_t)s * stage_bytes);
          for (int sub = 0; sub < p.nsub; ++sub) {
            const int c_lo = sub * p.NS;
            const int ns = min(p.NS, nc - c_lo);
            if (ns <= 0) break;
            const uint32_t lbo = (uint32_t)ns * 16u;
            const uint32_t w_addr = p.w_stream ? a_addr + (uint32_t)a_stage_bytes
                                               : smem_u32(sW) + (uint32_t)(p.K / EPC) * (uint32_t)c_lo * 16u +
                                                     (uint32_t)(k0 / EPC) * lbo;
            const uint32_t d_addr = tmem_base + a * (uint32_t)p.acc_stride + (uint32_t)(sub * p.sub_stride);
            const uint32_t idesc = make_idesc<TOp>(ns);
            for (int ks = 0; ks < kc / (2 * EPC); ++ks) {
              const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)(2 * ks) * a_lbo, a_lbo, 128);
              const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)(2 * ks) * lbo, lbo, 128);
              umma<TOp>(d_addr, adesc, bdesc, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&hdr->empty[s]));
          if (ch == nchunks - 1) umma_commit(smem_u32(&hdr->tmem_full[a]));
        }
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ===============================
    // warp e owns TMEM lane quarter e (hardware restriction: warp id % 4) == tile rows [32e, 32e+32)
    const int quarter = warp & 3;
    float* mystg = stg + warp * 32 * STG_LD;
    TY* yout = reinterpret_cast<TY*>(p.y);
    uint32_t j = 0;
    for (long long tile = t_begin; tile < t_end; ++tile, ++j) {
      const int b = (int)(tile / p.tiles_per_img);
      const int p0 = (int)(tile - (long long)b * p.tiles_per_img) * TM;
      const int valid = min(TM, p.HW - p0);
      const long long rowbase = (long long)b * p.HW;
      const uint32_t a = j % (uint32_t)p.nacc;
      EpiCtx<TY> ec;
      ec.r = p.r; ec.ldr = p.ldr; ec.y = yout; ec.ldy = p.ldy; ec.bias = p.bias;
      ec.row0 = rowbase + p0 + quarter * 32; ec.rows_valid = valid - quarter * 32; ec.lane = lane; ec.stg = mystg;
      ec.relu = p.relu; ec.sign = p.acc_sign == 0.f ? 1.f : p.acc_sign;
      ec.o_mode = p.o_mode; ec.b = b; ec.H = p.H; ec.W = p.W; ec.pix0 = p0 + quarter * 32;
      ec.n_valid = p.n_valid > 0 ? p.n_valid : p.N;
      // the residual of the tile's first column group does not depend on the MMA: fetch it before waiting
      float4 rr[8];
      const int ns0 = min(p.NS, nc);
      if (p.r && ec.rows_valid > 0 && p.o_mode == O_NHWC) {
        if (ns0 >= 32) fetch_residual<32>(ec, n0, rr); else fetch_residual<16>(ec, n0, rr);
      }
      mbar_wait(smem_u32(&hdr->tmem_full[a]), (j / (uint32_t)p.nacc) & 1u);
      tc_fence_after();
      for (int sub = 0; sub < p.nsub && ec.rows_valid > 0; ++sub) {
        const int c_lo = sub * p.NS;
        const int ns = min(p.NS, nc - c_lo);
        if (ns <= 0) break;
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * (uint32_t)p.acc_stride +
                               (uint32_t)(sub * p.sub_stride);
        const int nb = n0 + c_lo;                  // first global column of this sub-chunk
        const int n32 = ns >> 5;
        const bool tail16 = (ns & 31) != 0;
        if (p.o_mode != O_NHWC) {
          for (int g = 0; g < n32; ++g) epi_group_scatter<TY, 32>(ec, tbase + (uint32_t)(g * 32), nb + g * 32);
          if (tail16) epi_group_scatter<TY, 16>(ec, tbase + (uint32_t)(n32 * 32), nb + n32 * 32);
          continue;
        }
        int g_first = 0;
        if constexpr (sizeof(TY) == 2) {
          if (!p.r && p.ldy % 8 == 0) {         // 64-column fp16 groups; the remainder goes through the generic path
            const int n64 = ns >> 6;
            for (int g = 0; g < n64; ++g) epi_group_h64(ec, tbase + (uint32_t)(g * 64), nb + g * 64);
            g_first = n64 * 2;
          }
        }
        if (p.r && sub > 0) { if (n32 > 0) fetch_residual<32>(ec, nb, rr); else fetch_residual<16>(ec, nb, rr); }
        for (int g = g_first; g < n32; ++g) {
          const int c0 = g * 32;
          if (p.r) {
            epi_group<TY, 32, true>(ec, tbase + (uint32_t)c0, nb + c0, rr);
            if (g + 1 < n32) fetch_residual<32>(ec, nb + c0 + 32, rr);
            else if (tail16) fetch_residual<16>(ec, nb + c0 + 32, rr);
          } else {
            epi_group<TY, 32, false>(ec, tbase + (uint32_t)c0, nb + c0, rr);
          }
        }
        if (tail16) {
          const int c0 = n32 * 32;
          if (p.r) epi_group<TY, 16, true>(ec, tbase + (uint32_t)c0, nb + c0, rr);
          else     epi_group<TY, 16, false>(ec, tbase + (uint32_t)c0, nb + c0, rr);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&hdr->tmem_empty[a]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

}  // namespace

static int next_pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

// Choose the N-chunk, K-chunk, stage count and lane mapping; returns dynamic shared memory bytes (0 if unsupported).
size_t tc_gemm_configure(TcGemmParams& p) {
  const int a_es = p.a_half ? 2 : 4;                    // bytes per A element in global memory
  const int op_es = p.op_half ? 2 : 4;                  // bytes per tensor-core operand element
  if (op_es > a_es) return 0;
  const int epc = 16 / op_es;
  if (p.K <= 0 || p.N <= 0 || p.K % (2 * epc) != 0 || p.N % 16 != 0) return 0;
  if (p.k1 % (16 / a_es) != 0 || (p.k2 != 0 && p.k2 % (16 / a_es) != 0)) return 0;
  if (p.a_mode == 1 ? p.K != 9 * p.k1 : p.k1 + p.k2 != p.K) return 0;
  if (p.k2 != 0 && p.k1 % epc != 0) return 0;           // a smem chunk must not straddle the two sources
  if (p.a_mode == 1 && (p.k2 != 0 || p.k1 % epc != 0 || p.ln_mode != LN_NONE)) return 0;   // 3x3: a chunk stays inside one tap
  const int kc_max = 512 / op_es;                       // K elements per stage: 128 (tf32) / 256 (f16) = 64 KB of A
  p.KC = p.K <= kc_max ? p.K : kc_max / 2;
  if (p.ln_mode != LN_NONE && (p.KC != p.K || p.k2 != 0 || p.a_half)) return 0;   // LayerNorm needs the whole fp32 row
  const int units = p.KC / epc;
  int lpp = 1;
  while (lpp < 32 && (units + lpp - 1) / lpp > (epc == 4 ? 3 : 2)) lpp <<= 1;
  if ((units + lpp - 1) / lpp > 3) return 0;
  p.lpp = lpp;
  p.upl = (units + lpp - 1) / lpp;
  const int pp = PG_THREADS / lpp;
  p.unr = pp >= TM ? 1 : (TM / pp >= 4 ? 4 : TM / pp);
  if (epc == 8 && p.upl > 2) return 0;
  if (epc == 8 && p.unr > 2) p.unr = 2;                 // register budget: unr * upl * epc floats in flight
  const size_t fixed = HDR_BYTES + STG_BYTES;
  const size_t budget = 227 * 1024;
  auto round16 = [](int v) { return (v + 15) / 16 * 16; };
  // (1) CTA-resident weights.  A CTA may own up to 512 output columns as nsub sub-chunks of NS <= 256 (one MMA
  // each): fewer N-chunks means the A tile is read (and normalised) fewer times.  Fewest chunks wins; the pad row of
  // the A slabs is dropped when that saves a chunk.
  int best_chunks = 0, best_pad = p.a_pad, best_ns = 0, best_nsub = 0;
  for (int chunks = 1; chunks <= p.N / 16 && !best_chunks; ++chunks) {
    const int nc_cta = round16((p.N + chunks - 1) / chunks);
    const int nsub = (nc_cta + 255) / 256;
    const int ns = round16((nc_cta + nsub - 1) / nsub);
    if (nsub * ((ns + 31) / 32 * 32) > 512) continue;              // TMEM: one accumulator set must fit
    for (int pad = p.a_pad; pad >= 0 && !best_chunks; --pad) {
      const size_t a_st = (size_t)units * (TM + pad) * 16;
      if (fixed + (size_t)ns * nsub * p.K * op_es + 2 * a_st <= budget) {
        best_chunks = chunks; best_pad = pad; best_ns = ns; best_nsub = nsub;
      }
    }
  }
  // (2) streamed weights (3x3 convolutions, wide 1x1 at the low-resolution levels): every stage carries its own
  // K-slice of the weights, so N-chunks can stay 256 wide however long K is
  int nc_str = 0;
  const size_t a_stage_pad = (size_t)units * (TM + p.a_pad) * 16;
  if (p.K > p.KC) {
    for (int chunks = 1; chunks <= p.N / 16; ++chunks) {
      int nc = round16((p.N + chunks - 1) / chunks);
      if (nc > 256) continue;
      if (fixed + 2 * (a_stage_pad + (size_t)units * nc * 16) <= budget) { nc_str = nc; break; }
    }
  }
  const bool stream = nc_str != 0 && (best_chunks == 0 || best_ns * best_nsub < std::min(p.N, 128));
  if (!stream && !best_chunks) return 0;
  p.w_stream = stream ? 1 : 0;
  if (stream) { p.NS = nc_str; p.nsub = 1; }
  else { p.NS = best_ns; p.nsub = best_nsub; p.a_pad = best_pad; }
  p.NC = p.NS * p.nsub;
  const size_t a_stage = (size_t)units * (TM + p.a_pad) * 16;
  const size_t stage_bytes = a_stage + (stream ? (size_t)units * p.NC * 16 : 0);
  const size_t resident = stream ? 0 : (size_t)p.NC * p.K * op_es;
  int stages = (int)((budget - fixed - resident) / stage_bytes);
  p.stages = std::min(stages, MAX_STAGES);
  p.sub_stride = (p.NS + 31) / 32 * 32;
  p.acc_stride = p.nsub * p.sub_stride;
  p.nacc = 2 * p.acc_stride <= 512 ? 2 : 1;
  p.tmem_cols = next_pow2_cols(p.nacc * p.acc_stride);
  if (p.tmem_cols > 512) return 0;
  return fixed + resident + (size_t)p.stages * stage_bytes;
}

template <typename TA, typename TOp, typename TY, int MODE>
static int launch_one(const TcGemmParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(tc_gemm_kernel<TA, TOp, TY, MODE>, optin));
  IRB_CUDA(launch_pdl(tc_gemm_kernel<TA, TOp, TY, MODE>, grid, dim3(NTHREADS), smem, s, p));
  return IR_OK;
}

// producer mode: LayerNorm needs an fp32 source; the 3x3 gather reads / writes the fp32 streams
template <typename TA, typename TOp, typename TY>
static int launch_typed(const TcGemmParams& p, dim3 grid, size_t smem, cudaStream_t s) {
  if (p.a_mode == 1) {
    if constexpr (sizeof(TA) == 4 && sizeof(TY) == 4) return launch_one<TA, TOp, TY, 2>(p, grid, smem, s);
    else { set_error("invalid argument: tc_gemm 3x3 mode is fp32 in / fp32 out"); return IR_ERR_INVALID; }
  }
  if (p.ln_mode != LN_NONE) {
    if constexpr (sizeof(TA) == 4) return launch_one<TA, TOp, TY, 1>(p, grid, smem, s);
    else { set_error("invalid argument: tc_gemm LayerNorm prologue needs an fp32 source"); return IR_ERR_INVALID; }
  }
  return launch_one<TA, TOp, TY, 0>(p, grid, smem, s);
}

int launch_gemm_tc(TcGemmParams p, cudaStream_t s) {
  size_t smem = tc_gemm_configure(p);
  IRB_REQUIRE(smem != 0, "tc_gemm: unsupported shape");
  IRB_REQUIRE(p.a_mode == 0 || (p.H > 0 && p.W > 0 && p.H * p.W == p.HW), "tc_gemm: 3x3 mode needs H*W == HW");
  IRB_REQUIRE(p.o_mode == O_NHWC || (p.r == nullptr && p.H * p.W == p.HW), "tc_gemm: scatter epilogue takes no residual");
  IRB_REQUIRE(p.o_mode != O_UNSHUFFLE || (p.H % 2 == 0 && p.W % 2 == 0), "tc_gemm: unshuffle needs even H, W");
  const int a_vec = p.a_half ? 8 : 4;
  IRB_REQUIRE(p.lda1 % a_vec == 0 && (p.k2 == 0 || p.lda2 % a_vec == 0) && p.ldy % 4 == 0 &&
                  (p.r == nullptr || p.ldr % 4 == 0),
              "tc_gemm: leading dimensions must keep 16-byte (A) / 4-element (y, r) alignment");
  p.tiles_per_img = cdiv(p.HW, TM);
  p.ntiles = p.tiles_per_img * p.B;
  const int nchunks_n = cdiv(p.N, p.NC);
  // one CTA per SM (TMEM: two accumulators of up to 256 columns): pad smem so that two can never co-reside
  smem = std::max<size_t>(smem, 116 * 1024);
  dim3 grid;
  if (p.w_bstride != 0) {
    int gx = std::max(1, 148 / (nchunks_n * p.B));
    gx = std::min(gx, p.tiles_per_img);
    grid = dim3(gx, nchunks_n, p.B);
  } else {
    int gx = std::max(1, 148 / nchunks_n);
    gx = std::min(gx, p.ntiles);
    grid = dim3(gx, nchunks_n, 1);
  }
  const double rows = (double)p.B * p.HW;
  const double a_es = p.a_half ? 2.0 : 4.0, y_es = p.y_half ? 2.0 : 4.0;
  const double a_elems = p.a_mode == 1 ? p.k1 : p.K;     // 3x3: the nine shifted reads of a pixel hit cache, not HBM
  ProfScope prof(p.tag, rows * (a_elems * a_es + p.N * (y_es + (p.r ? 4.0 : 0.0))), 2.0 * rows * p.N * p.K, s);
  if (!p.a_half && !p.op_half && !p.y_half) return launch_typed<float, float, float>(p, grid, smem, s);
  if (!p.a_half && p.op_half && p.y_half) return launch_typed<float, __half, __half>(p, grid, smem, s);
  if (!p.a_half && p.op_half && !p.y_half) return launch_typed<float, __half, float>(p, grid, smem, s);
  if (p.a_half && p.op_half && !p.y_half) return launch_typed<__half, __half, float>(p, grid, smem, s);
  if (p.a_half && p.op_half && p.y_half) return launch_typed<__half, __half, __half>(p, grid, smem, s);
  IRB_REQUIRE(false, "tc_gemm: unsupported type combination");
  return IR_OK;
}

}  // namespace irb
